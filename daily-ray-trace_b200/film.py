"""Film planes and their combination across sample-sharded ranks.

Every rank renders a disjoint range of global sample indices for ALL pixels, so after rendering each rank holds, per
pixel and wavelength, a partial (count, sum, mean, M2).  The exact combination is Chan et al.'s parallel update:

    n    = sum_g n_g                     sum  = sum_g sum_g               mean = sum / n
    M2   = sum_g [ M2_g + n_g * (mean_g - mean)^2 ]

which needs the global mean before the M2 terms can be reduced: two collectives (all_reduce of sum and count, then a
reduce of the corrected M2 terms), nothing else crosses NVLink.  The arithmetic below is backend-agnostic torch so the
same code runs under gloo on CPU tensors (tests/test_film_merge.py) and under NCCL on the B200s.
"""
import torch
import torch.distributed as dist


class FilmPlanes:
    """The four f32 film planes of include/drt_cuda.h as torch tensors (allocator + NCCL buffers, nothing more)."""

    def __init__(self, width, height, n, device):
        self.width, self.height, self.n = width, height, n
        npix = width * height
        self.sum = torch.zeros(npix, n, dtype=torch.float32, device=device)
        self.filter = torch.zeros(npix, dtype=torch.float32, device=device)
        self.mean = torch.zeros(npix, n, dtype=torch.float32, device=device)
        self.m2 = torch.zeros(npix, n, dtype=torch.float32, device=device)

    def as_drt_film(self):
        from . import cuda
        return cuda.film_from_tensors(self.sum, self.filter, self.mean, self.m2)

    def nbytes(self):
        return sum(t.numel() * 4 for t in (self.sum, self.filter, self.mean, self.m2))


def merge_pair_(dst, src):
    """dst <- dst (+) src for two films over disjoint samples (the arithmetic of drt_cuda_film_merge)."""
    na, nb = dst.filter, src.filter
    nab = na + nb
    wb = torch.where(nab > 0, nb / nab.clamp_min(1e-30), torch.zeros_like(nab)).unsqueeze(1)
    delta = src.mean - dst.mean
    dst.m2 += src.m2 + delta * delta * na.unsqueeze(1) * wb
    dst.mean += delta * wb
    dst.sum += src.sum
    dst.filter += src.filter
    return dst


def merge_distributed_(film, root=0, group=None):
    """In place: after the call rank `root` holds the film of all ranks' samples (sum, filter and mean are valid on
    every rank).  Returns the number of collectives issued."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    n_local = film.filter.clone()
    dist.all_reduce(film.filter, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(film.sum, op=dist.ReduceOp.SUM, group=group)
    n_tot = film.filter.clamp_min(1.0).unsqueeze(1)
    mean_tot = film.sum / n_tot
    delta = film.mean - mean_tot
    film.m2 += n_local.unsqueeze(1) * delta * delta
    dist.reduce(film.m2, dst=root, op=dist.ReduceOp.SUM, group=group)
    film.mean.copy_(mean_tot)
    return 3
