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


def slice_partition(npix, world):
    """Pixel slices of the scattered exchange: (slice, [(p0, p1) per rank]) with slice = ceil(npix / world); rank r owns
    [r * slice, (r + 1) * slice) clipped to the image, so trailing ranks of a ragged image may own fewer pixels or none."""
    per = -(-npix // world)
    return per, [(min(npix, r * per), min(npix, (r + 1) * per)) for r in range(world)]


class PeerFilmGroup:
    """The fused multi-GPU film exchange over peer memory.  Every rank owns a library-allocated STAGING film; all staging films,
    the root's merged film and the root's three BGRA images are mapped into every process through CUDA IPC once.

    scatter=True (default): the render kernel itself does the scatter half -- each rank writes every finished pixel straight
    into the staging film of the rank that owns the pixel's slice (drt_cuda_render_device_scatter: peer stores over NVLink
    that hide under the render), then every rank merges the partial films of its slice from LOCAL memory and writes merged
    planes + images into the root's memory (drt_cuda_film_merge_slices).
    scatter=False: plain render into the local film, then one kernel per rank that reads all ranks' rows of its slice over
    NVLink (drt_cuda_film_merge_many).
    torch.distributed is used only to exchange the 64-byte handles and for the two host barriers around the merge kernel."""

    def __init__(self, ctx, width, height, root=0, group=None, scatter=True):
        self.ctx, self.width, self.height, self.root, self.group, self.scatter = ctx, width, height, root, group, scatter
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        npix = width * height
        self.slice, parts = slice_partition(npix, self.world)      # pixels per owner; staging holds world * slice pixels
        rows = -(-self.slice * self.world // width) if scatter else height
        self.mine = ctx.film_alloc(width, rows)
        self.merged = ctx.film_alloc(width, height) if self.rank == root else None
        self.images = ctx.buffer_alloc(3 * npix * 4) if self.rank == root else None
        handles = [None] * self.world
        dist.all_gather_object(handles, ctx.film_ipc_export(self.mine), group=group)
        rooted = [(ctx.film_ipc_export(self.merged), ctx.buffer_ipc_export(self.images)) if self.rank == root else None]
        dist.broadcast_object_list(rooted, src=root, group=group)
        self._opened = []
        self.films = []
        for r in range(self.world):
            if r == self.rank:
                self.films.append(self.mine)
            else:
                f = ctx.film_ipc_open(handles[r])
                self._opened.append(f)
                self.films.append(f)
        if self.rank == root:
            self.dst, self.img_base = self.merged, self.images
        else:
            self.dst = ctx.film_ipc_open(rooted[0][0])
            self._opened.append(self.dst)
            self.img_base = ctx.buffer_ipc_open(rooted[0][1])
        if scatter:
            self.p0, self.p1 = parts[self.rank]
        else:
            self.p0, self.p1 = self.rank * npix // self.world, (self.rank + 1) * npix // self.world
        self.bgra = [self.img_base + i * npix * 4 for i in range(3)]
        self.timing = None
        dist.barrier(group=group)

    def render(self, params, stream_ptr=None):
        """This rank's samples of every pixel, into the exchange's staging memory."""
        if self.scatter:
            self.ctx.render_device_scatter(params, self.films, self.rank, self.slice, stream=stream_ptr)
        else:
            self.ctx.render_device(params, self.mine, accumulate=False, stream=stream_ptr)

    def merge(self, stream_ptr=None, sync=None):
        """Call after this rank's render was enqueued.  `sync` = callable that waits for this rank's stream.
        self.timing (if set to a list) receives (wait for own render, barrier, merge kernel, barrier) in seconds."""
        import time
        t0 = time.perf_counter()
        sync()
        t1 = time.perf_counter()
        dist.barrier(group=self.group)          # every rank's partial film is complete (and, scattered, has arrived)
        t2 = time.perf_counter()
        if self.scatter:
            self.ctx.film_merge_slices(self.dst, self.mine, self.world, self.slice, self.width, self.height, self.p0, self.p1,
                                       bgra=self.bgra, stream=stream_ptr)
        else:
            self.ctx.film_merge_many(self.dst, self.films, self.width, self.height, self.p0, self.p1, bgra=self.bgra, stream=stream_ptr)
        sync()
        t3 = time.perf_counter()
        dist.barrier(group=self.group)          # the root's merged film and images are complete
        if self.timing is not None:
            self.timing.append((t1 - t0, t2 - t1, t3 - t2, time.perf_counter() - t3))
        return 1

    def close(self):
        for f in self._opened:
            try:
                self.ctx.film_ipc_close(f)
            except Exception:
                pass
        if self.rank != self.root:
            try:
                self.ctx.buffer_ipc_close(self.img_base)
            except Exception:
                pass
        self.ctx.film_free(self.mine)
        if self.merged is not None:
            self.ctx.film_free(self.merged)
            self.ctx.buffer_free(self.images)
