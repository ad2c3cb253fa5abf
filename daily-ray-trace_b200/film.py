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


class ShardedFilmGroup:
    """The multi-GPU film exchange with NO host synchronisation and a SHARDED result (csrc/drt_exchange.cu).

    Per step every rank enqueues, on one stream and without waiting for anything:
        render + scatter   each finished pixel goes straight into the staging film of the rank that owns its slice (NVLink)
        signal             epoch -> this rank's arrival word in every owner's flag block
        wait               until all ranks' arrival words in the LOCAL flag block hold the epoch
        merge              the N partial films of the own slice (local memory) -> the own merged slice; the three 8-bit
                           images of the slice -> the root's image buffer (the only thing that travels after the render)
        signal             epoch -> this rank's completion word at the root;  root only: wait for all completion words
    The merged spectral film stays distributed over the owners' slices; read_back() copies a rank's slice into a whole host
    film over the rank's own PCIe link.  Staging films are double-buffered by epoch parity: step e+1 scatters into the other
    buffer while slow owners may still be merging step e, and a rank can only reach step e+2 after every rank finished
    rendering e+1, i.e. after every owner finished merging e.  torch.distributed only exchanges the IPC handles at set-up."""

    ARRIVE, DONE, MERGED, WORDS = 0, 32, 48, 64   # flag block: arrive[parity][rank] at parity*16 + rank, done[rank] at 32 + rank,
                                                  # merged[owner] at 48 + owner (banded steps: owner has merged the whole step)
    BAND_CUTS = (0, 4, 8, 12, 14, 15, 16)         # band boundaries in 1/16 of a slice: ever smaller bands, so that the merge and
                                                  # read-back left exposed after the last render is 1/16 of the slice

    def __init__(self, ctx, width, height, root=0, group=None):
        self.ctx, self.width, self.height, self.root, self.group = ctx, width, height, root, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        npix = width * height
        self.slice, parts = slice_partition(npix, self.world)
        self.p0, self.p1 = parts[self.rank]
        rows = -(-self.slice * self.world // width)
        self.staging = [ctx.film_alloc(width, rows) for _ in range(2)]
        self.mine = ctx.film_alloc(width, max(1, -(-self.slice // width)))       # this rank's merged slice
        self.flags = ctx.buffer_alloc(self.WORDS * 4)
        self.images = ctx.buffer_alloc(3 * npix * 4) if self.rank == root else None
        mine = (ctx.film_ipc_export(self.staging[0]), ctx.film_ipc_export(self.staging[1]), ctx.buffer_ipc_export(self.flags),
                ctx.buffer_ipc_export(self.images) if self.rank == root else None)
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        self._films, self._bufs = [], []
        self.peer_staging = [[], []]
        self.peer_flags = []
        for r in range(self.world):
            if r == self.rank:
                self.peer_staging[0].append(self.staging[0]); self.peer_staging[1].append(self.staging[1])
                self.peer_flags.append(self.flags)
                continue
            for par in range(2):
                f = ctx.film_ipc_open(handles[r][par])
                self._films.append(f)
                self.peer_staging[par].append(f)
            b = ctx.buffer_ipc_open(handles[r][2])
            self._bufs.append(b)
            self.peer_flags.append(b)
        if self.rank == root:
            self.img_base = self.images
        else:
            self.img_base = ctx.buffer_ipc_open(handles[root][3])
            self._bufs.append(self.img_base)
        self.bgra = [self.img_base + i * npix * 4 for i in range(3)]
        self.epoch = 0
        self.steps = 0            # banded steps issued (their parity picks the staging buffer)
        dist.barrier(group=group)

    def step(self, params, stream_ptr=None, mark=None):
        """One render of this rank's samples of every pixel + the exchange, all enqueued on `stream_ptr`.  `mark(name)` (optional)
        is called between the phases so that a caller can record events."""
        self.epoch += 1
        e, par = self.epoch, self.epoch & 1
        ctx = self.ctx
        ctx.render_device_scatter(params, self.peer_staging[par], self.rank, self.slice, stream=stream_ptr)
        if mark: mark("render")
        ctx.flags_signal([f + 4 * (self.ARRIVE + 16 * par + self.rank) for f in self.peer_flags], e, stream=stream_ptr)
        ctx.flags_wait(self.flags + 4 * (self.ARRIVE + 16 * par), self.world, e, stream=stream_ptr)
        if mark: mark("arrived")
        ctx.film_merge_slices_local(self.mine, self.staging[par], self.world, self.slice, self.width, self.height, self.p0, self.p1,
                                    bgra=self.bgra, stream=stream_ptr)
        if mark: mark("merged")
        ctx.flags_signal([self.peer_flags[self.root] + 4 * (self.DONE + self.rank)], e, stream=stream_ptr)
        if self.rank == self.root:
            ctx.flags_wait(self.flags + 4 * self.DONE, self.world, e, stream=stream_ptr)
        if mark: mark("done")
        return 5 + (1 if self.rank == self.root else 0)      # kernels of this library launched by the step

    def step_bands_to_host(self, params, host_film, render_stream_ptr, copy_stream_ptr):
        """One render + exchange + read-back into the whole host film `host_film`, pipelined in bands: the render kernel walks the
        same part of EVERY owner's slice per band (drt_cuda_render_device_scatter_band) on the render stream; on the copy stream
        every owner waits for the band's arrival flags, merges the band's part of its slice and copies it to the host over its own
        PCIe link -- while the next band renders.  Returns (kernels launched, bytes this rank copied to the host).  The caller
        synchronises the copy stream (then the host film holds this rank's slice)."""
        ctx = self.ctx
        self.steps += 1
        k, par = self.steps, self.steps & 1
        launches = 0
        # staging[par] was last written in banded step k - 2: every owner must have merged that step before it is overwritten
        if k > 2:
            ctx.flags_wait(self.flags + 4 * self.MERGED, self.world, k - 2, stream=render_stream_ptr)
            launches += 1
        cuts = self.BAND_CUTS if params.sample_end - params.sample_begin >= 32 and self.slice >= 64 else (0, 16)
        copied = 0
        first = True
        for b in range(len(cuts) - 1):
            b0, b1 = self.slice * cuts[b] // 16, self.slice * cuts[b + 1] // 16
            if b1 <= b0:
                continue
            self.epoch += 1
            e = self.epoch
            if len(cuts) > 2:
                ctx.render_device_scatter_band(params, self.peer_staging[par], self.rank, self.slice, b0, b1, keep_stats=not first, stream=render_stream_ptr)
            else:
                ctx.render_device_scatter(params, self.peer_staging[par], self.rank, self.slice, stream=render_stream_ptr)
            first = False
            ctx.flags_signal([f + 4 * (self.ARRIVE + 16 * par + self.rank) for f in self.peer_flags], e, stream=render_stream_ptr)
            ctx.flags_wait(self.flags + 4 * (self.ARRIVE + 16 * par), self.world, e, stream=copy_stream_ptr)
            q0, q1 = min(self.p1, self.p0 + b0), min(self.p1, self.p0 + b1)
            launches += 3
            if q1 > q0:
                ctx.film_merge_slices_local(self.mine, self.staging[par], self.world, self.slice, self.width, self.height, q0, q1,
                                            bgra=self.bgra, stream=copy_stream_ptr)
                ctx.film_read_slice(self.mine, self.p0, q0, q1, host_film, stream=copy_stream_ptr)
                copied += (q1 - q0) * (3 * ctx.n + 1) * 4
                launches += 1
        ctx.flags_signal([f + 4 * (self.MERGED + self.rank) for f in self.peer_flags], k, stream=copy_stream_ptr)
        return launches + 1, copied

    def read_back(self, host_film, stream_ptr=None):
        """This rank's merged slice -> its place in a whole host film (four stream-ordered copies over this rank's PCIe link)."""
        self.ctx.film_read_slice(self.mine, self.p0, self.p0, self.p1, host_film, stream=stream_ptr)
        return (self.p1 - self.p0) * (3 * self.ctx.n + 1) * 4

    def check(self):
        t = self.ctx.flags_timeouts()
        if t:
            raise RuntimeError(f"rank {self.rank}: {t} flag waits gave up (a peer never arrived)")

    def close(self):
        for f in self._films:
            try:
                self.ctx.film_ipc_close(f)
            except Exception:
                pass
        for b in self._bufs:
            try:
                self.ctx.buffer_ipc_close(b)
            except Exception:
                pass
        for f in self.staging:
            self.ctx.film_free(f)
        self.ctx.film_free(self.mine)
        self.ctx.buffer_free(self.flags)
        if self.images is not None:
            self.ctx.buffer_free(self.images)
