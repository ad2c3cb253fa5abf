"""Host logic of the multi-GPU path (sample sharding + film merge), on CPU with gloo, world_size 2.

Each rank accumulates the oracle's per-path spectra of ITS sample range exactly as the device does (sum, count,
Welford mean/M2); merge_distributed_ must then reproduce the film of the whole sample range."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
import oracledriver

film_mod = importlib.import_module("daily-ray-trace_b200.film")

W, H, SPP, DEPTH, SEED = 12, 10, 14, 4, 31


def _partial_film(paths):
    """paths: [npix, k, n] f64 -> FilmPlanes-like object with the accumulation of the k samples in order."""
    npix, k, n = paths.shape
    f = film_mod.FilmPlanes(npix, 1, n, torch.device("cpu"))
    mean = np.zeros((npix, n)); m2 = np.zeros((npix, n))
    for s in range(k):
        delta = paths[:, s] - mean
        mean += delta / (s + 1)
        m2 += delta * (paths[:, s] - mean)
    f.sum.copy_(torch.from_numpy(paths.sum(axis=1)).float())
    f.filter.fill_(float(k))
    f.mean.copy_(torch.from_numpy(mean).float())
    f.m2.copy_(torch.from_numpy(m2).float())
    return f


def _worker(rank, world, port, paths, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = SPP // world
    mine = paths[:, rank * per:(rank + 1) * per]
    f = _partial_film(mine)
    ncoll = film_mod.merge_distributed_(f, root=0)
    if rank == 0:
        out.put((ncoll, f.sum.numpy(), f.filter.numpy(), f.mean.numpy(), f.m2.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _oracle_paths():
    cfg, tables, scene, cam = common.load("cornell_plane_light", W, H, SPP, DEPTH)
    prm = oracledriver.params(W, H, 0, SPP, DEPTH, cfg.pixel_scheme, SEED)
    o_sum, o_avg, o_m2, paths, _ = oracledriver.render_tile(scene, cam, prm, 0, 0, W, H, want_paths=True)
    return paths, o_sum, o_avg, o_m2


def _check(merged, o_sum, o_avg, o_m2):
    s, f, mean, m2 = merged
    n = o_avg.shape[1]
    assert np.array_equal(f, np.full(W * H, SPP, np.float32))
    for got, ref in ((s, o_sum[:, :n]), (mean, o_avg), (m2, o_m2)):
        floor = 1e-5 * max(np.abs(ref).max(), 1e-30)
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), floor)
        assert rel.max() < 5e-5, rel.max()


def test_merge_distributed_gloo_world2():
    paths, o_sum, o_avg, o_m2 = _oracle_paths()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, paths, out)) for r in range(2)]
    for p in procs:
        p.start()
    ncoll, *merged = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ncoll == 3
    _check(merged, o_sum, o_avg, o_m2)


def test_merge_pair_matches_single_pass():
    paths, o_sum, o_avg, o_m2 = _oracle_paths()
    a, b = _partial_film(paths[:, :5]), _partial_film(paths[:, 5:])
    film_mod.merge_pair_(a, b)
    _check((a.sum.numpy(), a.filter.numpy(), a.mean.numpy(), a.m2.numpy()), o_sum, o_avg, o_m2)


def test_single_rank_is_noop():
    f = film_mod.FilmPlanes(2, 2, 3, torch.device("cpu"))
    assert film_mod.merge_distributed_(f) == 0
