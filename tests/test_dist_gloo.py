"""world_size-2 gloo test of the N>1 host logic (frame sharding + the single detection gather); CPU only."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pigo_b200 import dist as pd


def test_shard_range_covers_all_frames_once():
    for n in (0, 1, 7, 8, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = pd.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nframes, cap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = pd.shard_range(nframes, rank, world)
    per = -(-nframes // world)
    rng = np.random.default_rng(100)
    all_counts = rng.integers(0, cap + 1, size=nframes).astype(np.int32)
    all_dets = rng.integers(0, 1000, size=(nframes, cap, 4)).astype(np.int32)
    dets = torch.zeros((per, cap, 4), dtype=torch.int32)
    counts = torch.zeros(per, dtype=torch.int32)
    dets[:hi - lo] = torch.from_numpy(all_dets[lo:hi])
    counts[:hi - lo] = torch.from_numpy(all_counts[lo:hi])
    res = pd.gather_detections(dets, counts, dst=0)
    if rank == 0:
        d, c = pd.merge_gathered(res[0], res[1], nframes)
        q.put((np.array_equal(d.numpy(), all_dets), np.array_equal(c.numpy(), all_counts)))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_restores_single_gpu_order_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok == (True, True)


def _worker_pipeline(rank, world, port, nframes, cap, ncalls, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = pd.shard_range(nframes, rank, world)
    per = -(-nframes // world)
    rng = np.random.default_rng(7)
    all_n = rng.integers(0, cap + 1, size=nframes).astype(np.int32)
    all_f = rng.integers(0, 2000, size=(nframes, cap, 4)).astype(np.int32)
    all_p = rng.integers(-5, 2000, size=(nframes, cap, 2 + ncalls, 4)).astype(np.int32)
    f = torch.zeros((per, cap, 4), dtype=torch.int32)
    n = torch.zeros(per, dtype=torch.int32)
    p = torch.zeros((per, cap, 2 + ncalls, 4), dtype=torch.int32)
    f[:hi - lo] = torch.from_numpy(all_f[lo:hi]); n[:hi - lo] = torch.from_numpy(all_n[lo:hi]); p[:hi - lo] = torch.from_numpy(all_p[lo:hi])
    res = pd.gather_pipeline(f, n, p, dst=0)
    if rank == 0:
        mf, mn, mp_ = pd.merge_pipeline(*res, nframes)
        q.put((np.array_equal(mf.numpy(), all_f), np.array_equal(mn.numpy(), all_n), np.array_equal(mp_.numpy(), all_p)))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_pipeline_gather_restores_frame_order_world2():
    """The landmark gather of BASELINE configs[4] (faces + 2 eyes + 15 landmark points per face slot), world size 2 over gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_pipeline, args=(r, 2, port, 5, 3, 15, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok == (True, True, True)
