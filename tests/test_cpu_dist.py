"""world_size-2 gloo run of the N>1 host logic (stream assignment + the max-over-ranks timing reduction that
bench.py uses under NCCL).  CPU only."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import vbload
    multi = vbload.load_submodule("multi")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = multi.assign_streams(5, world, rank)
    # rank 0 "takes" 1.5 s for 3 streams of 10 s audio, rank 1 takes 2.5 s for 2 streams
    elapsed = 1500.0 if rank == 0 else 2500.0
    agg = multi.aggregate_rtf(10.0 * len(mine), elapsed, dist)
    mx = multi.reduce_max([elapsed, float(rank)], dist)
    out[rank] = (mine, agg, mx)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    (s0, a0, m0), (s1, a1, m1) = out[0], out[1]
    assert s0 == [0, 1, 2] and s1 == [3, 4]                      # contiguous, balanced, disjoint, complete
    assert a0 == a1                                              # every rank sees the same aggregate
    assert a0["audio_s"] == 50.0 and a0["elapsed_ms"] == 2500.0 and abs(a0["rtf"] - 20.0) < 1e-12
    assert m0 == m1 == [2500.0, 1.0]


def test_assign_streams_properties():
    import vbload
    multi = vbload.load_submodule("multi")
    for n in (0, 1, 7, 8, 9, 64):
        for w in (1, 2, 4, 8):
            parts = [multi.assign_streams(n, w, r) for r in range(w)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert multi.reduce_max([1.0, 2.0]) == [1.0, 2.0]            # no process group: identity


def test_shard_plan_properties():
    """Host logic of the sequence-sharded encoder: contiguous 4-aligned ranges, complete, balanced; halo sizes."""
    import vbload
    sh = vbload.load_submodule("sharded")
    for P in (748, 1696, 30196, 180196, 180199):
        for w in (1, 2, 4, 8):
            plan = sh.plan_shards(P, w)
            assert plan[0][0] == 0 and plan[-1][1] == 4 * (P // 4)
            for (a, b), (c, d) in zip(plan, plan[1:]):
                assert b == c
            assert all(a % 4 == 0 and b % 4 == 0 for a, b in plan)
            sizes = [b - a for a, b in plan]
            assert max(sizes) - min(sizes) <= 4
    assert sh.halo_rows(0) == 0 and sh.halo_rows(300) == 300 and sh.halo_rows(22524) == 750
    # 1-hour config: 8 ranks x ~22.5k positions, each far larger than the 750-row window
    assert min(b - a for a, b in sh.plan_shards(180196, 8)) >= 22520


def test_c_shard_plan_equals_the_python_plan():
    """vox_cuda_shard_plan (host C, what vox_cuda_encode_sharded uses on the GPU box) == sharded.plan_shards / halo_rows."""
    import ctypes as C
    import vbload
    sh = vbload.load_submodule("sharded")
    L = vbload.load().lib()
    L.vox_cuda_shard_plan.argtypes = [C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
    for P in (0, 3, 748, 1696, 3196, 30196, 180196, 180199):
        for w in (1, 2, 3, 4, 8):
            plan = sh.plan_shards(P, w)
            for r in range(w):
                a, b, h = C.c_int(), C.c_int(), C.c_int()
                assert L.vox_cuda_shard_plan(P, w, r, C.byref(a), C.byref(b), C.byref(h)) == 0
                assert (a.value, b.value) == plan[r] and h.value == sh.halo_rows(plan[r][0])
    assert L.vox_cuda_shard_plan(100, 2, 2, None, None, None) == -1


def _halo_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import vbload
    sh = vbload.load_submodule("sharded")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = 4 * 900 * world + 8                                     # ~3600 positions per rank: more than one 750-row window
    shards = sh.plan_shards(P, world)
    p0, p1 = shards[rank]
    M, h = p1 - p0, sh.halo_rows(p0)
    # K row of global position g is filled with g, V with -g; halo rows start as garbage
    own = torch.arange(p0, p1, dtype=torch.float32)[:, None].expand(M, 8).contiguous()
    kb = torch.full((h + M, 8), 12345.0); vv = torch.full((h + M, 8), 12345.0)
    kb[h:] = own; vv[h:] = -own
    for _ in range(2):                                           # two "layers": the exchange is repeatable
        sh.exchange_halo(dist, rank, world, shards, kb, vv, h, M)
    a_r = torch.zeros((max((b - a) // 4 for a, b in shards), 4))
    a_r[:M // 4] = torch.arange(p0 // 4, p1 // 4, dtype=torch.float32)[:, None]
    adapter = sh.gather_adapter(dist, world, shards, a_r)
    out[rank] = (kb[:, 0].tolist(), vv[:, 0].tolist(), adapter[:, 0].tolist(), p0, p1, h)
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_and_adapter_gather_three_ranks():
    """The sharded encoder's only data movement, on CPU tensors over gloo: after the exchange every rank's halo area holds
    exactly the 750 positions that precede its shard, and the gathered adapter rows are complete and in order."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 3
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_halo_worker, args=(world, port, out), nprocs=world, join=True)
    total_tokens = None
    for r in range(world):
        k, v, adapter, p0, p1, h = out[r]
        assert h == (0 if r == 0 else 750)
        assert k == [float(g) for g in range(p0 - h, p1)]
        assert v == [-float(g) for g in range(p0 - h, p1)]
        total_tokens = out[world - 1][4] // 4
        assert adapter == [float(t) for t in range(total_tokens)]
