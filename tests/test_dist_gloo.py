"""world_size-2 checks of the multi-GPU decomposition on CPU (gloo): the block-cyclic marker shards of
saige_gpu_b200/sharding.py, summed with one allreduce, reproduce the full GRM product and the LOCO product; the
chromosome range of every shard is contiguous."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from saige_gpu_b200 import sharding
    bed, N0, M0, chrs = O.read_bed(os.path.join(ROOT, "tests", "golden", "grm10k"))
    o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(bed, N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    b = np.random.default_rng(0).normal(size=(N0, 2))
    mine = sharding.local_markers(o.M, rank, world)
    # partial sum over this rank's blocks (the oracle works on contiguous ranges)
    part = np.zeros((N0, 2))
    blocks = np.unique(mine // sharding.SHARD_BLOCK)
    for blk in blocks:
        lo, hi = blk * sharding.SHARD_BLOCK, min((blk + 1) * sharding.SHARD_BLOCK, o.M)
        part += o._crossprod_range(int(lo), int(hi), b)
    # LOCO: leave out global markers [3000, 5200]
    s, e = 3000, 5200
    lo_l, hi_l = sharding.local_range(mine, s, e)
    assert np.all((mine[lo_l:hi_l] >= s) & (mine[lo_l:hi_l] <= e))
    assert (lo_l == 0 or mine[lo_l - 1] < s) and (hi_l == len(mine) or mine[hi_l] > e)
    part_loco = part.copy()
    inside = mine[lo_l:hi_l]
    for blk in np.unique(inside // sharding.SHARD_BLOCK):
        lo = max(blk * sharding.SHARD_BLOCK, s); hi = min((blk + 1) * sharding.SHARD_BLOCK, e + 1)
        part_loco -= o._crossprod_range(int(lo), int(hi), b)
    t = torch.from_numpy(np.concatenate([part.ravel(), part_loco.ravel(), [float(len(mine))]]))
    dist.all_reduce(t)                                    # the one sum-allreduce per product
    tot = t.numpy()
    full = tot[:N0 * 2].reshape(N0, 2) / o.M
    loco = tot[N0 * 2:N0 * 4].reshape(N0, 2) / (o.M - (e - s + 1))
    assert int(tot[-1]) == o.M
    want = o.getCrossprodMatAndKin(b)
    o.setStartEndIndex(s, e, 0)
    want_loco = o.getCrossprodMatAndKin_LOCO(b)
    ok = (np.max(np.abs(full - want)) / np.max(np.abs(want)) < 1e-12 and
          np.max(np.abs(loco - want_loco)) / np.max(np.abs(want_loco)) < 1e-12)
    # control plane used by bench.py: NCCL-id style byte-string broadcast and max over ranks
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    tm = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ok = ok and ids[0] == bytes(range(128)) and float(tm[0]) == float(world)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_marker_sharding_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shards_are_balanced_per_chromosome():
    from saige_gpu_b200 import sharding, synth
    M = 500_000
    chrs = synth.chromosomes(M)
    for world in (2, 4, 8):
        for c in (1, 13, 22):
            per = [int((chrs[sharding.local_markers(M, r, world)] == c).sum()) for r in range(world)]
            assert max(per) - min(per) <= sharding.SHARD_BLOCK
        sizes = [len(sharding.local_markers(M, r, world)) for r in range(world)]
        assert sum(sizes) == M and max(sizes) - min(sizes) <= sharding.SHARD_BLOCK
