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


def test_step2_variant_sharding_is_a_partition(tmp_path):
    """Config 5 shards variants over ranks with no collective: SPAGMMATtest(rank, world) tests one contiguous range per
    rank and the parts, concatenated in rank order, are the single-rank table.  Host logic only: the GPU call is replaced
    by a stub that tags every row with its marker index."""
    sys.path.insert(0, ROOT)
    from saige_gpu_b200 import step2
    gd = os.path.join(ROOT, "tests", "golden")
    p = os.path.join(gd, "step2_100markers")
    n_fam = len(open(p + ".fam").read().splitlines())
    B0 = (n_fam + 3) // 4
    body = np.fromfile(p + ".bed", dtype=np.uint8)[3:]

    class Stub:
        STEP2_COLUMNS = ("tested", "AC_Allele2", "AF_Allele2", "MissingRate", "BETA", "SE", "Tstat", "var", "p.value", "p.value.NA",
                         "Is.SPA", "AF_case", "AF_ctrl", "N_case", "N_ctrl", "N_case_hom", "N_case_het", "N_ctrl_hom", "N_ctrl_het",
                         "var2", "Is.Firth", "Firth.converged")
        calls = []

        def setSAIGEobjInCPP(self, *a):
            pass

        def setFirth(self, *a, **k):
            pass

        def setMaxMACforER(self, *a):
            pass

        def setCondition(self, *a):
            pass

        def mainMarkerInCPP(self, rows, nf, nm, *a):
            flat = np.asarray(rows).reshape(-1)[:nm * B0]
            # recover where this contiguous chunk starts in the file body; keep every third marker "untested"
            m0 = next(k for k in range(100 - nm + 1) if np.array_equal(body[k * B0:(k + nm) * B0], flat))
            out = np.zeros((nm, len(self.STEP2_COLUMNS)))
            for j in range(nm):
                out[j, 0] = 0.0 if (m0 + j) % 3 == 0 else 1.0
                out[j, 1] = m0 + j
            self.calls.append(nm)
            return out

    def run(rank, world, chunk):
        return step2.SPAGMMATtest(Stub(), p + ".bed", p + ".bim", p + ".fam", os.path.join(gd, "example_binary.rda"),
                                  os.path.join(gd, "example_binary.varianceRatio.txt"), chrom="1", LOCO=True,
                                  markers_per_chunk=chunk, rank=rank, world=world)
    full = run(0, 1, 7)
    assert [r["MarkerID"] for r in full] == ["rs%d" % (m + 1) for m in range(100) if m % 3]
    for world in (2, 3, 8):
        parts = [run(r, world, 7) for r in range(world)]
        ids = [x["MarkerID"] for part in parts for x in part]
        assert ids == [r["MarkerID"] for r in full]
        sizes = [len(part) for part in parts]
        assert max(sizes) - min(sizes) <= (100 + world - 1) // world      # contiguous, near-equal ranges
    # streamed to files, no table in memory: the rank files concatenated (header once) are the single-rank file
    def run_file(rank, world, path):
        return step2.SPAGMMATtest(Stub(), p + ".bed", p + ".bim", p + ".fam", os.path.join(gd, "example_binary.rda"),
                                  os.path.join(gd, "example_binary.varianceRatio.txt"), SAIGEOutputFile=path, chrom="1", LOCO=True,
                                  markers_per_chunk=7, rank=rank, world=world, return_rows=False)
    assert run_file(0, 1, str(tmp_path / "all.txt")) == len(full)
    whole = open(str(tmp_path / "all.txt")).read().splitlines()
    assert len(whole) == len(full) + 1 and whole[0].split("\t")[:5] == ["CHR", "POS", "MarkerID", "Allele1", "Allele2"]
    got = [whole[0]]
    for r in range(3):
        assert run_file(r, 3, str(tmp_path / ("part%d.txt" % r))) == len(parts_of(full, r, 3))
        got += open(str(tmp_path / ("part%d.txt" % r))).read().splitlines()[1:]
    assert got == whole
    n = step2.merge_rank_outputs([str(tmp_path / ("part%d.txt" % r)) for r in range(3)], str(tmp_path / "merged.txt"))
    assert n == len(full) and open(str(tmp_path / "merged.txt")).read() == open(str(tmp_path / "all.txt")).read()


def parts_of(full, rank, world, n=100):
    per = (n + world - 1) // world
    lo, hi = rank * per, min(n, (rank + 1) * per)
    return [r for r in full if lo <= int(r["MarkerID"][2:]) - 1 < hi]


def test_rank_map_is_defined_on_raw_markers():
    """A rank stores the QC-passing markers of its raw 1024-marker blocks: the QC indices of all ranks partition 0..M-1,
    stay ascending per rank, and coincide with `local_markers` when every marker passes."""
    from saige_gpu_b200 import sharding
    rng = np.random.default_rng(0)
    qc = rng.uniform(size=10_000) < 0.9
    M = int(qc.sum())
    for world in (1, 2, 3, 8):
        parts = [sharding.local_markers_qc(qc, r, world) for r in range(world)]
        assert all(np.all(np.diff(p) > 0) for p in parts if len(p) > 1)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(M))
        raw = np.nonzero(qc)[0]
        for r, p in enumerate(parts):
            assert np.all((raw[p] // sharding.SHARD_BLOCK) % world == r)
    ones = np.ones(5000, dtype=bool)
    assert np.array_equal(sharding.local_markers_qc(ones, 1, 4), sharding.local_markers(5000, 1, 4))
