"""Dense-GRM timing (BASELINE config 4 shape by default): a bounded sample of the tcgen05 build and, when the whole
matrix fits the time budget, the stored-GRM product.  usage: dense_bench.py N M [limbs] [sample block-rows | full]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from saige_gpu_b200 import SaigeB200, synth
N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100_000, 500_000)
limbs = int(sys.argv[3]) if len(sys.argv) > 3 else 7
what = sys.argv[4] if len(sys.argv) > 4 else "8"
g = SaigeB200()
_, t0, t1 = synth.thresholds(M, 1)
g.setminMAFforGRM(0.01); g.setgeno_synth(N, M, 1, t0, t1)
nbr = (N + 127) // 128
if what != "full":
    n = int(what)
    for first in (max(nbr // 2 - n // 2, 0), max(nbr - n, 0)):
        info = g.bench_dense_build(limbs, first, n)
        print("block-rows [%d, %d) of %d, %d limbs: %.1f ms, %.3e int8 ops -> %.0f TOPS" % (
            first, first + n, nbr, limbs, info["build_ms"], info["int8_ops"], info["int8_ops"] / info["build_ms"] / 1e9))
    full_ops = limbs * 2.0 * M * 128 * 128 * nbr * (nbr + 1) / 2
    print("whole build at that rate: %.2f s (%.3e ops)" % (full_ops / (info["int8_ops"] / info["build_ms"] * 1e3), full_ops))
else:
    t = time.time(); info = g.buildDenseGRM(limbs); wall = time.time() - t
    print("full build: %.1f ms device (%.2f s wall), %.1f GB stored, %.0f TOPS" % (
        info["build_ms"], wall, info["stored_bytes"] / 1e9, info["int8_ops"] / info["build_ms"] / 1e9))
    rng = np.random.default_rng(0)
    for k in (1, 2, 4, 8, 31):
        B = rng.normal(size=(N, k))
        want = g.getCrossprodMatAndKin(B)
        g.setGRMMode("dense")
        got = g.getCrossprodMatAndKin(B)
        ms, _ = g.bench_crossprod_device(k, 3); ms, _ = g.bench_crossprod_device(k, 5)
        g.setGRMMode("packed")
        mp, _ = g.bench_crossprod_device(k, 3); mp, _ = g.bench_crossprod_device(k, 5)
        print("k=%2d stored-GRM product %.3f ms (%.0f GB/s of stored matrix), packed product %.3f ms, max rel diff %.2e" % (
            k, ms.mean(), info["stored_bytes"] * (1 if k <= 2 else (k + 7) // 8) / ms.mean() / 1e6, mp.mean(), np.abs(got - want).max() / np.abs(want).max()))
