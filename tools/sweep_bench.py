"""Quick device-resident timing of k-column GRM products (tuning aid; prints ms per product and per sweep)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from saige_gpu_b200 import SaigeB200, synth
N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200_000, 500_000)
ks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8, 31]
g = SaigeB200(engine=os.environ.get('SGB_ENGINE', 'tensor'))
_, t0, t1 = synth.thresholds(M, 1)
g.setminMAFforGRM(0.01); g.setgeno_synth(N, M, 1, t0, t1)
g.set_rhs_limbs(int(os.environ.get('SGB_DIGITS', '7')))
bytes_sweep = g.Mloc * ((N + 3) // 4)
for k in ks:
    g.bench_crossprod_device(k, 2)
    ms, mk = g.bench_crossprod_device(k, 5)
    s1, s2 = mk[:, 0].mean(), mk[:, 1].mean()
    print("k=%2d  product %8.3f ms  (%.3f ms/col)  sweep1 %7.3f  sweep2 %7.3f  -> %.0f GB/s per sweep-pass, %.2f of 6455"
          % (k, ms.mean(), ms.mean() / k, s1, s2, 2 * bytes_sweep / ((s1 + s2) * 1e-3) / 1e9, 2 * bytes_sweep / ((s1 + s2) * 1e-3) / 1e9 / 6455.3))
