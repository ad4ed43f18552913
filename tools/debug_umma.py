import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from saige_gpu_b200 import SaigeB200, synth
N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3000, 5000)
ks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 3, 4, 16, 31]
g = SaigeB200()
_, t0, t1 = synth.thresholds(M, 1)
g.setminMAFforGRM(0.01); g.setgeno_synth(N, M, 1, t0, t1)
rng = np.random.default_rng(0)
for k in ks:
    B = rng.normal(size=(N, k))
    g.set_engine("tensor"); Yt = g.getCrossprodMatAndKin(B)
    g.set_engine("umma"); Yu = g.getCrossprodMatAndKin(B)
    err = np.max(np.abs(Yu - Yt), axis=0) / np.max(np.abs(Yt), axis=0)
    print("k=%d max rel err per column: %s" % (k, np.array2string(err, precision=2)), flush=True)
    if k <= 4 and err.max() > 1e-8:
        print(" sample Yt", Yt[:4, 0], "\n sample Yu", Yu[:4, 0], "\n ratio", (Yu[:6, 0] / Yt[:6, 0]))
