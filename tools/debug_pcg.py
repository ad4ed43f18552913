import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import oracle as O
from saige_gpu_b200 import SaigeB200
p = os.path.join(ROOT, "tests/golden/grm10k")
bed, N0, M0, _ = O.read_bed(p)
o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.01, 0.15
o.setgeno(bed, N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
g = SaigeB200(); g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15)
g.setgeno(p + ".bed", p + ".bim", p + ".fam", np.arange(1, N0 + 1), np.ones(N0, np.uint8))
rng = np.random.default_rng(9)
w = rng.uniform(0.02, 0.25, size=o.N); tau = np.array([1.0, 0.35])
B = np.column_stack([rng.normal(size=o.N), rng.integers(0, 2, size=o.N) * 2.0 - 1, np.ones(o.N), np.zeros(o.N), 1e-4 * rng.normal(size=o.N)])
def rel(a, b): return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
for eng in ("tensor", "f64"):
    g.set_engine(eng)
    X, it = g.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
    Xo, ito = o.pcg_multi(w, tau, B, 500, 1e-5)
    print(eng, "batch iters", list(it), "oracle", ito, "rel", [rel(X[:, c], Xo[:, c]) for c in range(5)])
    for c in range(5):
        x1, i1 = g.getPCG1ofSigmaAndVector(w, tau, B[:, c], 500, 1e-5, return_iter=True)
        print("   single col", c, "iters", i1, "rel vs oracle", rel(x1, Xo[:, c]), "rel vs batch", rel(x1, X[:, c]))
    for sub in ([0, 1], [0, 2], [0, 4], [0, 3], [1, 0]):
        Xs, its = g.getPCG1ofSigmaAndVector(w, tau, B[:, sub], 500, 1e-5, return_iter=True)
        print("   subset", sub, list(its), [rel(Xs[:, j], Xo[:, c]) for j, c in enumerate(sub)])
