"""Where the step-1 wall time goes on the host side: time inside the C-ABI calls (GPU work + transfers) vs the Python mirror of
the R driver between them (IRLS algebra, score-test matrices).  usage: profile_step1_host.py [N M]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from saige_gpu_b200 import SaigeB200, synth, step1
N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200_000, 500_000)
SEED = 20260117
g = SaigeB200()
_, t0, t1 = synth.thresholds(M, SEED)
g.setminMAFforGRM(0.01); g.setgeno_synth(N, M, SEED, t0, t1)
acc = {}
def wrap(name):
    fn = getattr(g, name)
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); acc.setdefault(name, [0, 0.0]); acc[name][0] += 1; acc[name][1] += time.perf_counter() - t
        return r
    setattr(g, name, w)
for n in ("getCoefficients", "getAIScore", "fitglmmaiRPCG", "set_Diagof_StdGeno_LOCO", "setStartEndIndex"):
    wrap(n)
y, _, X = synth.phenotype(N, SEED)
probes = step1.ProbeStream(N, nmax=70, seed=200)
fit0 = step1.glm_fit(y, X, step1.Binomial)
loco = step1.set_loco_ranges(g, synth.chromosomes(M)[g.getQCdMarkerIndex()])
for rep in range(2):
    acc.clear(); tim = {}
    t = time.perf_counter()
    step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", timings=tim, LOCO=loco)
    wall = time.perf_counter() - t
    inside = sum(v[1] for v in acc.values())
    print("run %d: wall %.3f s (fit %.3f, loco %.3f); inside ABI calls %.3f s; python mirror between calls %.3f s" % (rep, wall, tim["fit_s"], tim["loco_s"], inside, wall - inside))
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print("   %-28s %4d calls %8.3f s  (%.2f ms per call)" % (k, v[0], v[1], 1e3 * v[1] / v[0]))
c = g.counters()
print("products", c["n_crossprod_calls"], "columns", c["n_crossprod_columns"], "pcg iterations", c["n_pcg_iterations"])
