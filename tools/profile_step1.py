"""Host-side time breakdown of a step-1 fit (which ABI calls take the wall time)."""
import os, sys, time, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from saige_gpu_b200 import SaigeB200, synth, step1
N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200_000, 500_000)
g = SaigeB200()
_, t0, t1 = synth.thresholds(M, 1)
g.setminMAFforGRM(0.01); g.setgeno_synth(N, M, 1, t0, t1)
acc = collections.defaultdict(lambda: [0, 0.0])
def wrap(name):
    f = getattr(g, name)
    def w(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); g.sync(); acc[name][0] += 1; acc[name][1] += time.perf_counter() - t; return r
    setattr(g, name, w)
for n in ("getCoefficients", "getAIScore", "fitglmmaiRPCG", "set_Diagof_StdGeno_LOCO", "setStartEndIndex", "Get_OneSNP_StdGeno"):
    wrap(n)
rng = np.random.default_rng(5)
gterm = np.zeros(N)
for m in rng.choice(g.M, 100, replace=False):
    gterm += rng.normal() * g.Get_OneSNP_StdGeno(int(m))
gterm *= 1.2 / gterm.std()
y, _, X = synth.phenotype(N, 1, gterm=gterm)
probes = step1.ProbeStream(N, 70, 200)
fit0 = step1.glm_fit(y, X, step1.Binomial)
loco = step1.set_loco_ranges(g, synth.chromosomes(M)[g.getQCdMarkerIndex()])
g.reset_counters()
t = time.perf_counter(); tim = {}
model = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", timings=tim, LOCO=loco)
tot = time.perf_counter() - t
print("total %.3f s  fit %.3f  loco %.3f  tau %s" % (tot, tim["fit_s"], tim["loco_s"], model["theta"]))
for k, (n, s) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("  %-28s calls %4d  %.3f s  (%.2f ms/call)" % (k, n, s, 1e3 * s / max(n, 1)))
print("  python-side remainder %.3f s" % (tot - sum(v[1] for k, v in acc.items() if k != "Get_OneSNP_StdGeno")))
print(g.counters())
