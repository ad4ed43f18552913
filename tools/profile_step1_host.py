"""Where the step-1 wall time goes on the host side: time inside the C-ABI calls (GPU work + transfers) vs the Python mirror of
the R driver between them (IRLS algebra, score-test matrices), for the per-export mirror and for the driver loops run inside the
library (native_loops).  usage: profile_step1_host.py [N M]     (M = 62500 gives one rank's share of the 8-GPU run)"""
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
for n in ("getCoefficients", "getAIScore", "fitglmmaiRPCG", "set_Diagof_StdGeno_LOCO", "setStartEndIndex", "Get_Coef", "Get_Coef_LOCO_all",
          "varianceRatioMarkers", "getSigma_X", "getSigma_G", "Get_OneSNP_Geno", "glmmkin_ai_PCG"):
    wrap(n)
# the bench's phenotype: polygenic liability (h2 ~ 0.3) from 200 causal markers
rngc = np.random.default_rng(SEED + 5)
causal = np.sort(rngc.choice(g.M, size=200, replace=False))
gterm = np.zeros(N)
for m_idx in causal:
    gterm += rngc.normal() * g.Get_OneSNP_StdGeno(int(m_idx))
gterm *= np.sqrt(0.3 / 0.7) * 1.8 / max(gterm.std(), 1e-12)
y, _, X = synth.phenotype(N, SEED, gterm=gterm)
probes = step1.ProbeStream(N, nmax=70, seed=200)
fit0 = step1.glm_fit(y, X, step1.Binomial)
loco = step1.set_loco_ranges(g, synth.chromosomes(M)[g.getQCdMarkerIndex()])
order = np.random.default_rng(SEED + 6).permutation(g.M)[:400]
def show():
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print("   %-28s %4d calls %8.3f s  (%.2f ms per call)" % (k, v[0], v[1], 1e3 * v[1] / v[0]))
taus = {}
for native in ((True,) if os.environ.get("SGB_PROFILE") else (False, "calls", True, False, "calls", True)):
    acc.clear(); tim = {}
    g.reset_counters()
    t = time.perf_counter()
    model = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", timings=tim, LOCO=loco, native_loops=native)
    wall = time.perf_counter() - t
    inside = sum(v[1] for v in acc.values())
    c = g.counters()
    print("native_loops=%s: wall %.3f s (fit %.3f, loco %.3f); inside ABI calls %.3f s; python between calls %.3f s; products %d columns %d "
          "pcg iterations %d h2d %.0f MB d2h %.0f MB" % (native, wall, tim["fit_s"], tim["loco_s"], inside, wall - inside, c["n_crossprod_calls"],
                                                         c["n_crossprod_columns"], c["n_pcg_iterations"], c["bytes_h2d"] / 1e6, c["bytes_d2h"] / 1e6))
    show()
    taus[native] = model["theta"].copy()
    acc.clear()
    t = time.perf_counter()
    vr, lst = step1.extractVarianceRatio(g, model, step1.Binomial, order, native_loops=bool(native))
    wall = time.perf_counter() - t
    print("  variance ratio %.12f from %d markers: wall %.3f s, inside ABI calls %.3f s" % (vr, len(lst), wall, sum(v[1] for v in acc.values())))
    show()
if False in taus:
    print("tau mirror %s native %s abs diff %.2e" % (taus[False], taus[True], np.max(np.abs(taus[False] - taus[True]))))
