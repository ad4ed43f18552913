"""Step-2 throughput: variants/s of the batched score-test + SPA kernel (BASELINE config 5 shape: 200k samples)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import oracle as O
from saige_gpu_b200 import SaigeB200
N, nm = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200_000, 20_000)
flo, fhi = (float(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (0.05, 0.5)     # alt-allele frequency range
spa_cutoff = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
rng = np.random.default_rng(1)
f = np.random.default_rng(4).uniform(flo, fhi, size=nm)
t0 = np.floor((1 - f) ** 2 * 4294967296.0).astype(np.uint64).clip(0, 4294967295).astype(np.uint32)
t1 = np.floor((1 - f * f) * 4294967296.0).astype(np.uint64).clip(0, 4294967295).astype(np.uint32)
bed = np.zeros(((N + 3) // 4) * nm, dtype=np.uint8)
O.lib().orc_synth_bed(O._ptr(bed), N, nm, 4, O._ptr(t0), O._ptr(t1), int(0.005 * 4294967296.0))
X = np.column_stack([np.ones(N), rng.normal(size=(N, 2))])
mu = 1 / (1 + np.exp(-(X @ np.array([-2.2, 0.4, -0.3]) + rng.normal(scale=0.3, size=N))))
y = (rng.uniform(size=N) < mu).astype(np.float64)
mu2 = mu * (1 - mu); res = y - mu
XV = (X * mu2[:, None]).T; XVX = X.T @ XV.T; XVXi = np.linalg.inv(XVX)
M = dict(mu=mu, res=res, mu2=mu2, tau=np.array([1.0, 0.3]), trait="binary", y=y, X=X, XVX=XVX, XXVX_inv=X @ XVXi,
         XVX_inv_XV=(X @ XVXi) * mu2[:, None], S_a=(X * res[:, None]).sum(0))
g = SaigeB200()
g.setSAIGEobjInCPP(M, 0.95, spa_cutoff, np.arange(N, dtype=np.int32))
g.mainMarkerInCPP(bed[: ((N + 3) // 4) * 256], N, 256)
import torch
pinned = torch.from_numpy(bed).pin_memory().numpy()          # the same rows in page-locked memory (no staging copy in the library)
for label, rows in (("pageable rows", bed), ("pinned rows", pinned)):
    for batched in (True, False):
        g.setStep2Batched(batched)
        g.mainMarkerInCPP(rows, N, nm)          # sizes the staging buffers
        t = time.time(); out = g.mainMarkerInCPP(rows, N, nm); dt = time.time() - t
        print("AF in [%g, %g], SPA cutoff %g, %s, %s: " % (flo, fhi, spa_cutoff, label, "batched GEMM sums" if batched else "per-variant kernel"), end="")
        print("N=%d variants=%d : %.3f s -> %.0f variants/s (%.1f GB/s of genotype bytes), SPA-adjusted %d, tested %d"
              % (N, nm, dt, nm / dt, bed.nbytes / dt / 1e9, int(out[:, 10].sum()), int(out[:, 0].sum())))
