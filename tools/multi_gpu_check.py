"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Every rank loads the bundled 10k-marker set (markers sharded block-cyclically inside the library), then GRM products,
LOCO products, the GRM diagonal, a multi-RHS PCG and a full binary step-1 fit are compared with the CPU oracle."""
import os, sys
import numpy as np
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import oracle as O
from saige_gpu_b200 import SaigeB200, step1

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo", rank=rank, world_size=world)
ids = [SaigeB200.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
g = SaigeB200(device=lrank, rank=rank, world=world, nccl_id=ids[0])
p = os.path.join(ROOT, "tests/golden/grm10k")
bed, N0, M0, chrs = O.read_bed(p)
o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.01, 0.15
o.setgeno(bed, N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15)
g.setgeno(p + ".bed", p + ".bim", p + ".fam", np.arange(1, N0 + 1), np.ones(N0, np.uint8))
def rel(a, b): return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))
res = {}
assert g.M == o.M and 0 < g.Mloc < g.M, (g.M, g.Mloc)
rng = np.random.default_rng(1)
B = rng.normal(size=(N0, 5))
res["crossprod"] = rel(g.getCrossprodMatAndKin(B), o.getCrossprodMatAndKin(B))
res["diag"] = rel(g.get_DiagofKin(), o.get_DiagofKin())
for idx in (0, 1023, 1024, 5000, o.M - 1):
    assert np.array_equal(g.Get_OneSNP_Geno(idx), o.Get_OneSNP_Geno(idx)), idx
chrq = np.array([int(c) for c in chrs])[o.qc_mask]
LOCO, s, e = O.updateChrStartEndIndexVec(chrq)
o.setStartEndIndexVec(s, e); step1.set_loco_ranges(g, chrq)
o.set_Diagof_StdGeno_LOCO(); g.set_Diagof_StdGeno_LOCO()
w = rng.uniform(0.05, 0.25, size=N0); tau = np.array([1.0, 0.4])
for j in range(22):
    if s[j] < 0: continue
    o.setStartEndIndex(s[j], e[j], j); g.setStartEndIndex(s[j], e[j], j)
    res["loco%d" % j] = rel(g.getCrossprodMatAndKin_LOCO(B), o.getCrossprodMatAndKin_LOCO(B))
    res["locodiag%d" % j] = rel(g.getDiagOfSigma_LOCO(w, tau), o.getDiagOfSigma(w, tau, loco=True))
X, it = g.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
Xo, ito = o.pcg_multi(w, tau, B, 500, 1e-5)
assert list(it) == list(ito)
res["pcg"] = rel(X, Xo)
rows = [l.split() for l in open(os.path.join(ROOT, "tests/golden/pheno_1000samples.txt"))]
col = {h: i for i, h in enumerate(rows[0])}
y = np.array([float(r[col["y_binary"]]) for r in rows[1:]])
Xc = np.column_stack([np.ones(N0), [float(r[col["x1"]]) for r in rows[1:]], [float(r[col["x2"]]) for r in rows[1:]]])
probes = step1.ProbeStream(N0, 130, 200)
mo = O.glmmkin_ai_PCG(o, O.glm_fit(y, Xc, O.Binomial), (0, 0), probes.U, trait="binary")
mg = step1.glmmkin_ai_PCG(g, step1.glm_fit(y, Xc, step1.Binomial), probes, trait="binary")
res["tau"] = rel(mg["theta"], mo["theta"]); res["alpha"] = rel(mg["coefficients"], mo["coefficients"])
# dense GRM: block-rows dealt over the ranks, every rank contracts over all marker shards (ncclBroadcast), products
# from the stored matrix end in the same allreduce
Z = np.stack([o.Get_OneSNP_StdGeno(m) for m in range(o.M)], axis=1)
Kor = Z @ Z.T / o.M
info = g.buildDenseGRM()
res["denseK"] = rel(g.getDenseGRMBlock(0, N0, 0, N0), Kor)
res["denseK_window"] = rel(g.getDenseGRMBlock(100, 300, 250, 500), Kor[100:400, 250:750])
g.setGRMMode("dense")
res["dense_product"] = rel(g.getCrossprodMatAndKin(B), Kor @ B)
Xd, itd = g.getPCG1ofSigmaAndVector(w, tau, B, 500, 1e-5, return_iter=True)
g.setGRMMode("packed")
assert list(itd) == list(ito)
res["dense_pcg"] = rel(Xd, Xo)
assert info["stored_bytes"] <= 8 * 128 * 128 * 8 * 9 / 2 and (world == 1 or info["stored_bytes"] < 8 * 128 * 128 * 8 * 9 / 2), info
pcg_keys = ("pcg", "tau", "alpha", "dense_pcg")
worst_mv = max(v for k, v in res.items() if k not in pcg_keys)
ok = worst_mv < 1e-10 and res["dense_pcg"] < 1e-6 and res["pcg"] < 1e-6 and res["tau"] < 1e-6 and res["alpha"] < 1e-6
print("rank %d/%d Mloc=%d worst product err %.2e pcg %.2e tau %.2e alpha %.2e allreduces %d -> %s"
      % (rank, world, g.Mloc, worst_mv, res["pcg"], res["tau"], res["alpha"], g.counters()["n_allreduce"], "OK" if ok else "FAIL"), flush=True)
g.close(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
