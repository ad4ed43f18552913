"""Multi-GPU parity worker, launched by tests/test_gpu_multi.py under torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py
Every rank loads the bundled 10k-marker set (markers sharded block-cyclically inside the library), then GRM products,
LOCO products, the GRM diagonal, a multi-RHS PCG, a full binary step-1 fit, the dense-GRM build / product / PCG are
compared with the CPU oracle, a larger synthetic set is checked against the oracle at production row lengths, and a
rank-sharded step-2 scan is compared with the single-rank table (BASELINE config 5: variants sharded, no collective)."""
import os, sys
import numpy as np
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ.pop("OMP_NUM_THREADS", None)          # torchrun pins it to 1; the oracle may use the host's cores
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
# the same fit with the R driver loops run inside the library (sgb_get_coef, sgb_get_coef_loco_all, resident probes) and the
# 22 leave-one-chromosome-out refits; then the variance-ratio marker loop (genotype columns of markers other ranks own arrive by
# an allreduce) -- against the mirror of the R loops on the same ranks and against the oracle
mm = step1.glmmkin_ai_PCG(g, step1.glm_fit(y, Xc, step1.Binomial), probes, trait="binary", LOCO=True)
mn = step1.glmmkin_ai_PCG(g, step1.glm_fit(y, Xc, step1.Binomial), probes, trait="binary", LOCO=True, native_loops=True)
g.setProbeStreamFixed(False)
native_same = max([rel(mn[k], mm[k]) for k in ("theta", "coefficients", "fitted_values")] +
                  [rel(a[k], b[k]) for a, b in zip(mn["LOCOResult"], mm["LOCOResult"]) if a["isLOCO"] for k in ("coefficients", "fitted_values", "Y", "cov")])
res["native_tau"] = rel(mn["theta"], mo["theta"]); res["native_alpha"] = rel(mn["coefficients"], mo["coefficients"])
order = np.random.default_rng(1).permutation(o.M)[:200]
vo, lo_ = O.extractVarianceRatio(o, mo, O.Binomial, order)
vm, lm_ = step1.extractVarianceRatio(g, mg, step1.Binomial, order)
vn, ln_ = step1.extractVarianceRatio(g, mg, step1.Binomial, order, native_loops=True)
res["vr"] = rel(vm, vo); res["native_vr"] = rel(vn, vo)
native_same = max(native_same, rel(ln_, lm_)) if len(ln_) == len(lm_) else 1.0
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
# ---- a synthetic set at production row lengths (20k samples x 60k markers), same generator on both sides ----
from saige_gpu_b200 import synth, step2
Ns, Ms, SEED = 20_000, 60_000, 20260117
_, t0, t1 = synth.thresholds(Ms, SEED)
g.setgeno_synth(Ns, Ms, SEED, t0, t1)
beds = O.synth_bed(Ns, Ms, SEED)
os_ = O.OracleGeno(); os_.minMAF, os_.maxMissing = 0.01, 0.15
os_.setgeno(beds, Ns, Ms, np.arange(1, Ns + 1), np.ones(Ns, np.uint8))
assert np.array_equal(g.getAlleleCountVec(), os_.ACVec)
Bs = rng.normal(size=(Ns, 3))
res["synth_crossprod"] = rel(g.getCrossprodMatAndKin(Bs), os_.getCrossprodMatAndKin(Bs))
res["synth_crossprod_k1"] = rel(g.getCrossprodMatAndKin(Bs[:, 0]), os_.getCrossprodMatAndKin(Bs[:, 0]))
res["synth_diag"] = rel(g.get_DiagofKin(), os_.get_DiagofKin())
chrs_s = synth.chromosomes(Ms)[g.getQCdMarkerIndex()]
_, ss, es = O.updateChrStartEndIndexVec(chrs_s)
os_.setStartEndIndexVec(ss, es); step1.set_loco_ranges(g, chrs_s)
for j in (0, 10, 21):
    os_.setStartEndIndex(ss[j], es[j], j); g.setStartEndIndex(ss[j], es[j], j)
    res["synth_loco%d" % j] = rel(g.getCrossprodMatAndKin_LOCO(Bs[:, 0]), os_.getCrossprodMatAndKin_LOCO(Bs[:, 0]))
ws = rng.uniform(0.05, 0.25, size=Ns)
Xs, its = g.getPCG1ofSigmaAndVector(ws, tau, Bs, 500, 1e-5, return_iter=True)
Xso, itso = os_.pcg_multi(ws, tau, Bs, 500, 1e-5)
assert list(its) == list(itso), (its, itso)
res["synth_pcg"] = rel(Xs, Xso)

# ---- sharded ingest with everything the QC path has: 2 % missing calls, a phenotyped subset in shuffled order, a variance-ratio
# hold-out set (its rows live on ONE rank's device and are replicated by an allreduce), 12,345 raw markers over `world` ranks ----
N1, M1 = 1237, 12_345
bed1 = O.synth_bed(N1, M1, seed=77, miss_rate=0.02)
rng1 = np.random.default_rng(4)
keep = np.sort(rng1.choice(N1, size=1001, replace=False))
sub = rng1.permutation(keep) + 1
ind = np.zeros(N1, np.uint8); ind[keep] = 1
vr = np.unique(rng1.integers(0, M1, size=400))
o1 = O.OracleGeno(); o1.minMAF, o1.maxMissing, o1.isVarRatio = 0.06, 0.03, True
o1.setgeno(bed1, N1, M1, sub, ind, vr_rand_idx=vr)
g.setminMAFforGRM(0.06); g.setmaxMissingRateforGRM(0.03); g.setminMAC_VarianceRatio(20, -1, True)
g.reset_counters()
g.setgeno_mem(bed1, N1, M1, sub, ind, vr_rand_idx=vr)
ingest_ok = ((g.N, g.M, g.Mvr) == (o1.N, o1.M, o1.Mvr) and o1.Mvr > 50 and np.array_equal(g.getAlleleCountVec(), o1.ACVec)
             and np.array_equal(g.getQCdMarkerIndex(), o1.qc_mask) and np.array_equal(g.getIndexVec_forVarRatio(), o1.markerIndexVec_forVarRatio)
             and all(np.array_equal(g.Get_OneSNP_Geno(i), o1.Get_OneSNP_Geno(i)) for i in range(0, o1.M, 397))
             and all(np.array_equal(g.Get_OneSNP_Geno_forVarRatio(i), o1.Get_OneSNP_Geno(i, vr=True)) for i in range(0, o1.Mvr, 7)))
# every rank read only its blocks of the file (one pass), not the whole body
ingest_ok = ingest_ok and (world == 1 or g.counters()["bytes_h2d"] < 0.75 * bed1.nbytes)
b1 = rng1.normal(size=o1.N)
res["ingest_crossprod"] = rel(g.getCrossprodMatAndKin(b1), o1.getCrossprodMatAndKin(b1))
g.setminMAC_VarianceRatio(20, -1, False)

# ---- step 2: the rank's contiguous slice of the variants, no collective; slices concatenated == single-rank table ----
gd = os.path.join(ROOT, "tests", "golden")
p2 = os.path.join(gd, "step2_100markers")
def scan(r, w):
    return step2.SPAGMMATtest(g, p2 + ".bed", p2 + ".bim", p2 + ".fam", os.path.join(gd, "example_binary.rda"),
                              os.path.join(gd, "example_binary.varianceRatio.txt"), chrom="1", LOCO=True,
                              markers_per_chunk=7, rank=r, world=w)
mine = scan(rank, world)
parts = [None] * world
dist.all_gather_object(parts, mine)
step2_ok = True
if rank == 0:
    full = scan(0, 1)
    cat = [r for part in parts for r in part]
    step2_ok = len(cat) == len(full) and all(a.keys() == b.keys() and all(
        (a[k] == b[k]) or (isinstance(a[k], float) and isinstance(b[k], float) and np.isnan(a[k]) and np.isnan(b[k]))
        for k in a) for a, b in zip(cat, full))
    assert len(full) > 20

pcg_keys = ("pcg", "tau", "alpha", "dense_pcg", "synth_pcg", "native_tau", "native_alpha", "vr", "native_vr")
worst_mv = max(v for k, v in res.items() if k not in pcg_keys)
ok = (worst_mv < 1e-10 and res["dense_pcg"] < 1e-6 and res["pcg"] < 1e-6 and res["tau"] < 1e-6 and res["alpha"] < 1e-6
      and res["synth_pcg"] < 1e-6 and step2_ok and ingest_ok and native_same < 1e-9
      and max(res["native_tau"], res["native_alpha"], res["vr"], res["native_vr"]) < 1e-6)
print("rank %d/%d Mloc=%d worst product err %.2e pcg %.2e/%.2e tau %.2e alpha %.2e native loops vs mirror %.2e vr %.2e step2 %s ingest %s allreduces %d -> %s"
      % (rank, world, g.Mloc, worst_mv, res["pcg"], res["synth_pcg"], res["tau"], res["alpha"], native_same, res["native_vr"], step2_ok, ingest_ok,
         g.counters()["n_allreduce"], "OK" if ok else "FAIL"), flush=True)
if not ok:
    print({k: v for k, v in res.items() if v > 1e-10}, flush=True)
g.close(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
