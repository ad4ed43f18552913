#!/usr/bin/env python
"""bench.py -- GRM matvecs/s of the step-1 null-GLMM hot path on synthetic genotypes (BASELINE.json metric).

A "step" is one k=1 GRM.vector product y = K b over the whole marker set (the unit of work of one PCG iteration,
getCrossprodMatAndKin, FG.cpp:1953).  Workload at every GPU count: BASELINE.json configs[2], synthetic
200,000 samples x 500,000 markers (the configuration the metric is quoted on; 2 x 25 GB packed copies fit one
180 GB B200), markers sharded block-cyclically over the ranks => "strong" scaling, one NCCL allreduce per product.

  value   : whole-job matvecs/s with genotypes and vectors resident in HBM, CUDA-event timed on the library's stream
  e2e     : the same through the public C-ABI call with HOST (pinned) vectors, H2D + D2H inside the timed region
  roofline: the dominant kernel (pk2_gemm_kernel, 2 launches per product) against the measured HBM peak
  cpu_baseline / --impl reference : the reference's CPU matvec (fp32, marker loop of parallelCrossProdOpenMP,
            FG.cpp:1576-1598, restated in oracle/saige_oracle.c) on all host cores, on a bounded marker sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N samples, M markers)
    "c3_200kx500k": (200_000, 500_000),
    "c2_50kx500k": (50_000, 500_000),
    "small_20kx50k": (20_000, 50_000),
}
SEED = 20260117
# ||K b||_2 of the fixed e2e probe vector on one GPU (profiles/r02_bench_1gpu.json): every GPU count must reproduce it -- the
# sharded sums differ only in the order of exact integer additions and one fp64 allreduce
Y_NORM2 = {"c3_200kx500k": 528.53549922164}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")] + [time.time()])

    def mark(self):
        """Samples before this call (warm-up: the sampler is started early so that its fork does not land in the timed steps)
        are discarded."""
        self.t_mark = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.rows = [r for r in self.rows if r[-1] >= getattr(self, "t_mark", 0.0)] or self.rows[-1:]
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline_reference(N, M, target_s=12.0, msample=8192, steps=None, warmup=1):
    """The reference's OWN CPU matvec: genoClass + parallelCrossProd (the OpenMP marker loop, SAIGE_fitGLMM_fast.cpp:37-1183,
    1576-1708) compiled unmodified into oracle/_ref/libfg_refcpu.so (oracle/Makefile), fed a PLINK fileset of `msample` synthetic
    markers x all N samples through its own reader, on all host cores.  The full-workload figure is the sample time scaled linearly
    in M (the loop is a sum over markers) and says so."""
    import tempfile
    from oracle import oracle as O
    from oracle import ref_solver as R
    bed = O.synth_bed(N, msample, SEED)
    d = tempfile.mkdtemp(prefix="saige_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    prefix = os.path.join(d, "sample")
    try:
        with open(prefix + ".bed", "wb") as f:
            f.write(bytes([0x6C, 0x1B, 0x01]))
            f.write(np.ascontiguousarray(bed, dtype=np.uint8).tobytes())
        with open(prefix + ".bim", "w") as f:
            f.writelines("1\tsnp%d\t0\t%d\tA\tC\n" % (m, m + 1) for m in range(msample))
        with open(prefix + ".fam", "w") as f:
            f.writelines("f%d i%d 0 0 1 -9\n" % (i, i) for i in range(N))
        del bed
        r = R.RefCPU()
        r.L.fgref_set_num_threads(host_cores())
        cores = r.L.fgref_num_threads()
        t_in = time.time()
        r.setgeno(prefix + ".bed", prefix + ".bim", prefix + ".fam", np.arange(1, N + 1), np.ones(N, np.uint8), minMAF=0.01, maxMissing=0.15)
        t_in = time.time() - t_in
    finally:
        for ext in (".bed", ".bim", ".fam"):
            if os.path.exists(prefix + ext):
                os.unlink(prefix + ext)
        os.rmdir(d)
    b = np.random.default_rng(1).integers(0, 2, N) * 2.0 - 1.0
    for _ in range(max(1, warmup)):
        r.getCrossprodMatAndKin(b)
    times = []
    t0 = time.time()
    while True:
        t1 = time.perf_counter()
        r.getCrossprodMatAndKin(b)
        times.append(time.perf_counter() - t1)
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.time() - t0 > target_s or len(times) >= 50:
            break
    per_sample = float(np.sum(times)) / len(times)
    scale = M / r.M
    return {"value": 1.0 / (per_sample * scale), "unit": "matvecs/s", "cores": int(cores), "kind": "reference",
            "sample": "%d of %d markers x %d samples through the reference's own genoClass + parallelCrossProd (fp32, OpenMP marker loop, "
                      "%d threads; compiled unmodified from SAIGE_fitGLMM_fast.cpp into oracle/_ref/libfg_refcpu.so), %d timed sample steps, "
                      "scaled linearly in M (x%.1f)" % (r.M, M, N, cores, len(times), scale),
            "extrapolated": True, "scale_in_markers": scale, "sample_steps": len(times), "reference_setgeno_s": t_in,
            "s_per_sample_matvec": per_sample, "s_per_sample_matvec_median": float(np.median(times)),
            "s_timed_total": float(np.sum(times))}


def cpu_baseline(N, M, target_s=12.0, msample=8192, steps=None, warmup=1):
    """CPU baseline on a bounded marker sample of the same workload, on ALL host cores (torchrun exports OMP_NUM_THREADS=1 to its
    workers, so the thread count is set explicitly).  kind "reference": the reference's own compiled CPU path when
    oracle/_ref/libfg_refcpu.so is there (it travels with the snapshot); else kind "port": the oracle's fp32 reference-order matvec
    (the marker loop of parallelCrossProdOpenMP, FG.cpp:1576-1598, restated in oracle/saige_oracle.c).  One "sample step" = one
    matvec over `msample` markers x all N samples; the full-workload figure is that time scaled linearly in M and says so."""
    from oracle import ref_solver as R
    if R.available("cpu"):
        # the reference's own PLINK reader is single-threaded (~45 M genotypes/s: genoClass::setGenoObj decodes, counts and re-packs
        # marker by marker), so its marker sample is kept to ~1.6e9 genotypes: ~35 s of ingest before the timed matvecs
        ref_sample = max(1024, min(msample, int(1.6e9 // max(N, 1)) // 1024 * 1024))
        return cpu_baseline_reference(N, M, target_s=target_s, msample=ref_sample, steps=steps, warmup=warmup)
    from oracle import oracle as O
    O.lib().orc_set_num_threads(host_cores())
    cores = O.lib().orc_num_threads()
    bed = O.synth_bed(N, msample, SEED)
    g = O.OracleGeno(mode=O.REF32)
    g.minMAF, g.maxMissing = 0.01, 0.15
    g.setgeno(bed, N, msample, np.arange(1, N + 1), np.ones(N, np.uint8))
    b = np.random.default_rng(1).integers(0, 2, N) * 2.0 - 1.0
    for _ in range(max(1, warmup)):
        g.getCrossprodMatAndKin(b)
    times = []
    t0 = time.time()
    while True:
        t1 = time.perf_counter()
        g.getCrossprodMatAndKin(b)
        times.append(time.perf_counter() - t1)
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.time() - t0 > target_s or len(times) >= 50:
            break
    per_sample = float(np.sum(times)) / len(times)
    scale = M / g.M
    per_matvec = per_sample * scale
    return {"value": 1.0 / per_matvec, "unit": "matvecs/s", "cores": int(cores), "kind": "port",
            "sample": "%d of %d markers x %d samples, fp32 reference-order matvec (oracle ref32 mode, OpenMP, %d threads), "
                      "%d timed sample steps, scaled linearly in M (x%.1f)" % (g.M, M, N, cores, len(times), scale),
            "extrapolated": True, "scale_in_markers": scale, "sample_steps": len(times),
            "s_per_sample_matvec": per_sample, "s_per_sample_matvec_median": float(np.median(times)),
            "s_timed_total": float(np.sum(times))}


def reference_gpu_kernel(N, M, msample=4096):
    """The reference's OWN GPU matvec (gpuSymMatMult::sym_sgemv: dense fp32 standardised A on the device, two
    cublasSgemv, H2D/D2H of the vectors per call -- gpuSymMatMult.cu:246-287), compiled unmodified into oracle/_ref,
    timed on this B200 on a marker sample and scaled linearly in M (both sgemv are bandwidth-bound in the columns)."""
    import ctypes as C
    so = os.path.join(ROOT, "oracle", "_ref", "libgpusymmatmult_ref.so")
    if not os.path.exists(so):
        return None
    from oracle import oracle as O
    ref = C.CDLL(so)
    ref.ref_set_matrix.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p]
    ref.ref_sym_sgemv.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p]
    bed = O.synth_bed(N, msample, SEED)
    o = O.OracleGeno(mode=O.REF32)
    o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(bed, N, msample, np.arange(1, N + 1), np.ones(N, np.uint8))
    A = np.empty((N, o.M), dtype=np.float32, order="F")
    for m in range(o.M):
        A[:, m] = o.Get_OneSNP_StdGeno(m)
    if ref.ref_set_matrix(N, o.M, A.ctypes.data) != 0:
        return None
    x = (np.random.default_rng(1).integers(0, 2, N) * 2.0 - 1.0).astype(np.float32)
    z = np.zeros(N, dtype=np.float32)
    for _ in range(3):
        ref.ref_sym_sgemv(N, x.ctypes.data, z.ctypes.data)
    reps, t0 = 20, time.time()
    for _ in range(reps):
        ref.ref_sym_sgemv(N, x.ctypes.data, z.ctypes.data)
    per = (time.time() - t0) / reps
    ref.ref_free()
    per_full = per * (M / o.M)
    return {"value": 1.0 / per_full, "unit": "matvecs/s", "kind": "reference GPU kernel (gpuSymMatMult.cu, unmodified) on this B200",
            "sample": "%d of %d markers x %d samples dense fp32 (%.1f GB), scaled linearly in M; the full matrix would need %.0f GB"
                      % (o.M, M, N, A.nbytes / 1e9, 4.0 * N * M / 1e9),
            "s_per_sample_matvec": per, "dense_bytes_per_matvec_full": 8.0 * N * M}


def step1_c1_beside(device):
    """BASELINE config 1 (bundled 1000 samples x 10k markers, binary trait): the full step-1 null-GLMM fit on the CPU
    oracle (fp64 numpy + C matvec, all cores) and on the GPU, same probes; wall seconds and the relative tau gap."""
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200, step1
    pre = os.path.join(ROOT, "tests", "golden", "grm10k")
    bed, N0, M0, _ = O.read_bed(pre)
    rows = [l.split() for l in open(os.path.join(ROOT, "tests", "golden", "pheno_1000samples.txt"))]
    col = {h: i for i, h in enumerate(rows[0])}
    y = np.array([float(r[col["y_binary"]]) for r in rows[1:]])
    X = np.column_stack([np.ones(N0), [float(r[col["x1"]]) for r in rows[1:]], [float(r[col["x2"]]) for r in rows[1:]]])
    probes = step1.ProbeStream(N0, 130, 200)
    o = O.OracleGeno(); o.minMAF, o.maxMissing = 0.01, 0.15
    t = time.time()
    o.setgeno(bed, N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    mo = O.glmmkin_ai_PCG(o, O.glm_fit(y, X, O.Binomial), (0, 0), probes.U, trait="binary")
    t_cpu = time.time() - t
    gg = SaigeB200(device=device)
    gg.setminMAFforGRM(0.01); gg.setmaxMissingRateforGRM(0.15)
    t = time.time()
    gg.setgeno(pre + ".bed", pre + ".bim", pre + ".fam", np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    mg = step1.glmmkin_ai_PCG(gg, step1.glm_fit(y, X, step1.Binomial), probes, trait="binary")
    t_gpu = time.time() - t
    gg.close()
    return {"workload": "c1_bundled_1000x10k (9650 markers pass QC), binary trait, incl. genotype load", "cpu_oracle_wall_s": t_cpu,
            "gpu_wall_s": t_gpu, "tau_cpu": float(mo["theta"][1]), "tau_gpu": float(mg["theta"][1]),
            "tau_rel_gap": abs(float(mg["theta"][1]) - float(mo["theta"][1])) / abs(float(mo["theta"][1]))}


def run_reference(args, N, M, rank, world):
    """--impl reference: the reference's CPU matvec on the host cores (rank 0 only; other ranks exit without work).  A step is one
    matvec over a 32,768-marker sample of the workload (all N samples); exactly --warmup + --steps of them run."""
    if rank != 0:
        return
    cb = cpu_baseline(N, M, msample=32768 if N <= 200_000 else 8192, steps=max(1, args.steps), warmup=max(1, args.warmup))
    line = {"impl": "reference", "metric": "grm_matvecs_per_s", "value": cb["value"], "unit": "matvecs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n_samples": N, "n_markers": M, "k": 1},
            "extrapolated": True,
            "ms_per_sample_step": 1e3 * cb["s_per_sample_matvec"], "timed_region_s": cb["s_timed_total"],
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "matvecs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_200kx500k", choices=sorted(WORKLOADS))
    ap.add_argument("--engine", default="tensor", choices=["tensor", "f64"])
    ap.add_argument("--k-batch", type=int, default=31, help="width of the extra batched-product measurement")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-GRM build sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-step1", action="store_true", help="skip the step-1 wall-time extra")
    ap.add_argument("--no-ingest", action="store_true", help="skip the full-size setgeno leg (host-resident .bed in /dev/shm)")
    ap.add_argument("--no-step2", action="store_true", help="skip the step-2 leg")
    ap.add_argument("--c4-full", action="store_true", help="run BASELINE config 4 at its named shape (default when --gpus >= 4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    N, M = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, N, M, rank, world)

    import torch
    import torch.distributed as dist
    if os.environ.get("SGB_PROFILE_RANK0") and int(os.environ.get("RANK", "0")) == 0:
        os.environ["SGB_PROFILE"] = "1"            # library phase timer (stderr) on rank 0 only: a diagnostic run, not a bench value
    from saige_gpu_b200 import SaigeB200, synth, step1

    if world > 1:
        # control plane only (barrier, max over ranks, NCCL id exchange); the data path uses the library's own NCCL
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ids = [SaigeB200.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        g = SaigeB200(device=local_rank, rank=rank, world=world, nccl_id=ids[0], engine=args.engine)
    else:
        g = SaigeB200(device=local_rank, engine=args.engine)
    torch.cuda.set_device(local_rank)

    def new_context():
        """A second library context on the same rank layout (its own NCCL communicator)."""
        if world > 1:
            ids2 = [SaigeB200.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids2, src=0)
            return SaigeB200(device=local_rank, rank=rank, world=world, nccl_id=ids2[0], engine=args.engine)
        return SaigeB200(device=local_rank, engine=args.engine)

    def barrier():
        g.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- workload: synthetic genotypes generated straight into the device store ----
    t_load0 = time.time()
    _, t0, t1 = synth.thresholds(M, SEED)
    g.setminMAFforGRM(0.01)
    g.setmaxMissingRateforGRM(0.15)
    g.setgeno_synth(N, M, SEED, t0, t1)
    t_load = time.time() - t_load0
    Bbytes = (N + 3) // 4
    Mloc = g.Mloc

    # ---- device-resident leg ("value") ----
    W, K = args.warmup, args.steps
    barrier()
    sampler = ClockSampler(local_rank)        # sampled from the first timed step to the end of the e2e leg (all GPU-busy)
    if rank == 0:
        sampler.start()                       # started before the warm-up: the fork of nvidia-smi must not perturb rank 0's first timed step
    g.bench_crossprod_device(1, W)                         # warm-up (also sizes every scratch buffer)
    g.reset_counters()
    barrier()
    sampler.mark()
    ms, mk = g.bench_crossprod_device(1, K)
    barrier()
    launches = g.counters()["n_kernel_launches"]
    total_ms = max_over_ranks(float(ms.sum()))
    value = K / (total_ms * 1e-3)
    # per-step list = max over ranks of every step (one straggler step must be visible, not own the number silently)
    step_list = [max_over_ranks(float(v)) for v in ms]
    step_median = float(np.median(step_list))
    sweep_ms = float(mk.sum()) / (2 * K)                   # average duration of one pk2_gemm launch
    bytes_launch = Mloc * Bbytes                           # algorithmic bytes of one sweep: the packed shard, once
    peak, peak_src = measured_hbm_peak()
    # dram__bytes_read + write of ONE launch of the sweep kernel from the committed ncu --set full capture of this command
    # (profiles/r02_traffic.json names the capture); null when no capture exists for this workload / GPU count
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get("%s@%d" % (args.workload, world))
        if tj and args.engine == "tensor":
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    achieved = bytes_launch / (sweep_ms * 1e-3) / 1e9
    bytes_alg_product = 2 * Mloc * Bbytes + 16 * N + 16 * Mloc * 2

    # ---- batched products (the Hutchinson probes + phenotype ride in ONE multi-vector product) ----
    batched = []
    for kb in sorted(set([4, 8, args.k_batch])):
        row = {"k": kb, "kernel": "pk2_umma_kernel (tcgen05)" if args.engine == "tensor" else args.engine}
        for digits in (7, 5):
            g.set_rhs_limbs(digits)
            g.bench_crossprod_device(kb, 1)
            barrier()
            msb, _ = g.bench_crossprod_device(kb, max(3, K // 2))
            barrier()
            bms = max_over_ranks(float(msb.mean()))
            npad = (digits * kb + 15) // 16 * 16
            alg_bytes = 2 * Mloc * Bbytes + 16 * N * kb + 16 * Mloc * (kb + 1)
            row["digits%d" % digits] = {
                "ms_per_product": bms, "columns_per_s": kb / (bms * 1e-3),
                "hbm_frac_algorithmic": alg_bytes / (bms * 1e-3) / 1e9 / measured_hbm_peak()[0],
                "algorithmic_int8_ops": 2 * 2.0 * Mloc * N * kb,            # one pass, one int8 MAC pair per genotype and column
                "issued_int8_ops": 2 * 2.0 * Mloc * N * npad,               # what the limb split makes the tensor pipe do
                "issued_tops": 2 * 2.0 * Mloc * N * npad / (bms * 1e-3) / 1e12}
        g.set_rhs_limbs(7)
        row["ms_per_product"] = row["digits7"]["ms_per_product"]
        row["note"] = "digits7 = exact 55-bit right-hand sides (bit-identical to the k=1 kernel); digits5 = sgb_set_product_tolerance(1e-10)"
        batched.append(row)

    # ---- end-to-end leg: the public C-ABI call on pinned host vectors ----
    hb = torch.empty(N, dtype=torch.float64).pin_memory()
    hy = torch.empty(N, dtype=torch.float64).pin_memory()
    hb.copy_(torch.from_numpy(np.random.default_rng(3).integers(0, 2, N) * 2.0 - 1.0))
    import ctypes as C
    L, h = g._L, g._h
    pb, py = C.c_void_p(hb.data_ptr()), C.c_void_p(hy.data_ptr())
    for _ in range(W):
        g._ck(L.sgb_get_crossprod_mat_and_kin(h, pb, 1, py))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e2e0 = time.perf_counter()
    for _ in range(K):
        g._ck(L.sgb_get_crossprod_mat_and_kin(h, pb, 1, py))      # returns after the D2H of y completed
    g.sync()
    t_e2e = max_over_ranks(time.perf_counter() - t_e2e0)
    barrier()
    e2e_value = K / t_e2e
    y_norm2 = float(np.sqrt(float((hy.double() ** 2).sum())))      # ||K b||_2 of the fixed probe b: must agree across N
    clocks = sampler.stop() if rank == 0 else None

    # ---- extras on rank 0: CPU baseline beside it, step-1 wall time ----
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(N, M)
        cb["step1_c1"] = step1_c1_beside(local_rank)
        try:
            cb["reference_gpu_kernel"] = reference_gpu_kernel(N, M)
        except Exception as e:                      # the reference class prints and returns codes; never fail the bench on it
            cb["reference_gpu_kernel"] = {"error": str(e)}
    # ---- setgeno at full size from a host-resident .bed: the file body lives in /dev/shm, generated once by all
    # ranks together (device generator, bit-identical to the genotypes of the timed store); every rank then runs the sharded
    # ingest (count pass over its 1/world of the file, int32 allreduce, QC, second read of the rows it owns) ----
    def all_ok(ok):
        """Collective agreement (min over ranks): a leg is skipped on EVERY rank when any rank cannot run it."""
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t[0] > 0.5)

    ingest_info = None
    if not args.no_ingest:
        shm = "/dev/shm/sgb_bench_%s.bed" % os.environ.get("MASTER_PORT", str(os.getpid()))
        nbytes = Bbytes * M
        why, mm, gi = None, None, None
        try:
            if rank == 0:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize < nbytes + (8 << 30):
                    raise RuntimeError("/dev/shm has %.0f GB free, the .bed body needs %.0f" % (st.f_bavail * st.f_frsize / 1e9, nbytes / 1e9))
                with open(shm, "wb") as f:
                    f.truncate(nbytes)
        except Exception as e:
            why = "%s: %s" % (type(e).__name__, e)
        if all_ok(why is None):
            tg = time.time()
            try:
                mm = np.memmap(shm, dtype=np.uint8, mode="r+")
                ma, mb = M * rank // world, M * (rank + 1) // world
                g.synth_bed_rows(N, ma, mb, SEED, t0, t1, 0.0, out=mm[ma * Bbytes:mb * Bbytes])
            except Exception as e:
                why = "%s: %s" % (type(e).__name__, e)
            tg = time.time() - tg
            if all_ok(why is None):
                try:
                    gi = new_context()
                    gi.setminMAFforGRM(0.01); gi.setmaxMissingRateforGRM(0.15)
                    ones = np.ones(N, np.uint8); ids = np.arange(1, N + 1)
                    gi.setgeno_mem(mm, N, min(M, 4096), ids, ones)          # warm-up: allocations, first kernel loads
                    barrier()
                    ti = time.time()
                    gi.setgeno_mem(mm, N, M, ids, ones)
                    gi.sync()
                    ti = max_over_ranks(time.time() - ti)
                    same = bool(gi.M == g.M and np.array_equal(gi.getAlleleCountVec(), g.getAlleleCountVec()))
                    yi = np.asarray(gi.getCrossprodMatAndKin(hb.numpy())).ravel()
                    ingest_info = {"bed_gbytes": nbytes / 1e9, "seconds": ti, "gb_per_s": nbytes / 1e9 / ti, "n_gpus": world,
                                   "h2d_gbytes_this_rank": gi.counters()["bytes_h2d"] / 1e9, "generate_s": tg,
                                   "allele_counts_equal_synth_store": same,
                                   "product_rel_diff_vs_synth_store": float(np.max(np.abs(yi - hy.numpy())) / np.max(np.abs(hy.numpy()))),
                                   "sample": "FULL workload: %d samples x %d markers, host-resident .bed body in /dev/shm (pageable), "
                                             "QC + imputation + re-pack + transpose on the GPU, sharded over %d rank(s)" % (N, M, world)}
                except Exception as e:                       # a failure inside the collective ingest cannot be agreed on; report it
                    why = "%s: %s" % (type(e).__name__, e)
                if gi is not None:
                    gi.close()
        if ingest_info is None:
            ingest_info = {"error": why or "skipped: another rank could not run this leg"}
        del mm
        if world > 1:
            dist.barrier()
        if rank == 0 and os.path.exists(shm):
            os.unlink(shm)

    # ---- BASELINE config 5 row: single-variant score test + SPA, variants sharded over the ranks (no collective): every rank
    # tests its own 65,536 synthetic variants x N samples from pinned host rows through the C ABI ----
    step2_info = None
    if not args.no_step2:
        n2, m2 = N, 131072 if N >= 100_000 else 262144      # 8 ranks x 131,072 = 1.05 M variants of config 5 per timed call
        rng2 = np.random.default_rng(SEED + 9)
        f2 = np.random.default_rng(SEED + 10 + rank).uniform(0.05, 0.5, size=m2)
        s0 = np.floor((1 - f2) ** 2 * 4294967296.0).astype(np.uint64).clip(0, 4294967295).astype(np.uint32)
        s1 = np.floor((1 - f2 * f2) * 4294967296.0).astype(np.uint64).clip(0, 4294967295).astype(np.uint32)
        rows2 = torch.empty(((n2 + 3) // 4) * m2, dtype=torch.uint8).pin_memory()
        g.synth_bed_rows(n2, 0, m2, SEED + 11 + rank, s0, s1, 0.005, out=rows2.numpy())
        X2 = np.column_stack([np.ones(n2), rng2.normal(size=(n2, 2))])
        mu_ = 1 / (1 + np.exp(-(X2 @ np.array([-2.2, 0.4, -0.3]) + rng2.normal(scale=0.3, size=n2))))
        y2 = (rng2.uniform(size=n2) < mu_).astype(np.float64)
        v2 = mu_ * (1 - mu_)
        XVXi = np.linalg.inv(X2.T @ (X2 * v2[:, None]))
        mdl = dict(mu=mu_, res=y2 - mu_, mu2=v2, tau=np.array([1.0, 0.3]), trait="binary", y=y2, X=X2, XVX=X2.T @ (X2 * v2[:, None]),
                   XXVX_inv=X2 @ XVXi, XVX_inv_XV=(X2 @ XVXi) * v2[:, None], S_a=(y2 - mu_) @ X2)
        g2 = SaigeB200(device=local_rank)
        g2.setSAIGEobjInCPP(mdl, 0.95, 2.0, np.arange(n2, dtype=np.int32))
        g2.mainMarkerInCPP(rows2.numpy(), n2, m2)                       # sizes the staging buffers
        barrier()
        t2 = time.time(); o2 = g2.mainMarkerInCPP(rows2.numpy(), n2, m2); t2 = max_over_ranks(time.time() - t2)
        g2.setSAIGEobjInCPP(mdl, 0.95, 1e9, np.arange(n2, dtype=np.int32))      # no variant takes the saddle-point branch
        g2.mainMarkerInCPP(rows2.numpy(), n2, m2)
        barrier()
        t3 = time.time(); g2.mainMarkerInCPP(rows2.numpy(), n2, m2); t3 = max_over_ranks(time.time() - t3)
        step2_info = {"variants_per_s": world * m2 / t2, "gb_per_s_raw_rows": world * rows2.numel() / t2 / 1e9,
                      "variants_per_s_without_spa": world * m2 / t3, "gb_per_s_raw_rows_without_spa": world * rows2.numel() / t3 / 1e9,
                      "spa_adjusted_rank0": int(o2[:, 10].sum()), "variants_per_rank": m2, "n_gpus": world,
                      "sample": "%d samples x %d variants PER RANK (AF ~ U(0.05, 0.5), 0.5%% missing), binary trait, 3 covariate columns, "
                                "SPA cutoff 2 (~5%% of the variants take the saddle-point branch), pinned host PLINK rows; score sums as one "
                                "tensor-engine GEMM per 1 GB chunk, per-variant kernel for flagged variants; 'without_spa' = same rows, cutoff 1e9 "
                                "(PCIe-bound)" % (n2, m2)}
        g2.close()
        del rows2
    step1_info = None
    if not args.no_step1:
        # polygenic liability (h2 ~ 0.3) from 200 causal markers read back through Get_OneSNP_StdGeno
        rngc = np.random.default_rng(SEED + 5)
        causal = np.sort(rngc.choice(g.M, size=200, replace=False))
        gterm = np.zeros(N)
        for m_idx in causal:
            gterm += rngc.normal() * g.Get_OneSNP_StdGeno(int(m_idx))
        gterm *= np.sqrt(0.3 / 0.7) * 1.8 / max(gterm.std(), 1e-12)
        y, _, X = synth.phenotype(N, SEED, gterm=gterm)
        probes = step1.ProbeStream(N, nmax=70, seed=200)
        fit0 = step1.glm_fit(y, X, step1.Binomial)
        # time spent INSIDE the C-ABI calls (device work + the host<->device copies of the exports' arguments and results) vs in
        # the Python mirror of the R driver between them (IRLS algebra, score-test matrices): the latter is the reference's
        # unchanged R code in a deployment
        abi = {"s": 0.0, "calls": 0}
        for name in ("getCoefficients", "getAIScore", "fitglmmaiRPCG", "set_Diagof_StdGeno_LOCO", "glmmkin_ai_PCG"):
            def timed(*a, _f=getattr(g, name), **k):
                t_ = time.perf_counter(); r_ = _f(*a, **k); abi["s"] += time.perf_counter() - t_; abi["calls"] += 1
                return r_
            setattr(g, name, timed)
        order = np.random.default_rng(SEED + 6).permutation(g.M)[:400]

        def run_fit(native):
            """One whole step-1 fit (22 LOCO refits included) + the variance ratio; native = the R driver loops inside the library
            (sgb_glmmkin_ai_pcg, sgb_variance_ratio_markers), else through the per-export mirror of the unchanged R code."""
            abi["s"], abi["calls"] = 0.0, 0
            g.reset_counters()
            barrier()
            ts = time.time()
            tim = {}
            loco = step1.set_loco_ranges(g, synth.chromosomes(M)[g.getQCdMarkerIndex()])   # config 3: LOCO on, 22 chromosomes
            model = step1.glmmkin_ai_PCG(g, fit0, probes, trait="binary", timings=tim, LOCO=loco, native_loops=native)
            g.sync()
            wall = max_over_ranks(time.time() - ts)
            abi_s, abi_calls = abi["s"], abi["calls"]
            c = g.counters()
            # variance ratio (FG.R:2152-2423): markers with MAC >= 20 in a fixed random order, the
            # getSigma_G solves of a round as ONE multi-column PCG; timed apart from the fit
            tv = time.time()
            vr, vr_list = step1.extractVarianceRatio(g, model, step1.Binomial, order, native_loops=native)
            g.sync()
            vr_s = max_over_ranks(time.time() - tv)
            return dict(model=model, wall=wall, tim=tim, abi_s=abi_s, abi_calls=abi_calls, c=c, vr=vr, vr_n=len(vr_list), vr_s=vr_s,
                        loco=loco)

        # the fit is timed on its second run: the first one also pays the library's scratch allocations for these batch widths, the
        # first NCCL collectives of their sizes and whatever state the legs above left on the host (reported as first_call_wall_s)
        rn_first = run_fit(True)
        rn = run_fit(True)
        model, c, tim = rn["model"], rn["c"], rn["tim"]
        step1_info = {"wall_s": rn["wall"], "first_call_wall_s": rn_first["wall"], "load_synth_s": t_load, "tau": [float(v) for v in model["theta"]],
                      "converged": bool(model["converged"]), "outer_iterations": int(model["n_outer"]),
                      "pcg_solves": c["n_pcg_solves"], "pcg_iterations": c["n_pcg_iterations"],
                      "product_columns": c["n_crossprod_columns"], "products": c["n_crossprod_calls"], "LOCO": bool(rn["loco"]),
                      "inside_abi_calls_s": rn["abi_s"], "abi_calls": rn["abi_calls"],
                      "host_between_calls_s": max(0.0, rn["wall"] - rn["abi_s"]),
                      "h2d_mbytes": c["bytes_h2d"] / 1e6, "d2h_mbytes": c["bytes_d2h"] / 1e6,
                      "variance_ratio": float(rn["vr"]), "variance_ratio_markers": int(rn["vr_n"]), "variance_ratio_s": rn["vr_s"],
                      "wall_with_variance_ratio_s": rn["wall"] + rn["vr_s"],
                      "driver": "R driver loops inside the library: glmmkin.ai_PCG_Rcpp_Binary as one call (sgb_glmmkin_ai_pcg: Get_Coef "
                                "IRLS, AI-REML steps, 22 LOCO refits on device-resident vectors, first probe batch resident), "
                                "extractVarianceRatio's marker loop as one call per round (sgb_variance_ratio_markers); the score-test "
                                "matrices of the result list on host threads",
                      "note": "binary trait, 3 fixed-effect columns, nrun=30 probes, tolPCG=1e-5, full GRM, "
                              "22 leave-one-chromosome-out refits included; genotypes already resident (load_synth_s apart); "
                              "wide batches at 7 digits (exact 55-bit right-hand sides)"}
        if ingest_info and "seconds" in ingest_info:
            step1_info["setgeno_s"] = ingest_info["seconds"]
            step1_info["wall_with_setgeno_s"] = rn["wall"] + ingest_info["seconds"]
        # the same fit through the per-export mirror of the UNCHANGED R driver (one ABI call per Rcpp export, the IRLS algebra and the
        # variance-ratio marker loop in the host language between them)
        rm = run_fit(False)
        mm = rm["model"]
        step1_info["r_mirror"] = {"wall_s": rm["wall"], "fit_s": rm["tim"].get("fit_s"), "loco_refits_s": rm["tim"].get("loco_s"),
                                  "inside_abi_calls_s": rm["abi_s"], "abi_calls": rm["abi_calls"],
                                  "host_mirror_between_calls_s": max(0.0, rm["wall"] - rm["abi_s"]),
                                  "h2d_mbytes": rm["c"]["bytes_h2d"] / 1e6, "d2h_mbytes": rm["c"]["bytes_d2h"] / 1e6,
                                  "variance_ratio_s": rm["vr_s"], "tau": [float(v) for v in mm["theta"]],
                                  "pcg_iterations": rm["c"]["n_pcg_iterations"],
                                  "tau_abs_diff_vs_library_loops": float(np.max(np.abs(mm["theta"] - model["theta"]))),
                                  "alpha_rel_diff_vs_library_loops": float(np.max(np.abs(mm["coefficients"] - model["coefficients"]) /
                                                                                  np.abs(mm["coefficients"]))),
                                  "variance_ratio_rel_diff_vs_library_loops": float(abs(rm["vr"] - rn["vr"]) / abs(rm["vr"]))}
        # the same fit with the tolerance-driven digit count of wide batches (sgb_set_product_tolerance(1e-10) = 5 digits)
        g.set_product_tolerance(1e-10)
        r5 = run_fit(True)
        g.set_rhs_limbs(7)
        model5 = r5["model"]
        step1_info["digits5"] = {"wall_s": r5["wall"], "variance_ratio_s": r5["vr_s"],
                                 "tau": [float(v) for v in model5["theta"]],
                                 "tau_rel_diff_vs_digits7": float(abs(model5["theta"][1] - model["theta"][1]) / abs(model["theta"][1])),
                                 "alpha_rel_diff_vs_digits7": float(np.max(np.abs(model5["coefficients"] - model["coefficients"]) /
                                                                           np.abs(model["coefficients"])))}

    dense_info = None
    if not args.no_dense and N < (1 << 18):
        # BASELINE config 4 row: a bounded sample of the dense-GRM build (the last 8 block-rows per rank, i.e. full-height
        # panels) on the tcgen05 int8 path, against 2 x the measured dense bf16 rate (int8 runs at twice bf16)
        nbr = (N + 127) // 128
        nsample = min(nbr, 8 * world)
        barrier()
        info = g.bench_dense_build(7, nbr - nsample, nsample)
        dms = max_over_ranks(float(info["build_ms"]))
        dops = float(info["int8_ops"]) * world
        g.freeDenseGRM()
        tpeak, tsrc = 2 * 2250.0, "nominal (2 x 2.25 PFLOP/s bf16)"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            tpeak, tsrc = 2 * float(mp["bf16_tflops"]), "2 x measured dense bf16 burst (MEASURED_PEAKS.json)"
        except Exception:
            pass
        dense_info = {"kernel": "pk2_umma_kernel (tcgen05 kind::i8, M=128 N=128 K=32)", "weight_limbs": 7,
                      "sample": "last %d of %d block-rows (128 samples each) x %d markers" % (nsample, nbr, M),
                      "ms": dms, "int8_tops": dops / (dms * 1e-3) / 1e12,
                      "roofline": {"bound": "tensor", "achieved": dops / (dms * 1e-3) / 1e12 / world, "peak": tpeak, "unit": "TOP/s per GPU",
                                   "frac": dops / (dms * 1e-3) / 1e12 / world / tpeak, "peak_source": tsrc},
                      "full_build_estimate_s": 7 * 2.0 * M * 128 * 128 * nbr * (nbr + 1) / 2 / (dops / (dms * 1e-3))}

    c4_info = None
    if not args.no_dense and (args.c4_full or world >= 4):
        # BASELINE config 4 at its named shape: 100,000 samples x 500,000 markers, the fp64 N x N GRM built on the tcgen05 int8
        # path, block-rows dealt over the ranks (marker shards exchanged once with ncclBroadcast), then PCG on the stored matrix
        N4, M4 = 100_000, 500_000
        g4 = new_context()
        _, u0, u1 = synth.thresholds(M4, SEED + 4)
        g4.setminMAFforGRM(0.01); g4.setmaxMissingRateforGRM(0.15)
        g4.setgeno_synth(N4, M4, SEED + 4, u0, u1)
        rng4 = np.random.default_rng(SEED + 4)
        B4 = np.asfortranarray(rng4.normal(size=(N4, 4)))
        w4 = rng4.uniform(0.05, 0.25, size=N4)
        want = g4.getCrossprodMatAndKin(B4)                      # on-the-fly product from the 2-bit store
        X4p, it4p = g4.getPCG1ofSigmaAndVector(w4, np.array([1.0, 0.3]), B4, 500, 1e-5, return_iter=True)
        c4_info = {"workload": "c4_100kx500k", "n_gpus": world, "builds": []}
        # weight limbs: 7 = the weights s_m^2 to 2^-47 of the largest (exact for every practical purpose), 5 = 2^-33 (the
        # tolerance-driven choice: still <= 1e-10 on the matrix, 2/7 less tensor work)
        for limbs in (7, 6, 5):
            g4.freeDenseGRM()
            barrier()
            tb = time.time()
            info4 = g4.buildDenseGRM(limbs)
            g4.sync()
            tb = max_over_ranks(time.time() - tb)
            g4.setGRMMode("dense")
            got = g4.getCrossprodMatAndKin(B4)
            prod = {}
            for kk in (1, 4):
                g4.bench_crossprod_device(kk, 2)
                barrier()
                msd, _ = g4.bench_crossprod_device(kk, 5)
                prod["k%d_ms" % kk] = max_over_ranks(float(msd.mean()))
            barrier()
            tp = time.time()
            X4, it4 = g4.getPCG1ofSigmaAndVector(w4, np.array([1.0, 0.3]), B4, 500, 1e-5, return_iter=True)
            tp = max_over_ranks(time.time() - tp)
            g4.setGRMMode("packed")
            c4_info["builds"].append({
                "weight_limbs": limbs, "build_wall_s": tb, "build_device_ms_this_rank": info4["build_ms"],
                "stored_gbytes_this_rank": info4["stored_bytes"] / 1e9,
                "int8_tops_aggregate": world * info4["int8_ops"] / (info4["build_ms"] * 1e-3) / 1e12,
                "stored_product_ms": prod, "product_rel_diff_vs_on_the_fly": float(np.max(np.abs(got - want)) / np.max(np.abs(want))),
                "pcg_on_stored_grm": {"columns": 4, "iterations": [int(v) for v in it4], "wall_s": tp,
                                      "iterations_on_the_fly": [int(v) for v in it4p],
                                      "solution_rel_diff_vs_on_the_fly": float(np.max(np.abs(X4 - X4p)) / np.max(np.abs(X4p)))}})
        g4.close()
    if rank == 0:
        line = {
            "metric": "grm_matvecs_per_s", "value": value, "unit": "matvecs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "ms_per_step_median": step_median, "ms_per_step_list": step_list, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int8 tensor-core limbs, exact int32 accumulate, fp64 recombine" if args.engine == "tensor" else "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "n_samples": N, "n_markers": M, "markers_per_gpu": int(Mloc), "k": 1,
                       "engine": args.engine, "l2_policy": "inputs (%.1f GB packed genotypes per sweep) >> 126 MB L2" % (bytes_launch / 1e9),
                       "sharding": "block-cyclic markers, 1 NCCL allreduce of N fp64 per product" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "matvecs/s", "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
                    "api": "sgb_get_crossprod_mat_and_kin (host pinned vectors)", "y_norm2": y_norm2,
                    "y_norm2_single_gpu": Y_NORM2.get(args.workload),
                    "y_norm2_equal_across_gpu_counts": (abs(y_norm2 - Y_NORM2[args.workload]) <= 1e-11 * Y_NORM2[args.workload]
                                                        if args.workload in Y_NORM2 else None)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "pk2_stream_kernel<1,2,3>", "peak_source": peak_src,
                         "bytes_per_launch": int(bytes_launch), "avg_launch_ms": sweep_ms,
                         "sweep1_ms": float(mk[:, 0].mean()), "sweep2_ms": float(mk[:, 1].mean()),
                         "whole_product_frac": bytes_alg_product / (total_ms / K * 1e-3) / 1e9 / peak},
            "batched": batched,
            "cpu_baseline": cb,
            "ingest": ingest_info,
            "step1": step1_info,
            "dense_grm": dense_info,
            "c4_dense_grm": c4_info,
            "step2": step2_info,
        }
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
