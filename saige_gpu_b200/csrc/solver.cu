// GRM products, multi-right-hand-side PCG and the AI-REML entry points: the B200 replacement of
// getCrossprodMatAndKin / getPCG1ofSigmaAndVector / getCoefficients / GetTrace / getAIScore / fitglmmaiRPCG
// (FG.cpp:1953-2006, 2593-2809, 3113-3341, 3409-3662).  Everything stays device resident inside one entry point;
// the host sees only scalars (dot products, tau) and the p x p covariance algebra.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "sgb_internal.h"

#define NEED_LOADED(h) do { if (!(h)->loaded) return sgb_fail(h, "genotypes not loaded: call setgeno first"); } while (0)

// d_scal layout (doubles)
enum { SC_COLSUM = 0, SC_T = 1024, SC_MAX = 2048, SC_MULT1 = 3072, SC_MULT2 = 4096, SC_PCG = 5120 /* rzA,rzB,r2: 3x1024 */ };

static int ensure_zeroed_i32(sgb_ctx *h, int32_t **p, size_t *cur_elems, size_t need_elems)
{
    if (*p && *cur_elems >= need_elems) return 0;
    size_t bytes = *cur_elems * sizeof(int32_t);
    SGB_TRY(sgb_ensure(h, (void **)p, &bytes, need_elems * sizeof(int32_t)));
    *cur_elems = bytes / sizeof(int32_t);
    CUDA_OK(h, cudaMemsetAsync(*p, 0, bytes, h->stream));
    return 0;
}

int sgb_ensure_f64(sgb_ctx *h, double **p, size_t *cur_elems, size_t need_elems)
{
    if (*p && *cur_elems >= need_elems) return 0;
    size_t bytes = *cur_elems * sizeof(double);
    SGB_TRY(sgb_ensure(h, (void **)p, &bytes, need_elems * sizeof(double)));
    *cur_elems = bytes / sizeof(double);
    return 0;
}

static void loco_local_range(sgb_ctx *h, int64_t gs, int64_t ge, int64_t *lo, int64_t *hi)
{
    *lo = std::lower_bound(h->loc2glob.begin(), h->loc2glob.end(), gs) - h->loc2glob.begin();
    *hi = std::upper_bound(h->loc2glob.begin(), h->loc2glob.end(), ge) - h->loc2glob.begin();
}

// Y[N x k] (ld N) = K.B  or the LOCO product; dB, dY device pointers (ld N).  May be called with dY == dB.
int sgb_crossprod_device(sgb_ctx *h, const double *dB, int k, double *dY, int loco)
{
    NEED_LOADED(h);
    if (k < 1 || k > 1024) return sgb_fail(h, "crossprod: k=%d out of range", k);
    if (h->grm_mode == SGB_GRM_DENSE) {
        if (loco) return sgb_fail(h, "LOCO products are not available from the stored dense GRM (use the packed-genotype mode)");
        return sgb_dense_crossprod_device(h, dB, k, dY);
    }
    const int64_t N = h->N, rowsG = h->rowsG, rowsT = h->rowsT;
    int64_t lo = 0, hi = 0;
    double mdiv = (double)h->M;
    if (loco) {
        if (h->loco_start < 0) return sgb_fail(h, "LOCO product requested before setStartEndIndex");
        loco_local_range(h, h->loco_start, h->loco_end, &lo, &hi);
        mdiv = (double)(h->M - (h->loco_end - h->loco_start + 1));       // FG.cpp:1848
    }
    const int64_t nblkN = h->sG / SGB_KSTEP_BYTES, nblkM = h->sT / SGB_KSTEP_BYTES;
    // fp64 scratch: raw1 [rowsG*k] | D [rowsG*k] | raw2 [rowsT*k + k]
    SGB_TRY(sgb_ensure_f64(h, &h->d_tmp, &h->tmp_elems, (size_t)(2 * rowsG + rowsT + 1) * k));
    double *raw1 = h->d_tmp, *D = raw1 + rowsG * k, *raw2 = D + rowsG * k;
    double *sc = h->d_scal;
    const bool tensor = h->engine != SGB_ENGINE_F64;
    const bool umma = (h->engine == SGB_ENGINE_UMMA && k >= 2) || (h->engine == SGB_ENGINE_TENSOR && k >= 3);   // wide batches: tcgen05
    h->cnt.n_crossprod_calls++; h->cnt.n_crossprod_columns += k;
    SGB_RANGE("grm_product");
    if (tensor) {
        // ---- tensor engines: 7 launches per product (statistics, split, sweep, epilogue+statistics, split, sweep, epilogue) ----
        const int nl = umma ? h->rhs_limbs : 8;                           // digits per value: 8 x 7 bits (mma.sync) or nl x 8 bits (tcgen05)
        const int pad = umma ? k_umma_npad(k, nl) : k;                    // accumulator columns per row (IMMA: x 8)
        const size_t accw = umma ? (size_t)pad : (size_t)8 * k;
        size_t lb = umma ? std::max(k_umma_image_bytes(pad, h->sG), k_umma_image_bytes(pad, h->sT)) : (size_t)k * std::max(nblkN, nblkM) * 2048;
        SGB_TRY(sgb_ensure(h, (void **)&h->d_limb, &h->limb_bytes, lb));
        // accumulators: every epilogue zeroes what its sweep wrote, so they are all-zero between products
        SGB_TRY(ensure_zeroed_i32(h, &h->d_acc1, &h->acc1_elems, (size_t)rowsG * accw));
        SGB_TRY(ensure_zeroed_i32(h, &h->d_acc2, &h->acc2_elems, (size_t)rowsT * accw));
        int32_t *ls1 = h->d_limbsum, *ls2 = h->d_limbsum + 8192;
        SGB_TRY(k_col_stats(h, dB, N, N, k, sc + SC_COLSUM, ls1, nl));
        // ---- sweep 1: acc1[m,c] = sum_i (2 - g_mi) q_ic over the marker-major copy ----
        if (umma) SGB_TRY(k_split_limbs_umma(h, dB, N, N, k, h->d_limb, h->sG, sc + SC_MULT1, ls1, nl, 1));
        else SGB_TRY(k_split_limbs(h, dB, N, N, k, h->d_limb, nblkN, sc + SC_MULT1, ls1, 1));
        if (h->time_sweeps) CUDA_OK(h, cudaEventRecord(h->ev[0], h->stream));
        {
            SGB_RANGE("sweep1_marker_major");
            if (umma) SGB_TRY(k_pk2_umma(h, h->dG, h->sG, rowsG, h->sG, h->d_limb, k, h->d_acc1, SGB_PLANE_VALUE, nl));
            else SGB_TRY(k_pk2_gemm(h, h->dG, h->sG, rowsG, h->sG, h->d_limb, k, h->d_acc1, SGB_PLANE_VALUE));
        }
        if (h->time_sweeps) CUDA_OK(h, cudaEventRecord(h->ev[1], h->stream));
        // ---- D = s^2 (G b - 2f sum b), left-out chromosome zeroed; t = sum_m 2f D (also appended to raw2 for the allreduce) ----
        SGB_TRY(k_recomb_post1(h, nl, h->d_acc1, rowsG, k, pad, sc + SC_MULT1, ls1, sc + SC_COLSUM, lo, hi, D, rowsG, sc + SC_T,
                               h->world > 1 ? raw2 + rowsT * k : nullptr, ls2, nl));
        // ---- sweep 2: acc2[i,c] = sum_m (2 - g_mi) q'_mc over the sample-major copy ----
        if (umma) SGB_TRY(k_split_limbs_umma(h, D, h->Mloc, rowsG, k, h->d_limb, h->sT, sc + SC_MULT2, ls2, nl, 1));
        else SGB_TRY(k_split_limbs(h, D, h->Mloc, rowsG, k, h->d_limb, nblkM, sc + SC_MULT2, ls2, 1));
        if (h->time_sweeps) CUDA_OK(h, cudaEventRecord(h->ev[2], h->stream));
        {
            SGB_RANGE("sweep2_sample_major");
            if (umma) SGB_TRY(k_pk2_umma(h, h->dGt, h->sT, rowsT, h->sT, h->d_limb, k, h->d_acc2, SGB_PLANE_VALUE, nl));
            else SGB_TRY(k_pk2_gemm(h, h->dGt, h->sT, rowsT, h->sT, h->d_limb, k, h->d_acc2, SGB_PLANE_VALUE));
        }
        if (h->time_sweeps) CUDA_OK(h, cudaEventRecord(h->ev[3], h->stream));
        if (h->world > 1) {
            // one sum-allreduce per (multi-)product: the N x k partial and the k centring scalars travel together
            SGB_TRY(k_recomb_post2(h, nl, h->d_acc2, rowsT, k, pad, sc + SC_MULT2, ls2, nullptr, 0.0, nullptr, 0, raw2, rowsT));
            SGB_TRY(sgb_allreduce_sum(h, raw2, rowsT * k + k));
            SGB_TRY(k_sweep2_post(h, raw2, rowsT, k, raw2 + rowsT * k, 1.0 / mdiv, dY, N));
        } else {
            SGB_TRY(k_recomb_post2(h, nl, h->d_acc2, rowsT, k, pad, sc + SC_MULT2, ls2, sc + SC_T, 1.0 / mdiv, dY, N, nullptr, 0));
        }
    } else {
        SGB_TRY(k_colsum(h, dB, N, N, k, sc + SC_COLSUM));
        if (h->time_sweeps) CUDA_OK(h, cudaEventRecord(h->ev[0], h->stream));
        SGB_TRY(k_rowdot_f64(h, dB, N, k, raw1, rowsG));
        if (h->time_sweeps) CUDA_OK(h, cudaEventRecord(h->ev[1], h->stream));
        SGB_TRY(k_sweep1_post(h, raw1, rowsG, k, sc + SC_COLSUM, lo, hi, D, sc + SC_T));
        if (h->time_sweeps) CUDA_OK(h, cudaEventRecord(h->ev[2], h->stream));
        SGB_TRY(k_coldot_f64(h, D, nullptr, rowsG, k, raw2, rowsT));
        if (h->time_sweeps) CUDA_OK(h, cudaEventRecord(h->ev[3], h->stream));
        const double *tvec = sc + SC_T;
        if (h->world > 1) {
            CUDA_OK(h, cudaMemcpyAsync(raw2 + rowsT * k, sc + SC_T, sizeof(double) * k, cudaMemcpyDeviceToDevice, h->stream));
            SGB_TRY(sgb_allreduce_sum(h, raw2, rowsT * k + k));
            tvec = raw2 + rowsT * k;
        }
        SGB_TRY(k_sweep2_post(h, raw2, rowsT, k, tvec, 1.0 / mdiv, dY, N));
    }
    if (h->time_sweeps) {
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        cudaEventElapsedTime(&h->last_sweep_ms[0], h->ev[0], h->ev[1]);
        cudaEventElapsedTime(&h->last_sweep_ms[1], h->ev[2], h->ev[3]);
    }
    return 0;
}


// raw[i + c*rowsT] = sum_m g_mi D[m + c*rowsG]  over the local markers, on the tensor engine (the second sweep of a
// product on its own).  Exact whenever the entries of D are integers below 2^53 and the sums stay below 2^53.
int sgb_gt_times_cols(sgb_ctx *h, const double *D, int k, double *raw)
{
    NEED_LOADED(h);
    const int64_t rowsG = h->rowsG, rowsT = h->rowsT;
    const int kpad = (k + 1) & ~1;
    SGB_TRY(sgb_ensure(h, (void **)&h->d_limb, &h->limb_bytes, k_umma_limb_bytes(k, h->sT)));
    SGB_TRY(ensure_zeroed_i32(h, &h->d_acc2, &h->acc2_elems, (size_t)rowsT * 8 * kpad));
    double *sc = h->d_scal;
    SGB_TRY(k_split_limbs_umma(h, D, h->Mloc, rowsG, k, h->d_limb, h->sT, sc + SC_MULT2, h->d_limbsum + 8192));
    SGB_TRY(k_pk2_umma(h, h->dGt, h->sT, rowsT, h->sT, h->d_limb, k, h->d_acc2, SGB_PLANE_VALUE));
    SGB_TRY(k_recombine_umma(h, h->d_acc2, rowsT, k, sc + SC_MULT2, h->d_limbsum + 8192, SGB_PLANE_VALUE, raw, rowsT));
    return 0;
}

// sum_m z_mi^2 for `nc` local-row ranges at once -> out[N x nc] (ld N); summed over ranks
static int diag_ranges(sgb_ctx *h, int nc, const std::vector<int64_t> &lo, const std::vector<int64_t> &hi, double *out)
{
    const int64_t N = h->N, rowsG = h->rowsG, rowsT = h->rowsT;
    const int64_t nblkM = h->sT / SGB_KSTEP_BYTES;
    // D1 | D2 [rowsG*nc each] | r1 | r2 [rowsT*nc each] | const[nc]
    SGB_TRY(sgb_ensure_f64(h, &h->d_tmp, &h->tmp_elems, (size_t)(2 * rowsG + 2 * rowsT + 1) * nc));
    double *D1 = h->d_tmp, *D2 = D1 + rowsG * nc, *r1 = D2 + rowsG * nc, *r2 = r1 + rowsT * nc, *cst = r2 + rowsT * nc;
    int64_t *d_rng = reinterpret_cast<int64_t *>(h->d_idx);
    if ((size_t)nc * 2 * sizeof(int64_t) > 8192 * sizeof(int)) return sgb_fail(h, "too many chromosome ranges");
    std::vector<int64_t> both(lo);
    both.insert(both.end(), hi.begin(), hi.end());
    CUDA_OK(h, cudaMemcpyAsync(d_rng, both.data(), sizeof(int64_t) * 2 * nc, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    SGB_TRY(k_diag_prep(h, nc, d_rng, d_rng + nc, D1, D2, cst));
    const bool umma = (h->engine == SGB_ENGINE_UMMA && nc >= 2) || (h->engine == SGB_ENGINE_TENSOR && nc >= 3);   // all chromosomes at once: tcgen05
    if (umma) {
        const int kpad = (nc + 1) & ~1;
        SGB_TRY(sgb_ensure(h, (void **)&h->d_limb, &h->limb_bytes, k_umma_limb_bytes(nc, h->sT)));
        SGB_TRY(ensure_zeroed_i32(h, &h->d_acc2, &h->acc2_elems, (size_t)rowsT * 8 * kpad));
        SGB_TRY(k_split_limbs_umma(h, D1, h->Mloc, rowsG, nc, h->d_limb, h->sT, h->d_scal + SC_MULT1, h->d_limbsum));
        SGB_TRY(k_pk2_umma(h, h->dGt, h->sT, rowsT, h->sT, h->d_limb, nc, h->d_acc2, SGB_PLANE_VALUE));
        SGB_TRY(k_recombine_umma(h, h->d_acc2, rowsT, nc, h->d_scal + SC_MULT1, h->d_limbsum, SGB_PLANE_VALUE, r1, rowsT));
        SGB_TRY(k_split_limbs_umma(h, D2, h->Mloc, rowsG, nc, h->d_limb, h->sT, h->d_scal + SC_MULT2, h->d_limbsum + 8192));
        SGB_TRY(k_pk2_umma(h, h->dGt, h->sT, rowsT, h->sT, h->d_limb, nc, h->d_acc2, SGB_PLANE_IS2));   // [g==2] plane
        SGB_TRY(k_recombine_umma(h, h->d_acc2, rowsT, nc, h->d_scal + SC_MULT2, h->d_limbsum + 8192, SGB_PLANE_IS2, r2, rowsT));
    } else if (h->engine != SGB_ENGINE_F64) {
        SGB_TRY(sgb_ensure(h, (void **)&h->d_limb, &h->limb_bytes, (size_t)nc * nblkM * 2048));
        SGB_TRY(ensure_zeroed_i32(h, &h->d_acc2, &h->acc2_elems, (size_t)rowsT * 8 * nc));
        SGB_TRY(k_split_limbs(h, D1, h->Mloc, rowsG, nc, h->d_limb, nblkM, h->d_scal + SC_MULT1, h->d_limbsum));
        SGB_TRY(k_pk2_gemm(h, h->dGt, h->sT, rowsT, h->sT, h->d_limb, nc, h->d_acc2, SGB_PLANE_VALUE));
        SGB_TRY(k_recombine(h, h->d_acc2, rowsT, nc, nc, h->d_scal + SC_MULT1, h->d_limbsum, SGB_PLANE_VALUE, r1, rowsT));
        SGB_TRY(k_split_limbs(h, D2, h->Mloc, rowsG, nc, h->d_limb, nblkM, h->d_scal + SC_MULT2, h->d_limbsum + 8192));
        SGB_TRY(k_pk2_gemm(h, h->dGt, h->sT, rowsT, h->sT, h->d_limb, nc, h->d_acc2, SGB_PLANE_IS2));   // [g==2] plane
        SGB_TRY(k_recombine(h, h->d_acc2, rowsT, nc, nc, h->d_scal + SC_MULT2, h->d_limbsum + 8192, SGB_PLANE_IS2, r2, rowsT));
    } else {
        // weight of g==2 is 2*D1 + D2; fold it into one pass and leave r2 = 0
        SGB_TRY(k_axpby(h, 2.0, D1, 1.0, D2, rowsG * nc, D2));
        SGB_TRY(k_coldot_f64(h, D1, D2, rowsG, nc, r1, rowsT));
        CUDA_OK(h, cudaMemsetAsync(r2, 0, sizeof(double) * rowsT * nc, h->stream));
    }
    if (h->world > 1) SGB_TRY(sgb_allreduce_sum(h, r1, 2 * rowsT * nc + nc));
    SGB_TRY(k_diag_post(h, r1, r2, rowsT, nc, cst, out, N));
    return 0;
}

// Get_Diagof_StdGeno (FG.cpp:665-704): cached sum over all markers
int sgb_diag_device(sgb_ctx *h)
{
    NEED_LOADED(h);
    if (h->diag_ready) return 0;
    std::vector<int64_t> lo(1, 0), hi(1, h->Mloc);
    SGB_TRY(diag_ranges(h, 1, lo, hi, h->d_diag));
    h->diag_ready = true;
    return 0;
}

// set_Diagof_StdGeno_LOCO (FG.cpp:4934-4958): column c = full diag - diag over chromosome c
int sgb_diag_loco_device(sgb_ctx *h)
{
    NEED_LOADED(h);
    int nc = (int)h->startVec.size();
    if (nc == 0) return sgb_fail(h, "set_Diagof_StdGeno_LOCO: call setStartEndIndexVec first");
    SGB_PROF(h, "set_Diagof_StdGeno_LOCO");
    SGB_TRY(sgb_diag_device(h));
    std::vector<int64_t> lo(nc, 0), hi(nc, 0);
    h->msub_by_chr.assign(nc, 0);
    for (int c = 0; c < nc; c++)
        if (h->startVec[c] != -1 && h->endVec[c] != -1) {
            loco_local_range(h, h->startVec[c], h->endVec[c], &lo[c], &hi[c]);
            h->msub_by_chr[c] = h->endVec[c] - h->startVec[c] + 1;
        }
    {
        SGB_PROF(h, "diag_loco alloc");
        size_t bytes = h->diag_loco_elems * sizeof(double);
        SGB_TRY(sgb_ensure(h, (void **)&h->d_diag_loco, &bytes, sizeof(double) * h->N * nc));
        h->diag_loco_elems = bytes / sizeof(double);
    }
    {
        SGB_PROF(h, "diag_ranges (all chromosomes)");
        SGB_TRY(diag_ranges(h, nc, lo, hi, h->d_diag_loco));
    }
    for (int c = 0; c < nc; c++)
        SGB_TRY(k_axpby(h, 1.0, h->d_diag, -1.0, h->d_diag_loco + (int64_t)c * h->N, h->N, h->d_diag_loco + (int64_t)c * h->N));
    h->diag_loco_ready = true;
    return 0;
}

// diag(Sigma) on device (not inverted), FG.cpp:2322-2393
static int sigma_diag_device(sgb_ctx *h, const double *d_w, const double *tau, int loco, double *d_out)
{
    if (loco) {
        if (!h->diag_loco_ready) return sgb_fail(h, "LOCO solve requested before set_Diagof_StdGeno_LOCO");
        if (h->loco_chrom < 0 || h->loco_chrom >= (int)h->msub_by_chr.size()) return sgb_fail(h, "bad chromIndex %d", h->loco_chrom);
        double scale = 1.0 / (double)(h->M - h->msub_by_chr[h->loco_chrom]);
        return k_sigma_diag(h, h->d_diag_loco + (int64_t)h->loco_chrom * h->N, scale, 0, d_w, tau[0], tau[1], d_out);
    }
    if (!h->kinDiagOne) SGB_TRY(sgb_diag_device(h));
    return k_sigma_diag(h, h->d_diag, 1.0 / (double)h->M, h->kinDiagOne ? 1 : 0, d_w, tau[0], tau[1], d_out);
}

// Multi-RHS Jacobi-PCG.  dB, dX device [N x k]; every column reproduces the sequential recurrence of
// getPCG1ofSigmaAndVector (FG.cpp:2593-2809) and freezes once its own ||r||^2 <= tol.
int sgb_pcg_device(sgb_ctx *h, const double *d_w, const double *tau, const double *dB, int k, int maxiter, double tol,
                   int loco, double *dX, int32_t *iters)
{
    NEED_LOADED(h);
    if (k < 1 || k > 1024) return sgb_fail(h, "pcg: k=%d out of range", k);
    const int64_t N = h->N;
    // arena: R,Z,P [N*k] | Pp,KP [N*k] | dsig [N] | partA [k*PB] | partB [2*k*PB]
    size_t need = (size_t)N * k * 5 + N + (size_t)3 * k * SGB_PART_BLOCKS;
    SGB_TRY(sgb_ensure_f64(h, &h->d_pcg, &h->pcg_elems, need));
    double *R = h->d_pcg, *Z = R + N * k, *P = Z + N * k, *Pp = P + N * k, *KP = Pp + N * k, *dsig = KP + N * k;
    double *partA = dsig + N, *partB = partA + (size_t)k * SGB_PART_BLOCKS;
    double *rz[2] = {h->d_scal + SC_PCG, h->d_scal + SC_PCG + 1024}, *r2 = h->d_scal + SC_PCG + 2048;
    SGB_TRY(sigma_diag_device(h, d_w, tau, loco, dsig));
    SGB_TRY(k_pcg_init(h, dB, dsig, k, dX, R, Z, P, rz[0], r2));
    CUDA_OK(h, cudaMemcpyAsync(h->h_scal, r2, sizeof(double) * k, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    std::vector<int> act, it(k, 0);
    for (int c = 0; c < k; c++) if (h->h_scal[c] > tol) act.push_back(c);
    int cur = 0, iter = 0;
    SGB_RANGE("pcg_solve");
    SGB_PROF(h, "pcg_solve");
    while (!act.empty() && iter < maxiter) {
        SGB_RANGE("pcg_iteration");
        iter++;
        int na = (int)act.size();
        CUDA_OK(h, cudaMemcpyAsync(h->d_idx, act.data(), sizeof(int) * na, cudaMemcpyHostToDevice, h->stream));
        if (tau[1] != 0.0) {                                   // FG.cpp:2401-2404 short-circuit
            SGB_TRY(k_gather_cols(h, P, h->d_idx, na, Pp));
            SGB_TRY(sgb_crossprod_device(h, Pp, na, KP, loco));
        }
        SGB_TRY(k_pcg_step1(h, P, KP, d_w, tau[0], tau[1], h->d_idx, na, partA));
        SGB_TRY(k_pcg_step2(h, P, KP, dsig, h->d_idx, na, dX, R, Z, rz[cur], partA, partB));
        SGB_TRY(k_pcg_step3(h, P, Z, h->d_idx, na, rz[cur], rz[cur ^ 1], r2, partB));
        // columns that were inactive keep their rz in both buffers: copy forward the active ones only is implicit
        // (step3 writes rz_out for active columns); inactive columns are never read again.
        cur ^= 1;
        CUDA_OK(h, cudaMemcpyAsync(h->h_scal, r2, sizeof(double) * k, cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        std::vector<int> next;
        for (int c : act) {
            it[c] = iter;
            if (h->h_scal[c] > tol) next.push_back(c);
        }
        act.swap(next);
    }
    for (int c = 0; c < k; c++) { h->cnt.n_pcg_solves++; h->cnt.n_pcg_iterations += it[c]; if (iters) iters[c] = it[c]; }
    if (h->verbose) {
        // the reference's log lines, one per right-hand side in the order its sequential solves would print them
        // (FG.cpp:2794-2798); downstream log scrapers read them
        for (int c = 0; c < k; c++) {
            if (it[c] >= maxiter) printf("pcg did not converge. You may increase maxiter number.\n");
            printf("iter from getPCG1ofSigmaAndVector %d\n", it[c]);
        }
        fflush(stdout);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// small host linear algebra (p x p, p = number of fixed-effect columns)
// ---------------------------------------------------------------------------------------------------
// arma::inv_sympd(arma::symmatu(A)) with pinv fallback (FG.cpp:3185-3190).  A column-major p x p.
static void inv_sympd_or_pinv(std::vector<double> &A, int p)
{
    for (int i = 0; i < p; i++) for (int j = 0; j < i; j++) A[i + j * p] = A[j + i * p];     // symmatu
    std::vector<double> L(A);
    bool ok = true;
    for (int j = 0; j < p && ok; j++) {
        double d = L[j + j * p];
        for (int q = 0; q < j; q++) d -= L[j + q * p] * L[j + q * p];
        if (!(d > 0)) { ok = false; break; }
        d = sqrt(d);
        L[j + j * p] = d;
        for (int i = j + 1; i < p; i++) {
            double v = L[i + j * p];
            for (int q = 0; q < j; q++) v -= L[i + q * p] * L[j + q * p];
            L[i + j * p] = v / d;
        }
    }
    if (ok) {
        // inverse via forward/back substitution on the identity
        std::vector<double> inv(p * p, 0.0);
        for (int c = 0; c < p; c++) {
            std::vector<double> y(p, 0.0);
            for (int i = 0; i < p; i++) {
                double v = (i == c) ? 1.0 : 0.0;
                for (int q = 0; q < i; q++) v -= L[i + q * p] * y[q];
                y[i] = v / L[i + i * p];
            }
            for (int i = p - 1; i >= 0; i--) {
                double v = y[i];
                for (int q = i + 1; q < p; q++) v -= L[q + i * p] * inv[q + c * p];
                inv[i + c * p] = v / L[i + i * p];
            }
        }
        A = inv;
        return;
    }
    // pinv through cyclic Jacobi eigen-decomposition of the symmetric matrix
    std::vector<double> S(A), V(p * p, 0.0);
    for (int i = 0; i < p; i++) V[i + i * p] = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0;
        for (int i = 0; i < p; i++) for (int j = 0; j < p; j++) if (i != j) off += S[i + j * p] * S[i + j * p];
        if (off < 1e-300) break;
        for (int a = 0; a < p; a++)
            for (int b = a + 1; b < p; b++) {
                double apq = S[a + b * p];
                if (fabs(apq) < 1e-300) continue;
                double th = (S[b + b * p] - S[a + a * p]) / (2.0 * apq);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int q = 0; q < p; q++) {
                    double x = S[q + a * p], y = S[q + b * p];
                    S[q + a * p] = c * x - s * y; S[q + b * p] = s * x + c * y;
                }
                for (int q = 0; q < p; q++) {
                    double x = S[a + q * p], y = S[b + q * p];
                    S[a + q * p] = c * x - s * y; S[b + q * p] = s * x + c * y;
                }
                for (int q = 0; q < p; q++) {
                    double x = V[q + a * p], y = V[q + b * p];
                    V[q + a * p] = c * x - s * y; V[q + b * p] = s * x + c * y;
                }
            }
    }
    double smax = 0.0;
    for (int i = 0; i < p; i++) smax = std::max(smax, fabs(S[i + i * p]));
    double tolv = p * smax * 2.220446049250313e-16;          // arma::pinv default tolerance
    std::vector<double> inv(p * p, 0.0);
    for (int q = 0; q < p; q++) {
        double ev = S[q + q * p];
        if (fabs(ev) <= tolv) continue;
        for (int i = 0; i < p; i++) for (int j = 0; j < p; j++) inv[i + j * p] += V[i + q * p] * V[j + q * p] / ev;
    }
    A = inv;
}

// ---------------------------------------------------------------------------------------------------
// per-call device arena + staging helpers
// ---------------------------------------------------------------------------------------------------
struct arena {
    sgb_ctx *h; double *base; size_t off;
    double *take(size_t n) { double *p = base + off; off += n; return p; }
};

static int arena_begin(sgb_ctx *h, size_t elems, arena *a)
{
    SGB_TRY(sgb_ensure_f64(h, &h->d_ai, &h->ai_elems, elems));
    a->h = h; a->base = h->d_ai; a->off = 0;
    return 0;
}

static int up(sgb_ctx *h, double *dst, const double *src, size_t n)
{
    CUDA_OK(h, cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    h->cnt.bytes_h2d += sizeof(double) * n;
    return 0;
}
static int down(sgb_ctx *h, double *dst, const double *src, size_t n)
{
    CUDA_OK(h, cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    h->cnt.bytes_d2h += sizeof(double) * n;
    return 0;
}

// out[q] = A[:,ia[q]] . B[:,ib[q]] for npairs pairs; result copied to host
static int dots_to_host(sgb_ctx *h, const double *A, const double *B, const std::vector<int> &pairs, double *out)
{
    int np = (int)pairs.size() / 2;
    if (np == 0) return 0;
    if (np > 2048) return sgb_fail(h, "too many dot products in one batch (%d)", np);
    SGB_TRY(sgb_ensure(h, &h->ws, &h->ws_bytes, (size_t)np * SGB_PART_BLOCKS * sizeof(double)));
    CUDA_OK(h, cudaMemcpyAsync(h->d_idx, pairs.data(), sizeof(int) * pairs.size(), cudaMemcpyHostToDevice, h->stream));
    SGB_TRY(k_pair_dots(h, A, h->N, B, h->N, h->d_idx, np, h->d_scal + SC_COLSUM));
    CUDA_OK(h, cudaMemcpyAsync(h->h_scal, h->d_scal + SC_COLSUM, sizeof(double) * np, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    memcpy(out, h->h_scal, sizeof(double) * np);
    return 0;
}

// C = cov * V   (p x p times p x n, column-major)
static std::vector<double> matmul_pp(const std::vector<double> &cov, int p, const double *V, int n)
{
    std::vector<double> C((size_t)p * n, 0.0);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < p; i++) {
            double s = 0.0;
            for (int q = 0; q < p; q++) s += cov[i + q * p] * V[q + (size_t)j * p];
            C[i + (size_t)j * p] = s;
        }
    return C;
}

// ---------------------------------------------------------------------------------------------------
// C ABI: products, PCG
// ---------------------------------------------------------------------------------------------------
extern "C" int sgb_get_diag_of_kin(sgb_ctx *h, double *out)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    if (h->kinDiagOne) { for (int64_t i = 0; i < h->N; i++) out[i] = 1.0; return 0; }      // FG.cpp:4372
    SGB_TRY(sgb_diag_device(h));
    SGB_TRY(down(h, out, h->d_diag, h->N));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < h->N; i++) out[i] /= (double)h->M;
    return 0;
}

extern "C" int sgb_set_diag_of_stdgeno_loco(sgb_ctx *h)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    SGB_TRY(sgb_diag_loco_device(h));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

static int crossprod_host(sgb_ctx *h, const double *B, int k, double *Y, int loco)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    if (k < 1) return sgb_fail(h, "crossprod: k must be >= 1");
    SGB_TRY(sgb_ensure_f64(h, &h->d_io, &h->io_elems, (size_t)h->N * k));
    SGB_TRY(up(h, h->d_io, B, (size_t)h->N * k));
    SGB_TRY(sgb_crossprod_device(h, h->d_io, k, h->d_io, loco));
    SGB_TRY(down(h, Y, h->d_io, (size_t)h->N * k));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int sgb_get_crossprod_mat_and_kin(sgb_ctx *h, const double *B, int k, double *Y) { return crossprod_host(h, B, k, Y, 0); }
extern "C" int sgb_get_crossprod_mat_and_kin_loco(sgb_ctx *h, const double *B, int k, double *Y) { return crossprod_host(h, B, k, Y, 1); }

extern "C" int sgb_get_diag_of_sigma(sgb_ctx *h, const double *w, const double *tau, int loco, double *out)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    arena a;
    SGB_TRY(arena_begin(h, (size_t)2 * h->N, &a));
    double *dw = a.take(h->N), *dd = a.take(h->N);
    SGB_TRY(up(h, dw, w, h->N));
    SGB_TRY(sigma_diag_device(h, dw, tau, loco, dd));
    SGB_TRY(down(h, out, dd, h->N));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int sgb_get_crossprod(sgb_ctx *h, const double *B, int k, const double *w, const double *tau, int loco, double *Y)
{
    NEED_LOADED(h);
    // getCrossprod (FG.cpp:2397-2425): tau0 * b/w + tau1 * K b, short-circuit when tau1 == 0
    if (tau[1] != 0.0) SGB_TRY(crossprod_host(h, B, k, Y, loco));
    for (int c = 0; c < k; c++)
        for (int64_t i = 0; i < h->N; i++) {
            size_t o = (size_t)c * h->N + i;
            double v = tau[0] * (B[o] * (1.0 / w[i]));
            Y[o] = tau[1] != 0.0 ? v + tau[1] * Y[o] : v;
        }
    return 0;
}

extern "C" int sgb_get_pcg1_of_sigma_and_vector(sgb_ctx *h, const double *w, const double *tau, const double *B, int k,
                                                int maxiterPCG, double tolPCG, int loco, double *X, int32_t *iters_out)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    if (k < 1) return sgb_fail(h, "pcg: k must be >= 1");
    arena a;
    SGB_TRY(arena_begin(h, (size_t)h->N * (2 * k + 1), &a));
    double *dw = a.take(h->N), *dB = a.take((size_t)h->N * k), *dX = a.take((size_t)h->N * k);
    SGB_TRY(up(h, dw, w, h->N));
    SGB_TRY(up(h, dB, B, (size_t)h->N * k));
    SGB_TRY(sgb_pcg_device(h, dw, tau, dB, k, maxiterPCG, tolPCG, loco, dX, iters_out));
    SGB_TRY(down(h, X, dX, (size_t)h->N * k));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int sgb_get_sigma_x(sgb_ctx *h, const double *w, const double *tau, const double *X, int p, int maxiterPCG,
                               double tolPCG, int loco, double *Sigma_iX)
{
    return sgb_get_pcg1_of_sigma_and_vector(h, w, tau, X, p, maxiterPCG, tolPCG, loco, Sigma_iX, nullptr);
}
extern "C" int sgb_get_sigma_g(sgb_ctx *h, const double *w, const double *tau, const double *Gm, int k, int maxiterPCG,
                               double tolPCG, int loco, double *Sigma_iG)
{
    return sgb_get_pcg1_of_sigma_and_vector(h, w, tau, Gm, k, maxiterPCG, tolPCG, loco, Sigma_iG, nullptr);
}

// ---------------------------------------------------------------------------------------------------
// C ABI: AI-REML
// ---------------------------------------------------------------------------------------------------
extern "C" int sgb_get_coefficients(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, const double *tau,
                                    int maxiterPCG, double tolPCG, int loco, double *Sigma_iY, double *Sigma_iX,
                                    double *cov, double *alpha, double *eta)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    if (p < 1 || p > 30) return sgb_fail(h, "getCoefficients: p=%d out of range [1,30]", p);
    const int64_t N = h->N;
    arena a;
    SGB_TRY(arena_begin(h, (size_t)N * (2 * (1 + p) + 2) + 64, &a));
    double *dw = a.take(N), *dYX = a.take((size_t)N * (1 + p)), *dS = a.take((size_t)N * (1 + p)), *deta = a.take(N), *dal = a.take(64);
    SGB_TRY(up(h, dw, w, N));
    SGB_TRY(up(h, dYX, Y, N));
    SGB_TRY(up(h, dYX + N, X, (size_t)N * p));
    // Sigma^-1 [Y | X] as ONE (1+p)-column solve (the reference runs 1+p sequential solves, FG.cpp:3171-3178)
    SGB_TRY(sgb_pcg_device(h, dw, tau, dYX, 1 + p, maxiterPCG, tolPCG, loco, dS, nullptr));
    // X^T Sigma_iX (p x p) and Sigma_iX^T Y (p)
    std::vector<int> pairs;
    for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) { pairs.push_back(1 + i); pairs.push_back(1 + j); }   // X_i . SiX_j
    for (int i = 0; i < p; i++) { pairs.push_back(0); pairs.push_back(1 + i); }                                   // Y . SiX_i
    std::vector<double> d(pairs.size() / 2);
    SGB_TRY(dots_to_host(h, dYX, dS, pairs, d.data()));
    std::vector<double> covm(d.begin(), d.begin() + p * p);
    inv_sympd_or_pinv(covm, p);
    std::vector<double> al = matmul_pp(covm, p, d.data() + p * p, 1);
    SGB_TRY(up(h, dal, al.data(), p));
    SGB_TRY(k_eta(h, dYX, dS, dS + N, p, dal, dw, tau[0], deta));
    SGB_TRY(down(h, Sigma_iY, dS, N));
    SGB_TRY(down(h, Sigma_iX, dS + N, (size_t)N * p));
    SGB_TRY(down(h, eta, deta, N));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    memcpy(cov, covm.data(), sizeof(double) * p * p);
    memcpy(alpha, al.data(), sizeof(double) * p);
    return 0;
}

// Device scratch of the AI step: dB right-hand sides [PY | U...] (column 0 = Sigma_iY on entry), dK = K.[...], dS = Sigma^-1 [...],
// dPr projected; kb = widest batch in columns; dC 4096 doubles
struct ai_scratch { double *dB, *dK, *dS, *dPr, *dC; int kb; };
static size_t ai_scratch_elems(int64_t N, int nrun) { return (size_t)N * 4 * (std::max(nrun, 10) + 2) + 4096; }
static void ai_scratch_take(arena *a, int64_t N, int nrun, ai_scratch *sc)
{
    sc->kb = std::max(nrun, 10) + 2;               // widest batch: [U | APY | PY]
    sc->dB = a->take((size_t)N * sc->kb); sc->dK = a->take((size_t)N * sc->kb); sc->dS = a->take((size_t)N * sc->kb);
    sc->dPr = a->take((size_t)N * sc->kb); sc->dC = a->take(4096);
}
static int ai_score_core(sgb_ctx *h, bool quant, const double *dw, const double *dY, const double *dX, const double *dSiX, int p,
                         const double *tau, const double *cov_in, int nrun, int maxiterPCG, double tolPCG, double traceCVcutoff,
                         sgb_probe_fn probes, void *user, const ai_scratch &sc, double *out8, double *PY_out, bool announce);

// Shared body of getAIScore / getAIScore_q.  out: YPAPY, YPA0PY, Trace0, Trace1, AI00, AI01, AI11, nrun used.
static int ai_score_impl(sgb_ctx *h, bool quant, const double *Y, const double *X, int p, const double *w, const double *tau,
                         const double *Sigma_iY, const double *Sigma_iX, const double *cov_in, int nrun, int maxiterPCG,
                         double tolPCG, double traceCVcutoff, sgb_probe_fn probes, void *user, double *out8, double *PY_out)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    if (p < 1 || p > 30) return sgb_fail(h, "getAIScore: p=%d out of range [1,30]", p);
    if (nrun < 2 || nrun > 500) return sgb_fail(h, "getAIScore: nrun=%d out of range [2,500]", nrun);
    if (!probes) return sgb_fail(h, "getAIScore: probe callback is NULL (probes are drawn by the caller's RNG)");
    const int64_t N = h->N;
    arena a;
    SGB_TRY(arena_begin(h, (size_t)N * (2 + 2 * p) + ai_scratch_elems(N, nrun), &a));
    double *dw = a.take(N), *dY = a.take(N), *dX = a.take((size_t)N * p), *dSiX = a.take((size_t)N * p);
    ai_scratch sc;
    ai_scratch_take(&a, N, nrun, &sc);
    SGB_TRY(up(h, dw, w, N));
    SGB_TRY(up(h, dY, Y, N));
    SGB_TRY(up(h, dX, X, (size_t)N * p));
    SGB_TRY(up(h, dSiX, Sigma_iX, (size_t)N * p));
    SGB_TRY(up(h, sc.dB, Sigma_iY, N));            // dB[:,0] = Sigma_iY for now
    return ai_score_core(h, quant, dw, dY, dX, dSiX, p, tau, cov_in, nrun, maxiterPCG, tolPCG, traceCVcutoff, probes, user, sc, out8, PY_out, false);
}

// The AI step on device-resident inputs (the host wrapper above uploads them; sgb_glmmkin_ai_pcg takes them from its Get_Coef)
static int ai_score_core(sgb_ctx *h, bool quant, const double *dw, const double *dY, const double *dX, const double *dSiX, int p,
                         const double *tau, const double *cov_in, int nrun, int maxiterPCG, double tolPCG, double traceCVcutoff,
                         sgb_probe_fn probes, void *user, const ai_scratch &sc, double *out8, double *PY_out, bool announce)
{
    const int64_t N = h->N;
    const int kb = sc.kb;
    // several trace estimates share one callback inside sgb_glmmkin_ai_pcg: count = 0 tells it that a new estimate starts
    // (the reference re-seeds there, GetTrace, FG.cpp:3114)
    if (announce && probes(user, N, 0, nullptr)) return sgb_fail(h, "getAIScore: probe callback failed");
    double *dB = sc.dB, *dK = sc.dK, *dS = sc.dS, *dPr = sc.dPr, *dC = sc.dC;
    SGB_RANGE("ai_score");
    SGB_PROF(h, "AI step (getAIScore / fitglmmaiRPCG)");
    std::vector<double> covm(cov_in, cov_in + p * p);
    std::vector<int> pairs;
    std::vector<double> d;
    if (quant) {                                   // getAIScore_q recomputes cov from X^T Sigma_iX (FG.cpp:3488-3494)
        pairs.clear();
        for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) { pairs.push_back(i); pairs.push_back(j); }
        d.resize(p * p);
        SGB_TRY(dots_to_host(h, dX, dSiX, pairs, d.data()));
        covm.assign(d.begin(), d.end());
        inv_sympd_or_pinv(covm, p);
    }
    // PY = Sigma_iY - Sigma_iX (cov (Sigma_iX^T Y))                                  FG.cpp:3284
    pairs.clear();
    for (int i = 0; i < p; i++) { pairs.push_back(i); pairs.push_back(0); }
    d.resize(p);
    SGB_TRY(dots_to_host(h, dSiX, dY, pairs, d.data()));
    std::vector<double> C = matmul_pp(covm, p, d.data(), 1);
    SGB_TRY(up(h, dC, C.data(), p));
    SGB_TRY(k_project(h, dB, dSiX, p, dC, 1, dB));            // dB[:,0] = PY
    if (PY_out) SGB_TRY(down(h, PY_out, dB, N));

    std::vector<double> t1, t0;
    double YPAPY = 0, YPA0PY = 0, AI00 = 0, AI01 = 0, AI11 = 0;
    int nstart = 0, nend = nrun;
    std::vector<double> hostU;
    bool first = true;
    int skipped_cols = 0;                          // first-batch columns taken from the resident copy instead of the callback
    while (true) {
        const int nu = nend - nstart;
        // sgb_set_probe_stream_fixed: the first batch is the same in every call (set_seed(200), FG.cpp:3114) -> resident copy
        const bool resident = first && h->probe_stream_fixed && h->ku_cols == nu && h->d_ku;
        if (!resident) {
            if (skipped_cols) {                    // bring the callback's stream to where the reference's would be
                hostU.resize((size_t)N * skipped_cols);
                if (probes(user, N, skipped_cols, hostU.data())) return sgb_fail(h, "getAIScore: probe callback failed");
                skipped_cols = 0;
            }
            hostU.resize((size_t)N * nu);
            if (probes(user, N, nu, hostU.data())) return sgb_fail(h, "getAIScore: probe callback failed");
        }
        if (first) {
            // ---- one (1+nu)-column product K.[PY | U] ----                        FG.cpp:3285, 3140
            if (resident) {
                CUDA_OK(h, cudaMemcpyAsync(dB + N, h->d_ku, sizeof(double) * N * nu, cudaMemcpyDeviceToDevice, h->stream));
                skipped_cols = nu;
                h->cnt.n_probe_batches_resident++;
            } else SGB_TRY(up(h, dB + N, hostU.data(), (size_t)N * nu));
            // K.U does not change between outer iterations (same probe stream): reuse it when U is bitwise the same
            bool reuse = resident;
            if (!resident && h->ku_cols == nu && h->d_ku) {
                int *d_cnt = h->d_idx + 8000, ndiff = 1;
                SGB_TRY(k_count_diff(h, dB + N, h->d_ku, N * nu, d_cnt));
                CUDA_OK(h, cudaMemcpyAsync(&ndiff, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
                CUDA_OK(h, cudaStreamSynchronize(h->stream));
                reuse = ndiff == 0;
            }
            {
            SGB_PROF(h, "K.[PY | U]");
            if (reuse) {
                SGB_TRY(sgb_crossprod_device(h, dB, 1, dK, 0));
                CUDA_OK(h, cudaMemcpyAsync(dK + N, h->d_ku + (size_t)N * nu, sizeof(double) * N * nu, cudaMemcpyDeviceToDevice, h->stream));
                h->cnt.n_probe_product_reuse++;
            } else {
                SGB_TRY(sgb_crossprod_device(h, dB, 1 + nu, dK, 0));
                h->ku_cols = 0;
                SGB_TRY(sgb_ensure_f64(h, &h->d_ku, &h->ku_elems, (size_t)N * nu * 2));
                CUDA_OK(h, cudaMemcpyAsync(h->d_ku, dB + N, sizeof(double) * N * nu, cudaMemcpyDeviceToDevice, h->stream));
                CUDA_OK(h, cudaMemcpyAsync(h->d_ku + (size_t)N * nu, dK + N, sizeof(double) * N * nu, cudaMemcpyDeviceToDevice, h->stream));
                h->ku_cols = nu;
            }
            }
            pairs.assign({0, 0});
            if (quant) { pairs.push_back(0); pairs.push_back(0); }
            double dd[2];
            SGB_TRY(dots_to_host(h, dB, dK, pairs, dd));      // PY . APY
            YPAPY = dd[0];
            if (quant) {
                pairs.assign({0, 0});
                SGB_TRY(dots_to_host(h, dB, dB, pairs, dd));  // PY . PY (A0 = I)             FG.cpp:3499-3500
                YPA0PY = dd[0];
            }
            // ---- right-hand sides of the batched solve: [U | APY | (PY)] ----
            // dB currently [PY | U]; build dS-input in dPr: [U | APY | PY]
            CUDA_OK(h, cudaMemcpyAsync(dPr, dB + N, sizeof(double) * N * nu, cudaMemcpyDeviceToDevice, h->stream));
            CUDA_OK(h, cudaMemcpyAsync(dPr + (size_t)N * nu, dK, sizeof(double) * N, cudaMemcpyDeviceToDevice, h->stream));
            int ks = nu + 1;
            if (quant) { CUDA_OK(h, cudaMemcpyAsync(dPr + (size_t)N * (nu + 1), dB, sizeof(double) * N, cudaMemcpyDeviceToDevice, h->stream)); ks++; }
            SGB_TRY(sgb_pcg_device(h, dw, tau, dPr, ks, maxiterPCG, tolPCG, 0, dS, nullptr));     // FG.cpp:3138, 3290
            // coefficients of the projections
            //   probes : Sigma_iX^T u            (FG.cpp:3139)
            //   APY/PY : Sigma_iX^T (Sigma^-1 v) (FG.cpp:3291 -- the reference projects the SOLVED vector here)
            pairs.clear();
            for (int j = 0; j < nu; j++) for (int i = 0; i < p; i++) { pairs.push_back(i); pairs.push_back(j); }
            d.resize((size_t)p * ks);
            SGB_TRY(dots_to_host(h, dSiX, dPr, pairs, d.data()));
            pairs.clear();
            for (int j = nu; j < ks; j++) for (int i = 0; i < p; i++) { pairs.push_back(i); pairs.push_back(j); }
            SGB_TRY(dots_to_host(h, dSiX, dS, pairs, d.data() + (size_t)p * nu));
            C = matmul_pp(covm, p, d.data(), ks);
            if ((size_t)p * ks > 4096) return sgb_fail(h, "getAIScore: p*nrun too large");
            SGB_TRY(up(h, dC, C.data(), (size_t)p * ks));
            SGB_TRY(k_project(h, dS, dSiX, p, dC, ks, dS));   // dS = [Pu... | PAPY | PA0PY]
            // traces: Au.Pu (and u.Pu);  AI entries
            pairs.clear();
            for (int j = 0; j < nu; j++) { pairs.push_back(1 + j); pairs.push_back(j); }        // dK[:,1+j] . dS[:,j]
            d.resize(nu);
            SGB_TRY(dots_to_host(h, dK, dS, pairs, d.data()));
            t1.insert(t1.end(), d.begin(), d.end());
            if (quant) {
                pairs.clear();
                for (int j = 0; j < nu; j++) { pairs.push_back(j); pairs.push_back(j); }        // u . Pu
                SGB_TRY(dots_to_host(h, dPr, dS, pairs, d.data()));
                t0.insert(t0.end(), d.begin(), d.end());
            }
            pairs.assign({0, nu});                                                                // APY . PAPY
            double ai[3];
            SGB_TRY(dots_to_host(h, dK, dS, pairs, ai));
            AI11 = ai[0];
            if (quant) {
                pairs.assign({0, nu, 0, nu + 1});                                                 // A0PY.PAPY , A0PY.PA0PY
                SGB_TRY(dots_to_host(h, dB, dS, pairs, ai));
                AI01 = ai[0]; AI00 = ai[1];
            }
            first = false;
        } else {
            // ---- 10 more probes (FG.cpp:3148-3153) ----
            SGB_TRY(up(h, dPr, hostU.data(), (size_t)N * nu));
            SGB_TRY(sgb_crossprod_device(h, dPr, nu, dK, 0));
            SGB_TRY(sgb_pcg_device(h, dw, tau, dPr, nu, maxiterPCG, tolPCG, 0, dS, nullptr));
            pairs.clear();
            for (int j = 0; j < nu; j++) for (int i = 0; i < p; i++) { pairs.push_back(i); pairs.push_back(j); }
            d.resize((size_t)p * nu);
            SGB_TRY(dots_to_host(h, dSiX, dPr, pairs, d.data()));
            C = matmul_pp(covm, p, d.data(), nu);
            SGB_TRY(up(h, dC, C.data(), (size_t)p * nu));
            SGB_TRY(k_project(h, dS, dSiX, p, dC, nu, dS));
            pairs.clear();
            for (int j = 0; j < nu; j++) { pairs.push_back(j); pairs.push_back(j); }
            d.resize(nu);
            SGB_TRY(dots_to_host(h, dK, dS, pairs, d.data()));
            t1.insert(t1.end(), d.begin(), d.end());
            if (quant) {
                SGB_TRY(dots_to_host(h, dPr, dS, pairs, d.data()));
                t0.insert(t0.end(), d.begin(), d.end());
            }
        }
        double cv1 = sgb_cal_cv(t1.data(), (int)t1.size());
        double cv0 = quant ? sgb_cal_cv(t0.data(), (int)t0.size()) : 0.0;
        if (cv1 > traceCVcutoff || cv0 > traceCVcutoff) {
            nstart = nend; nend += 10;
            if (nend - nstart > kb) return sgb_fail(h, "internal: probe batch larger than arena");
            if (nend > 2000) return sgb_fail(h, "getAIScore: trace estimator did not reach CV cutoff within 2000 probes");
        } else break;
    }
    double m1 = 0, m0 = 0;
    for (double v : t1) m1 += v;
    m1 /= (double)t1.size();
    if (quant) { for (double v : t0) m0 += v; m0 /= (double)t0.size(); }
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    out8[0] = YPAPY; out8[1] = YPA0PY; out8[2] = m0; out8[3] = m1; out8[4] = AI00; out8[5] = AI01; out8[6] = AI11;
    out8[7] = (double)nend;
    return 0;
}

extern "C" int sgb_get_ai_score(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, const double *tau,
                                const double *Sigma_iY, const double *Sigma_iX, const double *cov, int nrun, int maxiterPCG,
                                double tolPCG, double traceCVcutoff, sgb_probe_fn probes, void *user, double *out4, double *PY)
{
    double o[8];
    SGB_TRY(ai_score_impl(h, false, Y, X, p, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff, probes, user, o, PY));
    out4[0] = o[0]; out4[1] = o[3]; out4[2] = o[6]; out4[3] = o[7];
    return 0;
}

extern "C" int sgb_get_ai_score_q(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, const double *tau,
                                  const double *Sigma_iY, const double *Sigma_iX, const double *cov, int nrun, int maxiterPCG,
                                  double tolPCG, double traceCVcutoff, sgb_probe_fn probes, void *user, double *out8, double *PY)
{
    return ai_score_impl(h, true, Y, X, p, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff, probes, user, out8, PY);
}

// the tau update of fitglmmaiRPCG (FG.cpp:3321-3340) from o = {YPAPY, YPA0PY, Trace0, Trace1, AI00, AI01, AI11, nrun}
static void tau_step_binary(const double *o, double *tau, double tol)
{
    double score1 = o[0] - o[3], AI1 = o[6];
    double Dtau = score1 / AI1;
    double tau0[2] = {tau[0], tau[1]};
    tau[1] = tau0[1] + Dtau;
    for (int i = 0; i < 2; i++) if (tau[i] < tol) tau[i] = 0;
    double step = 1.0;
    while (tau[1] < 0.0) { step *= 0.5; tau[1] = tau0[1] + step * Dtau; }
    for (int i = 0; i < 2; i++) if (tau[i] < tol) tau[i] = 0;
}

// the tau update of fitglmmaiRPCG_q (FG.cpp:3634-3661); zero[i] = tau[i] < tol before the AI step; false: singular AI matrix
static bool tau_step_q(const double *o, const bool *zero, double *tau, double tol)
{
    double s0 = o[1] - o[2], s1 = o[0] - o[3];
    double a00 = o[4], a01 = o[5], a11 = o[6];
    double det = a00 * a11 - a01 * a01;
    if (det == 0.0 || !std::isfinite(det)) return false;
    double D0 = (a11 * s0 - a01 * s1) / det, D1 = (-a01 * s0 + a00 * s1) / det;     // solve(AI, score)
    double tau0[2] = {tau[0], tau[1]};
    tau[0] = tau0[0] + D0; tau[1] = tau0[1] + D1;
    for (int i = 0; i < 2; i++) if (zero[i] && tau[i] < tol) tau[i] = 0;
    double step = 1.0;
    while (tau[0] < 0.0 || tau[1] < 0.0) {
        step *= 0.5;
        tau[0] = tau0[0] + step * D0; tau[1] = tau0[1] + step * D1;
        for (int i = 0; i < 2; i++) if (zero[i] && tau[i] < tol) tau[i] = 0;
    }
    for (int i = 0; i < 2; i++) if (tau[i] < tol) tau[i] = 0;
    return true;
}

// fitglmmaiRPCG (FG.cpp:3302-3341)
extern "C" int sgb_fit_glmmai_rpcg(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, double *tau,
                                   const double *Sigma_iY, const double *Sigma_iX, const double *cov, int nrun, int maxiterPCG,
                                   double tolPCG, double tol, double traceCVcutoff, sgb_probe_fn probes, void *user)
{
    std::vector<double> PY(h->N > 0 ? h->N : 1);
    double o[8];
    SGB_TRY(ai_score_impl(h, false, Y, X, p, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff, probes, user, o, PY.data()));
    tau_step_binary(o, tau, tol);
    return 0;
}

// fitglmmaiRPCG_q (FG.cpp:3610-3662)
extern "C" int sgb_fit_glmmai_rpcg_q(sgb_ctx *h, const double *Y, const double *X, int p, const double *w, double *tau,
                                     const double *Sigma_iY, const double *Sigma_iX, const double *cov, int nrun, int maxiterPCG,
                                     double tolPCG, double tol, double traceCVcutoff, sgb_probe_fn probes, void *user)
{
    std::vector<double> PY(h->N > 0 ? h->N : 1);
    double o[8];
    bool zero[2] = {tau[0] < tol, tau[1] < tol};
    SGB_TRY(ai_score_impl(h, true, Y, X, p, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff, probes, user, o, PY.data()));
    if (!tau_step_q(o, zero, tau, tol)) return sgb_fail(h, "fitglmmaiRPCG_q: singular AI matrix");
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// C ABI: driver loops of SAIGE_fitGLMM_fast.R with the N-vectors resident on the device
// ---------------------------------------------------------------------------------------------------
// device state of one Get_Coef loop: w = W, YX = [Y | X], S = Sigma^-1 [Y | X], eta (offset included), mu, y, offset
struct coef_dev { double *w, *YX, *S, *eta, *mu, *y, *off, *al; };

static int coef_dev_take(sgb_ctx *h, int p, size_t extra, arena *a, coef_dev *s)
{
    const int64_t N = h->N;
    SGB_TRY(arena_begin(h, (size_t)N * (2 * (1 + p) + 5) + 64 + extra, a));
    s->w = a->take(N); s->YX = a->take((size_t)N * (1 + p)); s->S = a->take((size_t)N * (1 + p));
    s->eta = a->take(N); s->mu = a->take(N); s->y = a->take(N); s->off = a->take(N); s->al = a->take(64);
    return 0;
}

// Get_Coef (FG.R:2-35) on device state: s.eta holds eta0 on entry and the final eta on return; alpha0 is updated to the
// last alpha that entered a convergence test, as the R loop leaves it
static int get_coef_device(sgb_ctx *h, int family, const coef_dev &s, int p, const double *tau, std::vector<double> &alpha0,
                           int maxiter, int maxiterPCG, double tolPCG, int loco, std::vector<double> &covm,
                           std::vector<double> &al, int32_t *n_iter)
{
    const int64_t N = h->N;
    const double tol_coef = 0.1;
    SGB_RANGE("get_coef");
    SGB_PROF(h, loco ? "Get_Coef_LOCO" : "Get_Coef");
    SGB_TRY(k_irls_update(h, family, s.eta, 0, s.y, s.off, s.eta, s.mu, s.YX, s.w));
    std::vector<int> pairs;
    for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) { pairs.push_back(1 + i); pairs.push_back(1 + j); }   // X_i . SiX_j
    for (int i = 0; i < p; i++) { pairs.push_back(0); pairs.push_back(1 + i); }                                   // Y . SiX_i
    std::vector<double> d(pairs.size() / 2);
    int it = 0;
    while (it < maxiter) {
        it++;
        // getCoefficients (FG.cpp:3160-3200): Sigma^-1 [Y | X] as one (1+p)-column solve
        SGB_TRY(sgb_pcg_device(h, s.w, tau, s.YX, 1 + p, maxiterPCG, tolPCG, loco, s.S, nullptr));
        SGB_TRY(dots_to_host(h, s.YX, s.S, pairs, d.data()));
        covm.assign(d.begin(), d.begin() + p * p);
        inv_sympd_or_pinv(covm, p);
        al = matmul_pp(covm, p, d.data() + p * p, 1);
        SGB_TRY(up(h, s.al, al.data(), p));
        SGB_TRY(k_eta(h, s.YX, s.S, s.S + N, p, s.al, s.w, tau[0], s.eta));                 // re.coef$eta
        SGB_TRY(k_irls_update(h, family, s.eta, 1, s.y, s.off, s.eta, s.mu, s.YX, s.w));    // eta + offset, mu, Y, W
        double worst = 0.0;
        for (int i = 0; i < p; i++) worst = std::max(worst, fabs(al[i] - alpha0[i]) / (fabs(al[i]) + fabs(alpha0[i]) + tol_coef));
        if (worst < tol_coef) break;
        alpha0 = al;
    }
    if (n_iter) *n_iter = it;
    return 0;
}

static int coef_args_ok(sgb_ctx *h, int family, int p, int maxiter)
{
    if (family != 0 && family != 1) return sgb_fail(h, "Get_Coef: family %d (0 = binomial, 1 = gaussian)", family);
    if (p < 1 || p > 30) return sgb_fail(h, "Get_Coef: p=%d out of range [1,30]", p);
    if (maxiter < 1) return sgb_fail(h, "Get_Coef: maxiter must be >= 1");
    return 0;
}

extern "C" int sgb_get_coef(sgb_ctx *h, int family, const double *y, const double *X, int p, const double *offset, const double *tau,
                            const double *alpha0, const double *eta0, int maxiter, int maxiterPCG, double tolPCG, int loco,
                            double *Y, double *alpha, double *eta, double *W, double *cov, double *Sigma_iY, double *Sigma_iX,
                            double *mu, int32_t *n_iter)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    SGB_TRY(coef_args_ok(h, family, p, maxiter));
    const int64_t N = h->N;
    arena a; coef_dev s;
    SGB_TRY(coef_dev_take(h, p, 0, &a, &s));
    SGB_TRY(up(h, s.y, y, N));
    SGB_TRY(up(h, s.off, offset, N));
    SGB_TRY(up(h, s.eta, eta0, N));
    SGB_TRY(up(h, s.YX + N, X, (size_t)N * p));
    std::vector<double> a0(alpha0, alpha0 + p), covm, al;
    SGB_TRY(get_coef_device(h, family, s, p, tau, a0, maxiter, maxiterPCG, tolPCG, loco, covm, al, n_iter));
    if (Y) SGB_TRY(down(h, Y, s.YX, N));
    if (eta) SGB_TRY(down(h, eta, s.eta, N));
    if (W) SGB_TRY(down(h, W, s.w, N));
    if (mu) SGB_TRY(down(h, mu, s.mu, N));
    if (Sigma_iY) SGB_TRY(down(h, Sigma_iY, s.S, N));
    if (Sigma_iX) SGB_TRY(down(h, Sigma_iX, s.S + N, (size_t)N * p));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    memcpy(cov, covm.data(), sizeof(double) * p * p);
    memcpy(alpha, al.data(), sizeof(double) * p);
    return 0;
}

extern "C" int sgb_get_coef_loco_all(sgb_ctx *h, int family, const double *y, const double *X, int p, const double *offset,
                                     const double *tau, const double *alpha0, const double *eta0, int maxiter, int maxiterPCG,
                                     double tolPCG, double *Y, double *alpha, double *eta, double *cov, double *mu, int32_t *n_iter,
                                     sgb_chrom_done_fn on_chrom, void *chrom_user)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    SGB_TRY(coef_args_ok(h, family, p, maxiter));
    if (!h->diag_loco_ready) return sgb_fail(h, "LOCO refits requested before set_Diagof_StdGeno_LOCO");
    const int64_t N = h->N;
    const int nchr = (int)h->startVec.size();
    arena a; coef_dev s;
    SGB_TRY(coef_dev_take(h, p, 0, &a, &s));
    SGB_TRY(up(h, s.y, y, N));
    SGB_TRY(up(h, s.off, offset, N));
    SGB_TRY(up(h, s.eta, eta0, N));
    SGB_TRY(up(h, s.YX + N, X, (size_t)N * p));
    std::vector<double> a0(alpha0, alpha0 + p), covm, al;
    for (int c = 0; c < nchr; c++) {
        if (h->startVec[c] == -1 || h->endVec[c] == -1) continue;
        SGB_TRY(sgb_set_start_end_index(h, h->startVec[c], h->endVec[c], c));
        int32_t it = 0;
        // chromosome c starts from chromosome c-1's (alpha, eta): FG.R:265, 272-273
        SGB_TRY(get_coef_device(h, family, s, p, tau, a0, maxiter, maxiterPCG, tolPCG, 1, covm, al, &it));
        a0 = al;
        SGB_TRY(down(h, Y + (size_t)c * N, s.YX, N));
        SGB_TRY(down(h, eta + (size_t)c * N, s.eta, N));
        SGB_TRY(down(h, mu + (size_t)c * N, s.mu, N));
        memcpy(cov + (size_t)c * p * p, covm.data(), sizeof(double) * p * p);
        memcpy(alpha + (size_t)c * p, al.data(), sizeof(double) * p);
        if (n_iter) n_iter[c] = it;
        if (on_chrom) {
            CUDA_OK(h, cudaStreamSynchronize(h->stream));
            on_chrom(chrom_user, c);
        }
    }
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// glmmkin.ai_PCG_Rcpp_Binary / _Quantitative after setgeno (FG.R:127-304, 340-549) as one call: every Get_Coef, the AI steps and
// (LOCO) the leave-one-chromosome-out refits on one device-resident state; the host sees tau, alpha, cov and the trace scalars.
extern "C" int sgb_glmmkin_ai_pcg(sgb_ctx *h, int quantitative, const double *y, const double *X, int p, const double *offset,
                                  const double *alpha_fit0, const double *eta_fit0, const double *tauInit, int maxiter, double tol,
                                  int nrun, double tolPCG, int maxiterPCG, double traceCVcutoff, int loco, sgb_probe_fn probes,
                                  void *user, double *tau_out, double *alpha, double *eta, double *mu, double *Yout, double *cov,
                                  int32_t *converged, int32_t *n_outer, double *Y_loco, double *alpha_loco, double *eta_loco,
                                  double *cov_loco, double *mu_loco, int32_t *n_iter_loco, sgb_chrom_done_fn on_chrom, void *chrom_user)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    const bool quant = quantitative != 0;
    const int family = quant ? 1 : 0;
    SGB_TRY(coef_args_ok(h, family, p, maxiter));
    if (nrun < 2 || nrun > 500) return sgb_fail(h, "glmmkin.ai_PCG: nrun=%d out of range [2,500]", nrun);
    if (!probes) return sgb_fail(h, "glmmkin.ai_PCG: probe callback is NULL (probes are drawn by the caller's RNG)");
    if (loco && !(Y_loco && alpha_loco && eta_loco && cov_loco && mu_loco)) return sgb_fail(h, "glmmkin.ai_PCG: LOCO outputs are NULL");
    const int64_t N = h->N;
    SGB_RANGE("glmmkin_ai_pcg");
    arena a; coef_dev s;
    SGB_TRY(coef_dev_take(h, p, (size_t)N + ai_scratch_elems(N, nrun), &a, &s));
    double *d_eta_loop = a.take(N);                 // the R loop's `eta`: fit0's until the first iteration ends (FG.R:139, 179, 196)
    ai_scratch sc;
    ai_scratch_take(&a, N, nrun, &sc);
    SGB_TRY(up(h, s.y, y, N));
    SGB_TRY(up(h, s.off, offset, N));
    SGB_TRY(up(h, d_eta_loop, eta_fit0, N));
    SGB_TRY(up(h, s.YX + N, X, (size_t)N * p));
    double tau[2] = {0.0, 0.0};
    if (!quant) { tau[0] = 1.0; tau[1] = tauInit[1] == 0.0 ? 0.1 : tauInit[1]; }                  // FG.R:145-156
    else if (tauInit[0] + tauInit[1] == 0.0) { tau[0] = 1.0; tau[1] = 0.0; }                        // FG.R:375-381
    else { tau[0] = tauInit[0]; tau[1] = tauInit[1]; }
    double tau0[2] = {tau[0], tau[1]};
    std::vector<double> a0(alpha_fit0, alpha_fit0 + p), covm, al, al_loop(alpha_fit0, alpha_fit0 + p);
    double o[8];
    auto coef = [&](const std::vector<double> &start) -> int {      // Get_Coef from (start, the loop's eta)
        a0 = start;
        CUDA_OK(h, cudaMemcpyAsync(s.eta, d_eta_loop, sizeof(double) * N, cudaMemcpyDeviceToDevice, h->stream));
        return get_coef_device(h, family, s, p, tau, a0, maxiter, maxiterPCG, tolPCG, 0, covm, al, nullptr);
    };
    auto ai = [&]() -> int {                                         // the AI step on the state Get_Coef left
        CUDA_OK(h, cudaMemcpyAsync(sc.dB, s.S, sizeof(double) * N, cudaMemcpyDeviceToDevice, h->stream));
        return ai_score_core(h, quant, s.w, s.YX, s.YX + N, s.S + N, p, tau, covm.data(), nrun, maxiterPCG, tolPCG, traceCVcutoff,
                             probes, user, sc, o, nullptr, true);
    };
    SGB_TRY(coef(al_loop));
    SGB_TRY(ai());
    tau[1] = std::max(0.0, tau0[1] + tau0[1] * tau0[1] * (o[0] - o[3]) / (double)N);                // FG.R:163, 388-389
    if (quant) tau[0] = std::max(0.0, tau0[0] + tau0[0] * tau0[0] * (o[1] - o[2]) / (double)N);
    if (!quant) al_loop = al;                       // binary: alpha0 = re.coef$alpha; quantitative keeps fit0's until the loop sets it
    int i = 0;
    for (i = 1; i <= maxiter; i++) {
        tau0[0] = tau[0]; tau0[1] = tau[1];
        SGB_TRY(coef(al_loop));
        bool zero[2] = {tau[0] < tol, tau[1] < tol};
        SGB_TRY(ai());
        if (!quant) tau_step_binary(o, tau, tol);
        else if (!tau_step_q(o, zero, tau, tol)) return sgb_fail(h, "fitglmmaiRPCG_q: singular AI matrix");
        al_loop = al;                                                                                // alpha = re.coef$alpha
        CUDA_OK(h, cudaMemcpyAsync(d_eta_loop, s.eta, sizeof(double) * N, cudaMemcpyDeviceToDevice, h->stream));   // eta = re.coef$eta
        if (quant && tau[0] <= 0.0) return sgb_fail(h, "ERROR! The first variance component parameter estimate is 0");
        if ((!quant && tau[1] == 0.0) || (quant && tau[1] <= 0.0)) break;
        double worst = 0.0;
        for (int q = 0; q < 2; q++) worst = std::max(worst, fabs(tau[q] - tau0[q]) / (fabs(tau[q]) + fabs(tau0[q]) + tol));
        if (worst < tol) break;
        if (std::max(tau[0], tau[1]) > 1.0 / (tol * tol)) { i = maxiter; break; }
    }
    if (i > maxiter) i = maxiter;                   // seq_len(maxiter) ran out: R leaves i == maxiter
    SGB_TRY(coef(al_loop));                         // FG.R:206
    SGB_TRY(down(h, Yout, s.YX, N));
    SGB_TRY(down(h, eta, s.eta, N));
    SGB_TRY(down(h, mu, s.mu, N));
    memcpy(cov, covm.data(), sizeof(double) * p * p);
    memcpy(alpha, al.data(), sizeof(double) * p);
    tau_out[0] = tau[0]; tau_out[1] = tau[1];
    if (converged) *converged = i < maxiter;
    if (n_outer) *n_outer = i;
    if (on_chrom) {
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        on_chrom(chrom_user, -1);
    }
    if (loco) {
        // FG.R:255-292: set_Diagof_StdGeno_LOCO, then every chromosome from the previous one's (alpha, eta), starting at the fit's
        SGB_TRY(sgb_diag_loco_device(h));
        a0 = al;
        const int nchr = (int)h->startVec.size();
        for (int c = 0; c < nchr; c++) {
            if (h->startVec[c] == -1 || h->endVec[c] == -1) continue;
            SGB_TRY(sgb_set_start_end_index(h, h->startVec[c], h->endVec[c], c));
            int32_t it = 0;
            SGB_TRY(get_coef_device(h, family, s, p, tau, a0, maxiter, maxiterPCG, tolPCG, 1, covm, al, &it));
            a0 = al;
            SGB_TRY(down(h, Y_loco + (size_t)c * N, s.YX, N));
            SGB_TRY(down(h, eta_loco + (size_t)c * N, s.eta, N));
            SGB_TRY(down(h, mu_loco + (size_t)c * N, s.mu, N));
            memcpy(cov_loco + (size_t)c * p * p, covm.data(), sizeof(double) * p * p);
            memcpy(alpha_loco + (size_t)c * p, al.data(), sizeof(double) * p);
            if (n_iter_loco) n_iter_loco[c] = it;
            if (on_chrom) {
                CUDA_OK(h, cudaStreamSynchronize(h->stream));
                on_chrom(chrom_user, c);
            }
        }
    }
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int sgb_set_probe_stream_fixed(sgb_ctx *h, int on)
{
    h->probe_stream_fixed = on ? 1 : 0;
    return 0;
}

// A x = b for `nrhs` right-hand sides by Gaussian elimination with partial pivoting (R's solve(); p <= 30).  A, B column-major.
static bool solve_pp(std::vector<double> A, int p, std::vector<double> &B, int nrhs)
{
    for (int c = 0; c < p; c++) {
        int piv = c;
        for (int r = c + 1; r < p; r++) if (fabs(A[r + c * p]) > fabs(A[piv + c * p])) piv = r;
        if (A[piv + c * p] == 0.0) return false;
        if (piv != c) {
            for (int q = 0; q < p; q++) std::swap(A[c + q * p], A[piv + q * p]);
            for (int j = 0; j < nrhs; j++) std::swap(B[c + (size_t)j * p], B[piv + (size_t)j * p]);
        }
        for (int r = c + 1; r < p; r++) {
            const double f = A[r + c * p] / A[c + c * p];
            if (f == 0.0) continue;
            for (int q = c; q < p; q++) A[r + q * p] -= f * A[c + q * p];
            for (int j = 0; j < nrhs; j++) B[r + (size_t)j * p] -= f * B[c + (size_t)j * p];
        }
    }
    for (int j = 0; j < nrhs; j++)
        for (int r = p - 1; r >= 0; r--) {
            double v = B[r + (size_t)j * p];
            for (int q = r + 1; q < p; q++) v -= A[r + q * p] * B[q + (size_t)j * p];
            B[r + (size_t)j * p] = v / A[r + r * p];
        }
    return true;
}

static int dots_chunked(sgb_ctx *h, const double *A, const double *B, const std::vector<int> &pairs, double *out)
{
    const size_t np = pairs.size() / 2;
    for (size_t o = 0; o < np; o += 2048) {
        const size_t n = std::min<size_t>(2048, np - o);
        std::vector<int> part(pairs.begin() + 2 * o, pairs.begin() + 2 * (o + n));
        SGB_TRY(dots_to_host(h, A, B, part, out + o));
    }
    return 0;
}

extern "C" int sgb_variance_ratio_markers(sgb_ctx *h, const int64_t *marker_idx, int nmark, int from_vr_store, const double *w,
                                          const double *tau, const double *X, int p, const double *XV, const double *XXVX_inv,
                                          const double *Sigma_iX, const double *mu2, int maxiterPCG, double tolPCG, double *var1,
                                          double *var2null, double *AC)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    if (p < 1 || p > 30) return sgb_fail(h, "variance ratio: p=%d out of range [1,30]", p);
    if (nmark < 1 || nmark > 128) return sgb_fail(h, "variance ratio: %d markers per call (1..128)", nmark);
    const int64_t N = h->N, Bv = (N + 3) / 4;
    const int64_t Mstore = from_vr_store ? h->Mvr : h->M;
    for (int j = 0; j < nmark; j++)
        if (marker_idx[j] < 0 || marker_idx[j] >= Mstore)
            return sgb_fail(h, "variance ratio: marker index %lld out of range [0,%lld)", (long long)marker_idx[j], (long long)Mstore);
    SGB_RANGE("variance_ratio_markers");
    arena a;
    SGB_TRY(arena_begin(h, (size_t)N * (2 + 4 * p + 3 * (size_t)nmark) + 4096, &a));
    double *dw = a.take(N), *dmu2 = a.take(N), *dX = a.take((size_t)N * p), *dXVt = a.take((size_t)N * p);
    double *dXX = a.take((size_t)N * p), *dSiX = a.take((size_t)N * p);
    double *dG = a.take((size_t)N * nmark), *dSiG = a.take((size_t)N * nmark), *dT = a.take((size_t)N * nmark), *dC = a.take(4096);
    SGB_TRY(up(h, dw, w, N));
    if (mu2) SGB_TRY(up(h, dmu2, mu2, N));
    SGB_TRY(up(h, dX, X, (size_t)N * p));
    SGB_TRY(up(h, dXX, XXVX_inv, (size_t)N * p));
    SGB_TRY(up(h, dSiX, Sigma_iX, (size_t)N * p));
    {   // XV is p x N (R's layout): its transpose is the N x p operand of the dot products
        std::vector<double> t((size_t)N * p);
        for (int64_t i = 0; i < N; i++) for (int q = 0; q < p; q++) t[(size_t)q * N + i] = XV[(size_t)i * p + q];
        SGB_TRY(up(h, dXVt, t.data(), (size_t)N * p));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
    }
    // ---- G0: the markers' genotype columns (Get_OneSNP_Geno[_forVarRatio], FG.R:2305-2309) ----
    int64_t *d_rows = reinterpret_cast<int64_t *>(h->d_idx);
    std::vector<int64_t> rows(nmark);
    // workspace: the packed hold-out rows first, then the partial sums of the column sums / dot products
    SGB_TRY(sgb_ensure(h, &h->ws, &h->ws_bytes, std::max((size_t)nmark * Bv, (size_t)2048 * SGB_PART_BLOCKS * sizeof(double))));
    if (from_vr_store) {
        for (int j = 0; j < nmark; j++) {
            CUDA_OK(h, cudaMemcpyAsync((uint8_t *)h->ws + (size_t)j * Bv, h->vr_packed.data() + (size_t)marker_idx[j] * Bv, (size_t)Bv,
                                       cudaMemcpyHostToDevice, h->stream));
            rows[j] = j;
        }
        CUDA_OK(h, cudaMemcpyAsync(d_rows, rows.data(), sizeof(int64_t) * nmark, cudaMemcpyHostToDevice, h->stream));
        SGB_TRY(k_decode_marker_cols(h, (const uint8_t *)h->ws, 0, Bv, d_rows, nmark, dG));
    } else {
        for (int j = 0; j < nmark; j++) {
            rows[j] = -1;
            if (sgb_owner_of(h, marker_idx[j]) == h->rank)
                rows[j] = std::lower_bound(h->loc2glob.begin(), h->loc2glob.end(), marker_idx[j]) - h->loc2glob.begin();
        }
        CUDA_OK(h, cudaMemcpyAsync(d_rows, rows.data(), sizeof(int64_t) * nmark, cudaMemcpyHostToDevice, h->stream));
        SGB_TRY(k_decode_marker_cols(h, h->dG, 1, h->sG, d_rows, nmark, dG));
        if (h->world > 1) SGB_TRY(sgb_allreduce_sum(h, dG, N * nmark));      // the owner's column + zeros
    }
    // ---- flip to the minor allele, AC (FG.R:2318-2322) ----
    std::vector<double> sums(nmark);
    SGB_TRY(k_colsum(h, dG, N, N, nmark, h->d_scal + SC_COLSUM));
    CUDA_OK(h, cudaMemcpyAsync(h->h_scal, h->d_scal + SC_COLSUM, sizeof(double) * nmark, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    std::vector<int> flip(nmark);
    for (int j = 0; j < nmark; j++) {
        sums[j] = h->h_scal[j];
        flip[j] = sums[j] / (double)(2 * N) > 0.5;
        AC[j] = flip[j] ? (double)(2 * N) - sums[j] : sums[j];
    }
    CUDA_OK(h, cudaMemcpyAsync(h->d_idx, flip.data(), sizeof(int) * nmark, cudaMemcpyHostToDevice, h->stream));
    SGB_TRY(k_flip_cols(h, dG, h->d_idx, nmark));
    // ---- G = G0 - XXVX_inv (XV G0)  (FG.R:2333) ----
    std::vector<int> pairs;
    for (int j = 0; j < nmark; j++) for (int q = 0; q < p; q++) { pairs.push_back(q); pairs.push_back(j); }
    std::vector<double> c((size_t)p * nmark);
    SGB_TRY(dots_chunked(h, dXVt, dG, pairs, c.data()));
    SGB_TRY(up(h, dC, c.data(), (size_t)p * nmark));
    SGB_TRY(k_project(h, dG, dXX, p, dC, nmark, dG));
    // ---- Sigma^-1 G: one nmark-column solve (getSigma_G per marker in the reference, FG.R:2341) ----
    SGB_TRY(sgb_pcg_device(h, dw, tau, dG, nmark, maxiterPCG, tolPCG, 0, dSiG, nullptr));
    // ---- var1, var2null (FG.R:2344-2345, 2367-2371) ----
    std::vector<double> gsg(nmark), gsx((size_t)p * nmark), xsg((size_t)p * nmark), xsx((size_t)p * p), v2(nmark);
    pairs.clear();
    for (int j = 0; j < nmark; j++) { pairs.push_back(j); pairs.push_back(j); }
    SGB_TRY(dots_chunked(h, dG, dSiG, pairs, gsg.data()));                   // G_j . Sigma_iG_j
    if (mu2) {
        SGB_TRY(k_rowscale_cols(h, dmu2, dG, nmark, dT));
        SGB_TRY(dots_chunked(h, dT, dG, pairs, v2.data()));                  // sum mu2 G_j^2
    } else SGB_TRY(dots_chunked(h, dG, dG, pairs, v2.data()));
    pairs.clear();
    for (int j = 0; j < nmark; j++) for (int q = 0; q < p; q++) { pairs.push_back(q); pairs.push_back(j); }
    SGB_TRY(dots_chunked(h, dSiX, dG, pairs, gsx.data()));                   // Sigma_iX_q . G_j
    SGB_TRY(dots_chunked(h, dX, dSiG, pairs, xsg.data()));                   // X_q . Sigma_iG_j
    pairs.clear();
    for (int r = 0; r < p; r++) for (int q = 0; q < p; q++) { pairs.push_back(q); pairs.push_back(r); }
    SGB_TRY(dots_chunked(h, dX, dSiX, pairs, xsx.data()));                   // (X' Sigma_iX)[q, r]
    if (!solve_pp(xsx, p, xsg, nmark)) return sgb_fail(h, "variance ratio: t(X) Sigma_iX is singular");
    for (int j = 0; j < nmark; j++) {
        double corr = 0.0;
        for (int q = 0; q < p; q++) corr += gsx[q + (size_t)j * p] * xsg[q + (size_t)j * p];
        var1[j] = (gsg[j] - corr) / AC[j];
        var2null[j] = v2[j] / AC[j];
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// device-resident benchmark hook
// ---------------------------------------------------------------------------------------------------
extern "C" int sgb_bench_crossprod_device(sgb_ctx *h, int k, int reps, uint64_t seed, float *ms_out, float *ms_kernel_out)
{
    NEED_LOADED(h);
    CUDA_OK(h, cudaSetDevice(h->device));
    const int64_t N = h->N;
    SGB_TRY(sgb_ensure_f64(h, &h->d_bench, &h->bench_elems, (size_t)2 * N * k));
    double *dB = h->d_bench, *dY = dB + (size_t)N * k;
    SGB_TRY(k_rademacher_fill(h, dB, N * k, seed));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < reps; r++) {
        h->time_sweeps = true;
        CUDA_OK(h, cudaEventRecord(e0, h->stream));
        int rc = sgb_crossprod_device(h, dB, k, dY, 0);
        if (rc) { h->time_sweeps = false; cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
        CUDA_OK(h, cudaEventRecord(e1, h->stream));
        CUDA_OK(h, cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms_out[r], e0, e1);
        if (ms_kernel_out) { ms_kernel_out[2 * r] = h->last_sweep_ms[0]; ms_kernel_out[2 * r + 1] = h->last_sweep_ms[1]; }
    }
    h->time_sweeps = false;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

extern "C" int sgb_bench_fetch_result(sgb_ctx *h, int k, double *Y, double *B)
{
    NEED_LOADED(h);
    const int64_t N = h->N;
    if (!h->d_bench || h->bench_elems < (size_t)2 * N * k) return sgb_fail(h, "no benchmark result of that width");
    if (B) SGB_TRY(down(h, B, h->d_bench, (size_t)N * k));
    if (Y) SGB_TRY(down(h, Y, h->d_bench + (size_t)N * k, (size_t)N * k));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
