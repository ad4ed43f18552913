// Internal declarations of libsaige_b200.so (not part of the ABI; the ABI is include/saige_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>     // header-only: ranges are no-ops unless a profiler injects itself (nsys / ncu --nvtx)
#include <stdint.h>
#include <string>
#include <vector>
#include "saige_b200.h"

// NVTX range around a host-side phase (sweeps, collective, PCG iteration, ingest passes): SGB_RANGE("name");
struct sgb_nvtx_range {
    explicit sgb_nvtx_range(const char *name) { nvtxRangePushA(name); }
    ~sgb_nvtx_range() { nvtxRangePop(); }
};
#define SGB_RANGE_CAT2(a, b) a##b
#define SGB_RANGE_CAT(a, b) SGB_RANGE_CAT2(a, b)
#define SGB_RANGE(name) sgb_nvtx_range SGB_RANGE_CAT(nvtx_range_, __LINE__)(name)

// Host-side phase timer, on when the environment has SGB_PROFILE=1: synchronises the stream on entry and exit of the scope and
// prints "[sgb] <name>: <ms>" to stderr (nested scopes indent).  Off: two predictable branches.  tools/profile_step1_host.py reads it.
struct sgb_ctx;
struct sgb_prof_scope {
    sgb_ctx *h; const char *name; double t0; bool on;
    sgb_prof_scope(sgb_ctx *h, const char *name);
    ~sgb_prof_scope();
};
#define SGB_PROF(h, name) sgb_prof_scope SGB_RANGE_CAT(prof_scope_, __LINE__)(h, name)

#define SGB_LIMBS 8          // signed base-128 digits per fp64 value (55-bit fixed point: round(v * 2^(53-E)))
#define SGB_KSTEP_BYTES 64   // packed bytes (256 genotypes) consumed per k-step of the tensor kernel
#define SGB_ROW_ALIGN 512    // row padding of both genotype copies (CTA tile of the tensor kernel)
#define SGB_SHARD_BLOCK 1024 // markers per block of the block-cyclic marker->rank map
#define SGB_PART_BLOCKS 256   // partial-sum slots per column of the deterministic reductions

// TILED genotype store (both copies).  A matrix of `rows` x `stride` packed bytes (rows a multiple of 128, stride a multiple
// of 64) is stored as 128-row PANELS; inside a panel the 64-byte k-slabs (256 genotypes of every row) follow each other, and
// inside a slab the 128 rows' 64 bytes are contiguous: one (panel, slab) block = 8 KB of consecutive addresses.  A CTA of the
// sweep kernels therefore streams long contiguous runs (cp.async.bulk of 16 KB per panel and stage) instead of 64-byte
// segments scattered over 128+ DRAM pages, which is what held the row-major store at 0.84 of the HBM peak.
#define SGB_PANEL_ROWS 128
#define SGB_SLAB_BYTES 8192  // SGB_PANEL_ROWS * SGB_KSTEP_BYTES
__host__ __device__ __forceinline__ int64_t sgb_tiled_off(int64_t row, int64_t byte, int64_t stride)
{
    return (row >> 7) * (stride << 7) + (byte >> 6) * SGB_SLAB_BYTES + ((row & 127) << 6) + (byte & 63);
}

struct sgb_dist;             // NCCL state (dist.cu)
struct sgb_step2;            // step-2 model state (step2.cu)
struct sgb_dense;            // stored dense GRM (dense_grm.cu)

struct sgb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int engine = SGB_ENGINE_TENSOR;
    int verbose = 0;                                  // print the reference's PCG log lines (sgb_set_verbose)
    int sm_count = 148;

    // configuration (FG.cpp:35,99,57-59)
    float minMAF = 0.f, maxMissing = 1.f;
    bool isVarRatio = false;
    float minMACvr = 20.f, maxMACvr = -1.f;
    bool kinDiagOne = false;

    // dimensions
    int64_t N0 = 0, M0 = 0, N = 0, M = 0, Mloc = 0, Mvr = 0;
    bool loaded = false;

    // host-side marker statistics over ALL QC'd markers (every rank has them, like the reference's SPMD ranks)
    std::vector<float> afreq, invstd, afreq_vr, invstd_vr;
    std::vector<int32_t> mac, ac, mac_vr, ac_vr, index_vr;
    std::vector<uint8_t> qc_mask;
    std::vector<uint8_t> vr_packed;          // Mvr x ceil(N/4), device coding (value bits), host resident
    std::vector<int64_t> loc2glob;           // local row -> global QC'd marker index (monotone)
    std::vector<int32_t> qc2raw;             // QC'd marker index -> raw (.bim) marker index: the rank map is defined on RAW markers

    // device genotype store, device coding: 2 bits per genotype = number of A1 copies (0,1,2), sample i of a
    // marker in the pair-ternary nibble coding of kernels.cu (sgb_pack4); all padding is genotype 0.
    // both copies in the TILED layout (sgb_tiled_off): element (row, byte) of a rows x stride matrix
    uint8_t *dG = nullptr;   int64_t sG = 0, rowsG = 0;   // marker-major  rowsG x sG,  rowsG>=Mloc
    uint8_t *dGt = nullptr;  int64_t sT = 0, rowsT = 0;   // sample-major  rowsT x sT,  rowsT>=N
    double *d_f2 = nullptr;  // 2*f_m      per local marker
    double *d_s = nullptr;   // 1/sqrt(2f(1-f)) per local marker
    double *d_s2 = nullptr;  // s_m^2

    // LOCO
    std::vector<int32_t> startVec, endVec;
    int loco_start = -1, loco_end = -1, loco_chrom = -1;
    double *d_diag = nullptr;        // sum_m z_mi^2 over all markers (N), lazily computed (FG.cpp:665-704)
    bool diag_ready = false;
    double *d_diag_loco = nullptr;   // N x nchr: full - per-chromosome (FG.cpp:4934-4958)
    size_t diag_loco_elems = 0;
    std::vector<int64_t> msub_by_chr;
    bool diag_loco_ready = false;

    // scratch
    void *ws = nullptr; size_t ws_bytes = 0;          // generic workspace
    int32_t *d_acc1 = nullptr; size_t acc1_elems = 0; // int32 limb sums of sweep 1  [rowsG][8k]
    int32_t *d_acc2 = nullptr; size_t acc2_elems = 0; // int32 limb sums of sweep 2  [rowsT][8k]
    int8_t *d_limb = nullptr; size_t limb_bytes = 0;  // limb fragments
    double *d_tmp = nullptr; size_t tmp_elems = 0;    // fp64 scratch (c/d vectors etc.)
    double *d_scal = nullptr;                         // small scalar scratch (4096 doubles)
    double *h_scal = nullptr;                         // pinned mirror
    double *d_io = nullptr; size_t io_elems = 0;      // staging for host<->device vectors
    double *d_bench = nullptr; size_t bench_elems = 0;
    double *d_pcg = nullptr; size_t pcg_elems = 0;    // PCG state arena
    double *d_ai = nullptr; size_t ai_elems = 0;      // per-call arena of the AI-REML entry points
    int *d_idx = nullptr;                             // small int scratch (8192 ints)
    // K.U of the Hutchinson probes: U is the same stream in every outer AI-REML iteration (set_seed(200) before each
    // GetTrace, FG.cpp:3114), so the nrun-column product is computed once per (genotypes, GRM mode) and checked bitwise
    double *d_ku = nullptr; size_t ku_elems = 0;      // [U | K.U]
    int64_t ku_cols = 0;                              // 0 = nothing cached
    int probe_stream_fixed = 0;                       // sgb_set_probe_stream_fixed: the cached first batch stands for the callback's
    int32_t *d_limbsum = nullptr;                     // [2][1024][8] column limb sums of the current split
    double *d_red = nullptr;                          // [2][1024][SGB_PART_BLOCKS] per-block partial (sums | maxima) of the fused statistics
    unsigned int *d_ticket = nullptr;                 // [1024] last-block tickets (zero between kernels)
    int rhs_limbs = 7;                                // base-256 digits per right-hand-side value on the tcgen05 engine (sgb_set_rhs_limbs)

    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float last_sweep_ms[2] = {0.f, 0.f};
    bool time_sweeps = false;

    sgb_dist *dist = nullptr;
    int rank = 0, world = 1;
    sgb_step2 *step2 = nullptr;
    sgb_dense *dense = nullptr;
    bool umma_accumulate = false;                     // k_pk2_umma adds to `out` instead of overwriting it
    int grm_mode = SGB_GRM_PACKED;                    // which GRM the products / PCG use

    sgb_counters cnt = {};
};

// ---- error helpers ----
int sgb_fail(sgb_ctx *h, const char *fmt, ...);
#define CUDA_OK(h, call)                                                                           \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) return sgb_fail(h, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
                                                cudaGetErrorString(e__));                          \
    } while (0)
#define SGB_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__) return rc__;         \
    } while (0)

int sgb_ensure(sgb_ctx *h, void **p, size_t *cur, size_t need_bytes);   // grow-only device buffer (sizes in BYTES)
int sgb_ensure_f64(sgb_ctx *h, double **p, size_t *cur_elems, size_t need_elems);   // same, sizes in doubles

// ---- kernels.cu launchers (all on h->stream) ----
int k_count_markers(sgb_ctx *h, const uint8_t *d_bed, int64_t B0, int64_t nmark, const uint8_t *d_indmask,
                    int32_t *d_ac, int32_t *d_nmiss);
int k_repack(sgb_ctx *h, const uint8_t *d_bed, int64_t B0, const int32_t *d_src_rows, const int32_t *d_fill,
             int64_t nrows, const int32_t *d_sub_idx, int identity, int64_t N, uint8_t *d_out, int64_t out_row0, int64_t out_stride,
             int tiled);   // tiled: d_out is a tiled store (sgb_tiled_off), else row-major with out_stride bytes per row
int k_gather_rows(sgb_ctx *h, const uint8_t *P, int64_t stride, const int64_t *d_rows, int nrows, int64_t nbytes, uint8_t *d_out);
int k_transpose(sgb_ctx *h);   // dG -> dGt
int k_synth(sgb_ctx *h, uint64_t seed, const uint32_t *d_t0, const uint32_t *d_t1, int32_t *d_ac);
int k_synth_bed(sgb_ctx *h, int64_t m0, int64_t nm, uint64_t seed, const uint32_t *d_t0, const uint32_t *d_t1, uint32_t miss_thr,
                int64_t N, uint8_t *d_out);

// tensor engine: out[r][c*8+l] += sum_k P[r][k] * L[c][k][l]  (int32, exact)
enum { SGB_PLANE_VALUE = 0, SGB_PLANE_IS2 = 1 };   // which function of the genotype the decode feeds the MMA
int k_pk2_gemm(sgb_ctx *h, const uint8_t *P, int64_t stride, int64_t rows_pad, int64_t kbytes, const int8_t *L,
               int ncol, int32_t *out, int plane);
// limb preparation: V[len x k] (ld) -> fragments [k][nblk][2048] + per-column multiplier (2^(E-53)) in d_mult[k]
int k_split_limbs(sgb_ctx *h, const double *V, int64_t len, int64_t ld, int k, int8_t *L, int64_t nblk, double *d_mult,
                  int32_t *d_limbsum, int have_stats = 0);
// fused product epilogues (nl = 8: mma.sync engine, 5..7: tcgen05 engine with nl digits)
int k_col_stats(sgb_ctx *h, const double *V, int64_t len, int64_t ld, int k, double *d_colsum, int32_t *d_limbsum, int nlimb);
int k_recomb_post1(sgb_ctx *h, int nl, int32_t *acc, int64_t rows_pad, int k, int pad, const double *d_mult, const int32_t *d_limbsum,
                   const double *d_colsum, int64_t mask_lo, int64_t mask_hi, double *D, int64_t ld, double *d_t, double *d_t2,
                   int32_t *d_limbsum2, int nlimb2);
int k_recomb_post2(sgb_ctx *h, int nl, int32_t *acc, int64_t rows_pad, int k, int pad, const double *d_mult, const int32_t *d_limbsum,
                   const double *d_t, double inv_m, double *Y, int64_t ldy, double *raw, int64_t ldr);
// raw[r + c*ld] = recombine(acc[r][c*8..]) * mult[c]; acc zeroed
int k_recombine(sgb_ctx *h, int32_t *acc, int64_t rows, int k, int kpad, const double *d_mult, const int32_t *d_limbsum,
                int plane, double *raw, int64_t ld);
// tcgen05 path (pk2_umma.cu)
size_t k_umma_limb_bytes(int k, int64_t kbytes);
size_t k_umma_image_bytes(int nrows, int64_t kbytes);   // image of `nrows` accumulator columns (dense-GRM panels: 128)
int k_pk2_umma_rows(sgb_ctx *h, const uint8_t *P, int64_t stride, int64_t rows_pad, int64_t kbytes, const int8_t *L, int nrows,
                    int32_t *out, int plane);
int k_recombine_umma(sgb_ctx *h, int32_t *acc, int64_t rows, int k, const double *d_mult, const int32_t *d_limbsum, int plane,
                     double *raw, int64_t ld);
int k_split_limbs_umma(sgb_ctx *h, const double *V, int64_t len, int64_t ld, int k, int8_t *L, int64_t kbytes, double *d_mult,
                       int32_t *d_limbsum, int nl = 7, int have_stats = 0);
int k_pk2_umma(sgb_ctx *h, const uint8_t *P, int64_t stride, int64_t rows_pad, int64_t kbytes, const int8_t *L, int k,
               int32_t *out, int plane, int nl = 7);
int k_umma_npad(int k, int nl);

// f64 engine
int k_rowdot_f64(sgb_ctx *h, const double *B, int64_t ldb, int k, double *out, int64_t ldo);   // out[m,c]=sum_i g_mi B[i,c]
int k_coldot_f64(sgb_ctx *h, const double *D1, const double *D2, int64_t ldd, int k, double *out, int64_t ldo);

// algebra around the sweeps
int k_colsum(sgb_ctx *h, const double *V, int64_t len, int64_t ld, int k, double *d_out);   // deterministic
int k_sweep1_post(sgb_ctx *h, const double *raw, int64_t ld, int k, const double *d_colsum, int64_t mask_lo,
                  int64_t mask_hi, double *D, double *d_t);  // D = s^2 (raw - 2f*sumb) (masked), t = sum 2f D
int k_sweep2_post(sgb_ctx *h, const double *raw, int64_t ldr, int k, const double *d_t, double inv_m, double *Y,
                  int64_t ldy);
int k_diag_prep(sgb_ctx *h, int nchr, const int64_t *h_lo, const int64_t *h_hi, double *D1, double *D2, double *d_const);
int k_diag_post(sgb_ctx *h, const double *raw1, const double *raw2, int64_t ld, int ncol, const double *d_const,
                double *out, int64_t ldo);

// PCG / BLAS-1 (N x k column-major, ld = N)
int k_sigma_diag(sgb_ctx *h, const double *diag, double diag_scale, int diag_one, const double *w, double tau0,
                 double tau1, double *out);
int k_pcg_init(sgb_ctx *h, const double *B, const double *minv, int k, double *X, double *R, double *Z, double *P,
               double *d_rz, double *d_r2);
int k_gather_cols(sgb_ctx *h, const double *src, const int *d_cols, int ncols, double *dst);
int k_scatter_cols(sgb_ctx *h, const double *src, const int *d_cols, int ncols, double *dst);
int k_pcg_step1(sgb_ctx *h, const double *P, double *KP, const double *w, double tau0, double tau1, const int *d_act,
                int nact, double *d_part);   // KP := Ap (packed);   part[j][blk] = partial p.Ap
int k_pcg_step2(sgb_ctx *h, const double *P, const double *AP, const double *minv, const int *d_act, int nact,
                double *X, double *R, double *Z, double *d_rz, const double *d_part_in, double *d_part_out);
int k_pcg_step3(sgb_ctx *h, double *P, const double *Z, const int *d_act, int nact, const double *rz_in, double *rz_out,
                double *d_r2, const double *d_part);
int k_pair_dots(sgb_ctx *h, const double *A, int64_t lda, const double *B, int64_t ldb, const int *d_pairs, int npairs,
                double *d_out);   // d_out[q] = A[:,pairs[2q]] . B[:,pairs[2q+1]]   (deterministic)
int k_project(sgb_ctx *h, const double *In, const double *SiX, int p, const double *d_C, int ncol, double *Out);
int k_eta(sgb_ctx *h, const double *Y, const double *SiY, const double *SiX, int p, const double *d_alpha,
          const double *w, double tau0, double *eta);
int k_irls_update(sgb_ctx *h, int family, const double *eta_in, int add_offset, const double *y, const double *offset, double *eta_out,
                  double *mu, double *Y, double *W);
int k_decode_marker_cols(sgb_ctx *h, const uint8_t *P, int tiled, int64_t stride, const int64_t *d_rows, int ncol, double *Out);
int k_flip_cols(sgb_ctx *h, double *G, const int *d_flip, int ncol);
int k_rowscale_cols(sgb_ctx *h, const double *v, const double *In, int ncol, double *Out);
int k_rademacher_fill(sgb_ctx *h, double *B, int64_t n, uint64_t seed);
int k_axpby(sgb_ctx *h, double a, const double *x, double b, const double *y, int64_t n, double *out);
int k_count_diff(sgb_ctx *h, const double *a, const double *b, int64_t n, int *d_count);
int k_grid_blocks(sgb_ctx *h, int64_t n);

// cudaFuncSetAttribute acts on the current device: every launch site that raises a kernel's dynamic shared-memory limit
// does so once per device ordinal (a process may hold handles on several devices), not once per process
enum sgb_attr_site { SGB_SITE_REPACK = 0, SGB_SITE_UMMA, SGB_SITE_STEP2, SGB_SITE_STREAM1, SGB_SITE_STREAM2, SGB_SITE_POST1_5, SGB_SITE_POST1_6, SGB_SITE_POST1_7, SGB_SITE_POST2_5, SGB_SITE_POST2_6, SGB_SITE_POST2_7, SGB_SITE_SPLIT_UMMA, SGB_SITE_SYMV_BASE /* + KC, KC <= 8 */, SGB_SITE_COUNT = SGB_SITE_SYMV_BASE + 9 };
inline bool sgb_first_on_device(int device, int site)
{
    static unsigned char done[SGB_SITE_COUNT][64];
    if (device < 0 || device >= 64) return true;
    if (done[site][device]) return false;
    done[site][device] = 1;
    return true;
}

// solver.cu / dist.cu / step2.cu / dense_grm.cu
int sgb_crossprod_device(sgb_ctx *h, const double *dB, int k, double *dY, int loco);
int sgb_gt_times_cols(sgb_ctx *h, const double *D, int k, double *raw);   // raw (ld rowsT) = G^T D (ld rowsG), tensor engine
int sgb_diag_device(sgb_ctx *h);
int sgb_diag_loco_device(sgb_ctx *h);
int sgb_pcg_device(sgb_ctx *h, const double *d_w, const double *tau, const double *dB, int k, int maxiter, double tol,
                   int loco, double *dX, int32_t *iters);
// rank that stores QC'd marker gidx: raw markers are dealt block-cyclically (SGB_SHARD_BLOCK), so a rank's rows of BOTH ingest
// passes come from the same 1/world of the file and chromosomes stay balanced over the ranks
inline int sgb_owner_of(const sgb_ctx *h, int64_t gidx) { return (int)(((int64_t)h->qc2raw[(size_t)gidx] / SGB_SHARD_BLOCK) % h->world); }
int sgb_allreduce_sum(sgb_ctx *h, double *d, int64_t n);
int sgb_allreduce_sum_i32(sgb_ctx *h, int32_t *d, int64_t n);
int sgb_dist_init(sgb_ctx *h, int rank, int world, const void *id128);
void sgb_dist_destroy(sgb_ctx *h);
int sgb_dist_unique_id(void *id128, std::string &err);
void sgb_step2_free(sgb_ctx *h);
void sgb_dense_free(sgb_ctx *h);
int sgb_dense_crossprod_device(sgb_ctx *h, const double *dB, int k, double *dY);
int sgb_broadcast_bytes(sgb_ctx *h, void *d, size_t bytes, int root);
