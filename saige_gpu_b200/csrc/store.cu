// Context life cycle and the packed-genotype store: the B200 replacement of genoClass (FG.cpp:37-1183).
//
// Ingest pipeline (setgeno, FG.cpp:739-1024): raw PLINK rows are streamed to the GPU in chunks; a count kernel
// produces per-marker allele/missing counts over the phenotyped samples; the QC decision is then taken on the host
// in fp32 with the reference's exact expression order (FG.cpp:438-493) so boundary markers cannot flip; kept
// markers owned by this rank are re-packed on the GPU (sample gather + best-guess imputation) straight into the
// marker-major device store; the sample-major copy used by the second sweep is produced by a transpose kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include <fstream>
#include <string>
#include <thread>
#include <vector>
#include "sgb_internal.h"

static std::string g_create_error;

int sgb_fail(sgb_ctx *h, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return 1;
}

extern "C" const char *sgb_last_error(sgb_ctx *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static bool sgb_prof_enabled()
{
    static const bool on = getenv("SGB_PROFILE") && atoi(getenv("SGB_PROFILE")) > 0;
    return on;
}
static int g_prof_depth = 0;
static double prof_now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
sgb_prof_scope::sgb_prof_scope(sgb_ctx *h_, const char *name_) : h(h_), name(name_), t0(0.0), on(sgb_prof_enabled())
{
    if (!on) return;
    cudaStreamSynchronize(h->stream);
    g_prof_depth++;
    t0 = prof_now();
}
sgb_prof_scope::~sgb_prof_scope()
{
    if (!on) return;
    cudaStreamSynchronize(h->stream);
    g_prof_depth--;
    fprintf(stderr, "[sgb] %*s%s: %.3f ms\n", 2 * g_prof_depth, "", name, prof_now() - t0);
}

int sgb_ensure(sgb_ctx *h, void **p, size_t *cur, size_t need)
{
    if (*cur >= need && *p) return 0;
    const bool prof = sgb_prof_enabled();
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    const double t0 = prof ? prof_now() : 0.0;
    const size_t had = *cur;
    if (*p) CUDA_OK(h, cudaFree(*p));
    *p = nullptr; *cur = 0;
    const double t1 = prof ? prof_now() : 0.0;
    size_t want = need + need / 8 + 256;
    CUDA_OK(h, cudaMalloc(p, want));
    *cur = want;
    if (prof) fprintf(stderr, "[sgb] %*s(re)allocation %.1f -> %.1f MB: cudaFree %.3f ms, cudaMalloc %.3f ms\n", 2 * g_prof_depth, "", had / 1048576.0,
                      want / 1048576.0, t1 - t0, prof_now() - t1);
    return 0;
}

static int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

static int create_common(int device, sgb_ctx **out)
{
    if (!out) return sgb_fail(nullptr, "sgb_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return sgb_fail(nullptr, "sgb_create: no CUDA device available (%s); this library has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return sgb_fail(nullptr, "sgb_create: device %d out of range (0..%d)", device, ndev - 1);
    sgb_ctx *h = new sgb_ctx();
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return sgb_fail(nullptr, "cudaSetDevice(%d) failed", device); }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    h->sm_count = prop.multiProcessorCount;
    if (prop.major < 8) { delete h; return sgb_fail(nullptr, "device %d (sm_%d%d) is not supported; built for sm_100a", device, prop.major, prop.minor); }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return sgb_fail(nullptr, "cudaStreamCreate failed"); }
    for (int i = 0; i < 4; i++) cudaEventCreate(&h->ev[i]);
    if (cudaMalloc((void **)&h->d_scal, sizeof(double) * 8192) != cudaSuccess ||
        cudaMalloc((void **)&h->d_idx, sizeof(int) * 8192) != cudaSuccess ||
        cudaMalloc((void **)&h->d_limbsum, sizeof(int32_t) * 16384) != cudaSuccess ||
        cudaMalloc((void **)&h->d_red, sizeof(double) * 2 * 1024 * SGB_PART_BLOCKS) != cudaSuccess ||
        cudaMalloc((void **)&h->d_ticket, sizeof(unsigned int) * 1024) != cudaSuccess ||
        cudaMemset(h->d_ticket, 0, sizeof(unsigned int) * 1024) != cudaSuccess ||
        cudaMallocHost((void **)&h->h_scal, sizeof(double) * 8192) != cudaSuccess) {
        delete h;
        return sgb_fail(nullptr, "scalar scratch allocation failed");
    }
    h->ws_bytes = 0;
    if (sgb_ensure(h, &h->ws, &h->ws_bytes, (size_t)8 << 20)) { std::string m = h->err; delete h; return sgb_fail(nullptr, "%s", m.c_str()); }
    *out = h;
    return 0;
}

extern "C" int sgb_create(int device, sgb_ctx **out) { return create_common(device, out); }

extern "C" int sgb_create_dist(int device, int rank, int world, const void *id128, sgb_ctx **out)
{
    SGB_TRY(create_common(device, out));
    (*out)->rank = 0; (*out)->world = 1;
    if (world > 1) {
        int rc = sgb_dist_init(*out, rank, world, id128);
        if (rc) { g_create_error = (*out)->err; sgb_destroy(*out); *out = nullptr; return rc; }
    }
    return 0;
}

extern "C" int sgb_nccl_unique_id(void *id128)
{
    std::string err;
    int rc = sgb_dist_unique_id(id128, err);
    if (rc) g_create_error = err;
    return rc;
}

static void free_store(sgb_ctx *h)
{
    cudaSetDevice(h->device);
    sgb_dense_free(h);          // a stored GRM belongs to the genotypes it was built from
    void **ptrs[] = {(void **)&h->dG, (void **)&h->dGt, (void **)&h->d_f2, (void **)&h->d_s, (void **)&h->d_s2,
                     (void **)&h->d_diag, (void **)&h->d_diag_loco};
    for (auto p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
    h->diag_loco_elems = 0;
    h->loaded = false; h->diag_ready = false; h->diag_loco_ready = false; h->ku_cols = 0;
    h->Mloc = h->M = h->Mvr = 0;
}

extern "C" void sgb_destroy(sgb_ctx *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    sgb_dist_destroy(h);
    sgb_step2_free(h);
    free_store(h);
    void *ptrs[] = {h->ws, h->d_acc1, h->d_acc2, h->d_limb, h->d_tmp, h->d_scal, h->d_io, h->d_bench, h->d_pcg, h->d_ai, h->d_idx, h->d_limbsum, h->d_ku, h->d_red, h->d_ticket};
    for (auto p : ptrs) if (p) cudaFree(p);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int sgb_set_engine(sgb_ctx *h, int engine)
{
    if (engine < SGB_ENGINE_TENSOR || engine > SGB_ENGINE_IMMA) return sgb_fail(h, "unknown engine %d", engine);
    h->engine = engine;
    h->diag_ready = false; h->diag_loco_ready = false; h->ku_cols = 0;
    return 0;
}
extern "C" int sgb_set_verbose(sgb_ctx *h, int verbose) { h->verbose = verbose; return 0; }
extern "C" int sgb_set_rhs_limbs(sgb_ctx *h, int n)
{
    if (n < 5 || n > 7) return sgb_fail(h, "sgb_set_rhs_limbs: %d digits not supported (5, 6 or 7)", n);
    if (n != h->rhs_limbs) h->ku_cols = 0;          // the cached K.U was computed at another precision
    h->rhs_limbs = n;
    return 0;
}
extern "C" int sgb_set_product_tolerance(sgb_ctx *h, double rel_tol)
{
    if (!(rel_tol >= 0)) return sgb_fail(h, "sgb_set_product_tolerance: bad tolerance");
    int n = 7;
    for (int c = 5; c <= 7; c++) if (ldexp(16.0, -(8 * c - 2)) <= rel_tol) { n = c; break; }
    return sgb_set_rhs_limbs(h, n);
}
extern "C" int sgb_device_sync(sgb_ctx *h) { CUDA_OK(h, cudaSetDevice(h->device)); CUDA_OK(h, cudaStreamSynchronize(h->stream)); return 0; }
extern "C" int sgb_set_min_maf_for_grm(sgb_ctx *h, float v) { h->minMAF = v; return 0; }
extern "C" int sgb_set_max_missing_rate_for_grm(sgb_ctx *h, float v) { h->maxMissing = v; return 0; }
extern "C" int sgb_set_min_mac_variance_ratio(sgb_ctx *h, float mn, float mx, int is)
{
    h->minMACvr = mn; h->maxMACvr = mx; h->isVarRatio = is != 0;
    return 0;
}

// ---- QC decision for one marker, fp32, reference expression order (FG.cpp:438-548) ----------------
struct qc_out { int ac; int mac; int fill; float afreq, invstd; bool passQC, passVR; };

static qc_out qc_marker(const sgb_ctx *h, int64_t N, int alleleCount, int numMissing, bool in_vr_set)
{
    qc_out o;
    float altFreq = alleleCount / float((N - numMissing) * 2);
    float missingRate = numMissing / float(N);
    o.fill = int(roundf(2 * altFreq));
    if (numMissing > 0) alleleCount = alleleCount + o.fill * numMissing;
    altFreq = alleleCount / float(N * 2);
    float maf = std::min(altFreq, 1 - altFreq);
    o.mac = std::min(alleleCount, int(N) * 2 - alleleCount);
    o.passQC = (maf >= h->minMAF && missingRate <= h->maxMissing);
    o.passVR = false;
    if (h->isVarRatio) {
        if (h->maxMACvr != -1) {
            if (o.mac >= h->minMACvr && o.mac < h->maxMACvr) o.passVR = true;
            else if (o.mac >= h->maxMACvr) o.passVR = in_vr_set;
        } else if (o.mac >= h->minMACvr) o.passVR = in_vr_set;
        if (o.passVR) o.passQC = false;
    }
    float Std = sqrtf(2 * altFreq * (1 - altFreq));
    o.invstd = (Std == 0) ? 0.f : 1 / Std;
    o.afreq = altFreq; o.ac = alleleCount;
    return o;
}

static bool owns_raw(const sgb_ctx *h, int64_t m) { return (m / SGB_SHARD_BLOCK) % h->world == h->rank; }     // raw marker m
static bool owns(const sgb_ctx *h, int64_t gidx) { return sgb_owner_of(h, gidx) == h->rank; }                    // QC'd marker

// allocate the device store for Mloc local markers and upload the per-marker fp64 constants
static int alloc_store(sgb_ctx *h)
{
    h->sG = round_up((h->N + 3) / 4, SGB_KSTEP_BYTES);
    h->rowsG = round_up(std::max<int64_t>(h->Mloc, 1), SGB_ROW_ALIGN);
    h->sT = round_up((h->rowsG + 3) / 4, SGB_KSTEP_BYTES);
    h->rowsT = round_up(h->N, SGB_ROW_ALIGN);
    CUDA_OK(h, cudaMalloc((void **)&h->dG, (size_t)h->rowsG * h->sG));
    CUDA_OK(h, cudaMalloc((void **)&h->dGt, (size_t)h->rowsT * h->sT));
    CUDA_OK(h, cudaMemsetAsync(h->dG, 0, (size_t)h->rowsG * h->sG, h->stream));
    CUDA_OK(h, cudaMemsetAsync(h->dGt, 0, (size_t)h->rowsT * h->sT, h->stream));
    CUDA_OK(h, cudaMalloc((void **)&h->d_f2, sizeof(double) * h->rowsG));
    CUDA_OK(h, cudaMalloc((void **)&h->d_s, sizeof(double) * h->rowsG));
    CUDA_OK(h, cudaMalloc((void **)&h->d_s2, sizeof(double) * h->rowsG));
    CUDA_OK(h, cudaMalloc((void **)&h->d_diag, sizeof(double) * h->N));
    std::vector<double> f2(h->rowsG, 0.0), s(h->rowsG, 0.0), s2(h->rowsG, 0.0);
    for (int64_t r = 0; r < h->Mloc; r++) {
        // fp64 definitions the GPU path is graded on (DESIGN.md "parity definition"): exact integer allele count
        double f = (double)h->ac[h->loc2glob[r]] / (double)(2 * h->N);
        double v = 2.0 * f * (1.0 - f);
        f2[r] = 2.0 * f;
        s[r] = v > 0 ? 1.0 / sqrt(v) : 0.0;
        s2[r] = s[r] * s[r];
    }
    CUDA_OK(h, cudaMemcpyAsync(h->d_f2, f2.data(), sizeof(double) * h->rowsG, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h, cudaMemcpyAsync(h->d_s, s.data(), sizeof(double) * h->rowsG, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h, cudaMemcpyAsync(h->d_s2, s2.data(), sizeof(double) * h->rowsG, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    h->cnt.bytes_h2d += 3 * sizeof(double) * h->rowsG;
    return 0;
}

struct chunk_reader {     // yields raw marker rows [m0, m1) either from memory or from a .bed file
    const uint8_t *mem = nullptr;
    FILE *fp = nullptr;
    int64_t B0 = 0;
    std::vector<uint8_t> buf;
    const uint8_t *get(int64_t m0, int64_t m1)
    {
        if (mem) return mem + m0 * B0;
        buf.resize((size_t)(m1 - m0) * B0);
        if (fseeko(fp, 3 + (off_t)m0 * B0, SEEK_SET)) return nullptr;                     // FG.cpp:902
        if (fread(buf.data(), 1, buf.size(), fp) != buf.size()) return nullptr;
        return buf.data();
    }
};

// Host -> device streaming of raw .bed rows through two pinned staging buffers: the host fill of piece i+1 (fread, or
// a multi-thread memcpy from the caller's buffer) overlaps the asynchronous H2D copy of piece i.
struct stager {
    sgb_ctx *h = nullptr;
    uint8_t *pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    size_t cap = 0;
    int cur = 0;
    int init(sgb_ctx *hh, size_t bytes)
    {
        h = hh; cap = bytes;
        for (int i = 0; i < 2; i++) {
            CUDA_OK(h, cudaMallocHost((void **)&pin[i], cap));
            CUDA_OK(h, cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        return 0;
    }
    void destroy()
    {
        for (int i = 0; i < 2; i++) { if (pin[i]) cudaFreeHost(pin[i]); if (ev[i]) cudaEventDestroy(ev[i]); pin[i] = nullptr; ev[i] = nullptr; }
    }
    // four staging threads per rank: eight were measured no faster on one GPU (21.2 against 20.9 GB/s) and slower on two
    // (14.5 against 24.7 GB/s: the ranks' threads then compete for the host's memory bandwidth)
    int host_threads() const { return 4; }
    void par_copy(uint8_t *dst, const uint8_t *src, size_t n) const
    {
        const int nt = n > ((size_t)8 << 20) ? host_threads() : 1;
        if (nt == 1) { memcpy(dst, src, n); return; }
        std::vector<std::thread> th;
        size_t per = (n + nt - 1) / nt;
        for (int t = 0; t < nt; t++) {
            size_t o = t * per, len = o < n ? std::min(per, n - o) : 0;
            if (len) th.emplace_back([=] { memcpy(dst + o, src + o, len); });
        }
        for (auto &x : th) x.join();
    }
    // rows [m0, m1) of the .bed body -> d_dst (device), asynchronously on h->stream
    int push(chunk_reader &rd, int64_t m0, int64_t m1, uint8_t *d_dst)
    {
        const int64_t B0 = rd.B0;
        const size_t total = (size_t)(m1 - m0) * B0;
        for (size_t off = 0; off < total; off += cap) {
            const size_t n = std::min(cap, total - off);
            CUDA_OK(h, cudaEventSynchronize(ev[cur]));            // the previous H2D out of this buffer has finished
            if (rd.mem) par_copy(pin[cur], rd.mem + (size_t)m0 * B0 + off, n);
            else {
                // marker i starts at byte 3 + B0*i of the file (FG.cpp:902); the staging threads pread disjoint slices
                const off_t fo = 3 + (off_t)m0 * B0 + (off_t)off;
                const int fd = fileno(rd.fp), nt = n > ((size_t)8 << 20) ? host_threads() : 1;
                std::vector<std::thread> th;
                std::vector<int> okv(nt, 1);
                const size_t per = (n + nt - 1) / nt;
                uint8_t *dstp = pin[cur];
                for (int t = 0; t < nt; t++) {
                    const size_t o = t * per, len = o < n ? std::min(per, n - o) : 0;
                    if (!len) continue;
                    th.emplace_back([=, &okv] {
                        size_t done = 0;
                        while (done < len) {
                            ssize_t got = pread(fd, dstp + o + done, len - done, fo + (off_t)(o + done));
                            if (got <= 0) { okv[t] = 0; return; }
                            done += (size_t)got;
                        }
                    });
                }
                for (auto &x : th) x.join();
                for (int v : okv) if (!v) return sgb_fail(h, "setgeno: short read of .bed at marker %lld", (long long)m0);
            }
            CUDA_OK(h, cudaMemcpyAsync(d_dst + off, pin[cur], n, cudaMemcpyHostToDevice, h->stream));
            CUDA_OK(h, cudaEventRecord(ev[cur], h->stream));
            h->cnt.bytes_h2d += n;
            cur ^= 1;
        }
        return 0;
    }
};

static int setgeno_impl(sgb_ctx *h, chunk_reader &rd, int64_t N0, int64_t M0, const int32_t *sub, int64_t N,
                        const uint8_t *indicator, int diagOne, const int32_t *vr_idx, int64_t n_vr)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    free_store(h);
    if (N <= 0 || N0 <= 0 || M0 < 0) return sgb_fail(h, "setgeno: bad dimensions N0=%lld M0=%lld N=%lld", (long long)N0, (long long)M0, (long long)N);
    h->kinDiagOne = diagOne != 0;
    h->N0 = N0; h->M0 = M0; h->N = N;
    const int64_t B0 = (N0 + 3) / 4, B = (N + 3) / 4;
    bool identity = (N == N0);
    for (int64_t k = 0; k < N; k++) {
        if (sub[k] < 1 || sub[k] > N0) return sgb_fail(h, "setgeno: subSampleInGeno[%lld]=%d out of range", (long long)k, sub[k]);
        if (sub[k] != k + 1) identity = false;
    }
    std::vector<uint8_t> indmask(B0, 0), in_vr(M0 + 1, 0);
    for (int64_t i = 0; i < N0; i++) if (indicator[i]) indmask[i >> 2] |= (uint8_t)(1u << ((i & 3) << 1));
    for (int64_t j = 0; j < n_vr; j++) if (vr_idx[j] >= 0 && vr_idx[j] < M0) in_vr[vr_idx[j]] = 1;

    uint8_t *d_indmask = nullptr; int32_t *d_sub = nullptr;
    CUDA_OK(h, cudaMalloc((void **)&d_indmask, B0));
    CUDA_OK(h, cudaMalloc((void **)&d_sub, sizeof(int32_t) * N));
    CUDA_OK(h, cudaMemcpy(d_indmask, indmask.data(), B0, cudaMemcpyHostToDevice));
    CUDA_OK(h, cudaMemcpy(d_sub, sub, sizeof(int32_t) * N, cudaMemcpyHostToDevice));

    // ---- pass 1: counts for every raw marker ----
    nvtxRangePushA("setgeno_pass1_count");
    // The rank map is block-cyclic over RAW markers (blocks of SGB_SHARD_BLOCK), so the rows a rank will keep lie inside the
    // same 1/world of the file it counts: every rank reads its blocks ONCE, keeps them on the device for the re-pack, and the
    // (allele count, missing count) vectors meet in one int32 allreduce -- the reference's SPMD ranks each read the whole
    // file (FG.cpp:897-953).  Only when the raw rows do not fit beside the two packed copies they are read a second time.
    const int64_t chunk = std::max<int64_t>(SGB_SHARD_BLOCK, std::min<int64_t>(M0, ((int64_t)256 << 20) / B0) / SGB_SHARD_BLOCK * SGB_SHARD_BLOCK);
    int64_t M0mine = 0;                                                    // raw markers of this rank's blocks
    for (int64_t b0 = 0; b0 < M0; b0 += SGB_SHARD_BLOCK) if (owns_raw(h, b0)) M0mine += std::min<int64_t>(SGB_SHARD_BLOCK, M0 - b0);
    size_t free_b = 0, total_b = 0;
    CUDA_OK(h, cudaMemGetInfo(&free_b, &total_b));
    const size_t raw_mine = (size_t)M0mine * B0;
    const bool keep_raw = raw_mine * 3 + ((size_t)4 << 30) < free_b;      // raw + marker-major + sample-major copies + slack
    uint8_t *d_raw = nullptr; int32_t *d_ac = nullptr, *d_nm = nullptr;
    CUDA_OK(h, cudaMalloc((void **)&d_raw, keep_raw ? std::max<size_t>(raw_mine, 1) : (size_t)chunk * B0));
    CUDA_OK(h, cudaMalloc((void **)&d_ac, sizeof(int32_t) * std::max<int64_t>(M0, 1)));
    CUDA_OK(h, cudaMalloc((void **)&d_nm, sizeof(int32_t) * std::max<int64_t>(M0, 1)));
    if (h->world > 1) {
        CUDA_OK(h, cudaMemsetAsync(d_ac, 0, sizeof(int32_t) * M0, h->stream));
        CUDA_OK(h, cudaMemsetAsync(d_nm, 0, sizeof(int32_t) * M0, h->stream));
    }
    stager st;
    SGB_TRY(st.init(h, (size_t)64 << 20));
    std::vector<int32_t> ac_raw(M0), nm_raw(M0);
    // runs of consecutive owned raw markers inside a chunk (one run = the whole chunk on a single rank)
    auto owned_runs = [&](int64_t m0, int64_t m1, std::vector<std::pair<int64_t, int64_t>> &runs) {
        runs.clear();
        for (int64_t b0 = m0; b0 < m1; b0 += SGB_SHARD_BLOCK) {
            if (!owns_raw(h, b0)) continue;
            const int64_t b1 = std::min(m1, b0 + SGB_SHARD_BLOCK);
            if (!runs.empty() && runs.back().second == b0) runs.back().second = b1; else runs.emplace_back(b0, b1);
        }
    };
    std::vector<std::pair<int64_t, int64_t>> runs;
    std::vector<int64_t> raw_slot(keep_raw ? (size_t)((M0 + SGB_SHARD_BLOCK - 1) / SGB_SHARD_BLOCK) : 0, -1);   // raw block -> first row in d_raw
    int64_t kept = 0;
    for (int64_t m0 = 0; m0 < M0; m0 += chunk) {
        const int64_t m1 = std::min(M0, m0 + chunk);
        owned_runs(m0, m1, runs);
        for (auto &r : runs) {
            uint8_t *dst = keep_raw ? d_raw + (size_t)kept * B0 : d_raw + (size_t)(r.first - m0) * B0;
            int rc = st.push(rd, r.first, r.second, dst);
            if (rc) { st.destroy(); return rc; }
            SGB_TRY(k_count_markers(h, dst, B0, r.second - r.first, d_indmask, d_ac + r.first, d_nm + r.first));
            if (keep_raw) {
                for (int64_t b0 = r.first; b0 < r.second; b0 += SGB_SHARD_BLOCK) raw_slot[(size_t)(b0 / SGB_SHARD_BLOCK)] = kept + (b0 - r.first);
                kept += r.second - r.first;
            }
        }
        if (!keep_raw) CUDA_OK(h, cudaStreamSynchronize(h->stream));      // the chunk buffer is reused
    }
    if (h->world > 1) {
        SGB_TRY(sgb_allreduce_sum_i32(h, d_ac, M0));
        SGB_TRY(sgb_allreduce_sum_i32(h, d_nm, M0));
    }
    CUDA_OK(h, cudaMemcpyAsync(ac_raw.data(), d_ac, sizeof(int32_t) * M0, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaMemcpyAsync(nm_raw.data(), d_nm, sizeof(int32_t) * M0, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    nvtxRangePop();
    // ---- host QC (fp32, reference order) ----
    h->afreq.clear(); h->invstd.clear(); h->mac.clear(); h->ac.clear();
    h->afreq_vr.clear(); h->invstd_vr.clear(); h->mac_vr.clear(); h->ac_vr.clear(); h->index_vr.clear();
    h->qc_mask.assign(M0, 0); h->loc2glob.clear(); h->qc2raw.clear();
    std::vector<int32_t> fill_raw(M0, 0);
    std::vector<int8_t> kind(M0, 0);      // 1 = GRM store, 2 = VR store
    for (int64_t m = 0; m < M0; m++) {
        qc_out q = qc_marker(h, N, ac_raw[m], nm_raw[m], in_vr[m] != 0);
        fill_raw[m] = q.fill;
        if (q.passQC) {
            int64_t gidx = (int64_t)h->afreq.size();
            h->afreq.push_back(q.afreq); h->invstd.push_back(q.invstd); h->mac.push_back(q.mac); h->ac.push_back(q.ac);
            h->qc_mask[m] = 1; kind[m] = 1;
            h->qc2raw.push_back((int32_t)m);
            if (owns_raw(h, m)) h->loc2glob.push_back(gidx);
        }
        if (h->isVarRatio && q.passVR) {
            h->afreq_vr.push_back(q.afreq); h->invstd_vr.push_back(q.invstd); h->mac_vr.push_back(q.mac); h->ac_vr.push_back(q.ac);
            h->index_vr.push_back((int32_t)m); kind[m] = 2;
        }
    }
    h->M = (int64_t)h->afreq.size();
    h->Mvr = (int64_t)h->index_vr.size();
    h->Mloc = (int64_t)h->loc2glob.size();
    SGB_TRY(alloc_store(h));
    h->vr_packed.assign((size_t)h->Mvr * B, 0);

    // ---- pass 2: re-pack kept rows ----
    SGB_RANGE("setgeno_pass2_repack");
    uint8_t *d_vr = nullptr; int32_t *d_rows = nullptr, *d_fill = nullptr;
    CUDA_OK(h, cudaMalloc((void **)&d_rows, sizeof(int32_t) * chunk));
    CUDA_OK(h, cudaMalloc((void **)&d_fill, sizeof(int32_t) * chunk));
    int64_t lrow = 0, vrow = 0;
    std::vector<int32_t> rows, fills, vrows, vfills;
    // A hold-out marker of the variance ratio is needed by every rank (the host-side store is replicated) but its raw row sits
    // on ONE rank's device: the owner re-packs it and the packed rows are summed over the ranks below (zeros elsewhere).
    for (int64_t m0 = 0; m0 < M0; m0 += chunk) {
        int64_t m1 = std::min(M0, m0 + chunk);
        owned_runs(m0, m1, runs);
        for (auto &r : runs) {
            rows.clear(); fills.clear(); vrows.clear(); vfills.clear();
            for (int64_t m = r.first; m < r.second; m++) {
                if (kind[m] == 1) { rows.push_back((int32_t)(m - r.first)); fills.push_back(fill_raw[m]); }
                else if (kind[m] == 2) { vrows.push_back((int32_t)(m - r.first)); vfills.push_back(fill_raw[m]); }
            }
            if (rows.empty() && vrows.empty()) continue;
            const uint8_t *d_chunk;
            if (keep_raw) d_chunk = d_raw + (size_t)raw_slot[(size_t)(r.first / SGB_SHARD_BLOCK)] * B0;
            else {                                     // second read of this run (raw rows too large to keep on the device)
                int rc = st.push(rd, r.first, r.second, d_raw);
                if (rc) { st.destroy(); return rc; }
                d_chunk = d_raw;
            }
            if (!rows.empty()) {
                CUDA_OK(h, cudaMemcpyAsync(d_rows, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, h->stream));
                CUDA_OK(h, cudaMemcpyAsync(d_fill, fills.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, h->stream));
                SGB_TRY(k_repack(h, d_chunk, B0, d_rows, d_fill, (int64_t)rows.size(), d_sub, identity, N, h->dG, lrow, h->sG, 1));
                CUDA_OK(h, cudaStreamSynchronize(h->stream));
                lrow += (int64_t)rows.size();
            }
            if (!vrows.empty()) {
                if (!d_vr) CUDA_OK(h, cudaMalloc((void **)&d_vr, (size_t)std::min<int64_t>(chunk, M0) * B));
                CUDA_OK(h, cudaMemcpyAsync(d_rows, vrows.data(), sizeof(int32_t) * vrows.size(), cudaMemcpyHostToDevice, h->stream));
                CUDA_OK(h, cudaMemcpyAsync(d_fill, vfills.data(), sizeof(int32_t) * vrows.size(), cudaMemcpyHostToDevice, h->stream));
                SGB_TRY(k_repack(h, d_chunk, B0, d_rows, d_fill, (int64_t)vrows.size(), d_sub, identity, N, d_vr, 0, B, 0));
                // position of these hold-out markers in the (replicated) host store: their rank among all kind == 2 markers
                for (size_t j = 0; j < vrows.size(); j++) {
                    const int64_t m = r.first + vrows[j];
                    const int64_t slot = std::lower_bound(h->index_vr.begin(), h->index_vr.end(), (int32_t)m) - h->index_vr.begin();
                    CUDA_OK(h, cudaMemcpyAsync(h->vr_packed.data() + (size_t)slot * B, d_vr + j * B, (size_t)B, cudaMemcpyDeviceToHost, h->stream));
                }
                CUDA_OK(h, cudaStreamSynchronize(h->stream));
                vrow += (int64_t)vrows.size();
            }
        }
    }
    (void)vrow;
    if (h->world > 1 && h->Mvr > 0) {
        // replicate the hold-out store: byte-wise sum over the ranks (every row is non-zero on exactly one rank)
        const size_t nb = h->vr_packed.size(), nw = (nb + 3) / 4;
        int32_t *d_sum = nullptr;
        CUDA_OK(h, cudaMalloc((void **)&d_sum, nw * 4));
        CUDA_OK(h, cudaMemsetAsync(d_sum, 0, nw * 4, h->stream));
        CUDA_OK(h, cudaMemcpyAsync(d_sum, h->vr_packed.data(), nb, cudaMemcpyHostToDevice, h->stream));
        SGB_TRY(sgb_allreduce_sum_i32(h, d_sum, (int64_t)nw));
        CUDA_OK(h, cudaMemcpyAsync(h->vr_packed.data(), d_sum, nb, cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        cudaFree(d_sum);
    }
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    st.destroy();
    cudaFree(d_raw); d_raw = nullptr;               // release the raw copy before the transpose needs its scratch
    SGB_TRY(k_transpose(h));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_ac); cudaFree(d_nm); cudaFree(d_rows); cudaFree(d_fill); cudaFree(d_indmask); cudaFree(d_sub);
    if (d_vr) cudaFree(d_vr);
    h->loaded = true;
    return 0;
}

static int64_t count_lines(const char *path)
{
    std::ifstream f(path);
    if (!f.is_open()) return -1;
    int64_t n = 0;
    std::string junk;
    while (std::getline(f, junk)) n++;          // FG.cpp:767-771, 782-786
    return n;
}

extern "C" int sgb_setgeno(sgb_ctx *h, const char *bed, const char *bim, const char *fam, const int32_t *sub, int64_t n_sub,
                           const uint8_t *indicator, int64_t n_fam, int diagOne, const int32_t *vr_idx, int64_t n_vr)
{
    int64_t N0 = count_lines(fam);
    if (N0 < 0) return sgb_fail(h, "Error! fam file not open! (%s)", fam);         // FG.cpp:762-765 (but as an error)
    int64_t M0 = count_lines(bim);
    if (M0 < 0) return sgb_fail(h, "Error! bim file not open! (%s)", bim);
    if (n_fam != N0) return sgb_fail(h, "setgeno: indicator length %lld != %lld samples in %s", (long long)n_fam, (long long)N0, fam);
    chunk_reader rd;
    rd.fp = fopen(bed, "rb");
    if (!rd.fp) return sgb_fail(h, "Error! bed file not open! (%s)", bed);
    unsigned char magic[3] = {0, 0, 0};
    if (fread(magic, 1, 3, rd.fp) != 3 || magic[0] != 0x6C || magic[1] != 0x1B || magic[2] != 0x01) {
        fclose(rd.fp);
        return sgb_fail(h, "%s is not a SNP-major PLINK .bed (bad magic bytes)", bed);
    }
    rd.B0 = (N0 + 3) / 4;
    int rc = setgeno_impl(h, rd, N0, M0, sub, n_sub, indicator, diagOne, vr_idx, n_vr);
    fclose(rd.fp);
    return rc;
}

extern "C" int sgb_setgeno_mem(sgb_ctx *h, const uint8_t *bed_body, int64_t n_fam, int64_t n_bim, const int32_t *sub,
                               int64_t n_sub, const uint8_t *indicator, int diagOne, const int32_t *vr_idx, int64_t n_vr)
{
    chunk_reader rd;
    rd.mem = bed_body; rd.B0 = (n_fam + 3) / 4;
    return setgeno_impl(h, rd, n_fam, n_bim, sub, n_sub, indicator, diagOne, vr_idx, n_vr);
}

// Bench / test input: raw PLINK .bed rows of the synthetic markers [m0, m1) written to a HOST buffer (generated on the device in
// 256 MB pieces, bit-identical to the oracle's generator).  t0 / t1 hold the thresholds of those markers only.
extern "C" int sgb_synth_bed_rows(sgb_ctx *h, int64_t N, int64_t m0, int64_t m1, uint64_t seed, const uint32_t *t0, const uint32_t *t1,
                                  double miss_rate, uint8_t *out)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    if (N <= 0 || m1 < m0) return sgb_fail(h, "synth_bed_rows: bad dimensions");
    const int64_t nm = m1 - m0, B0 = (N + 3) / 4;
    if (nm == 0) return 0;
    const int64_t piece = std::max<int64_t>(1, ((int64_t)256 << 20) / B0);
    uint32_t *d_t0 = nullptr, *d_t1 = nullptr; uint8_t *d_buf = nullptr, *pin = nullptr;
    CUDA_OK(h, cudaMalloc((void **)&d_t0, sizeof(uint32_t) * nm));
    CUDA_OK(h, cudaMalloc((void **)&d_t1, sizeof(uint32_t) * nm));
    CUDA_OK(h, cudaMalloc((void **)&d_buf, (size_t)std::min(piece, nm) * B0));
    CUDA_OK(h, cudaMallocHost((void **)&pin, (size_t)std::min(piece, nm) * B0));
    CUDA_OK(h, cudaMemcpy(d_t0, t0, sizeof(uint32_t) * nm, cudaMemcpyHostToDevice));
    CUDA_OK(h, cudaMemcpy(d_t1, t1, sizeof(uint32_t) * nm, cudaMemcpyHostToDevice));
    const uint32_t thr = (uint32_t)(miss_rate * 4294967296.0);
    int rc = 0;
    for (int64_t p0 = 0; p0 < nm && !rc; p0 += piece) {
        const int64_t np = std::min(piece, nm - p0);
        // the hash is keyed by the GLOBAL marker index; the threshold arrays are local to [m0, m1)
        rc = k_synth_bed(h, m0 + p0, np, seed, d_t0 - m0, d_t1 - m0, thr, N, d_buf);
        if (rc) break;
        if (cudaMemcpyAsync(pin, d_buf, (size_t)np * B0, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = sgb_fail(h, "synth_bed_rows: copy failed"); break; }
        memcpy(out + (size_t)p0 * B0, pin, (size_t)np * B0);
    }
    cudaFree(d_t0); cudaFree(d_t1); cudaFree(d_buf); cudaFreeHost(pin);
    return rc;
}

extern "C" int sgb_setgeno_synth(sgb_ctx *h, int64_t N, int64_t M0, uint64_t seed, const uint32_t *t0, const uint32_t *t1)
{
    CUDA_OK(h, cudaSetDevice(h->device));
    free_store(h);
    if (N <= 0 || M0 <= 0) return sgb_fail(h, "setgeno_synth: bad dimensions");
    h->kinDiagOne = false;
    h->N0 = N; h->M0 = M0; h->N = N;
    uint32_t *d_t0 = nullptr, *d_t1 = nullptr; int32_t *d_ac = nullptr;
    CUDA_OK(h, cudaMalloc((void **)&d_t0, sizeof(uint32_t) * M0));
    CUDA_OK(h, cudaMalloc((void **)&d_t1, sizeof(uint32_t) * M0));
    CUDA_OK(h, cudaMalloc((void **)&d_ac, sizeof(int32_t) * M0));
    CUDA_OK(h, cudaMemcpy(d_t0, t0, sizeof(uint32_t) * M0, cudaMemcpyHostToDevice));
    CUDA_OK(h, cudaMemcpy(d_t1, t1, sizeof(uint32_t) * M0, cudaMemcpyHostToDevice));
    CUDA_OK(h, cudaMemsetAsync(d_ac, 0, sizeof(int32_t) * M0, h->stream));
    SGB_TRY(k_synth(h, seed, d_t0, d_t1, d_ac));
    std::vector<int32_t> ac_raw(M0);
    CUDA_OK(h, cudaMemcpyAsync(ac_raw.data(), d_ac, sizeof(int32_t) * M0, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    h->afreq.clear(); h->invstd.clear(); h->mac.clear(); h->ac.clear();
    h->afreq_vr.clear(); h->invstd_vr.clear(); h->mac_vr.clear(); h->ac_vr.clear(); h->index_vr.clear();
    h->qc_mask.assign(M0, 0); h->loc2glob.clear(); h->vr_packed.clear(); h->qc2raw.clear();
    bool save_vr = h->isVarRatio; h->isVarRatio = false;
    std::vector<int64_t> rows;
    for (int64_t m = 0; m < M0; m++) {
        qc_out q = qc_marker(h, N, ac_raw[m], 0, false);
        if (!q.passQC) continue;
        int64_t gidx = (int64_t)h->afreq.size();
        h->afreq.push_back(q.afreq); h->invstd.push_back(q.invstd); h->mac.push_back(q.mac); h->ac.push_back(q.ac);
        h->qc_mask[m] = 1;
        h->qc2raw.push_back((int32_t)m);
        if (owns_raw(h, m)) { h->loc2glob.push_back(gidx); rows.push_back(m); }
    }
    h->isVarRatio = save_vr;
    h->M = (int64_t)h->afreq.size(); h->Mvr = 0; h->Mloc = (int64_t)h->loc2glob.size();
    SGB_TRY(alloc_store(h));
    SGB_TRY(sgb_ensure(h, &h->ws, &h->ws_bytes, sizeof(int64_t) * (rows.size() + 1)));
    CUDA_OK(h, cudaMemcpyAsync(h->ws, rows.data(), sizeof(int64_t) * rows.size(), cudaMemcpyHostToDevice, h->stream));
    SGB_TRY(k_synth(h, seed, d_t0, d_t1, nullptr));
    SGB_TRY(k_transpose(h));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_t0); cudaFree(d_t1); cudaFree(d_ac);
    h->loaded = true;
    return 0;
}

// ---- getters -----------------------------------------------------------------------------------------
#define NEED_LOADED(h) do { if (!(h)->loaded) return sgb_fail(h, "genotypes not loaded: call setgeno first"); } while (0)

extern "C" int64_t sgb_get_total_marker(sgb_ctx *h) { return h->M0; }
extern "C" int64_t sgb_get_num_qc_markers(sgb_ctx *h) { return h->M; }
extern "C" int64_t sgb_get_num_local_markers(sgb_ctx *h) { return h->Mloc; }
extern "C" int64_t sgb_get_nnomissing(sgb_ctx *h) { return h->N; }
extern "C" int64_t sgb_get_num_vr_markers(sgb_ctx *h) { return h->Mvr; }
extern "C" int sgb_get_is_var_ratio_geno(sgb_ctx *h) { return h->isVarRatio ? 1 : 0; }
extern "C" int sgb_get_allele_freq_vec(sgb_ctx *h, double *out) { NEED_LOADED(h); for (int64_t i = 0; i < h->M; i++) out[i] = h->afreq[i]; return 0; }
extern "C" int sgb_get_mac_vec(sgb_ctx *h, int32_t *out) { NEED_LOADED(h); std::copy(h->mac.begin(), h->mac.end(), out); return 0; }
extern "C" int sgb_get_allele_count_vec(sgb_ctx *h, int32_t *out) { NEED_LOADED(h); std::copy(h->ac.begin(), h->ac.end(), out); return 0; }
extern "C" int sgb_get_mac_vec_for_var_ratio(sgb_ctx *h, int32_t *out) { NEED_LOADED(h); std::copy(h->mac_vr.begin(), h->mac_vr.end(), out); return 0; }
extern "C" int sgb_get_index_vec_for_var_ratio(sgb_ctx *h, int32_t *out) { NEED_LOADED(h); std::copy(h->index_vr.begin(), h->index_vr.end(), out); return 0; }
extern "C" int sgb_get_qcd_marker_index(sgb_ctx *h, uint8_t *out) { NEED_LOADED(h); std::copy(h->qc_mask.begin(), h->qc_mask.end(), out); return 0; }

// pair-ternary nibble coding of the device store (kernels.cu: sgb_pack4): nibble = A + 3B
static inline int decode_sample(const uint8_t *row, int64_t i)
{
    unsigned n = (row[i >> 2] >> ((i & 2) << 1)) & 15u;
    unsigned b = (n * 11u) >> 5;
    return (i & 1) ? (int)b : (int)(n - 3 * b);
}
static void decode_row(const uint8_t *row, int64_t N, int32_t *out)
{
    for (int64_t i = 0; i < N; i++) out[i] = decode_sample(row, i);
}

// Get_OneSNP_Geno (FG.cpp:223-272).  In a multi-rank run the row lives on one rank; the owner broadcasts it.
extern "C" int sgb_get_one_snp_geno(sgb_ctx *h, int64_t idx, int32_t *out)
{
    NEED_LOADED(h);
    if (idx < 0 || idx >= h->M) return sgb_fail(h, "Get_OneSNP_Geno: index %lld out of range [0,%lld)", (long long)idx, (long long)h->M);
    CUDA_OK(h, cudaSetDevice(h->device));
    const int64_t B = (h->N + 3) / 4;
    std::vector<uint8_t> row(B, 0);
    bool mine = owns(h, idx);
    if (mine) {
        int64_t r = std::lower_bound(h->loc2glob.begin(), h->loc2glob.end(), idx) - h->loc2glob.begin();
        // row r of the tiled store: 64 bytes in every 8 KB (panel, slab) block of its panel
        row.resize((size_t)h->sG);
        CUDA_OK(h, cudaMemcpy2DAsync(row.data(), SGB_KSTEP_BYTES, h->dG + sgb_tiled_off(r, 0, h->sG), SGB_SLAB_BYTES, SGB_KSTEP_BYTES,
                                     (size_t)(h->sG / SGB_KSTEP_BYTES), cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        h->cnt.bytes_d2h += B;
    }
    if (h->world > 1) {
        // sum-allreduce of the decoded row (non-owners contribute zeros)
        std::vector<double> tmp(h->N, 0.0);
        if (mine) for (int64_t i = 0; i < h->N; i++) tmp[i] = decode_sample(row.data(), i);
        SGB_TRY(sgb_ensure_f64(h, &h->d_io, &h->io_elems, (size_t)h->N));
        CUDA_OK(h, cudaMemcpyAsync(h->d_io, tmp.data(), sizeof(double) * h->N, cudaMemcpyHostToDevice, h->stream));
        SGB_TRY(sgb_allreduce_sum(h, h->d_io, h->N));
        CUDA_OK(h, cudaMemcpyAsync(tmp.data(), h->d_io, sizeof(double) * h->N, cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        for (int64_t i = 0; i < h->N; i++) out[i] = (int32_t)tmp[i];
        return 0;
    }
    decode_row(row.data(), h->N, out);
    return 0;
}

extern "C" int sgb_get_one_snp_geno_for_var_ratio(sgb_ctx *h, int64_t idx, int32_t *out)
{
    NEED_LOADED(h);
    if (idx < 0 || idx >= h->Mvr) return sgb_fail(h, "Get_OneSNP_Geno_forVarRatio: index %lld out of range [0,%lld)", (long long)idx, (long long)h->Mvr);
    decode_row(h->vr_packed.data() + (size_t)idx * ((h->N + 3) / 4), h->N, out);
    return 0;
}

extern "C" int sgb_get_one_snp_stdgeno(sgb_ctx *h, int64_t idx, double *out)
{
    std::vector<int32_t> g(h->N > 0 ? h->N : 1);
    SGB_TRY(sgb_get_one_snp_geno(h, idx, g.data()));
    double f = (double)h->ac[idx] / (double)(2 * h->N), v = 2.0 * f * (1.0 - f);
    double s = v > 0 ? 1.0 / sqrt(v) : 0.0;
    for (int64_t i = 0; i < h->N; i++) out[i] = ((double)g[i] - 2.0 * f) * s;
    return 0;
}

// ---- LOCO bookkeeping (FG.cpp:3067-3091) -------------------------------------------------------------
extern "C" int sgb_set_start_end_index_vec(sgb_ctx *h, const int32_t *start, const int32_t *end, int n)
{
    h->startVec.assign(start, start + n);
    h->endVec.assign(end, end + n);
    h->diag_loco_ready = false;
    if (h->loaded && n > 0) {
        // the N x nchr buffer of set_Diagof_StdGeno_LOCO: allocated here, next to setgeno's allocations, so that the first fit
        // does not pay a cudaMalloc between its solves
        cudaSetDevice(h->device);
        size_t bytes = h->diag_loco_elems * sizeof(double);
        SGB_TRY(sgb_ensure(h, (void **)&h->d_diag_loco, &bytes, sizeof(double) * (size_t)h->N * n));
        h->diag_loco_elems = bytes / sizeof(double);
    }
    return 0;
}

extern "C" int sgb_set_start_end_index(sgb_ctx *h, int start, int end, int chromIndex)
{
    NEED_LOADED(h);
    if (start < 0 || end < start || end >= h->M) return sgb_fail(h, "setStartEndIndex: bad range [%d,%d] for %lld markers", start, end, (long long)h->M);
    h->loco_start = start; h->loco_end = end; h->loco_chrom = chromIndex;
    return 0;
}

extern "C" int sgb_get_counters(sgb_ctx *h, sgb_counters *out) { *out = h->cnt; return 0; }
extern "C" int sgb_reset_counters(sgb_ctx *h) { h->cnt = sgb_counters(); return 0; }

extern "C" double sgb_cal_cv(const double *x, int n)
{
    // calCV (FG.cpp:3104-3110): (sd/mean)/n with the n-1 standard deviation
    double mean = 0.0;
    for (int i = 0; i < n; i++) mean += x[i];
    mean /= n;
    double ss = 0.0;
    for (int i = 0; i < n; i++) ss += (x[i] - mean) * (x[i] - mean);
    double sd = n > 1 ? sqrt(ss / (n - 1)) : 0.0;
    return (sd / mean) / n;
}

extern "C" double sgb_inner_product(const double *x, const double *y, int64_t n)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += x[i] * y[i];
    return s;
}
