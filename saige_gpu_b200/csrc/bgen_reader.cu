// Host-side BGEN v1.2 reader for step 2: variant blocks -> rows of dosages for sgb_step2_test_dosages.
//
// Replaces BgenClass::setBgenObj / getOneMarker / Parse2 (src/SAIGE/src/BGEN.cpp:25-130, 360-520, 132-345): layout 2,
// zlib-compressed or plain probability blocks, unphased diploid biallelic variants, 8- or 16-bit probabilities.  The
// stored pair is P(AA), P(AB) of the FIRST allele; dosage of the first allele = 2 P(AA) + P(AB) with P = byte / 255
// (BGEN.cpp:198-223, the same lut arithmetic), dosage of the second = 2 - that; missing samples (ploidy byte bit 7) -> -1.
// What is new against the reference: the blocks of a batch are read sequentially and then inflated and decoded by several
// host threads (one variant per task), because at biobank size the inflate is what a dosage scan waits for.
// zlib is resolved with dlopen at first use, so the CUDA library itself does not depend on it.  No statistics here.
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>
#include <thread>
#include <vector>
#include "sgb_internal.h"

struct sgb_bgen {
    FILE *f = nullptr;
    std::string path;
    int64_t n_samples = 0, n_variants = 0, next_variant = 0;
    int compression = 0;
    std::vector<std::string> sample_ids;
};

typedef int (*uncompress_fn)(unsigned char *, unsigned long *, const unsigned char *, unsigned long);
static uncompress_fn g_uncompress = nullptr;

static int load_zlib()
{
    if (g_uncompress) return 0;
    const char *names[] = {"libz.so.1", "libz.so"};
    void *lib = nullptr;
    for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return sgb_fail(nullptr, "bgen: cannot load zlib (libz.so.1): %s", dlerror());
    g_uncompress = (uncompress_fn)dlsym(lib, "uncompress");
    if (!g_uncompress) return sgb_fail(nullptr, "bgen: libz.so.1 lacks uncompress");
    return 0;
}

static bool rd(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n; }
static bool rd_str(FILE *f, int len_bytes, std::string &s)
{
    uint32_t l = 0;
    if (!rd(f, &l, len_bytes)) return false;
    s.resize(l);
    return l == 0 || rd(f, &s[0], l);
}

extern "C" int sgb_bgen_open(const char *path, sgb_bgen **out, int64_t *n_samples, int64_t *n_variants, int *has_sample_ids)
{
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return sgb_fail(nullptr, "bgen: cannot open %s", path);
    uint32_t offset, hlen, M, N, flags;
    char magic[4];
    if (!rd(f, &offset, 4) || !rd(f, &hlen, 4) || !rd(f, &M, 4) || !rd(f, &N, 4) || !rd(f, magic, 4) || hlen < 20 ||
        (memcmp(magic, "bgen", 4) != 0 && memcmp(magic, "\0\0\0\0", 4) != 0)) {
        fclose(f);
        return sgb_fail(nullptr, "bgen: %s is not a BGEN file", path);
    }
    fseek(f, 4 + (long)hlen - 4, SEEK_SET);
    if (!rd(f, &flags, 4)) { fclose(f); return sgb_fail(nullptr, "bgen: %s: truncated header", path); }
    const int compression = flags & 3, layout = (flags >> 2) & 15;
    if (layout != 2) { fclose(f); return sgb_fail(nullptr, "bgen: %s has layout %d (only v1.2 layout 2 is read, as in BGEN.cpp)", path, layout); }
    if (compression > 1) { fclose(f); return sgb_fail(nullptr, "bgen: %s is zstd-compressed: not read", path); }
    if (compression == 1 && load_zlib()) { fclose(f); return 1; }
    sgb_bgen *b = new sgb_bgen();
    b->f = f; b->path = path; b->n_samples = N; b->n_variants = M; b->compression = compression;
    if (flags >> 31) {
        uint32_t blen, n;
        fseek(f, 4 + (long)hlen, SEEK_SET);
        bool ok = rd(f, &blen, 4) && rd(f, &n, 4) && n == N;
        for (uint32_t i = 0; ok && i < n; i++) { std::string s; ok = rd_str(f, 2, s); b->sample_ids.push_back(s); }
        if (!ok) { fclose(f); delete b; return sgb_fail(nullptr, "bgen: %s: bad sample identifier block", path); }
    }
    fseek(f, 4 + (long)offset, SEEK_SET);
    *out = b; *n_samples = N; *n_variants = M; *has_sample_ids = b->sample_ids.empty() ? 0 : 1;
    return 0;
}

extern "C" int sgb_bgen_sample_id(sgb_bgen *b, int64_t i, char *buf, int buflen)
{
    if (!b || i < 0 || i >= (int64_t)b->sample_ids.size()) return sgb_fail(nullptr, "bgen: no sample identifier %lld", (long long)i);
    snprintf(buf, buflen, "%s", b->sample_ids[(size_t)i].c_str());
    return 0;
}

struct bgen_block { std::vector<unsigned char> z; uint32_t raw_len = 0; std::string rsid; };

// one variant: inflate (when compressed) and decode into dst[n_samples]; returns an error text or nullptr
// in_model / info: when info is given, the imputation INFO score over the flagged, non-missing samples (all samples when
// in_model is NULL): theta = sum e / 2n, INFO = 1 - sum(f - e^2) / (2 n theta (1 - theta)) with e = 2 P(AA) + P(AB),
// f = 4 P(AA) + P(AB); 1 when theta is 0 or 1 (BGEN.cpp:275-345)
static const char *decode_block(const sgb_bgen *b, bgen_block &blk, int alt_first, double *dst, std::vector<unsigned char> &scratch,
                                const uint8_t *in_model, double *info)
{
    const unsigned char *p = blk.z.data();
    size_t len = blk.z.size();
    if (b->compression) {
        scratch.resize(blk.raw_len);
        unsigned long dl = blk.raw_len;
        if (g_uncompress(scratch.data(), &dl, blk.z.data(), (unsigned long)blk.z.size()) != 0 || dl != blk.raw_len) return "inflate failed";
        p = scratch.data(); len = dl;
    }
    const size_t N = (size_t)b->n_samples;
    if (len < 10 + N) return "probability block too short";
    uint32_t n; uint16_t k;
    memcpy(&n, p, 4); memcpy(&k, p + 4, 2);
    if (n != N || k != 2) return "probability block inconsistent with the header";
    const unsigned char *pm = p + 8;
    const int phased = p[8 + N], bits = p[9 + N];
    if (phased) return "phased data";
    const unsigned char *pr = p + 10 + N;
    if (bits != 8 && bits != 16) return "probabilities are neither 8 nor 16 bits";
    if (len < 10 + N + 2 * N * (size_t)(bits / 8)) return "probability block too short";
    const double scale = bits == 8 ? 255.0 : 65535.0;
    double sum_e = 0.0, sum_f = 0.0, cnt = 0.0;
    for (size_t i = 0; i < N; i++) {
        if (pm[i] & 0x80) { dst[i] = -1.0; continue; }
        if ((pm[i] & 63) != 2) return "not diploid";
        double paa, pab;
        if (bits == 8) { paa = pr[2 * i] / scale; pab = pr[2 * i + 1] / scale; }
        else { uint16_t a, c; memcpy(&a, pr + 4 * i, 2); memcpy(&c, pr + 4 * i + 2, 2); paa = a / scale; pab = c / scale; }
        const double first = 2.0 * paa + pab;
        dst[i] = alt_first ? first : 2.0 - first;
        if (info && (!in_model || in_model[i])) { sum_e += first; sum_f += (4.0 * paa + pab) - first * first; cnt += 1.0; }
    }
    if (info) {
        const double theta = cnt > 0 ? sum_e / (2.0 * cnt) : 0.0;
        *info = (theta == 0.0 || theta == 1.0) ? 1.0 : 1.0 - sum_f / (2.0 * cnt * theta * (1.0 - theta));
    }
    return nullptr;
}

extern "C" int sgb_bgen_read(sgb_bgen *b, int64_t max_variants, int alt_first, int n_threads, const uint8_t *in_model,
                             double *dosages, double *info_scores, char *info_buf, int64_t info_len, int64_t *n_read)
{
    *n_read = 0;
    if (!b || !b->f) return sgb_fail(nullptr, "bgen: not open");
    const int64_t want = std::min<int64_t>(max_variants, b->n_variants - b->next_variant);
    if (want <= 0) { if (info_len > 0) info_buf[0] = 0; return 0; }
    std::vector<bgen_block> blocks((size_t)want);
    std::string info;
    for (int64_t v = 0; v < want; v++) {                // sequential: identifying data + the raw block
        std::string vid, rsid, chrom, a0, a1;
        uint32_t pos, C, D = 0; uint16_t K;
        bool ok = rd_str(b->f, 2, vid) && rd_str(b->f, 2, rsid) && rd_str(b->f, 2, chrom) && rd(b->f, &pos, 4) && rd(b->f, &K, 2);
        if (ok && rsid == ".") rsid = vid;            // BGEN.cpp: RSID = (rsID == ".") ? snpID : rsID
        if (ok && K != 2) return sgb_fail(nullptr, "bgen: %s: variant %s has %d alleles", b->path.c_str(), rsid.c_str(), (int)K);
        ok = ok && rd_str(b->f, 4, a0) && rd_str(b->f, 4, a1) && rd(b->f, &C, 4);
        if (ok && b->compression) { ok = rd(b->f, &D, 4) && C >= 4; C -= 4; }
        if (!ok) return sgb_fail(nullptr, "bgen: %s: truncated at variant %lld", b->path.c_str(), (long long)(b->next_variant + v));
        blocks[(size_t)v].z.resize(C);
        blocks[(size_t)v].raw_len = b->compression ? D : C;
        blocks[(size_t)v].rsid = rsid;
        if (C && !rd(b->f, blocks[(size_t)v].z.data(), C)) return sgb_fail(nullptr, "bgen: %s: truncated block of %s", b->path.c_str(), rsid.c_str());
        // ref-first (the reader's default, BGEN.cpp:485-486): first allele = REF, tested allele = second
        const std::string &ref = alt_first ? a1 : a0, &alt = alt_first ? a0 : a1;
        info += chrom + "\t" + std::to_string(pos) + "\t" + rsid + "\t" + ref + "\t" + alt + "\n";
    }
    if ((int64_t)info.size() + 1 > info_len) return sgb_fail(nullptr, "bgen: info buffer of %lld bytes is too small (%zu needed)", (long long)info_len, info.size() + 1);
    memcpy(info_buf, info.c_str(), info.size() + 1);
    // parallel: inflate + decode, one variant per task
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, want));
    std::atomic<int64_t> next(0);
    std::vector<std::string> errs((size_t)nt);
    auto work = [&](int t) {
        std::vector<unsigned char> scratch;
        for (;;) {
            const int64_t v = next.fetch_add(1);
            if (v >= want) break;
            const char *e = decode_block(b, blocks[(size_t)v], alt_first, dosages + (size_t)v * (size_t)b->n_samples, scratch,
                                         in_model, info_scores ? info_scores + v : nullptr);
            if (e) { errs[(size_t)t] = std::string(e) + " (variant " + blocks[(size_t)v].rsid + ")"; break; }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    for (auto &e : errs) if (!e.empty()) return sgb_fail(nullptr, "bgen: %s: %s", b->path.c_str(), e.c_str());
    b->next_variant += want;
    *n_read = want;
    return 0;
}

extern "C" void sgb_bgen_close(sgb_bgen *b)
{
    if (!b) return;
    if (b->f) fclose(b->f);
    delete b;
}
