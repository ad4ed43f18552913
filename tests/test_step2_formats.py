"""Step-2 inputs other than PLINK: VCF (GT / DS) and BGEN v1.2 read on the host (saige_gpu_b200/genoio.py), tested as dosage
rows by `sgb_step2_test_dosages`.

Pins: the reference ships the same 100 markers as .bed, .vcf.gz and .bgen together with the result tables its step 2
produced from the VCF copy (LOCO off) and from the BGEN copy (AlleleOrder alt-first on a ref-first file: alleles exchanged).
CPU: the readers give the .bed's genotypes bit for bit, and the driver (device calls answered by the oracle) turns the real
files into those tables.  GPU: the dosage kernel against the oracle on fractional dosages (three imputation methods, zeroing
of small dosages, flips, the exact test, Firth), the real files through the kernel, and dosage rows of hard calls against
the 2-bit kernel."""
import os

import numpy as np
import pytest

from conftest import OracleDevice

TOL_PRINT = 2e-5          # the golden tables print 6 significant digits
NUMERIC = ["AC_Allele2", "AF_Allele2", "MissingRate", "BETA", "SE", "Tstat", "var", "p.value", "p.value.NA", "AF_case", "AF_ctrl",
           "N_case", "N_ctrl", "N_case_hom", "N_case_het", "N_ctrl_hom", "N_ctrl_het"]


def bed_genotypes(prefix):
    from oracle import oracle as O
    bed, N0, M0, _ = O.read_bed(prefix)
    B0 = (N0 + 3) // 4
    codes = ((bed.reshape(M0, B0)[:, :, None] >> np.array([0, 2, 4, 6])) & 3).reshape(M0, -1)[:, :N0]
    return np.array([2.0, -1.0, 1.0, 0.0])[codes]            # copies of A1; -1 missing


def test_readers_reproduce_the_bed(golden_dir):
    from saige_gpu_b200 import genoio
    g = bed_genotypes(os.path.join(golden_dir, "step2_100markers"))
    fam = [l.split()[1] for l in open(os.path.join(golden_dir, "step2_100markers.fam"))]
    bim = [l.split() for l in open(os.path.join(golden_dir, "step2_100markers.bim"))]
    vcf = os.path.join(golden_dir, "step2_100markers.vcf.gz")
    assert genoio.vcf_samples(vcf) == fam
    chunks = list(genoio.iter_vcf(vcf, "GT", chunk=33))
    assert [len(i) for i, _ in chunks] == [33, 33, 33, 1]
    info = [x for i, _ in chunks for x in i]
    D = np.vstack([d for _, d in chunks])
    assert np.array_equal(D, g)                                # ALT copies of the VCF = A1 copies of the .bed
    assert [(c, p, i) for c, p, i, _, _ in info] == [(b[0], b[3], b[1]) for b in bim]
    assert [(r, a) for _, _, _, r, a in info] == [(b[5], b[4]) for b in bim]          # REF = A2, ALT = A1
    with pytest.raises(ValueError):
        next(genoio.iter_vcf(vcf, "DS"))                       # the file has no DS field
    bg = genoio.BgenFile(os.path.join(golden_dir, "step2_100markers.bgen"))
    assert (bg.M, bg.N, bg.layout, bg.compression, bg.samples) == (100, len(fam), 2, 1, None)
    i2, D2 = next(bg.variants("ref-first", chunk=1000))
    assert np.array_equal(D2, g) and i2 == info                # second allele of the BGEN = A1 of the .bed
    bg = genoio.BgenFile(os.path.join(golden_dir, "step2_100markers.bgen"))
    i3, D3 = next(bg.variants("alt-first", chunk=1000))
    assert np.array_equal(D3, 2 - g) and [(r, a) for _, _, _, r, a in i3] == [(a, r) for _, _, _, r, a in info]
    # missing calls: ./. in the VCF = ploidy-byte flag in the BGEN
    _, Dv = next(genoio.iter_vcf(os.path.join(golden_dir, "missing_10markers.vcf.gz"), "GT"))
    _, Db = next(genoio.BgenFile(os.path.join(golden_dir, "missing_10markers.bgen")).variants("ref-first"))
    assert np.array_equal(Dv, Db) and (Dv < 0).sum() == 6
    for f in ("step2_100markers.bgen", "missing_10markers.bgen"):          # native reader == Python reader on the reference's files
        for order in ("ref-first", "alt-first"):
            ia, da = next(genoio.BgenNative(os.path.join(golden_dir, f)).variants(order, chunk=1000))
            ib, db = next(genoio.BgenFile(os.path.join(golden_dir, f)).variants(order, chunk=1000))
            assert ia == ib and np.array_equal(da, db)
    _, Dd = next(genoio.iter_vcf(os.path.join(golden_dir, "dosage_10markers.vcf.gz"), "DS"))
    assert Dd.shape == (10, 1000) and Dd.min() == 0 and Dd.max() == 2
    # hard calls pack back into the raw PLINK rows of the .bed
    raw = np.fromfile(os.path.join(golden_dir, "step2_100markers.bed"), dtype=np.uint8)[3:]
    assert np.array_equal(genoio.hardcalls_to_bed_rows(D), raw)
    with pytest.raises(ValueError):
        genoio.hardcalls_to_bed_rows(np.array([[0.0, 0.4]]))


def write_bgen(path, D1, bits=8, compress=True, sample_ids=None, missing=None):
    """A BGEN v1.2 layout-2 file (unphased diploid biallelic) from first-allele dosages that are exactly representable:
    D1[m, i] in {0, 1, 2} or, with fractions, P(AA) = 0, P(AB) = d for d <= 1 / P(AA) = d - 1, P(AB) = 2 - d for d > 1."""
    import struct
    import zlib
    nm, n = D1.shape
    scale = (1 << bits) - 1
    flags = (1 if compress else 0) | (2 << 2) | ((1 << 31) if sample_ids else 0)
    blocks = b""
    if sample_ids:
        body = b"".join(struct.pack("<H", len(x)) + x.encode() for x in sample_ids)
        blocks = struct.pack("<II", 8 + len(body), n) + body
    var = b""
    for m in range(nm):
        d = D1[m]
        paa = np.where(d > 1, d - 1, 0.0)
        pab = np.where(d > 1, 2 - d, d)
        pm = np.full(n, 2, dtype=np.uint8)
        if missing is not None:
            pm[missing[m]] |= 0x80
        pr = np.empty(2 * n, dtype="<u2" if bits == 16 else np.uint8)
        pr[0::2] = np.rint(paa * scale)
        pr[1::2] = np.rint(pab * scale)
        if missing is not None:
            pr[0::2][missing[m]] = 0
            pr[1::2][missing[m]] = 0
        data = struct.pack("<IHBB", n, 2, 2, 2) + pm.tobytes() + bytes([0, bits]) + pr.tobytes()
        ident = ("v%d" % m).encode()
        head = (struct.pack("<H", len(ident)) + ident) * 2 + struct.pack("<H", 1) + b"7" + struct.pack("<IH", 100 + m, 2)
        head += struct.pack("<I", 1) + b"G" + struct.pack("<I", 1) + b"T"
        if compress:
            z = zlib.compress(data)
            var += head + struct.pack("<II", len(z) + 4, len(data)) + z
        else:
            var += head + struct.pack("<I", len(data)) + data
    header = struct.pack("<IIII", 20 + len(blocks), 20, nm, n) + b"bgen" + struct.pack("<I", flags)
    open(path, "wb").write(header + blocks + var)


@pytest.mark.parametrize("bits,compress,with_ids", [(8, True, False), (16, True, True), (16, False, True), (8, False, False)])
def test_bgen_variants_of_the_format(tmp_path, bits, compress, with_ids):
    """16-bit probabilities, uncompressed blocks, embedded sample identifiers, fractional dosages, missing samples."""
    from saige_gpu_b200 import genoio
    rng = np.random.default_rng(bits + compress)
    n, nm = 57, 9
    scale = (1 << bits) - 1
    D1 = rng.integers(0, 2 * scale + 1, size=(nm, n)) / scale                 # first-allele dosages on the file's grid
    missing = rng.uniform(size=(nm, n)) < 0.05
    ids = ["id%d" % i for i in range(n)] if with_ids else None
    path = str(tmp_path / "t.bgen")
    write_bgen(path, D1, bits, compress, ids, missing)
    bg = genoio.BgenFile(path)
    assert (bg.M, bg.N, bg.compression, bg.samples) == (nm, n, int(compress), ids)
    got = [d for _, d in bg.variants("alt-first", chunk=4)]
    assert [len(d) for d in got] == [4, 4, 1]
    A = np.vstack(got)
    assert np.array_equal(A < 0, missing) and np.allclose(A[~missing], D1[~missing], rtol=0, atol=2e-16 * 4)
    info, B = next(genoio.BgenFile(path).variants("ref-first", chunk=100))
    assert np.allclose(B[~missing], 2 - D1[~missing], rtol=0, atol=1e-15) and info[0] == ("7", "100", "v0", "G", "T")
    # imputation INFO score over a subset of samples (BGEN.cpp:275-345), checked against the formula written out
    mask = rng.uniform(size=n) < 0.7
    bgp, bgn = genoio.BgenFile(path), genoio.BgenNative(path, n_threads=2)
    for (_, _), (_, _) in zip(bgp.variants("ref-first", chunk=nm, info_for=mask), bgn.variants("ref-first", chunk=nm, info_for=mask)):
        assert np.allclose(bgp.last_info, bgn.last_info, rtol=1e-12, atol=1e-14) and bgp.last_info.shape == (nm,)
        for m in range(nm):
            use = mask & ~missing[m]
            e = D1[m, use]
            paa = np.where(e > 1, e - 1, 0.0)
            f = 4 * paa + np.where(e > 1, 2 - e, e)
            theta = e.sum() / (2 * use.sum())
            assert abs(bgn.last_info[m] - (1 - (f - e * e).sum() / (2 * use.sum() * theta * (1 - theta)))) < 1e-9
        assert np.all(bgn.last_info <= 1 + 1e-12)
    # the library's native reader (multi-threaded inflate / decode) against the pure-Python one: bit for bit
    for order in ("alt-first", "ref-first"):
        nat = genoio.BgenNative(path, n_threads=3)
        assert (nat.M, nat.N, nat.samples) == (nm, n, ids)
        ref_chunks = list(genoio.BgenFile(path).variants(order, chunk=4))
        nat_chunks = list(nat.variants(order, chunk=4))
        assert len(nat_chunks) == len(ref_chunks) == 3
        for (ia, da), (ib, db) in zip(nat_chunks, ref_chunks):
            assert ia == ib and np.array_equal(da, db)
        nat.close()
    open(path, "r+b").write(b"\x00\x00")
    with pytest.raises(Exception):
        list(genoio.BgenFile(path).variants())
    with pytest.raises(ValueError):
        list(genoio.BgenNative(path).variants())           # a damaged offset field: refused, with a message, not read as data


def _compare_with_golden(path, golden_path):
    mine = [l.split("\t") for l in open(path).read().splitlines()]
    gold = [l.split("\t") for l in open(golden_path).read().splitlines()]
    assert len(mine) == len(gold) == 33 and mine[0] == gold[0]
    for a, b in zip(mine[1:], gold[1:]):
        for name, x, y in zip(gold[0], a, b):
            if name in NUMERIC:
                assert abs(float(x) - float(y)) <= TOL_PRINT * abs(float(y)) + 1e-300, (name, a[2], x, y)
            else:
                assert x == y, (name, a[2], x, y)


def _sample_file(golden_dir, tmp_path):
    p = str(tmp_path / "bgen_samples.txt")
    with open(p, "w") as f:
        f.write("".join(l.split()[1] + "\n" for l in open(os.path.join(golden_dir, "step2_100markers.fam"))))
    return p


CASES = [("vcf", dict(vcfField="GT", LOCO=False), "step2_100markers_golden_noLOCO.txt"),
         ("bgen", dict(AlleleOrder="alt-first", LOCO=True), "step2_100markers_golden_flipped.txt"),
         ("bgen", dict(AlleleOrder="ref-first", LOCO=True), "step2_100markers_golden.txt"),
         ("plink", dict(AlleleOrder="ref-first", LOCO=True), "step2_100markers_golden_flipped.txt")]


def _run_case(device, golden_dir, tmp_path, kind, kw, out):
    from saige_gpu_b200 import step2
    p = os.path.join(golden_dir, "step2_100markers")
    src = dict(vcf=dict(vcfFile=p + ".vcf.gz"), bgen=dict(bgenFile=p + ".bgen", sampleFile=_sample_file(golden_dir, tmp_path)),
               plink=dict(bedFile=p + ".bed", bimFile=p + ".bim", famFile=p + ".fam"))[kind]
    return step2.SPAGMMATtest(device, GMMATmodelFile=os.path.join(golden_dir, "example_binary.rda"),
                              varianceRatioFile=os.path.join(golden_dir, "example_binary.varianceRatio.txt"), SAIGEOutputFile=out,
                              chrom="1", min_MAC=20, markers_per_chunk=17, return_rows=False, **src, **kw)


@pytest.mark.parametrize("kind,kw,golden", CASES)
def test_driver_turns_the_reference_files_into_the_reference_tables(golden_dir, tmp_path, kind, kw, golden):
    out = str(tmp_path / "out.txt")
    assert _run_case(OracleDevice(), golden_dir, tmp_path, kind, kw, out) == 32
    _compare_with_golden(out, os.path.join(golden_dir, golden))


COND_NUM = ["BETA_c", "SE_c", "Tstat_c", "var_c", "p.value_c", "p.value.NA_c"]


def _compare_cond_golden(path, golden_path):
    """The reference's conditional table (--condition=1:13:A:C,1:79:A:C).  Tstat_c / BETA_c are differences of nearly equal
    numbers for some rows: compared on the scale of the marginal score; the conditioning marker itself (rs79) is 0 / 0."""
    mine = [l.split("\t") for l in open(path).read().splitlines()]
    gold = [l.split("\t") for l in open(golden_path).read().splitlines()]
    assert len(mine) == len(gold) == 33 and mine[0] == gold[0]
    for a, b in zip(mine[1:], gold[1:]):
        da, db = dict(zip(gold[0], a)), dict(zip(gold[0], b))
        for name in gold[0]:
            x, y = da[name], db[name]
            if name in NUMERIC:
                assert abs(float(x) - float(y)) <= TOL_PRINT * abs(float(y)) + 1e-300, (name, da["MarkerID"], x, y)
            elif name in COND_NUM:
                if da["MarkerID"] == "rs79":
                    continue
                scale = {"Tstat_c": abs(float(db["Tstat"])), "BETA_c": abs(float(db["Tstat"])) / float(db["var_c"])}.get(name, abs(float(y)))
                assert abs(float(x) - float(y)) <= TOL_PRINT * scale + 1e-300, (name, da["MarkerID"], x, y)
            else:
                assert x == y, (name, da["MarkerID"], x, y)


def _run_cond(device, golden_dir, out):
    from saige_gpu_b200 import step2
    p = os.path.join(golden_dir, "step2_100markers")
    return step2.SPAGMMATtest(device, vcfFile=p + ".vcf.gz", vcfField="GT", GMMATmodelFile=os.path.join(golden_dir, "example_binary.rda"),
                              varianceRatioFile=os.path.join(golden_dir, "example_binary.varianceRatio.txt"), SAIGEOutputFile=out,
                              chrom="1", LOCO=True, min_MAC=20, markers_per_chunk=40, return_rows=False,
                              condition="1:13:A:C,1:79:A:C")


def test_conditional_analysis_reproduces_the_reference_table(golden_dir, tmp_path):
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import step2
    out = str(tmp_path / "cond.txt")
    assert _run_cond(OracleDevice(), golden_dir, out) == 32
    _compare_cond_golden(out, os.path.join(golden_dir, "step2_100markers_golden_cond.txt"))
    # the same condition given through the PLINK copy finds the same markers; an unknown marker is an error
    p = os.path.join(golden_dir, "step2_100markers")
    a = step2._find_markers(p + ".bed", p + ".bim", 10000, "", "", "", "alt-first", ["1:13:A:C", "1:79:A:C"])
    b = step2._find_markers("", "", 0, p + ".vcf.gz", "GT", "", "alt-first", ["1:13:A:C", "1:79:A:C"])
    assert set(a) == set(b) == {"1:13:A:C", "1:79:A:C"} and all(np.array_equal(a[k], b[k]) for k in a)
    c = step2._find_markers("", "", 0, "", "", p + ".bgen", "ref-first", ["1:13:A:C", "1:79:A:C"])       # other blocks are skipped, not inflated
    assert set(c) == set(a) and all(np.array_equal(a[k], c[k]) for k in a)
    assert step2._find_markers("", "", 0, "", "", p + ".bgen", "alt-first", ["1:13:A:C"]) == {}             # alleles the other way round
    with pytest.raises(ValueError):
        step2.SPAGMMATtest(OracleDevice(), vcfFile=p + ".vcf.gz", vcfField="GT", GMMATmodelFile=os.path.join(golden_dir, "example_binary.rda"),
                           varianceRatioFile=os.path.join(golden_dir, "example_binary.varianceRatio.txt"), chrom="1", condition="1:13:A:T")
    # the product's host-side factors (step2.condition_factors) against the oracle's
    model = step2.ReadModel(os.path.join(golden_dir, "example_binary.rda"), "1", True)
    pos = np.arange(1000)
    rows = [a["1:13:A:C"][pos], a["1:79:A:C"][pos]]
    f = step2.condition_factors(model, 0.94, rows)
    M = dict(model, trait="binary", varRatio=0.94, XV=(model["X"] * model["mu2"][:, None]).T)
    fo = S2.condition_factors(M, rows)
    assert np.allclose(f["P2"], fo["P2"], rtol=1e-12) and np.allclose(f["VarInv"], fo["VarInv"], rtol=1e-10)
    assert np.allclose(f["Tstat_cond"], fo["Tstat"], rtol=1e-10) and np.allclose(f["XtP2"], model["XXVX_inv"].T @ fo["P2"], rtol=1e-10)


def test_imputed_data_columns_and_info_filter(golden_dir, tmp_path):
    """is_imputed_data / minInfo: hard calls have INFO = 1 (every variant kept, `imputationInfo` column = 1); a file with
    uncertain calls loses its low-INFO variants."""
    from saige_gpu_b200 import step2
    p = os.path.join(golden_dir, "step2_100markers")
    base = dict(GMMATmodelFile=os.path.join(golden_dir, "example_binary.rda"), chrom="1", LOCO=True, min_MAC=20,
                varianceRatioFile=os.path.join(golden_dir, "example_binary.varianceRatio.txt"))
    out = str(tmp_path / "imp.txt")
    rows = step2.SPAGMMATtest(OracleDevice(), bgenFile=p + ".bgen", sampleFile=_sample_file(golden_dir, tmp_path), AlleleOrder="ref-first",
                              is_imputed_data=True, minInfo=0.3, SAIGEOutputFile=out, **base)
    assert len(rows) == 32 and all(abs(r["imputationInfo"] - 1.0) < 1e-12 for r in rows)
    hdr = open(out).readline().rstrip("\n").split("\t")
    assert hdr[7] == "imputationInfo" and "MissingRate" not in hdr and open(out).read().splitlines()[1].split("\t")[7] == "1"
    # uncertain calls: blur half of the variants of a small synthetic file, keep the model's samples in front
    rng = np.random.default_rng(3)
    ids = [l.split()[1] for l in open(p + ".fam")][:1000]
    g = rng.binomial(2, 0.3, size=(12, 1000)).astype(np.float64)
    blur = g.copy()
    blur[::2] = np.clip(np.rint((g[::2] * 0.3 + 0.6 * 0.7) * 255) / 255, 0, 2)           # shrunk towards the mean: low INFO
    path = str(tmp_path / "imp.bgen")
    write_bgen(path, blur, 8, True, ids)
    kept = step2.SPAGMMATtest(OracleDevice(), bgenFile=path, AlleleOrder="alt-first", is_imputed_data=True, minInfo=0.8,
                              **{**base, "min_MAC": 1})
    allv = step2.SPAGMMATtest(OracleDevice(), bgenFile=path, AlleleOrder="alt-first", is_imputed_data=True, minInfo=0.0,
                              **{**base, "min_MAC": 1})
    assert [r["MarkerID"] for r in kept] == ["v%d" % m for m in range(1, 12, 2)] and len(allv) == 12
    assert all(r["imputationInfo"] < 0.8 for r in allv[::2]) and all(r["imputationInfo"] > 0.99 for r in allv[1::2])


def test_variant_sharding_of_dosage_inputs(golden_dir, tmp_path):
    from saige_gpu_b200 import step2
    p = os.path.join(golden_dir, "step2_100markers")
    common = dict(GMMATmodelFile=os.path.join(golden_dir, "example_binary.rda"), vcfFile=p + ".vcf.gz", vcfField="GT",
                  varianceRatioFile=os.path.join(golden_dir, "example_binary.varianceRatio.txt"), chrom="1", min_MAC=20, markers_per_chunk=9)
    full = [r["MarkerID"] for r in step2.SPAGMMATtest(OracleDevice(), **common)]
    parts = [[r["MarkerID"] for r in step2.SPAGMMATtest(OracleDevice(), rank=r, world=3, **common)] for r in range(3)]
    assert sum(parts, []) == full and len(full) == 32
    with pytest.raises(ValueError):
        step2.SPAGMMATtest(OracleDevice(), bedFile=p + ".bed", **common)           # two genotype sources
    with pytest.raises(ValueError):
        step2.SPAGMMATtest(OracleDevice(), impute_method="median", **common)


def test_plink_rows_with_mean_or_minor_imputation_go_through_the_dosage_entry(tmp_path):
    """A model written with rdata.save_rda + a synthetic .bed with missing calls: the driver decodes the rows for the two
    imputation methods that give fractional genotypes; best_guess keeps the raw 2-bit rows."""
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import step2
    from saige_gpu_b200.rdata import RList, save_rda
    from test_step2_rare_exact import rare_variant_set
    n_fam, N = 300, 260
    M, pos, bed, nm = rare_variant_set(41, n_fam, N, identity=False)
    ids = ["s%d" % i for i in range(n_fam)]
    p = str(tmp_path / "syn")
    open(p + ".bed", "wb").write(b"\x6c\x1b\x01" + bed.tobytes())
    open(p + ".fam", "w").write("".join("f %s 0 0 0 -9\n" % i for i in ids))
    open(p + ".bim", "w").write("".join("1\tv%d\t0\t%d\tA\tG\n" % (m, m + 1) for m in range(nm)))
    noK = RList([(k, np.asarray(M[k])) for k in ("XV", "XVX", "XXVX_inv")] + [("XVX_inv", np.linalg.inv(M["XVX"])), ("S_a", M["S_a"]),
                ("XVX_inv_XV", M["XVX_inv_XV"]), ("V", M["mu2"])], r_class=["SA_NULL"])
    save_rda(p + ".rda", {"modglmm": dict([("theta", M["tau"]), ("fitted.values", M["mu"].reshape(-1, 1)), ("residuals", M["res"].reshape(-1, 1)),
                                          ("sampleID", [ids[k] for k in pos]), ("obj.noK", noK), ("y", M["y"]), ("X", M["X"]),
                                          ("traitType", "binary"), ("LOCO", False), ("offset", M["offset"].reshape(-1, 1))])})
    open(p + ".varianceRatio.txt", "w").write("%.15g null 1\n" % M["varRatio"])
    calls = []

    class Dev(OracleDevice):
        def mainMarkerInCPP(self, *a, **k):
            calls.append("2bit")
            return super().mainMarkerInCPP(*a, **k)

        def mainMarkerInCPP_dosage(self, *a, **k):
            calls.append("dosage")
            return super().mainMarkerInCPP_dosage(*a, **k)

    for method in ("best_guess", "mean", "minor"):
        calls.clear()
        rows = step2.SPAGMMATtest(Dev(), p + ".bed", p + ".bim", p + ".fam", p + ".rda", p + ".varianceRatio.txt", LOCO=False,
                                  impute_method=method, markers_per_chunk=100, dosage_zerod_cutoff=0.0)
        assert set(calls) == ({"2bit"} if method == "best_guess" else {"dosage"}) and len(rows) == nm
        M2 = dict(M, varRatio=M["varRatio"])
        for m in (2, 7, 12, 57):                                  # markers with missing calls (m % 5 == 2)
            r = S2.test_marker(M2, S2.plink_marker(bed, n_fam, m, pos), impute_method=method, max_MAC_for_ER=4.0)
            assert abs(rows[m]["p.value"] - r["p_value"]) <= 1e-9 * r["p_value"] and abs(rows[m]["AC_Allele2"] - r["AC_Allele2"]) < 1e-9
    assert rows[2]["MissingRate"] > 0


def test_quantitative_trait_table_layout(tmp_path):
    """openOutfile_single / writeOutfile_single for a quantitative trait (Main.cpp:2392-2425, 2437-2560): no p.value.NA /
    Is.SPA / case-control columns, N at the end; with a condition the five _c columns precede N."""
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import step2
    from saige_gpu_b200.rdata import RList, save_rda
    from test_step2_rare_exact import pack_bed
    rng = np.random.default_rng(8)
    n_fam, N, nm = 240, 200, 40
    pos = rng.permutation(n_fam)[:N]
    G = rng.binomial(2, rng.uniform(0.1, 0.5, size=(nm, 1)), size=(nm, n_fam))
    X = np.column_stack([np.ones(N), rng.normal(size=N)])
    y = X @ np.array([0.2, 0.5]) + rng.normal(size=N)
    mu = X @ np.linalg.lstsq(X, y, rcond=None)[0]
    tau = np.array([0.8, 0.3])
    mu2 = np.full(N, 1 / tau[0])
    XV = (X * mu2[:, None]).T
    XVX = X.T @ XV.T
    XVXi = np.linalg.inv(XVX)
    ids = ["q%d" % i for i in range(n_fam)]
    p = str(tmp_path / "q")
    open(p + ".bed", "wb").write(b"\x6c\x1b\x01" + pack_bed(G).tobytes())
    open(p + ".fam", "w").write("".join("f %s 0 0 0 -9\n" % i for i in ids))
    open(p + ".bim", "w").write("".join("3\tq%d\t0\t%d\tT\tC\n" % (m, 10 * m + 5) for m in range(nm)))
    noK = RList([("XV", XV), ("XVX", XVX), ("XXVX_inv", X @ XVXi), ("XVX_inv", XVXi), ("S_a", (X * (y - mu)[:, None]).sum(0)),
                 ("XVX_inv_XV", (X @ XVXi) * mu2[:, None]), ("V", mu2)], r_class=["SA_NULL"])
    save_rda(p + ".rda", {"modglmm": dict([("theta", tau), ("fitted.values", mu.reshape(-1, 1)), ("residuals", (y - mu).reshape(-1, 1)),
                                          ("sampleID", [ids[k] for k in pos]), ("obj.noK", noK), ("y", y), ("X", X),
                                          ("traitType", "quantitative"), ("LOCO", False)])})
    open(p + ".varianceRatio.txt", "w").write("0.97 null 1\n")
    out = p + ".txt"
    n = step2.SPAGMMATtest(OracleDevice(), p + ".bed", GMMATmodelFile=p + ".rda", varianceRatioFile=p + ".varianceRatio.txt",
                           SAIGEOutputFile=out, LOCO=False, return_rows=False)           # .bim / .fam default to the .bed prefix
    lines = [l.split("\t") for l in open(out).read().splitlines()]
    assert n == nm and lines[0] == ["CHR", "POS", "MarkerID", "Allele1", "Allele2", "AC_Allele2", "AF_Allele2", "MissingRate", "BETA", "SE",
                                    "Tstat", "var", "p.value", "N"]
    M = dict(mu=mu, res=y - mu, mu2=mu2, tau=tau, trait="quantitative", y=y, X=X, XV=XV, XVX=XVX, XXVX_inv=X @ XVXi,
             XVX_inv_XV=(X @ XVXi) * mu2[:, None], S_a=(X * (y - mu)[:, None]).sum(0), varRatio=0.97)
    for m in (0, 17, 39):
        r = S2.test_marker(M, G[m, pos].astype(float))
        row = dict(zip(lines[0], lines[1 + m]))
        assert row["MarkerID"] == "q%d" % m and row["Allele1"] == "C" and row["Allele2"] == "T" and row["N"] == str(N)
        assert abs(float(row["p.value"]) - r["p_value"]) <= 2e-5 * r["p_value"] and abs(float(row["BETA"]) - r["BETA"]) <= 2e-5 * abs(r["BETA"])
    step2.SPAGMMATtest(OracleDevice(), p + ".bed", GMMATmodelFile=p + ".rda", varianceRatioFile=p + ".varianceRatio.txt",
                       SAIGEOutputFile=out, LOCO=False, return_rows=False, condition="3:5:C:T")
    hdr = open(out).readline().rstrip("\n").split("\t")
    assert hdr[12:] == ["p.value", "BETA_c", "SE_c", "Tstat_c", "var_c", "p.value_c", "N"]


def dosage_set(seed, n_file=900, N=800, nm=180):
    """Binary-trait model + fractional dosages: common and rare variants, major-allele-coded ones (flip), missing entries (as
    -1 and as NaN), rare variants whose small dosages get zeroed, carriers enriched among cases for the rare ones."""
    from test_step2_rare_exact import rare_variant_set
    rng = np.random.default_rng(seed)
    M, pos, _, _ = rare_variant_set(seed, n_file, N, identity=False)
    cases = np.nonzero(M["y"] == 1)[0]
    D = np.zeros((nm, n_file))
    for m in range(nm):
        kind = m % 6
        if kind in (0, 1):                                    # common, imputed-looking dosages
            f = rng.uniform(0.05, 0.5)
            g = rng.binomial(2, f, size=n_file).astype(np.float64)
            D[m] = np.clip(g + rng.normal(scale=0.08, size=n_file) * (rng.uniform(size=n_file) < 0.5), 0, 2)
        elif kind == 2:                                       # major allele tested: flip
            g = rng.binomial(2, rng.uniform(0.6, 0.95), size=n_file).astype(np.float64)
            D[m] = np.clip(g - np.abs(rng.normal(scale=0.05, size=n_file)), 0, 2)
        else:                                                 # rare: a few carriers (mostly cases) + dust below the zeroing cutoff
            k = 1 + m % 5
            car = pos[rng.choice(cases, size=k, replace=False)]
            D[m, car] = rng.uniform(0.6, 1.0, size=k) * rng.choice([1.0, 2.0], size=k, p=[0.8, 0.2])
            dust = rng.choice(n_file, size=12, replace=False)
            D[m, dust] = np.maximum(D[m, dust], rng.uniform(0.01, 0.15, size=12))
        if m % 4 == 1:
            miss = rng.choice(n_file, size=9, replace=False)
            D[m, miss[:5]] = -1.0
            D[m, miss[5:]] = np.nan
    return M, pos, D


def test_oracle_dosage_semantics():
    """imputeGenoAndFlip on dosages (UTIL.cpp:58-135): the three imputation values, zeroing only for rare variants."""
    from oracle import step2_oracle as S2
    M, pos, D = dosage_set(31)
    n_zeroed = n_er = 0
    for m in range(D.shape[0]):
        G = np.where(np.isnan(D[m, pos]), -1.0, D[m, pos])
        r0 = S2.test_marker(M, G, max_MAC_for_ER=4.0)
        rz = S2.test_marker(M, G, max_MAC_for_ER=4.0, dosage_zerod_cutoff=0.2, dosage_zerod_MAC_cutoff=10.0)
        assert r0 is not None and rz is not None
        n_er += bool(rz["Is_ER"])
        if rz["AC_Allele2"] != r0["AC_Allele2"]:
            n_zeroed += 1
            assert r0["AC_Allele2"] <= 10 + 1e-9 and rz["AC_Allele2"] < r0["AC_Allele2"]
        if (G < 0).any():
            acs = [S2.test_marker(M, G, impute_method=k)["AC_Allele2"] for k in ("best_guess", "mean", "minor")]
            assert len(set(np.round(acs, 9))) >= 2                 # the imputed value enters the allele count
    assert n_zeroed > 60 and n_er > 20


# ---- through the C ABI ------------------------------------------------------------------------------------------------------
ALL_COLS = (("AC_Allele2", 1), ("AF_Allele2", 2), ("MissingRate", 3), ("BETA", 4), ("SE", 5), ("Tstat", 6), ("var", 7), ("p.value", 8),
            ("p.value.NA", 9), ("Is.SPA", 10), ("AF_case", 11), ("AF_ctrl", 12), ("N_case", 13), ("N_ctrl", 14), ("N_case_hom", 15),
            ("N_case_het", 16), ("N_ctrl_hom", 17), ("N_ctrl_het", 18), ("Is.Firth", 20), ("Firth.converged", 21))


@pytest.mark.gpu
def test_gpu_dosage_kernel_vs_oracle():
    from saige_gpu_b200 import SaigeB200
    M, pos, D = dosage_set(32)
    g, o = SaigeB200(device=0), OracleDevice()
    for dev in (g, o):
        dev.setSAIGEobjInCPP(M, M["varRatio"], 2.0, pos)
        dev.setMaxMACforER(4.0)
    o.M["offset"] = M["offset"]
    for method in (1, 2, 3):
        for firth in (False, True):
            g.setFirth(firth, 0.05, M["offset"], se_from_fit=True)
            o.setFirth(firth, 0.05, M["offset"], se_from_fit=True)
            a = g.mainMarkerInCPP_dosage(D, 0.0, 0.5, 0.15, True, method, 0.2, 10.0)
            b = o.mainMarkerInCPP_dosage(D, 0.0, 0.5, 0.15, True, method, 0.2, 10.0)
            assert np.array_equal(a[:, 0], b[:, 0]) and a[:, 0].all()
            for name, c in ALL_COLS:
                err = np.abs(a[:, c] - b[:, c]) / np.maximum(np.abs(b[:, c]), 1e-300)
                err[(a[:, c] == b[:, c])] = 0.0
                assert np.nanmax(err) <= 1e-6, (method, firth, name, int(np.nanargmax(err)), a[np.nanargmax(err), c], b[np.nanargmax(err), c])
            assert (a[:, 10] == 1).sum() > 10 and (not firth or (a[:, 20] == 1).sum() > 10)
    # zeroing off / filters on (Firth and the exact test off: without zeroing a rare variant can have more than ten
    # fractional carriers, where the product falls back to the saddle point)
    for dev in (g, o):
        dev.setFirth(False)
        dev.setMaxMACforER(-1.0)
    a = g.mainMarkerInCPP_dosage(D, 0.01, 3.0, 0.005, True, 1, 0.0, 0.0)
    b = o.mainMarkerInCPP_dosage(D, 0.01, 3.0, 0.005, True, 1, 0.0, 0.0)
    assert np.array_equal(a[:, 0], b[:, 0]) and 0 < a[:, 0].sum() < len(a)
    t = a[:, 0] == 1
    assert np.allclose(a[t][:, [4, 6, 7, 9]], b[t][:, [4, 6, 7, 9]], rtol=1e-6, atol=0)
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,kw,golden", CASES)
def test_gpu_reference_files_reproduce_the_reference_tables(golden_dir, tmp_path, kind, kw, golden):
    from saige_gpu_b200 import SaigeB200
    g = SaigeB200(device=0)
    out = str(tmp_path / "out.txt")
    assert _run_case(g, golden_dir, tmp_path, kind, kw, out) == 32
    _compare_with_golden(out, os.path.join(golden_dir, golden))
    g.close()


@pytest.mark.gpu
def test_gpu_conditional_analysis(golden_dir, tmp_path):
    """The reference's conditional table through the kernel, then GPU vs oracle on synthetic sets: hard calls (identity and
    indexed sample maps) and fractional dosages, with the exact test and categorical variance ratios switched on."""
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import SaigeB200, step2
    from test_step2_rare_exact import rare_variant_set
    g = SaigeB200(device=0)
    out = str(tmp_path / "cond.txt")
    assert _run_cond(g, golden_dir, out) == 32
    _compare_cond_golden(out, os.path.join(golden_dir, "step2_100markers_golden_cond.txt"))
    cols = [("Tstat", 6), ("var", 7), ("p.value", 8), ("BETA_c", 22), ("SE_c", 23), ("Tstat_c", 24), ("var_c", 25), ("p.value_c", 26),
            ("p.value.NA_c", 27)]

    def check(a, b, skip):
        for name, c in cols:
            ok = np.ones(len(a), bool)
            ok[skip] = False
            scale = np.abs(b[:, 6]) if name == "Tstat_c" else (np.abs(b[:, 6]) / b[:, 25] if name == "BETA_c" else np.abs(b[:, c]))
            err = np.abs(a[:, c] - b[:, c]) / np.maximum(scale, 1e-300)
            err[(a[:, c] == b[:, c]) | ~ok] = 0.0
            assert np.nanmax(err) <= 1e-6, (name, int(np.nanargmax(err)), a[np.nanargmax(err), c], b[np.nanargmax(err), c])
        assert np.sum(np.abs(a[ok][:, 24]) ** 2 / a[ok][:, 25] > 4) > 10          # the conditional SPA ran on some variants

    for identity in (True, False):
        M, pos, bed, nm = rare_variant_set(51 + identity, 900, 800, identity)
        M.update(varRatio=[0.8, 1.05], cateVarRatioMinMACVecExclude=(2, 4.5), cateVarRatioMaxMACVecInclude=(4.5,))
        model = dict(M, res=M["res"], XVX_inv_XV=M["XVX_inv_XV"])
        cond_m = [6, 13]                                                         # markers with 7 minor alleles
        rows = [S2.plink_marker(bed, 900, m, pos) for m in cond_m]
        f = step2.condition_factors(model, M["varRatio"], rows, cate_lo=(2, 4.5), cate_hi=(4.5,))
        o = OracleDevice()
        for dev in (g, o):
            dev.setSAIGEobjInCPP(M, M["varRatio"], 2.0, pos)
            dev.setMaxMACforER(4.0)
            dev.setCondition(f["P2"], f["XtP2"], f["VarInv"], f["Tstat_cond"])
        check(g.mainMarkerInCPP(bed, 900, nm), o.mainMarkerInCPP(bed, 900, nm), cond_m)
    M, pos, D = dosage_set(53)
    rows = [np.where(np.isnan(D[m, pos]), -1.0, D[m, pos]) for m in (0, 6)]
    f = step2.condition_factors(dict(M), M["varRatio"], rows)
    o = OracleDevice()
    for dev in (g, o):
        dev.setSAIGEobjInCPP(M, M["varRatio"], 2.0, pos)
        dev.setMaxMACforER(4.0)
        dev.setCondition(f["P2"], f["XtP2"], f["VarInv"], f["Tstat_cond"])
    check(g.mainMarkerInCPP_dosage(D), o.mainMarkerInCPP_dosage(D), [0, 6])
    g.setCondition()
    assert np.isnan(g.mainMarkerInCPP_dosage(D)[:, 22:28]).all()
    g.close()


@pytest.mark.gpu
def test_gpu_dosage_rows_of_hard_calls_equal_the_two_bit_kernel(golden_dir):
    from saige_gpu_b200 import SaigeB200, step2
    p = os.path.join(golden_dir, "step2_100markers")
    g = SaigeB200(device=0)
    common = dict(GMMATmodelFile=os.path.join(golden_dir, "example_binary.rda"),
                  varianceRatioFile=os.path.join(golden_dir, "example_binary.varianceRatio.txt"), chrom="1", min_MAC=1,
                  is_Firth_beta=True, pCutoffforFirth=0.1)
    a = step2.SPAGMMATtest(g, p + ".bed", p + ".bim", p + ".fam", **common)
    b = step2.SPAGMMATtest(g, vcfFile=p + ".vcf.gz", vcfField="GT", **common)
    assert len(a) == len(b) > 60
    for x, y in zip(a, b):
        assert x["MarkerID"] == y["MarkerID"] and x["Allele2"] == y["Allele2"]
        for k in NUMERIC:
            assert abs(x[k] - y[k]) <= 1e-9 * abs(x[k]) + 1e-300, (x["MarkerID"], k, x[k], y[k])
        assert x["Is.SPA"] == y["Is.SPA"] and x["Is.Firth"] == y["Is.Firth"]
    g.close()


def test_chromosome_argument_is_parsed_like_the_reference():
    """getChromNumber (R/readInGLMM.R:1-20): CHR stripped case-insensitively, digits kept; not 1..22 -> no LOCO refit."""
    from saige_gpu_b200.step2 import chrom_number
    assert [chrom_number(c) for c in ("1", "chr7", "CHR22", "Chr03", 5)] == [1, 7, 22, 3, 5]
    assert [chrom_number(c) for c in ("X", "chrX", "MT", "")] == [0, 0, 0, 0]
