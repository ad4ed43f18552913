"""Step-2 single-variant score test + SPA pinned on the reference's OWN golden output:
extdata/output/genotype_100markers_marker_plink.txt (32 variants x 23 columns), produced by the reference from
extdata/output/example_binary.rda + example_binary.varianceRatio.txt + extdata/input/genotype_100markers.{bed,bim,fam}
(all committed under tests/golden/ by make_golden.py).  The table prints 6 significant digits, hence 2e-5.  Two more of the reference's tables on the same
markers and model are pinned as well: its run without LOCO (genotype_100markers_marker_vcf.txt) and its run with the
alleles exchanged (genotype_100markers_marker_bgen.txt).

Note: the golden SE of the two SPA-adjusted variants equals |BETA|/|qnorm(p/2)| while this fork's source computes
qnorm(p, upper tail) (SAIGE_test.cpp:523-526); the fixture wins (se_two_sided=True), the source's variant is kept
behind the flag and tested for self-consistency."""
import os

import numpy as np
import pytest

TOL_PRINT = 2e-5
NUMERIC = ["AC_Allele2", "AF_Allele2", "MissingRate", "BETA", "SE", "Tstat", "var", "p.value", "p.value.NA", "AF_case",
           "AF_ctrl", "N_case", "N_ctrl", "N_case_hom", "N_case_het", "N_ctrl_hom", "N_ctrl_het"]


def golden_rows(golden_dir, name="step2_100markers_golden.txt"):
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(golden_dir, name))]
    return [dict(zip(rows[0], r)) for r in rows[1:]]


def swap_alleles(bed):
    """The same genotypes with A1 and A2 exchanged: homozygote codes 00 <-> 11, het (10) and missing (01) unchanged."""
    lo, hi = bed & 0x55, (bed >> 1) & 0x55
    hom = ~(lo ^ hi) & 0x55
    return bed ^ (hom | (hom << 1))


def oracle_rows(golden_dir, LOCO=True, swapped=False, **kw):
    from oracle import oracle as O
    from oracle import step2_oracle as S2
    from saige_gpu_b200.rdata import load_rda
    mod = load_rda(os.path.join(golden_dir, "example_binary.rda"))["modglmm"]
    M = S2.read_model(mod, chrom=1, LOCO=LOCO)
    M["varRatio"] = float(open(os.path.join(golden_dir, "example_binary.varianceRatio.txt")).read().split()[0])
    bed, N0, M0, _ = O.read_bed(os.path.join(golden_dir, "step2_100markers"))
    if swapped:
        bed = swap_alleles(bed)
    fam = [l.split()[1] for l in open(os.path.join(golden_dir, "step2_100markers.fam"))]
    bim = [l.split() for l in open(os.path.join(golden_dir, "step2_100markers.bim"))]
    pos = np.array([fam.index(s) for s in M["sampleID"]])
    out = {}
    for m, b in enumerate(bim):
        r = S2.test_marker(M, S2.plink_marker(bed, N0, m, pos), min_mac=20, **kw)
        if r is not None:
            out[b[1]] = r
    return out


def test_rda_reader_on_reference_model(golden_dir):
    from saige_gpu_b200.rdata import load_rda
    m = load_rda(os.path.join(golden_dir, "example_binary.rda"))["modglmm"]
    assert abs(m["theta"][1] - 0.33267712593078613) < 1e-15 and m["theta"][0] == 1.0
    assert abs(float(np.ravel(m["coefficients"])[0]) - (-2.522796154022217)) < 1e-15
    assert m["X"].shape == (1000, 3) and m["obj.noK"]["XV"].shape == (3, 1000) and len(m["LOCOResult"]) == 22
    assert m["traitType"] == ["binary"] and len(m["sampleID"]) == 1000


def test_oracle_reproduces_reference_golden_table(golden_dir):
    gold = golden_rows(golden_dir)
    mine = oracle_rows(golden_dir)
    assert [g["MarkerID"] for g in gold] == list(mine.keys())            # same 32 of 100 markers pass MAC >= 20
    key = {"p.value": "p_value", "p.value.NA": "p_value_NA"}
    for g in gold:
        r = mine[g["MarkerID"]]
        for col in ("AC_Allele2", "AF_Allele2", "MissingRate", "BETA", "SE", "Tstat", "var", "p.value", "p.value.NA", "AF_case",
                    "AF_ctrl", "N_case", "N_ctrl"):
            gv, mv = float(g[col]), float(r[key.get(col, col)])
            assert abs(mv - gv) <= TOL_PRINT * max(abs(gv), 1e-300) + 1e-300, (g["MarkerID"], col, mv, gv)
        assert str(r["Is_SPA"]).lower() == g["Is.SPA"]
    assert sum(g["Is.SPA"] == "true" for g in gold) == 2                  # rs23, rs38 go through SPA_fast


def test_driver_writes_the_reference_table_as_a_file(golden_dir, tmp_path):
    """SPAGMMATtest's host side (model / ratio / .fam matching, mapped .bed, chunked writing, number formats) with the
    device call answered by the oracle: the file it writes IS the reference's golden table, header and all 32 rows."""
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import step2
    from conftest import OracleDevice

    p = os.path.join(golden_dir, "step2_100markers")
    path = str(tmp_path / "out.txt")
    n = step2.SPAGMMATtest(OracleDevice(), p + ".bed", p + ".bim", p + ".fam", os.path.join(golden_dir, "example_binary.rda"),
                           os.path.join(golden_dir, "example_binary.varianceRatio.txt"), SAIGEOutputFile=path, chrom="1", LOCO=True,
                           min_MAC=20, markers_per_chunk=9, return_rows=False)
    mine = [l.split("\t") for l in open(path).read().splitlines()]
    gold = [l.split("\t") for l in open(os.path.join(golden_dir, "step2_100markers_golden.txt")).read().splitlines()]
    assert n == 32 and len(mine) == len(gold) == 33 and mine[0] == gold[0]
    for a, b in zip(mine[1:], gold[1:]):
        for name, x, y in zip(gold[0], a, b):
            if name in ("CHR", "POS", "MarkerID", "Allele1", "Allele2", "Is.SPA", "N_case", "N_ctrl", "N_case_hom", "N_case_het",
                        "N_ctrl_hom", "N_ctrl_het", "AC_Allele2"):
                assert x == y, (name, a[2], x, y)
            else:
                assert abs(float(x) - float(y)) <= TOL_PRINT * abs(float(y)) + 1e-300, (name, a[2], x, y)


def test_restart_from_the_index_file(golden_dir, tmp_path):
    """is_overwrite_output = FALSE (R/Util.R:441-595): the .index file records finished chunks; an interrupted scan is resumed
    after the last recorded chunk and ends with the table an uninterrupted scan writes; a finished scan is left alone."""
    from saige_gpu_b200 import step2
    from conftest import OracleDevice
    p = os.path.join(golden_dir, "step2_100markers")
    common = dict(bedFile=p + ".bed", bimFile=p + ".bim", famFile=p + ".fam", GMMATmodelFile=os.path.join(golden_dir, "example_binary.rda"),
                  varianceRatioFile=os.path.join(golden_dir, "example_binary.varianceRatio.txt"), chrom="1", LOCO=True, min_MAC=20,
                  markers_per_chunk=16, return_rows=False)
    whole = str(tmp_path / "whole.txt")
    assert step2.SPAGMMATtest(OracleDevice(), SAIGEOutputFile=whole, **common) == 32
    idx = open(whole + ".index").read().splitlines()
    ref_idx = open(os.path.join(golden_dir, "..", "golden", "step2_100markers_golden.txt")).readline()      # (fixture present)
    assert idx[0].startswith("This is the output index file for SAIGE package") and idx[1] == "This is a Marker level analysis."
    assert idx[2] == "nEachChunk = 16" and idx[3:] == ["Have completed the analysis of chunk %d" % i for i in range(1, 8)] + [
        "Have completed the analyses of all chunks."] and ref_idx

    class Dies(OracleDevice):
        calls = 0

        def mainMarkerInCPP(self, *a, **k):
            Dies.calls += 1
            if Dies.calls == 4:
                raise RuntimeError("power cut")
            return super().mainMarkerInCPP(*a, **k)

    part = str(tmp_path / "part.txt")
    with pytest.raises(RuntimeError):
        step2.SPAGMMATtest(Dies(), SAIGEOutputFile=part, **common)
    assert open(part + ".index").read().splitlines()[-1] == "Have completed the analysis of chunk 3"
    n_before = len(open(part).read().splitlines())
    n_more = step2.SPAGMMATtest(OracleDevice(), SAIGEOutputFile=part, is_overwrite_output=False, **common)
    assert n_before - 1 + n_more == 32 and open(part).read() == open(whole).read()
    assert open(part + ".index").read() == open(whole + ".index").read()
    assert step2.SPAGMMATtest(OracleDevice(), SAIGEOutputFile=part, is_overwrite_output=False, **common) == 0       # finished: untouched
    assert open(part).read() == open(whole).read()
    with pytest.raises(ValueError):
        step2.SPAGMMATtest(OracleDevice(), SAIGEOutputFile=part, is_overwrite_output=False, **{**common, "markers_per_chunk": 10})
    os.remove(part + ".index")
    with pytest.raises(ValueError):
        step2.SPAGMMATtest(OracleDevice(), SAIGEOutputFile=part, is_overwrite_output=False, **common)


def test_marker_selection_files(golden_dir, tmp_path):
    """idstoIncludeFile / rangestoIncludeFile (R/Geno.R:282-335) for PLINK rows (gathered) and dosage rows (filtered)."""
    from saige_gpu_b200 import step2
    from conftest import OracleDevice
    p = os.path.join(golden_dir, "step2_100markers")
    ids, rng_ = str(tmp_path / "ids.txt"), str(tmp_path / "ranges.txt")
    open(ids, "w").write("rs14\n1:26:A:C\nrs_not_there\n")
    open(rng_, "w").write("1 50 60\n2 1 1000\n")
    base = dict(GMMATmodelFile=os.path.join(golden_dir, "example_binary.rda"), chrom="1", LOCO=True, min_MAC=20, markers_per_chunk=3,
                varianceRatioFile=os.path.join(golden_dir, "example_binary.varianceRatio.txt"))
    plink = dict(bedFile=p + ".bed", bimFile=p + ".bim", famFile=p + ".fam")
    full = {r["MarkerID"]: r for r in step2.SPAGMMATtest(OracleDevice(), **plink, **base)}
    want = [m for m in full if m in ("rs14", "rs26") or 50 <= int(m[2:]) <= 60]
    for src in (plink, dict(vcfFile=p + ".vcf.gz", vcfField="GT")):
        got = step2.SPAGMMATtest(OracleDevice(), idstoIncludeFile=ids, rangestoIncludeFile=rng_, **src, **base)
        assert [r["MarkerID"] for r in got] == want and len(want) >= 4
        for r in got:
            assert r["p.value"] == full[r["MarkerID"]]["p.value"] and r["Tstat"] == full[r["MarkerID"]]["Tstat"]
        two = [step2.SPAGMMATtest(OracleDevice(), idstoIncludeFile=ids, rangestoIncludeFile=rng_, rank=k, world=2, **src, **base) for k in (0, 1)]
        assert [r["MarkerID"] for part in two for r in part] == want
    assert step2.SPAGMMATtest(OracleDevice(), restrict_to_chrom=True, **plink, **{**base, "chrom": "2", "LOCO": False}) == []
    open(rng_, "w").write("1 50\n")
    with pytest.raises(ValueError):
        step2.SPAGMMATtest(OracleDevice(), rangestoIncludeFile=rng_, **plink, **base)


@pytest.mark.gpu
def test_gpu_step2_reproduces_reference_golden_table(golden_dir, tmp_path):
    from saige_gpu_b200 import SaigeB200, step2
    g = SaigeB200(device=0)
    p = os.path.join(golden_dir, "step2_100markers")
    out = str(tmp_path / "step2_out.txt")
    rows = step2.SPAGMMATtest(g, p + ".bed", p + ".bim", p + ".fam", os.path.join(golden_dir, "example_binary.rda"),
                              os.path.join(golden_dir, "example_binary.varianceRatio.txt"), SAIGEOutputFile=out, chrom="1",
                              LOCO=True, min_MAC=20)
    gold = golden_rows(golden_dir)
    assert [r["MarkerID"] for r in rows] == [x["MarkerID"] for x in gold]
    for r, x in zip(rows, gold):
        for col in ("CHR", "POS", "Allele1", "Allele2"):
            assert str(r[col]) == x[col]
        for col in NUMERIC:
            gv, mv = float(x[col]), float(r[col])
            assert abs(mv - gv) <= TOL_PRINT * max(abs(gv), 1e-300) + 1e-300, (r["MarkerID"], col, mv, gv)
        assert ("true" if r["Is.SPA"] else "false") == x["Is.SPA"]
    # the written file has the reference's header and as many lines
    lines = open(out).read().splitlines()
    assert lines[0].split("\t") == list(gold[0].keys()) and len(lines) == 33
    # GPU vs the fp64 oracle at full precision (both SE conventions)
    for two_sided in (True, False):
        ora = oracle_rows(golden_dir) if two_sided else None
        rows2 = step2.SPAGMMATtest(g, p + ".bed", p + ".bim", p + ".fam", os.path.join(golden_dir, "example_binary.rda"),
                                   os.path.join(golden_dir, "example_binary.varianceRatio.txt"), chrom="1", LOCO=True,
                                   min_MAC=20, se_two_sided=two_sided)
        if two_sided:
            for r in rows2:
                o = ora[r["MarkerID"]]
                for col, oc in (("BETA", "BETA"), ("SE", "SE"), ("Tstat", "Tstat"), ("var", "var"), ("p.value", "p_value"),
                                ("p.value.NA", "p_value_NA"), ("AF_case", "AF_case")):
                    assert abs(r[col] - o[oc]) <= 1e-9 * abs(o[oc]), (r["MarkerID"], col)
        else:
            from scipy import stats
            spa = [r for r in rows2 if r["Is.SPA"]]
            assert len(spa) == 2
            for r in spa:
                assert abs(r["SE"] - abs(r["BETA"]) / stats.norm.isf(r["p.value"])) < 1e-9 * r["SE"]
    g.close()


ALL_NUMERIC_ORACLE = (("AC_Allele2", "AC_Allele2"), ("AF_Allele2", "AF_Allele2"), ("MissingRate", "MissingRate"), ("BETA", "BETA"),
                      ("SE", "SE"), ("Tstat", "Tstat"), ("var", "var"), ("p.value", "p_value"), ("p.value.NA", "p_value_NA"),
                      ("AF_case", "AF_case"), ("AF_ctrl", "AF_ctrl"), ("N_case", "N_case"), ("N_ctrl", "N_ctrl"))


@pytest.mark.parametrize("name,kw", [("step2_100markers_golden_noLOCO.txt", dict(LOCO=False)),
                                     ("step2_100markers_golden_flipped.txt", dict(LOCO=True, swapped=True))])
def test_oracle_reproduces_two_more_reference_tables(golden_dir, name, kw):
    """genotype_100markers_marker_vcf.txt is the reference's run WITHOUT LOCO on the same markers and model;
    genotype_100markers_marker_bgen.txt is its run with Allele1 / Allele2 exchanged (AF_Allele2 up to 0.99: every row
    takes the flip branch of imputeGenoAndFlip, UTIL.cpp:58-135, and comes back with the sign of BETA / Tstat turned)."""
    gold = golden_rows(golden_dir, name)
    mine = oracle_rows(golden_dir, **kw)
    assert [g["MarkerID"] for g in gold] == list(mine.keys()) and len(gold) == 32
    for g in gold:
        r = mine[g["MarkerID"]]
        for col, oc in ALL_NUMERIC_ORACLE:
            gv, mv = float(g[col]), float(r[oc])
            assert abs(mv - gv) <= TOL_PRINT * max(abs(gv), 1e-300) + 1e-300, (g["MarkerID"], col, mv, gv)
        assert str(r["Is_SPA"]).lower() == g["Is.SPA"]
    if kw.get("swapped"):
        base = {g["MarkerID"]: g for g in golden_rows(golden_dir)}
        for g in gold:                      # the reference's own two tables are mirror images of each other
            b = base[g["MarkerID"]]
            assert abs(float(g["BETA"]) + float(b["BETA"])) <= TOL_PRINT * abs(float(b["BETA"]))
            assert abs(float(g["AF_Allele2"]) + float(b["AF_Allele2"]) - 1.0) < 1e-5
            assert (g["Allele1"], g["Allele2"]) == (b["Allele2"], b["Allele1"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,loco,swapped", [("step2_100markers_golden_noLOCO.txt", False, False),
                                               ("step2_100markers_golden_flipped.txt", True, True)])
def test_gpu_step2_reproduces_two_more_reference_tables(golden_dir, tmp_path, name, loco, swapped):
    import shutil
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200, step2
    p = os.path.join(golden_dir, "step2_100markers")
    q = str(tmp_path / "markers")
    shutil.copyfile(p + ".fam", q + ".fam")
    if swapped:
        bed, _, _, _ = O.read_bed(p)
        with open(q + ".bed", "wb") as f:
            f.write(bytes([0x6C, 0x1B, 0x01]))
            f.write(swap_alleles(bed).tobytes())
        with open(q + ".bim", "w") as f:
            for l in open(p + ".bim"):
                t = l.split()
                f.write("\t".join([t[0], t[1], t[2], t[3], t[5], t[4]]) + "\n")
    else:
        shutil.copyfile(p + ".bed", q + ".bed")
        shutil.copyfile(p + ".bim", q + ".bim")
    g = SaigeB200(device=0)
    try:
        rows = step2.SPAGMMATtest(g, q + ".bed", q + ".bim", q + ".fam", os.path.join(golden_dir, "example_binary.rda"),
                                  os.path.join(golden_dir, "example_binary.varianceRatio.txt"), chrom="1", LOCO=loco, min_MAC=20)
    finally:
        g.close()
    gold = golden_rows(golden_dir, name)
    assert [r["MarkerID"] for r in rows] == [x["MarkerID"] for x in gold]
    for r, x in zip(rows, gold):
        for col in ("CHR", "POS", "Allele1", "Allele2"):
            assert str(r[col]) == x[col]
        for col in NUMERIC:
            gv, mv = float(x[col]), float(r[col])
            assert abs(mv - gv) <= TOL_PRINT * max(abs(gv), 1e-300) + 1e-300, (r["MarkerID"], col, mv, gv)
        assert ("true" if r["Is.SPA"] else "false") == x["Is.SPA"]


def positive_signal_inputs(golden_dir):
    """The reference's genome-wide-significant example: a one-marker VCF over 10,000 samples of which the model uses the
    first 1,000 (the identity fast path with a .fam longer than the model), example_binary_positive_signal.rda and its
    result row."""
    lines = [l for l in open(os.path.join(golden_dir, "positive_signal_1marker.vcf")) if not l.startswith("##")]
    ids = lines[0].rstrip("\n").split("\t")[9:]
    rec = lines[1].rstrip("\n").split("\t")

    def alt_count(gt):
        a = gt.replace("|", "/").split("/")
        return -1 if "." in a else sum(int(x) for x in a)
    g = np.array([alt_count(x) for x in rec[9:]], dtype=np.int64)
    gold_lines = open(os.path.join(golden_dir, "positive_signal_golden.txt")).read().splitlines()
    gold = dict(zip(gold_lines[0].split("\t"), gold_lines[1].split("\t")))
    vr = float(open(os.path.join(golden_dir, "positive_signal.varianceRatio.txt")).read().split()[0])
    return ids, rec[:5], g, gold, vr


# BETA / SE of this row come from the reference's Firth refit (is_Firth_beta = TRUE, pCutoffforFirth 0.01): bias-reduced
# effect size 1.05265 against the score-test estimate 1.15015, SE = the fit's own standard error
POSITIVE_COLS = (("AC_Allele2", "AC_Allele2"), ("AF_Allele2", "AF_Allele2"), ("MissingRate", "MissingRate"), ("BETA", "BETA"), ("SE", "SE"),
                 ("Tstat", "Tstat"),
                 ("var", "var"), ("p.value", "p_value"), ("p.value.NA", "p_value_NA"), ("AF_case", "AF_case"), ("AF_ctrl", "AF_ctrl"),
                 ("N_case", "N_case"), ("N_ctrl", "N_ctrl"))


def test_oracle_reproduces_the_positive_signal_row(golden_dir):
    from oracle import step2_oracle as S2
    from saige_gpu_b200.rdata import load_rda
    ids, _, g, gold, vr = positive_signal_inputs(golden_dir)
    mod = load_rda(os.path.join(golden_dir, "positive_signal.rda"))["modglmm"]
    M = S2.read_model(mod, chrom=1, LOCO=True)
    M["varRatio"] = vr
    where = {s: i for i, s in enumerate(ids)}
    pos = np.array([where[s] for s in M["sampleID"]])
    assert len(ids) == 10000 and np.array_equal(pos, np.arange(1000))      # the model's samples are a prefix of the file's
    r = S2.test_marker(M, g[pos].astype(np.float64), min_mac=0.5, is_Firth_beta=True, pCutoffforFirth=0.01)
    assert r["Is_SPA"] and gold["Is.SPA"] == "true" and r["Is_Firth"] and r["Firth_converged"]
    plain = S2.test_marker(M, g[pos].astype(np.float64), min_mac=0.5)
    assert abs(plain["BETA"] - 1.15015) < 1e-5 and not plain["Is_Firth"]            # Tstat / var, what the row would hold without Firth
    assert float(gold["p.value"]) < 5e-7 < float(gold["p.value.NA"]) * 10          # SPA moves the p-value by 3x at 1e-7
    for col, oc in POSITIVE_COLS:
        gv, mv = float(gold[col]), float(r[oc])
        assert abs(mv - gv) <= TOL_PRINT * max(abs(gv), 1e-300) + 1e-300, (col, mv, gv)


@pytest.mark.gpu
def test_gpu_reproduces_the_positive_signal_row(golden_dir, tmp_path):
    from saige_gpu_b200 import SaigeB200, step2
    ids, rec, g, gold, _ = positive_signal_inputs(golden_dir)
    q = str(tmp_path / "one")
    n = len(ids)
    code = np.array([3, 2, 0, 1], dtype=np.uint8)[np.where(g < 0, 3, g)]       # A1 = ALT: 0 copies -> 11, 1 -> 10, 2 -> 00, missing -> 01
    code = np.concatenate([code, np.zeros((-n) % 4, dtype=np.uint8)]).reshape(-1, 4)
    row = (code[:, 0] | (code[:, 1] << 2) | (code[:, 2] << 4) | (code[:, 3] << 6)).astype(np.uint8)
    with open(q + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]) + row.tobytes())
    with open(q + ".bim", "w") as f:
        f.write("\t".join([rec[0], rec[2], "0", rec[1], rec[4], rec[3]]) + "\n")      # A1 = ALT, A2 = REF
    with open(q + ".fam", "w") as f:
        for s in ids:
            f.write("%s %s 0 0 0 -9\n" % (s, s))
    gpu = SaigeB200(device=0)
    try:
        rows = step2.SPAGMMATtest(gpu, q + ".bed", q + ".bim", q + ".fam", os.path.join(golden_dir, "positive_signal.rda"),
                                  os.path.join(golden_dir, "positive_signal.varianceRatio.txt"), chrom="1", LOCO=True, min_MAC=0.5,
                                  is_Firth_beta=True, pCutoffforFirth=0.01)
    finally:
        gpu.close()
    assert len(rows) == 1 and rows[0]["MarkerID"] == gold["MarkerID"] and rows[0]["Is.SPA"]
    assert rows[0]["Is.Firth"] and rows[0]["Firth.converged"]
    assert (str(rows[0]["Allele1"]), str(rows[0]["Allele2"])) == (gold["Allele1"], gold["Allele2"])
    for col, _ in POSITIVE_COLS:
        gv, mv = float(gold[col]), float(rows[0][col])
        assert abs(mv - gv) <= TOL_PRINT * max(abs(gv), 1e-300) + 1e-300, (col, mv, gv)


@pytest.mark.gpu
def test_gpu_step2_synthetic_vs_oracle():
    """Bigger synthetic check with missing calls, allele flips, sample subset/reorder, binary + quantitative traits."""
    from oracle import oracle as O
    from oracle import step2_oracle as S2
    from saige_gpu_b200 import SaigeB200
    rng = np.random.default_rng(5)
    n_fam, nm, N, p = 1503, 400, 1200, 4
    bed = O.synth_bed(n_fam, nm, seed=9, miss_rate=0.02)
    # make a third of the markers major-allele coded (alt freq > 0.5 -> flip) by complementing hom codes 00 <-> 11
    B0 = (n_fam + 3) // 4
    rows = bed.reshape(nm, B0).copy()
    for m in range(0, nm, 3):
        r = rows[m]
        lo, hi = r & 0x55, (r >> 1) & 0x55
        hom = ~(lo ^ hi) & 0x55                     # 00 or 11
        rows[m] = r ^ (hom | (hom << 1))
    bed = rows.reshape(-1)
    pos = rng.permutation(n_fam)[:N].astype(np.int32)
    X = np.column_stack([np.ones(N), rng.normal(size=(N, p - 1))])
    for trait in ("binary", "quantitative"):
        if trait == "binary":
            mu = 1 / (1 + np.exp(-(X @ np.array([-1.5, 0.4, -0.3, 0.2]) + rng.normal(scale=0.3, size=N))))
            y = (rng.uniform(size=N) < mu).astype(np.float64)
            tau = np.array([1.0, 0.4]); mu2 = mu * (1 - mu)
        else:
            y = X @ np.array([0.3, 0.4, -0.3, 0.2]) + rng.normal(size=N)
            mu = X @ np.linalg.lstsq(X, y, rcond=None)[0]
            tau = np.array([0.7, 0.3]); mu2 = np.full(N, 1 / tau[0])
        res = y - mu
        V = mu2
        XV = (X * V[:, None]).T
        XVX = X.T @ XV.T
        XVX_inv = np.linalg.inv(XVX)
        M = dict(mu=mu, res=res, mu2=mu2, tau=tau, trait=trait, y=y, X=X, XV=XV, XVX=XVX, XXVX_inv=X @ XVX_inv,
                 XVX_inv_XV=(X @ XVX_inv) * V[:, None], S_a=(X * res[:, None]).sum(0), varRatio=0.93)
        g = SaigeB200(device=0)
        g.setSAIGEobjInCPP(M, 0.93, 2.0, pos)
        out = g.mainMarkerInCPP(bed, n_fam, nm, 0.0, 5.0, 0.15)
        ntest = nspa = nflip = 0
        for m in range(nm):
            r = S2.test_marker(M, S2.plink_marker(bed, n_fam, m, pos), min_mac=5.0)
            assert (r is not None) == (out[m, 0] == 1.0), m
            if r is None:
                continue
            ntest += 1; nspa += bool(r["Is_SPA"]); nflip += r["AF_Allele2"] > 0.5
            got = dict(zip(g.STEP2_COLUMNS, out[m]))
            for col, oc in (("AC_Allele2", "AC_Allele2"), ("AF_Allele2", "AF_Allele2"), ("MissingRate", "MissingRate"),
                            ("BETA", "BETA"), ("SE", "SE"), ("Tstat", "Tstat"), ("var", "var"), ("p.value", "p_value"),
                            ("p.value.NA", "p_value_NA")):
                assert abs(got[col] - r[oc]) <= 1e-6 * abs(r[oc]) + 1e-300, (trait, m, col, got[col], r[oc])
            assert bool(got["Is.SPA"]) == bool(r["Is_SPA"])
        assert ntest > 300 and nflip > 50 and (nspa > 5 or trait == "quantitative")
        # Firth's effect size for the variants with p <= 0.2 (a wide cutoff so that dozens of refits run), both SE forms
        offset = rng.normal(scale=0.2, size=N)
        M["offset"] = offset
        for from_fit in (True, False):
            g.setFirth(True, 0.2, offset, se_from_fit=from_fit)
            outf = g.mainMarkerInCPP(bed, n_fam, nm, 0.0, 5.0, 0.15)
            nfirth = 0
            for m in range(nm):
                r = S2.test_marker(M, S2.plink_marker(bed, n_fam, m, pos), min_mac=5.0, is_Firth_beta=True, pCutoffforFirth=0.2,
                                   firth_se_from_fit=from_fit)
                if r is None:
                    continue
                got = dict(zip(g.STEP2_COLUMNS, outf[m]))
                assert bool(got["Is.Firth"]) == bool(r["Is_Firth"]) == (trait == "binary" and r["p_value"] <= 0.2), (trait, m)
                if r["Is_Firth"]:
                    nfirth += 1
                    assert bool(got["Firth.converged"]) == bool(r["Firth_converged"])
                for col, oc in (("BETA", "BETA"), ("SE", "SE"), ("p.value", "p_value")):
                    assert abs(got[col] - r[oc]) <= 1e-6 * abs(r[oc]) + 1e-300, (trait, m, col, got[col], r[oc], from_fit)
            assert nfirth > 30 or trait == "quantitative"
        g.setFirth(False)
        assert np.array_equal(np.nan_to_num(g.mainMarkerInCPP(bed, n_fam, nm, 0.0, 5.0, 0.15)), np.nan_to_num(out))
        g.close()
