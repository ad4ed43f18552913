"""Step-2 driver: single-variant association tests from a step-1 null model, PLINK input.

Python mirror of the R side of `SPAGMMATtest` for single-variant tests (/root/reference/src/SAIGE/R/
SAIGE_Test_main.R:61-420, R/readInGLMM.R:39-170 `ReadModel`, R/SAIGE_SPATest_Marker.R `SAIGE.Marker`): read the
null model (.rda written by step 1), the variance ratio, the PLINK files; match model samples to .fam rows; stream
marker rows to the C ABI (`sgb_step2_test_markers` = the body of mainMarkerInCPP, Main.cpp:149-560); write the result
table with the reference's column names.  No numerical work happens here."""
import os

import re

import numpy as np

from . import genoio
from .rdata import load_rda

OUT_COLUMNS = ["CHR", "POS", "MarkerID", "Allele1", "Allele2", "AC_Allele2", "AF_Allele2", "MissingRate", "BETA", "SE",
               "Tstat", "var", "p.value", "p.value.NA", "Is.SPA", "AF_case", "AF_ctrl", "N_case", "N_ctrl", "N_case_hom",
               "N_case_het", "N_ctrl_hom", "N_ctrl_het"]


def ReadModel(GMMATmodelFile, chrom="", LOCO=True):
    """readInGLMM.R:39-170: the fields step 2 consumes; with LOCO the chromosome's refit replaces mu/res/obj.noK."""
    m = load_rda(GMMATmodelFile)["modglmm"]
    # per-chromosome files of isLowMemLOCO (FG.R:1205-1290) carry the chromosome's refit only: the main fields are NULL
    mu = None if m.get("fitted.values") is None else np.asarray(m["fitted.values"], dtype=np.float64).ravel()
    res = None if m.get("residuals") is None else np.asarray(m["residuals"], dtype=np.float64).ravel()
    noK = m.get("obj.noK")
    has_loco = bool(np.asarray(m.get("LOCO", [0])).ravel()[0])
    if LOCO:
        if not has_loco:
            raise ValueError("LOCO is TRUE but the null model file .rda does not contain LOCO results")
        if chrom == "":
            raise ValueError("chrom needs to be specified in order to apply Leave-one-chromosome-out")
        c = chrom_number(chrom)
        if 1 <= c <= 22:
            lr = m["LOCOResult"][c - 1]
            if isinstance(lr, dict) and "fitted.values" in lr:
                mu = np.asarray(lr["fitted.values"], dtype=np.float64).ravel()
                res = np.asarray(lr["residuals"], dtype=np.float64).ravel()
                noK = lr["obj.noK"]
    if mu is None or noK is None:
        raise ValueError("%s holds no fit for chromosome %s (a per-chromosome model file of another chromosome?)" % (GMMATmodelFile, chrom))
    trait = m["traitType"][0] if isinstance(m["traitType"], list) else str(m["traitType"])
    tau = np.asarray(m["theta"], dtype=np.float64).ravel()
    mu2 = mu * (1 - mu) if trait == "binary" else np.full(len(mu), 1.0 / tau[0])
    # offset of the Firth refit (readInGLMM.R:99-101, 134-160): the chromosome's own when LOCO stored one, else the model's
    offset = m.get("offset")
    if LOCO and has_loco and chrom != "":
        c = chrom_number(chrom)
        if 1 <= c <= 22 and isinstance(m["LOCOResult"][c - 1], dict) and m["LOCOResult"][c - 1].get("offset") is not None:
            offset = m["LOCOResult"][c - 1]["offset"]
    offset = np.zeros(len(mu)) if offset is None else np.asarray(offset, dtype=np.float64).ravel()
    return dict(mu=mu, res=res, mu2=mu2, tau=tau, trait=trait, offset=offset, y=np.asarray(m["y"], dtype=np.float64).ravel(),
                X=np.asarray(m["X"], dtype=np.float64), XVX=np.asarray(noK["XVX"], dtype=np.float64),
                XXVX_inv=np.asarray(noK["XXVX_inv"], dtype=np.float64),
                XVX_inv_XV=np.asarray(noK["XVX_inv_XV"], dtype=np.float64), S_a=np.asarray(noK["S_a"], dtype=np.float64).ravel(),
                sampleID=[str(s) for s in m["sampleID"]])


def chrom_number(chrom):
    """getChromNumber (R/readInGLMM.R:4-20): 'chr' stripped case-insensitively, digits only; anything that is not 1..22 (X, Y, MT,
    empty) gives 0 = no leave-one-chromosome-out refit, i.e. the genome-wide fit is used as the reference does."""
    digits = re.sub(r"[^0-9]", "", re.sub(r"(?i)chr", "", str(chrom)))
    return int(digits) if digits else 0


def Get_Variance_Ratio(varianceRatioFile, cateVarRatioMinMACVecExclude=(10, 20.5), cateVarRatioMaxMACVecInclude=(20.5,)):
    """readInGLMM.R:358-435: the 'null' rows of the step-1 variance-ratio file (3 columns: value, null / sparse, category;
    files of versions < 1.0.6 hold the values only).  One row: a float.  Several rows: categorical variance ratios, returned
    as a list whose length must match the MAC category bounds."""
    rows = [l.split() for l in open(varianceRatioFile) if l.strip()]
    if not rows:
        raise ValueError("variance ratio file %s is empty" % varianceRatioFile)
    if len(rows[0]) == 3:
        vals = [float(r[0]) for r in rows if r[1] == "null"]
        if any(r[1] == "sparse" and not (0.9999 <= float(r[0]) <= 1.0001) for r in rows):
            raise ValueError("sparse GRM is not specified but it was used for estimating variance ratios in Step 1")
    else:
        vals = [float(r[0]) for r in rows]
    if len(vals) == 1:
        return vals[0]
    if len(vals) != len(cateVarRatioMinMACVecExclude):
        raise ValueError("ERROR! The number of variance ratios are different from the length of cateVarRatioMinMACVecExclude")
    if len(cateVarRatioMinMACVecExclude) != len(cateVarRatioMaxMACVecInclude) + 1:
        raise ValueError("ERROR! The length of cateVarRatioMaxMACVecInclude does not match with the lenght of cateVarRatioMinMACVecExclude (-1)")
    return vals


IMPUTE_METHODS = {"best_guess": 1, "mean": 2, "minor": 3}
COND_COLUMNS = ["BETA_c", "SE_c", "Tstat_c", "var_c", "p.value_c", "p.value.NA_c"]


def _impute_and_flip(Graw, impute_method, zerod_cutoff, zerod_mac_cutoff):
    """Host copy of getOneMarker's counts + imputeGenoAndFlip (UTIL.cpp:58-135) for the handful of conditioning markers."""
    n = len(Graw)
    miss = ~(Graw >= 0)
    cnt = n - int(miss.sum())
    af = float(Graw[~miss].sum()) / cnt / 2 if cnt > 0 else 0.0
    mac = min(af, 1 - af) * n * (1 - miss.sum() / n) * 2
    G = np.where(miss, 0.0, Graw)
    if af > 0.5:
        G, af = 2 - G, 1 - af
    if miss.any():
        g0 = {"best_guess": np.floor(2 * af + 0.5), "mean": 2 * af, "minor": 0.0}[impute_method]
        G[miss] = g0
        mac += g0 * int(miss.sum())
    if zerod_cutoff > 0 and mac <= zerod_mac_cutoff:
        G[np.abs(G) <= zerod_cutoff] = 0.0
    return G, min(G.sum(), 2 * n - G.sum())


def _pick_ratio(ratio, mac, lo, hi):
    r = np.asarray(ratio, dtype=np.float64).reshape(-1)
    if len(r) == 1:
        return float(r[0])
    for i in range(len(hi)):
        if mac <= hi[i]:
            return float(r[i])
    return float(r[-1])


def condition_factors(model, ratio, cond_rows, impute_method="best_guess", zerod_cutoff=0.2, zerod_mac_cutoff=10.0,
                      cate_lo=(10, 20.5), cate_hi=(20.5,)):
    """assign_conditionMarkers_factors (Main.cpp:2002-2179) on the host, as in the reference: for each conditioning marker
    (dosage row in model-sample order) gtilde = G - XXVX_inv (XV G), P1 row = sqrt(vr) gtilde, P2 column = sqrt(vr) gtilde %
    mu2 tau0, its score T = (res.G - S_a.Z) / tau0; VarInv = pinv(P1 P2).  Returns what sgb_step2_set_condition takes."""
    X, mu2, tau0 = model["X"], model["mu2"], float(model["tau"][0])
    XV = (X * mu2[:, None]).T
    P1, P2, T = [], [], []
    for row in cond_rows:
        G, mac = _impute_and_flip(np.asarray(row, dtype=np.float64), impute_method, zerod_cutoff, zerod_mac_cutoff)
        if G.sum() == 0:
            raise ValueError("ERROR: Conditioning marker is monomorphic")
        vr = _pick_ratio(ratio, mac, cate_lo, cate_hi)
        gt = G - model["XXVX_inv"] @ (XV @ G)
        P1.append(np.sqrt(vr) * gt)
        P2.append(np.sqrt(vr) * gt * mu2 * tau0)
        T.append((float(model["res"] @ G) - float(model["S_a"] @ (model["XVX_inv_XV"].T @ G))) / tau0)
    P1, P2 = np.array(P1), np.array(P2).T
    return dict(P2=P2, XtP2=model["XXVX_inv"].T @ P2, VarInv=np.linalg.pinv(P1 @ P2), Tstat_cond=np.array(T))


def SPAGMMATtest(geno, bedFile="", bimFile="", famFile="", GMMATmodelFile="", varianceRatioFile="", SAIGEOutputFile=None, chrom="",
                 LOCO=True, min_MAF=0.0, min_MAC=0.5, max_missing=0.15, SPAcutoff=2.0, markers_per_chunk=10000,
                 is_output_moreDetails=True, se_two_sided=True, rank=0, world=1, is_Firth_beta=False, pCutoffforFirth=0.01,
                 firth_se_from_fit=True, max_MAC_for_ER=4.0, cateVarRatioMinMACVecExclude=(10, 20.5),
                 cateVarRatioMaxMACVecInclude=(20.5,), return_rows=True, vcfFile="", vcfField="DS", bgenFile="", sampleFile="",
                 AlleleOrder="alt-first", impute_method="best_guess", dosage_zerod_cutoff=0.2, dosage_zerod_MAC_cutoff=10.0,
                 condition="", is_overwrite_output=True, idstoIncludeFile="", rangestoIncludeFile="", restrict_to_chrom=False,
                 is_imputed_data=False, minInfo=0.0):
    """Returns the result table (list of dict rows; with return_rows=False only the number of tested variants, for scans
    whose table should not be held in memory); writes it tab-separated to SAIGEOutputFile when given, chunk by chunk.
    Genotypes: PLINK (bedFile / bimFile / famFile; raw 2-bit rows go to the device), or vcfFile (+ vcfField "DS" / "GT"), or
    bgenFile (+ sampleFile when the file holds no sample identifiers): rows of dosages go to the device (genoio.py).
    AlleleOrder applies to PLINK and BGEN as in the reference ("alt-first": the first allele is the tested one).
    condition = "chr:pos:ref:alt,..." (at most 4 markers of the same genotype file): conditional analysis, six more columns.
    idstoIncludeFile (one marker ID or chr:pos:ref:alt per line) / rangestoIncludeFile (chromosome, start, end per line):
    only these variants are tested (R/Geno.R:282-335); restrict_to_chrom=True also drops variants of other chromosomes than
    `chrom`, as the reference's PLINK branch does (Geno.R:178-180).
    is_imputed_data=True: the table holds `imputationInfo` in place of `MissingRate` and variants whose INFO score (BGEN:
    computed from the probabilities over the model's samples, BGEN.cpp:275-345; 1 for PLINK / VCF) is below minInfo are
    not tested (Main.cpp:351).
    is_overwrite_output=False: restart from `<SAIGEOutputFile>.index`, the reference's record of finished chunks
    (R/Util.R:441-595), appending to the existing table; a finished analysis is left alone.
    Multi-GPU (BASELINE config 5): variants are sharded, rank r of `world` tests the r-th contiguous slice of the variants
    and writes its own part; there is no collective, the parts are concatenated in rank order."""
    if impute_method not in IMPUTE_METHODS:
        raise ValueError("impute_method should be 'best_guess', 'mean' or 'minor'.")
    if AlleleOrder not in ("alt-first", "ref-first"):
        raise ValueError("AlleleOrder should be 'alt-first' or 'ref-first'")
    if sum(bool(x) for x in (bedFile, vcfFile, bgenFile)) != 1:
        raise ValueError("give exactly one of bedFile (+ bimFile, famFile), vcfFile, bgenFile")
    if bedFile:                                        # R/Geno.R:159-164
        bimFile = bimFile or bedFile[:-3] + "bim"
        famFile = famFile or bedFile[:-3] + "fam"
    model = ReadModel(GMMATmodelFile, chrom, LOCO)
    ratio = Get_Variance_Ratio(varianceRatioFile, cateVarRatioMinMACVecExclude, cateVarRatioMaxMACVecInclude)
    model["cateVarRatioMinMACVecExclude"], model["cateVarRatioMaxMACVecInclude"] = cateVarRatioMinMACVecExclude, cateVarRatioMaxMACVecInclude
    if bedFile:
        ids = [l.split()[1] for l in open(famFile)]
    elif vcfFile:
        ids = genoio.vcf_samples(vcfFile)
    else:
        bg = genoio.BgenNative(bgenFile)               # the library's multi-threaded host reader
        ids = genoio.read_sample_file(sampleFile) if sampleFile else bg.samples
        if ids is None:
            raise ValueError("%s holds no sample identifiers: give sampleFile" % bgenFile)
        if len(ids) != bg.N:
            raise ValueError("sampleFile lists %d samples, %s holds %d" % (len(ids), bgenFile, bg.N))
    where = {}
    for i, sid in enumerate(ids):
        where.setdefault(sid, i)
    missing = [sid for sid in model["sampleID"] if sid not in where]
    if missing:
        raise ValueError("%d samples of the null model are not in the genotype file" % len(missing))
    pos = np.array([where[sid] for sid in model["sampleID"]], dtype=np.int32)
    geno.setSAIGEobjInCPP(model, ratio, SPAcutoff, pos)
    geno.setFirth(is_Firth_beta, pCutoffforFirth, model["offset"], firth_se_from_fit)
    geno.setMaxMACforER(max_MAC_for_ER)                 # exact test of rare variants (step2_SPAtests.R:126 --max_MAC_for_ER, default 4)
    if not bedFile or impute_method != "best_guess":
        # dosage rows cost 8 bytes per sample (2-bit rows: 1/4 byte): chunks of at most 256 MB of rows
        markers_per_chunk = max(1, min(markers_per_chunk, (1 << 28) // (8 * len(ids))))
    if condition:
        wanted = [c.strip() for c in condition.split(",") if c.strip()]
        found = _find_markers(bedFile, bimFile, len(ids), vcfFile, vcfField, bgenFile, AlleleOrder, wanted)
        if any(w not in found for w in wanted):
            raise ValueError("conditioning marker(s) %s not found in the genotype file" % ", ".join(w for w in wanted if w not in found))
        f = condition_factors(model, ratio, [found[w][pos] for w in wanted], impute_method, dosage_zerod_cutoff,
                              dosage_zerod_MAC_cutoff, cateVarRatioMinMACVecExclude, cateVarRatioMaxMACVecInclude)
        geno.setCondition(f["P2"], f["XtP2"], f["VarInv"], f["Tstat_cond"])
    else:
        geno.setCondition()
    keep_marker = _marker_filter(idstoIncludeFile, rangestoIncludeFile, str(chrom) if restrict_to_chrom else "")
    if bedFile:
        # raw 2-bit rows are tested as they are (best-guess imputation is an integer); the other two imputation methods give
        # fractional genotypes, so those rows are decoded here and go through the dosage entry
        source = lambda skip: _plink_chunks(geno, bedFile, bimFile, len(ids), AlleleOrder, rank, world, markers_per_chunk,
                                            (min_MAF, min_MAC, max_missing, se_two_sided),
                                            None if impute_method == "best_guess" else (IMPUTE_METHODS[impute_method], dosage_zerod_cutoff,
                                                                                        dosage_zerod_MAC_cutoff), skip, keep_marker)
    else:
        if vcfFile:
            n_var = sum(1 for l in genoio._open_text(vcfFile) if not l.startswith("#"))
            it = genoio.iter_vcf(vcfFile, vcfField, markers_per_chunk)
        else:
            n_var = bg.M
            if is_imputed_data:
                in_model = np.zeros(bg.N, dtype=bool)
                in_model[pos] = True
                it = _with_info_scores(bg, bg.variants(AlleleOrder, markers_per_chunk, info_for=in_model), minInfo)
            else:
                it = bg.variants(AlleleOrder, markers_per_chunk)
        if keep_marker is not None:
            it = _filtered(it, keep_marker)              # (ranks then share the file's variants, not the selected ones)
        per_rank = (n_var + world - 1) // world
        source = lambda skip: _dosage_chunks(geno, it, min(n_var, rank * per_rank), min(n_var, (rank + 1) * per_rank),
                                             (min_MAF, min_MAC, max_missing, se_two_sided, IMPUTE_METHODS[impute_method],
                                              dosage_zerod_cutoff, dosage_zerod_MAC_cutoff), skip)
    # header of openOutfile_single (Main.cpp:2392-2425): binary traits carry p.value.NA, Is.SPA and the case / control
    # columns, quantitative traits end with N; the _c columns of a conditional analysis follow p.value (/ Is.SPA)
    if model["trait"] == "binary":
        cols = list(OUT_COLUMNS if is_output_moreDetails else OUT_COLUMNS[:19])
        k = cols.index("Is.SPA") + 1
        cols = cols[:k] + (COND_COLUMNS if condition else []) + cols[k:]
    else:
        cols = OUT_COLUMNS[:13] + (COND_COLUMNS[:5] if condition else []) + ["N"]
    if is_imputed_data:                                 # t_isImputation (Main.cpp:2395-2400)
        cols[cols.index("MissingRate")] = "imputationInfo"
    rows = [] if return_rows else None
    done_chunks, index_path = 0, (SAIGEOutputFile + ".index") if SAIGEOutputFile else None
    if SAIGEOutputFile and not is_overwrite_output and os.path.exists(SAIGEOutputFile):
        done_chunks, finished = _read_index(SAIGEOutputFile, index_path, markers_per_chunk)
        if finished:
            return [] if return_rows else 0              # "The analysis has been finished!" (SAIGE_SPATest_Marker.R:68)
    out = open(SAIGEOutputFile, "a" if done_chunks else "w") if SAIGEOutputFile else None
    n_tested, indexed = 0, done_chunks > 0
    try:
        if out and not done_chunks:
            out.write("\t".join(cols) + "\n")
        for i_chunk, (info, res) in enumerate(source(done_chunks), start=done_chunks + 1):      # info rows: (CHR, POS, MarkerID, Allele1, Allele2)
            keep = np.nonzero(res[:, 0] == 1.0)[0]      # the others were filtered: not written (Main.cpp:296 `continue`)
            n_tested += len(keep)
            if out and len(keep):
                out.write(_format_chunk(res[keep], [info[j] for j in keep], cols, geno.STEP2_COLUMNS))
            if return_rows:
                for j in keep:
                    r, b = res[j], info[j]
                    row = {"CHR": b[0], "POS": b[1], "MarkerID": b[2], "Allele1": b[3], "Allele2": b[4]}
                    for name, v in zip(geno.STEP2_COLUMNS[1:19], r[1:19]):
                        row[name] = v
                    row["Is.SPA"] = bool(r[10])
                    row["Is.Firth"], row["Firth.converged"] = bool(r[20]), bool(r[21])
                    if is_imputed_data:
                        row["imputationInfo"] = float(b[5]) if len(b) > 5 else 1.0
                    if condition:
                        for name, v in zip(COND_COLUMNS, r[22:28]):
                            row[name] = v
                    if len(r) > 29:
                        row["log.p.value"], row["log.p.value.NA"] = float(r[28]), float(r[29])
                    rows.append(row)
            if out:
                out.flush()
                _write_index(index_path, markers_per_chunk, i_chunk, start=(i_chunk == 1))
                indexed = True
        if out:
            if not indexed:                              # an empty slice of variants: header only
                with open(index_path, "w") as f:
                    f.write(_INDEX_MSG[0] + "\n" + _INDEX_MSG[1] + "\n" + (_INDEX_MSG[2] % markers_per_chunk) + "\n")
            with open(index_path, "a") as f:
                f.write(_INDEX_MSG[4] + "\n")
    finally:
        if out:
            out.close()
    return rows if return_rows else n_tested


def _find_markers(bedFile, bimFile, n_fam, vcfFile, vcfField, bgenFile, AlleleOrder, wanted):
    """Dosage rows (file sample order) of the markers named chr:pos:ref:alt (extract_genoIndex_condition, R/SAIGE_Test_main.R:357)."""
    wanted, found = set(wanted), {}
    if bedFile:
        B0 = (n_fam + 3) // 4
        body = np.memmap(bedFile, dtype=np.uint8, mode="r", offset=3)
        with open(bimFile) as f:
            for m, l in enumerate(f):
                b = l.split()
                ref, alt = (b[5], b[4]) if AlleleOrder == "alt-first" else (b[4], b[5])
                key = "%s:%s:%s:%s" % (b[0], b[3], ref, alt)
                if key in wanted:
                    raw = np.asarray(body[m * B0:(m + 1) * B0])
                    codes = ((raw[:, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(-1)[:n_fam]
                    d = np.array([2.0, -1.0, 1.0, 0.0])[codes]
                    found[key] = d if AlleleOrder == "alt-first" else np.where(d < 0, -1.0, 2.0 - d)
        return found
    it = (genoio.iter_vcf(vcfFile, vcfField, 256, only=wanted) if vcfFile
          else genoio.BgenFile(bgenFile).variants(AlleleOrder, 256, only=wanted))
    for info, D in it:
        for j, (c, p_, _, ref, alt) in enumerate(info):
            key = "%s:%s:%s:%s" % (c, p_, ref, alt)
            if key in wanted:
                found[key] = D[j]
        if len(found) == len(wanted):
            break
    return found


_INDEX_MSG = ["This is the output index file for SAIGE package to record the end point in case users want to restart the analysis. "
              "Please do not modify this file.", "This is a Marker level analysis.", "nEachChunk = %d",
              "Have completed the analysis of chunk %d", "Have completed the analyses of all chunks."]


def _read_index(out_path, index_path, n_each):
    """checkOutputFile (R/Util.R:441-520): (number of finished chunks, analysis finished?)."""
    if not os.path.exists(index_path):
        raise ValueError("'OutputFile' of '%s' has existed. Please use another 'OutputFile' or specify is_overwrite_output=TRUE "
                         "to overwrite the OutputFile." % out_path)
    lines = [l.rstrip("\n") for l in open(index_path) if l.strip()]
    if len(lines) < 3 or lines[0] != _INDEX_MSG[0] or lines[1] != _INDEX_MSG[1] or lines[2] != _INDEX_MSG[2] % n_each:
        raise ValueError("'OutputFileIndex' of '%s' is not as expected. Probably, it has been modified by user, which is not "
                         "permitted. Please remove the existing files of 'OutputFile' and 'OutputFileIndex' or specify "
                         "is_overwrite_output=TRUE to overwrite the OutputFile." % index_path)
    finished = lines[-1] == _INDEX_MSG[4]
    last = lines[-2] if finished else lines[-1]
    prefix = _INDEX_MSG[3][:-2]
    done = int(last[len(prefix):]) if last.startswith(prefix) else 0
    return done, finished


def _write_index(index_path, n_each, i_chunk, start):
    """writeOutputFileIndex (R/Util.R:570-595)."""
    with open(index_path, "w" if start else "a") as f:
        if start:
            f.write(_INDEX_MSG[0] + "\n" + _INDEX_MSG[1] + "\n" + (_INDEX_MSG[2] % n_each) + "\n")
        f.write((_INDEX_MSG[3] % i_chunk) + "\n")


def _marker_filter(idstoIncludeFile, rangestoIncludeFile, chrom):
    """None when every variant is tested, else keep(CHR, POS, ID, REF, ALT) (R/Geno.R:282-335; ID or chr:pos:ref:alt)."""
    ids = ranges = None
    if idstoIncludeFile:
        ids = {l.split()[0] for l in open(idstoIncludeFile) if l.strip()}
    if rangestoIncludeFile:
        ranges = []
        for l in open(rangestoIncludeFile):
            t = l.split()
            if t:
                if len(t) != 3:
                    raise ValueError("rangestoIncludeFile should only include three columns.")
                ranges.append((t[0], float(t[1]), float(t[2])))
    if ids is None and ranges is None and not chrom:
        return None

    def keep(c, pos, mid, ref, alt):
        if chrom and str(c) != chrom:
            return False
        if ids is None and ranges is None:
            return True
        if ids is not None and (mid in ids or "%s:%s:%s:%s" % (c, pos, ref, alt) in ids):
            return True
        return ranges is not None and any(str(c) == rc and lo <= float(pos) <= hi for rc, lo, hi in ranges)
    return keep


def _with_info_scores(bg, it, min_info):
    """Appends the INFO score of every variant to its info row and drops the variants below min_info (Main.cpp:351)."""
    for info, D in it:
        sc = bg.last_info
        sel = [j for j in range(len(info)) if not sc[j] < min_info]
        yield [tuple(info[j]) + (float(sc[j]),) for j in sel], (D if len(sel) == len(info) else D[sel])


def _filtered(it, keep):
    for info, D in it:
        sel = [j for j, x in enumerate(info) if keep(*x[:5])]
        if sel:
            yield [info[j] for j in sel], D[sel]
        else:
            yield [], D[:0]


def _plink_chunks(geno, bedFile, bimFile, n_fam, AlleleOrder, rank, world, markers_per_chunk, args, as_dosage=None, skip=0, keep=None):
    with open(bedFile, "rb") as f:
        magic = f.read(3)
    if magic != b"\x6c\x1b\x01":
        raise ValueError("%s is not a SNP-major PLINK .bed" % bedFile)
    B0 = (n_fam + 3) // 4
    n_bim = _count_lines(bimFile)
    # the .bed is mapped, not read: a chunk of raw rows is paged in when it is handed to the library, so a rank touches only
    # its own slice of a file that can be far larger than host memory (BASELINE config 5: 10M variants x 50 KB)
    body = np.memmap(bedFile, dtype=np.uint8, mode="r", offset=3)
    if body.size < n_bim * B0:
        raise ValueError("%s holds fewer than %d markers x %d bytes" % (bedFile, n_bim, B0))
    sel = None
    if keep is not None:                                # selected variants: their rows are gathered chunk by chunk
        sel, sel_bim = [], []
        with open(bimFile) as f:
            for m, l in enumerate(f):
                b = l.split()
                ref, alt = (b[5], b[4]) if AlleleOrder == "alt-first" else (b[4], b[5])
                if keep(b[0], b[3], b[1], ref, alt):
                    sel.append(m)
                    sel_bim.append(b)
        n_items = len(sel)
    else:
        n_items = n_bim
    per_rank = (n_items + world - 1) // world
    lo, hi = min(n_items, rank * per_rank), min(n_items, (rank + 1) * per_rank)
    lo = min(hi, lo + skip * markers_per_chunk)          # restart: the first `skip` chunks are already in the output file
    bim_iter = _bim_lines(bimFile, lo, hi) if sel is None else None
    for m0 in range(lo, hi, markers_per_chunk):
        m1 = min(hi, m0 + markers_per_chunk)
        if sel is None:
            bim = [next(bim_iter) for _ in range(m1 - m0)]
            raw = body[m0 * B0:m1 * B0]
        else:
            bim = sel_bim[m0:m1]
            raw = np.concatenate([body[m * B0:(m + 1) * B0] for m in sel[m0:m1]]) if m1 > m0 else body[:0]
        if AlleleOrder == "alt-first":                  # A1 of the .bim is the tested allele: Allele1 = A2, Allele2 = A1
            info = [(b[0], b[3], b[1], b[5], b[4]) for b in bim]
        else:                                           # ref-first: A2 is the tested allele; homozygote codes 00 <-> 11 exchanged
            lo_b, hi_b = raw & 0x55, (raw >> 1) & 0x55
            hom = ~(lo_b ^ hi_b) & 0x55
            raw = raw ^ (hom | (hom << 1))
            info = [(b[0], b[3], b[1], b[4], b[5]) for b in bim]
        if as_dosage is None:
            yield info, geno.mainMarkerInCPP(raw, n_fam, m1 - m0, *args)
        else:
            codes = ((np.asarray(raw).reshape(m1 - m0, B0)[:, :, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(m1 - m0, -1)
            D = np.array([2.0, -1.0, 1.0, 0.0])[codes[:, :n_fam]]          # PLINK.hpp:48-56: copies of A1, 01 = missing
            yield info, geno.mainMarkerInCPP_dosage(D, *args, *as_dosage)


def _dosage_chunks(geno, it, lo, hi, args, skip=0):
    seen = n_yield = 0
    for info, D in it:
        a, b = max(lo - seen, 0), min(hi - seen, len(info))
        seen += len(info)
        if a < b:
            n_yield += 1
            if n_yield > skip:                           # restart: chunks already in the output file are read past, not tested
                yield info[a:b], geno.mainMarkerInCPP_dosage(D[a:b], *args)
        if seen >= hi:
            break


def merge_rank_outputs(part_files, out_file):
    """Concatenates the per-rank result tables of a variant-sharded scan in rank order (header once): with contiguous
    slices per rank this is the table a single-rank scan writes.  Returns the number of result rows."""
    n = 0
    with open(out_file, "w") as out:
        for k, p in enumerate(part_files):
            with open(p) as f:
                header = f.readline()
                if k == 0:
                    out.write(header)
                elif header != first_header:
                    raise ValueError("%s has a different header than %s" % (p, part_files[0]))
                first_header = header
                for line in f:
                    out.write(line)
                    n += 1
    return n


def _count_lines(path):
    n, last = 0, b"\n"
    with open(path, "rb") as f:
        while True:
            buf = f.read(1 << 24)
            if not buf:
                break
            n += buf.count(b"\n")
            last = buf[-1:]
    return n + (0 if last == b"\n" else 1)


def _bim_lines(path, lo, hi):
    """.bim rows lo .. hi-1 (split), without holding the rest of the file."""
    with open(path) as f:
        for i, l in enumerate(f):
            if i >= hi:
                break
            if i >= lo:
                yield l.split()


_INFO_COLS = ["CHR", "POS", "MarkerID", "Allele1", "Allele2"]


def _format_chunk(res, bim, cols, table_cols):
    """Tab-separated lines of one chunk, numbers with 6 significant digits like the reference's ofstream (Main.cpp:2437-2560);
    formatted column-wise with numpy, one join per line."""
    idx = {c: i for i, c in enumerate(table_cols)}
    fields = []
    for c in cols:
        if c in _INFO_COLS:
            k = _INFO_COLS.index(c)
            fields.append([b[k] for b in bim])
        elif c == "imputationInfo":
            fields.append([("%.6g" % b[5]) if len(b) > 5 else "1" for b in bim])
        elif c == "N":                                  # quantitative traits: the number of model samples (Main.cpp:526-527)
            fields.append(np.char.mod("%.6g", res[:, idx["N_case"]] + res[:, idx["N_ctrl"]]).tolist())
        elif c == "Is.SPA":
            fields.append(np.where(res[:, idx[c]] != 0, "true", "false").tolist())
        elif c in _LOGP_OF and ("log." + c) in idx:
            # p-values that underflow a double are printed from their logarithm, mantissa and exponent, as the reference does
            # (SAIGE_test.cpp:273-283: "%.1fE%d")
            txt = np.char.mod("%.6g", res[:, idx[c]]).tolist()
            lp = res[:, idx["log." + c]]
            for j in np.nonzero((res[:, idx[c]] == 0) & np.isfinite(lp))[0]:
                txt[j] = format_logp(lp[j])
            fields.append(txt)
        else:
            fields.append(np.char.mod("%.6g", res[:, idx[c]]).tolist())
    return "".join("\t".join(t) + "\n" for t in zip(*fields))


_LOGP_OF = ("p.value", "p.value.NA", "p.value_c", "p.value.NA_c")


def format_logp(logp):
    """The reference's string for a p-value known by its natural log only (SAIGE_test.cpp:273-283, 553-561)."""
    log10p = logp / np.log(10.0)
    exponent = int(np.floor(log10p))
    fraction = 10.0 ** (log10p - exponent)
    if fraction >= 9.95:
        fraction, exponent = 1.0, exponent + 1
    return "%.1fE%d" % (fraction, exponent)


def _fmt(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float):
        return "%.6g" % v
    return str(v)
