import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so.12")
        n = ctypes.c_int(0)
        return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def grm10k():
    """The reference's 1000-sample x 10k-marker quick-test set (docs/installation.md:110)."""
    from oracle import oracle as O
    bed, N0, M0, chrs = O.read_bed(os.path.join(GOLDEN, "grm10k"))
    return dict(bed=bed, N0=N0, M0=M0, chrs=np.array([int(c) for c in chrs]), prefix=os.path.join(GOLDEN, "grm10k"))


@pytest.fixture(scope="session")
def chr22():
    from oracle import oracle as O
    bed, N0, M0, chrs = O.read_bed(os.path.join(GOLDEN, "chr22_1000"))
    return dict(bed=bed, N0=N0, M0=M0, chrs=np.array([int(c) for c in chrs]), prefix=os.path.join(GOLDEN, "chr22_1000"))


class OracleDevice:
    """TEST INFRASTRUCTURE: answers the device calls of saige_gpu_b200.step2.SPAGMMATtest from the CPU oracle, so the driver's
    host side (file formats, sample matching, chunking, sharding, output) runs without a GPU.  Never imported by the product."""
    STEP2_COLUMNS = ("tested", "AC_Allele2", "AF_Allele2", "MissingRate", "BETA", "SE", "Tstat", "var", "p.value", "p.value.NA",
                     "Is.SPA", "AF_case", "AF_ctrl", "N_case", "N_ctrl", "N_case_hom", "N_case_het", "N_ctrl_hom", "N_ctrl_het",
                     "var2", "Is.Firth", "Firth.converged", "BETA_c", "SE_c", "Tstat_c", "var_c", "p.value_c", "p.value.NA_c")

    def __init__(self):
        self.er, self.firth, self.cond = -1.0, dict(is_Firth_beta=False), None

    def setCondition(self, P2=None, XtP2=None, VarInv=None, Tstat_cond=None):
        self.cond = None if P2 is None else dict(P2=np.asarray(P2), VarInv=np.asarray(VarInv), Tstat=np.asarray(Tstat_cond))

    def setSAIGEobjInCPP(self, model, ratio, cutoff, pos):
        self.M = dict(model, varRatio=ratio)
        self.M["XV"] = (np.asarray(model["X"]) * np.asarray(model["mu2"])[:, None]).T
        self.pos, self.cutoff = np.asarray(pos), cutoff

    def setFirth(self, is_Firth_beta, pCutoffforFirth=0.01, offset=None, se_from_fit=True):
        self.firth = dict(is_Firth_beta=bool(is_Firth_beta), pCutoffforFirth=pCutoffforFirth, firth_se_from_fit=se_from_fit)

    def setMaxMACforER(self, v):
        self.er = v

    def _rows(self, G_of, nm, kw):
        from oracle import step2_oracle as S2
        out = np.full((nm, len(self.STEP2_COLUMNS)), np.nan)
        for j in range(nm):
            r = S2.test_marker(self.M, G_of(j), spa_cutoff=self.cutoff, max_MAC_for_ER=self.er, cond=self.cond, **self.firth, **kw)
            out[j, 0] = 0.0 if r is None else 1.0
            if r is None:
                continue
            out[j, 1:13] = [r["AC_Allele2"], r["AF_Allele2"], r["MissingRate"], r["BETA"], r["SE"], r["Tstat"], r["var"],
                            r["p_value"], r["p_value_NA"], float(r["Is_SPA"]), r["AF_case"], r["AF_ctrl"]]
            out[j, 13:19] = [r["N_case"], r["N_ctrl"], r["N_case_hom"], r["N_case_het"], r["N_ctrl_hom"], r["N_ctrl_het"]]
            if self.M["trait"] != "binary":               # the kernel counts y == 1 as "cases" and everybody else as "controls"
                out[j, 13:15] = [float(np.sum(self.M["y"] == 1)), float(np.sum(self.M["y"] != 1))]
            out[j, 20:22] = [float(r["Is_Firth"]), float(r["Firth_converged"])]
            if self.cond is not None:
                out[j, 22:28] = [r["BETA_c"], r["SE_c"], r["Tstat_c"], r["var_c"], r["p_value_c"], r["p_value_NA_c"]]
        return out

    def mainMarkerInCPP(self, rows, n_fam, nm, min_MAF=0.0, min_MAC=0.5, max_missing=0.15, se_two_sided=True):
        from oracle import step2_oracle as S2
        body = np.asarray(rows)
        return self._rows(lambda j: S2.plink_marker(body, n_fam, j, self.pos), nm,
                          dict(min_maf=min_MAF, min_mac=min_MAC, max_missing=max_missing, se_two_sided=se_two_sided))

    def mainMarkerInCPP_dosage(self, D, min_MAF=0.0, min_MAC=0.5, max_missing=0.15, se_two_sided=True, impute_method=1,
                               dosage_zerod_cutoff=0.2, dosage_zerod_MAC_cutoff=10.0):
        D = np.asarray(D, dtype=np.float64)
        name = {1: "best_guess", 2: "mean", 3: "minor"}[impute_method]
        return self._rows(lambda j: np.where(np.isnan(D[j, self.pos]), -1.0, D[j, self.pos]), D.shape[0],
                          dict(min_maf=min_MAF, min_mac=min_MAC, max_missing=max_missing, se_two_sided=se_two_sided,
                               impute_method=name, dosage_zerod_cutoff=dosage_zerod_cutoff,
                               dosage_zerod_MAC_cutoff=dosage_zerod_MAC_cutoff))
