"""CPU oracle for the SAIGE step-1 hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (saige_gpu_b200) never does.

Heavy loops (ingest/QC/repack, decode, GRM.vector, diag) live in saige_oracle.c; everything above the
matvec is restated here in numpy fp64, one function per reference function, citing
/root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp ("FG.cpp") and src/SAIGE/R/SAIGE_fitGLMM_fast.R ("FG.R").

Parity status: decode/allele counts/QC are pinned by the reference's .frq / .acount fixtures (tests/test_oracle_golden.py).
The solver / AI-REML functions below (getDiagOfSigma ... fitglmmaiRPCG_q) are pinned to the reference's OWN code: 18 functions of
FG.cpp:2322-3662 are cut out of the reference tree at build time and compiled unmodified into oracle/_ref/libfg_ref{64,32}.so
(oracle/Makefile, oracle/ref_fg/, binding oracle/ref_solver.py); tests/test_reference_solver.py holds identical PCG iteration
counts and <= 1e-9 on every output and on the whole fit.  What stays restated from the source text only: the R-level driver
(Get_Coef, the outer AI-REML loop, the LOCO loop, extractVarianceRatio; R is not available here).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FP64, REF32 = 0, 1


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsaige_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "libsaige_oracle.so"])
        L = C.CDLL(so)
        L.orc_new.restype = C.c_void_p
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_setgeno.restype = C.c_int
        L.orc_setgeno.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int64]
        for n in ("orc_get_M", "orc_get_M0", "orc_get_N", "orc_get_Mvr", "orc_get_B"):
            getattr(L, n).restype = C.c_int64
            getattr(L, n).argtypes = [C.c_void_p]
        for n in ("orc_afreq", "orc_invstd", "orc_mac", "orc_ac", "orc_packed"):
            getattr(L, n).restype = C.c_void_p
            getattr(L, n).argtypes = [C.c_void_p, C.c_int]
        L.orc_index_vr.restype = C.c_void_p
        L.orc_index_vr.argtypes = [C.c_void_p]
        L.orc_qc_mask.restype = C.c_void_p
        L.orc_qc_mask.argtypes = [C.c_void_p]
        L.orc_one_snp_geno.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.orc_one_snp_stdgeno.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.orc_crossprod_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_diag_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        L.orc_synth_bed.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _view(addr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


# ----------------------------------------------------------------------------------------------------
# synthetic genotypes -- shared bit-exactly with the CUDA library's generator
# ----------------------------------------------------------------------------------------------------
def synth_thresholds(M, seed):
    """Per-marker A1 frequency f ~ U(0.05, 0.5) and integer genotype thresholds (uint32)."""
    rng = np.random.default_rng(seed)
    f = rng.uniform(0.05, 0.5, size=M)
    t0 = np.floor((1 - f) ** 2 * 4294967296.0).astype(np.uint64).clip(0, 4294967295).astype(np.uint32)
    t1 = np.floor((1 - f * f) * 4294967296.0).astype(np.uint64).clip(0, 4294967295).astype(np.uint32)
    return f, t0, t1


def synth_bed(N0, M0, seed, miss_rate=0.0):
    """Body of a PLINK .bed (no magic bytes), SNP-major, ceil(N0/4) bytes per marker."""
    _, t0, t1 = synth_thresholds(M0, seed)
    bed = np.zeros(((N0 + 3) // 4) * M0, dtype=np.uint8)
    thr = int(miss_rate * 4294967296.0)
    lib().orc_synth_bed(_ptr(bed), N0, M0, seed, _ptr(t0), _ptr(t1), thr)
    return bed


def read_bed(prefix):
    """Returns (bed body uint8, N0, M0, chr per marker) from <prefix>.bed/.bim/.fam (FG.cpp:760-787,902)."""
    with open(prefix + ".fam") as f:
        N0 = sum(1 for _ in f)
    chrs = []
    with open(prefix + ".bim") as f:
        for line in f:
            chrs.append(line.split()[0])
    M0 = len(chrs)
    raw = np.fromfile(prefix + ".bed", dtype=np.uint8)
    assert raw[0] == 0x6C and raw[1] == 0x1B and raw[2] == 0x01, "not a SNP-major PLINK .bed"
    return raw[3:].copy(), N0, M0, chrs


# ----------------------------------------------------------------------------------------------------
# genoClass (FG.cpp:37-1183) and its exported accessors
# ----------------------------------------------------------------------------------------------------
class OracleGeno:
    """State of the reference's file-global `genoClass geno` (FG.cpp:1188) plus flags set through exports."""

    def __init__(self, mode=FP64):
        self.mode = mode
        self._g = C.c_void_p(lib().orc_new())
        self.minMAF = 0.0            # minMAFtoConstructGRM (FG.cpp:35, setminMAFforGRM FG.cpp:4923)
        self.maxMissing = 1.0        # geno.maxMissingRate (setmaxMissingRateforGRM FG.cpp:4928)
        self.isVarRatio = False      # setminMAC_VarianceRatio (FG.cpp:4970)
        self.minMACvr, self.maxMACvr = 20.0, -1.0
        self.setKinDiagtoOne = False
        self._diag = None
        self._diag_loco = None
        self.startIndexVec = self.endIndexVec = None
        self.startIndex = self.endIndex = self.chromIndex = None

    def __del__(self):
        try:
            lib().orc_free(self._g)
        except Exception:
            pass

    def setgeno(self, bed, N0, M0, subSampleInGeno, indicator, isDiagofKinSetAsOne=False, vr_rand_idx=None):
        """setgeno -> genoClass::setGenoObj (FG.cpp:2267, 739-1024)."""
        L = lib()
        self.setKinDiagtoOne = bool(isDiagofKinSetAsOne)
        sub = np.ascontiguousarray(subSampleInGeno, dtype=np.int32)
        ind = np.ascontiguousarray(indicator, dtype=np.uint8)
        vr = np.ascontiguousarray(vr_rand_idx if vr_rand_idx is not None else [], dtype=np.int32)
        bed = np.ascontiguousarray(bed, dtype=np.uint8)
        rc = L.orc_setgeno(self._g, _ptr(bed), N0, M0, _ptr(sub), len(sub), _ptr(ind), self.minMAF, self.maxMissing,
                           int(self.isVarRatio), self.minMACvr, self.maxMACvr, _ptr(vr), len(vr))
        assert rc == 0
        self.N, self.M, self.M0, self.Mvr, self.B = (L.orc_get_N(self._g), L.orc_get_M(self._g), L.orc_get_M0(self._g),
                                                     L.orc_get_Mvr(self._g), L.orc_get_B(self._g))
        self.alleleFreqVec = _view(L.orc_afreq(self._g, 0), self.M, np.float32)
        self.invstdvVec = _view(L.orc_invstd(self._g, 0), self.M, np.float32)
        self.MACVec = _view(L.orc_mac(self._g, 0), self.M, np.int32)
        self.ACVec = _view(L.orc_ac(self._g, 0), self.M, np.int32)
        self.qc_mask = _view(L.orc_qc_mask(self._g), self.M0, np.uint8).astype(bool)
        self.MACVec_forVarRatio = _view(L.orc_mac(self._g, 1), self.Mvr, np.int32)
        self.markerIndexVec_forVarRatio = _view(L.orc_index_vr(self._g), self.Mvr, np.int32)
        self.alleleFreqVec_forVarRatio = _view(L.orc_afreq(self._g, 1), self.Mvr, np.float32)
        self._diag = None
        self._diag_loco = None

    def packed(self, vr=False):
        L = lib()
        n = (self.Mvr if vr else self.M) * self.B
        return _view(L.orc_packed(self._g, int(vr)), n, np.uint8).reshape(-1, self.B)

    def Get_OneSNP_Geno(self, idx, vr=False):
        out = np.zeros(self.N, dtype=np.int32)
        lib().orc_one_snp_geno(self._g, idx, int(vr), _ptr(out))
        return out

    def Get_OneSNP_StdGeno(self, idx):
        out = np.zeros(self.N, dtype=np.float64)
        lib().orc_one_snp_stdgeno(self._g, idx, self.mode, _ptr(out))
        return out

    # -- GRM.vector (FG.cpp:1576-1598, 1669-1708, 1746-1851, 1953-2006) --
    def _crossprod_range(self, m0, m1, b):
        b = np.asarray(b, dtype=np.float64)
        k = 1 if b.ndim == 1 else b.shape[1]
        bf = np.asfortranarray(b.reshape(self.N, k))
        out = np.zeros((self.N, k), dtype=np.float64, order="F")
        lib().orc_crossprod_range(self._g, m0, m1, _ptr(bf), k, self.mode, _ptr(out))
        return out[:, 0].copy() if b.ndim == 1 else out

    def getCrossprodMatAndKin(self, b):
        """parallelCrossProd: sum over all QC'd markers / Mqc (FG.cpp:1669-1708)."""
        return self._crossprod_range(0, self.M, b) / self.M

    def getCrossprodMatAndKin_LOCO(self, b):
        """parallelCrossProd_LOCO: (full - [start,end]) / (Mfull - Mchr) (FG.cpp:1790-1851)."""
        full = self._crossprod_range(0, self.M, b)
        sub = self._crossprod_range(self.startIndex, self.endIndex + 1, b)
        return (full - sub) / (self.M - (self.endIndex - self.startIndex + 1))

    # -- diagonals (FG.cpp:665-704, 4358-4375, 4934-4958, 709-729) --
    def Get_Diagof_StdGeno(self):
        if self._diag is None:
            out = np.zeros(self.N, dtype=np.float64)
            lib().orc_diag_range(self._g, 0, self.M, self.mode, _ptr(out))
            self._diag = out
        return self._diag

    def get_DiagofKin(self):
        if self.setKinDiagtoOne:
            return np.ones(self.N)
        return self.Get_Diagof_StdGeno() / self.M

    def setStartEndIndexVec(self, start, end):
        self.startIndexVec = np.asarray(start, dtype=np.int64)
        self.endIndexVec = np.asarray(end, dtype=np.int64)

    def setStartEndIndex(self, start, end, chromIndex):
        self.startIndex, self.endIndex, self.chromIndex = int(start), int(end), int(chromIndex)

    def set_Diagof_StdGeno_LOCO(self):
        n = len(self.startIndexVec)
        self._diag_loco = np.zeros((self.N, n))
        self.Msub_byChr = np.zeros(n, dtype=np.int64)
        full = self.Get_Diagof_StdGeno()
        for k in range(n):
            s, e = self.startIndexVec[k], self.endIndexVec[k]
            if s != -1 and e != -1:
                out = np.zeros(self.N)
                lib().orc_diag_range(self._g, int(s), int(e) + 1, self.mode, _ptr(out))
                self._diag_loco[:, k] = full - out
                self.Msub_byChr[k] = e - s + 1

    # -- Sigma = tau0 diag(1/W) + tau1 K (FG.cpp:2322-2443) --
    def getDiagOfSigma(self, w, tau, loco=False):
        if loco:
            d = tau[1] * self._diag_loco[:, self.chromIndex] / (self.M - self.Msub_byChr[self.chromIndex]) + tau[0] / w
        elif not self.setKinDiagtoOne:
            d = tau[1] * self.Get_Diagof_StdGeno() / self.M + tau[0] / w
        else:
            d = tau[1] + tau[0] / w
        return np.maximum(d, 1e-4)          # FG.cpp:2355-2357

    def getCrossprod(self, b, w, tau, loco=False):
        wcol = w if b.ndim == 1 else w[:, None]
        if tau[1] == 0:                      # FG.cpp:2401-2404
            return tau[0] * (b / wcol)
        kb = self.getCrossprodMatAndKin_LOCO(b) if loco else self.getCrossprodMatAndKin(b)
        return tau[0] * (b / wcol) + tau[1] * kb

    # -- getPCG1ofSigmaAndVector[_LOCO] (FG.cpp:2593-2809, 2943-3034) --
    def getPCG1ofSigmaAndVector(self, w, tau, b, maxiterPCG, tolPCG, loco=False, return_iter=False):
        b = np.asarray(b, dtype=np.float64)
        r = b.copy()
        minv = 1.0 / self.getDiagOfSigma(w, tau, loco)
        z = minv * r
        sumr2 = float(r @ r)
        p = z.copy()
        x = np.zeros(self.N)
        it = 0
        while sumr2 > tolPCG and it < maxiterPCG:
            it += 1
            Ap = self.getCrossprod(p, w, tau, loco)
            a = float(r @ z) / float(p @ Ap)
            x = x + a * p
            r1 = r - a * Ap
            z1 = minv * r1
            bet = float(z1 @ r1) / float(z @ r)
            p = z1 + bet * p
            z, r = z1, r1
            sumr2 = float(r @ r)
        return (x, it) if return_iter else x

    def pcg_multi(self, w, tau, Bm, maxiterPCG, tolPCG, loco=False):
        Bm = np.asarray(Bm, dtype=np.float64).reshape(self.N, -1)
        out = np.zeros_like(Bm)
        iters = []
        for c in range(Bm.shape[1]):
            out[:, c], it = self.getPCG1ofSigmaAndVector(w, tau, Bm[:, c], maxiterPCG, tolPCG, loco, True)
            iters.append(it)
        return out, iters


# ----------------------------------------------------------------------------------------------------
# AI-REML pieces (FG.cpp:3104-3341, 3347-3405, 3409-3662)
# ----------------------------------------------------------------------------------------------------
def inv_sympd_or_pinv(A):
    """arma::inv_sympd(symmatu(A)) with pinv fallback (FG.cpp:3185-3190)."""
    A = np.triu(A) + np.triu(A, 1).T
    try:
        Lc = np.linalg.cholesky(A)
        Li = np.linalg.inv(Lc)
        return Li.T @ Li
    except np.linalg.LinAlgError:
        return np.linalg.pinv(A)


def calCV(x):
    """FG.cpp:3104-3110 (note the extra division by the length)."""
    x = np.asarray(x, dtype=np.float64)
    return (np.std(x, ddof=1) / np.mean(x)) / len(x)


def getCoefficients(g, Y, X, w, tau, maxiterPCG, tolPCG, loco=False):
    """FG.cpp:3167-3196 / _LOCO 3203-3233."""
    Sigma_iY = g.getPCG1ofSigmaAndVector(w, tau, Y, maxiterPCG, tolPCG, loco)
    Sigma_iX = np.column_stack([g.getPCG1ofSigmaAndVector(w, tau, X[:, i], maxiterPCG, tolPCG, loco)
                                for i in range(X.shape[1])])
    cov = inv_sympd_or_pinv(X.T @ Sigma_iX)
    alpha = cov @ (Sigma_iX.T @ Y)
    eta = Y - tau[0] * (Sigma_iY - Sigma_iX @ alpha) / w
    return dict(Sigma_iY=Sigma_iY, Sigma_iX=Sigma_iX, cov=cov, alpha=alpha, eta=eta)


def GetTrace(g, Sigma_iX, X, w, tau, cov1, nrun, maxiterPCG, tolPCG, traceCVcutoff, draw, quantitative=False):
    """GetTrace / GetTrace_q (FG.cpp:3113-3160, 3409-3472).  `draw(n)` returns the next n Rademacher
    probe vectors (N x n), i.e. 2*rbinom(N,1,0.5)-1 in the reference's R RNG stream (FG.cpp:3052-3054,3134-3137)."""
    t1, t0 = [], []
    nstart, nend = 0, nrun
    while True:
        U = draw(nend - nstart)
        for i in range(nend - nstart):
            u = U[:, i]
            Sigma_iu = g.getPCG1ofSigmaAndVector(w, tau, u, maxiterPCG, tolPCG)
            Pu = Sigma_iu - Sigma_iX @ (cov1 @ (Sigma_iX.T @ u))
            Au = g.getCrossprodMatAndKin(u)
            t1.append(float(Au @ Pu))
            t0.append(float(u @ Pu))
        cv1 = calCV(t1)
        cv0 = calCV(t0) if quantitative else 0.0
        if cv1 > traceCVcutoff or cv0 > traceCVcutoff:
            nstart, nend = nend, nend + 10
        else:
            break
    if quantitative:
        return np.array([np.mean(t0), np.mean(t1)]), nend
    return float(np.mean(t1)), nend


def getAIScore(g, Y, X, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff, draw):
    """FG.cpp:3279-3295."""
    PY = Sigma_iY - Sigma_iX @ (cov @ (Sigma_iX.T @ Y))
    APY = g.getCrossprodMatAndKin(PY)
    YPAPY = float(PY @ APY)
    Trace, nused = GetTrace(g, Sigma_iX, X, w, tau, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff, draw)
    PAPY_1 = g.getPCG1ofSigmaAndVector(w, tau, APY, maxiterPCG, tolPCG)
    PAPY = PAPY_1 - Sigma_iX @ (cov @ (Sigma_iX.T @ PAPY_1))
    AI = float(APY @ PAPY)
    return dict(YPAPY=YPAPY, Trace=Trace, PY=PY, AI=AI, nrun_used=nused)


def fitglmmaiRPCG(g, Y, X, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, tol, traceCVcutoff, draw):
    """FG.cpp:3302-3341."""
    re = getAIScore(g, Y, X, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff, draw)
    Dtau = (re["YPAPY"] - re["Trace"]) / re["AI"]
    tau0 = np.array(tau, dtype=np.float64)
    tau = tau0.copy()
    tau[1] = tau0[1] + Dtau
    tau[tau < tol] = 0
    step = 1.0
    while tau[1] < 0.0:
        step *= 0.5
        tau[1] = tau0[1] + step * Dtau
    tau[tau < tol] = 0
    return tau


def getAIScore_q(g, Y, X, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff, draw):
    """FG.cpp:3479-3531 (recomputes cov from X^T Sigma_iX)."""
    cov1 = inv_sympd_or_pinv(X.T @ Sigma_iX)
    PY = Sigma_iY - Sigma_iX @ (cov1 @ (Sigma_iX.T @ Y))
    APY = g.getCrossprodMatAndKin(PY)
    YPAPY = float(PY @ APY)
    A0PY = PY
    YPA0PY = float(PY @ A0PY)
    Trace, nused = GetTrace(g, Sigma_iX, X, w, tau, cov1, nrun, maxiterPCG, tolPCG, traceCVcutoff, draw, True)
    PA0PY_1 = g.getPCG1ofSigmaAndVector(w, tau, A0PY, maxiterPCG, tolPCG)
    PA0PY = PA0PY_1 - Sigma_iX @ (cov1 @ (Sigma_iX.T @ PA0PY_1))
    PAPY_1 = g.getPCG1ofSigmaAndVector(w, tau, APY, maxiterPCG, tolPCG)
    PAPY = PAPY_1 - Sigma_iX @ (cov1 @ (Sigma_iX.T @ PAPY_1))
    AI = np.array([[A0PY @ PA0PY, A0PY @ PAPY], [A0PY @ PAPY, APY @ PAPY]])
    return dict(YPAPY=YPAPY, YPA0PY=YPA0PY, Trace=Trace, PY=PY, AI=AI, nrun_used=nused)


def fitglmmaiRPCG_q(g, Y, X, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, tol, traceCVcutoff, draw):
    """FG.cpp:3610-3662."""
    tau = np.array(tau, dtype=np.float64)
    zeroVec = tau < tol
    re = getAIScore_q(g, Y, X, w, tau, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff, draw)
    score = np.array([re["YPA0PY"] - re["Trace"][0], re["YPAPY"] - re["Trace"][1]])
    Dtau = np.linalg.solve(re["AI"], score)
    tau0 = tau.copy()
    tau = tau0 + Dtau
    tau[zeroVec & (tau < tol)] = 0
    step = 1.0
    while tau[0] < 0.0 or tau[1] < 0.0:
        step *= 0.5
        tau = tau0 + step * Dtau
        tau[zeroVec & (tau < tol)] = 0
    tau[tau < tol] = 0
    return tau


def getSigma_X(g, w, tau, X, maxiterPCG, tolPCG, loco=False):
    """FG.cpp:3347-3371."""
    return np.column_stack([g.getPCG1ofSigmaAndVector(w, tau, X[:, i], maxiterPCG, tolPCG, loco)
                            for i in range(X.shape[1])])


def getSigma_G(g, w, tau, Gvec, maxiterPCG, tolPCG, loco=False):
    """FG.cpp:3374-3384."""
    return g.getPCG1ofSigmaAndVector(w, tau, Gvec, maxiterPCG, tolPCG, loco)


# ----------------------------------------------------------------------------------------------------
# R driver restated (FG.R) -- binomial family with logit link / gaussian identity
# ----------------------------------------------------------------------------------------------------
class Binomial:
    name = "binomial"
    linkinv = staticmethod(lambda eta: 1.0 / (1.0 + np.exp(-eta)))
    mu_eta = staticmethod(lambda eta: np.maximum(np.exp(-np.abs(eta)) / (1.0 + np.exp(-np.abs(eta))) ** 2,
                                                 np.finfo(float).eps))
    variance = staticmethod(lambda mu: mu * (1.0 - mu))


class Gaussian:
    name = "gaussian"
    linkinv = staticmethod(lambda eta: eta)
    mu_eta = staticmethod(lambda eta: np.ones_like(eta))
    variance = staticmethod(lambda mu: np.ones_like(mu))


def glm_fit(y, X, family, offset=None, maxit=25, epsilon=1e-8):
    """R's glm.fit IRLS (used at FG.R:1119 for the start values fit0)."""
    n = len(y)
    offset = np.zeros(n) if offset is None else offset
    if family.name == "binomial":
        mu = (y + 0.5) / 2.0
        eta = np.log(mu / (1 - mu))
    else:
        mu = y.copy()
        eta = mu.copy()
    devold = np.inf
    coef = np.zeros(X.shape[1])
    for _ in range(maxit):
        me = family.mu_eta(eta)
        z = (eta - offset) + (y - mu) / me
        wt = me ** 2 / family.variance(mu)
        sw = np.sqrt(wt)
        coef, *_ = np.linalg.lstsq(X * sw[:, None], z * sw, rcond=None)
        eta = X @ coef + offset
        mu = family.linkinv(eta)
        if family.name == "binomial":
            with np.errstate(divide="ignore", invalid="ignore"):
                d = 2 * (np.where(y > 0, y * np.log(y / mu), 0) + np.where(y < 1, (1 - y) * np.log((1 - y) / (1 - mu)), 0))
            dev = d.sum()
        else:
            dev = ((y - mu) ** 2).sum()
        if abs(dev - devold) / (abs(dev) + 0.1) < epsilon:
            break
        devold = dev
    return dict(coef=coef, eta=eta, mu=mu, y=y, offset=offset, family=family, X=X)


def Covariate_Transform(X1):
    """FG.R:1612-1647 (QR step only; collinearity screening is the caller's business)."""
    Q, R = np.linalg.qr(X1)
    return Q * np.sqrt(X1.shape[0]), R


def ScoreTest_NULL_Model(mu, mu2, y, X):
    """FG.R:579-591."""
    V = mu2
    res = y - mu
    XV = (X * V[:, None]).T
    XVX = X.T @ XV.T
    XVX_inv = np.linalg.inv(XVX)
    XXVX_inv = X @ XVX_inv
    return dict(XV=XV, XVX=XVX, XXVX_inv=XXVX_inv, XVX_inv=XVX_inv, S_a=(X * res[:, None]).sum(0),
                XVX_inv_XV=XXVX_inv * V[:, None], V=V)


def Get_Coef(g, y, X, tau, family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter, loco=False):
    """FG.R:2-35 / Get_Coef_LOCO 42-73."""
    tol_coef = 0.1
    mu = family.linkinv(eta0)
    me = family.mu_eta(eta0)
    Y = eta0 - offset + (y - mu) / me
    sqrtW = me / np.sqrt(family.variance(mu))
    W = sqrtW ** 2
    for _ in range(maxiter):
        rc = getCoefficients(g, Y, X, W, tau, maxiterPCG, tolPCG, loco)
        alpha = rc["alpha"]
        eta = rc["eta"] + offset
        mu = family.linkinv(eta)
        me = family.mu_eta(eta)
        Y = eta - offset + (y - mu) / me
        sqrtW = me / np.sqrt(family.variance(mu))
        W = sqrtW ** 2
        if np.max(np.abs(alpha - alpha0) / (np.abs(alpha) + np.abs(alpha0) + tol_coef)) < tol_coef:
            break
        alpha0 = alpha
    return dict(Y=Y, alpha=alpha, eta=eta, W=W, cov=rc["cov"], sqrtW=sqrtW, Sigma_iY=rc["Sigma_iY"],
                Sigma_iX=rc["Sigma_iX"], mu=mu)


def make_draw(U):
    """Probe source that hands out columns of a pre-drawn N x nmax Rademacher matrix in order, restarting
    at column 0 on every GetTrace call -- the reference resets the seed to 200 each call (FG.cpp:3114)."""
    def factory():
        pos = [0]

        def draw(n):
            out = U[:, pos[0]:pos[0] + n]
            assert out.shape[1] == n, "probe matrix exhausted"
            pos[0] += n
            return out
        return draw
    return factory


def glmmkin_ai_PCG(g, fit0, tau_init, U, trait="binary", maxiter=20, tol=0.02, nrun=30, tolPCG=1e-5,
                   maxiterPCG=500, traceCVcutoff=0.0025, LOCO=False, log=None):
    """glmmkin.ai_PCG_Rcpp_Binary (FG.R:79-304) / _Quantitative (FG.R:309-549), after setgeno."""
    y, X, offset, family = fit0["y"], fit0["X"], fit0["offset"], fit0["family"]
    n = len(y)
    eta = fit0["eta"]
    alpha0 = fit0["coef"]
    eta0 = eta
    draws = make_draw(U)
    tau = np.array([0.0, 0.0])
    quant = trait == "quantitative"
    if not quant:
        tau[0] = 1.0
        tau[1] = 0.1 if tau_init[1] == 0 else tau_init[1]       # FG.R:145-158
    else:
        if np.sum(tau_init) == 0:
            tau[:] = (1.0, 0.0)                                   # FG.R:389-398
        else:
            tau[:] = tau_init
    tau0 = tau.copy()
    rc = Get_Coef(g, y, X, tau, family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter)
    if not quant:
        re = getAIScore(g, rc["Y"], X, rc["W"], tau, rc["Sigma_iY"], rc["Sigma_iX"], rc["cov"], nrun, maxiterPCG,
                        tolPCG, traceCVcutoff, draws())
        tau[1] = max(0.0, tau0[1] + tau0[1] ** 2 * (re["YPAPY"] - re["Trace"]) / n)       # FG.R:165
    else:
        re = getAIScore_q(g, rc["Y"], X, rc["W"], tau, rc["Sigma_iY"], rc["Sigma_iX"], rc["cov"], nrun, maxiterPCG,
                          tolPCG, traceCVcutoff, draws())
        tau[1] = max(0.0, tau0[1] + tau0[1] ** 2 * (re["YPAPY"] - re["Trace"][1]) / n)    # FG.R:405-406
        tau[0] = max(0.0, tau0[0] + tau0[0] ** 2 * (re["YPA0PY"] - re["Trace"][0]) / n)
    if log is not None:
        log.append(("init", tau.copy()))
    alpha = fit0["coef"] if quant else rc["alpha"]    # FG.R:343 (quantitative keeps fit0's alpha until iteration 1)
    i = 0
    for i in range(1, maxiter + 1):
        alpha0 = rc["alpha"] if not quant else alpha
        tau0 = tau.copy()
        eta0 = eta
        rc = Get_Coef(g, y, X, tau, family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter)
        fit = (fitglmmaiRPCG_q if quant else fitglmmaiRPCG)(g, rc["Y"], X, rc["W"], tau, rc["Sigma_iY"], rc["Sigma_iX"],
                                                           rc["cov"], nrun, maxiterPCG, tolPCG, tol, traceCVcutoff, draws())
        tau = np.asarray(fit, dtype=np.float64)
        alpha, eta = rc["alpha"], rc["eta"]
        if log is not None:
            log.append((i, tau.copy()))
        if quant and tau[0] <= 0:
            raise RuntimeError("ERROR! The first variance component parameter estimate is 0")
        if (tau[1] == 0 and not quant) or (quant and tau[1] <= 0):
            break
        if np.max(np.abs(tau - tau0) / (np.abs(tau) + np.abs(tau0) + tol)) < tol:
            break
        if np.max(tau) > tol ** (-2):
            i = maxiter
            break
    rc = Get_Coef(g, y, X, tau, family, alpha, eta, offset, maxiterPCG, tolPCG, maxiter)
    alpha, eta, mu = rc["alpha"], rc["eta"], rc["mu"]
    res = y - mu
    mu2 = mu * (1 - mu) if not quant else np.full(n, 1.0 / tau[0])
    out = dict(theta=tau, coefficients=alpha, linear_predictors=eta, fitted_values=mu, Y=rc["Y"], residuals=res,
               cov=rc["cov"], converged=i < maxiter, obj_noK=ScoreTest_NULL_Model(mu, mu2, y, X), y=y, X=X,
               traitType=trait, LOCO=LOCO)
    if LOCO:
        g.set_Diagof_StdGeno_LOCO()
        out["LOCOResult"] = []
        for j in range(len(g.startIndexVec)):
            s, e = g.startIndexVec[j], g.endIndexVec[j]
            if s == -1 or e == -1:
                out["LOCOResult"].append(dict(isLOCO=False))
                continue
            g.setStartEndIndex(s, e, j)
            rl = Get_Coef(g, y, X, tau, family, alpha, eta, offset, maxiterPCG, tolPCG, maxiter, loco=True)
            alpha, eta, mu = rl["alpha"], rl["eta"], rl["mu"]      # FG.R:267-271: chained start values
            mu2 = mu * (1 - mu) if not quant else np.full(n, 1.0 / tau[0])
            out["LOCOResult"].append(dict(isLOCO=True, coefficients=alpha, linear_predictors=eta, fitted_values=mu,
                                          Y=rl["Y"], residuals=y - mu, cov=rl["cov"],
                                          obj_noK=ScoreTest_NULL_Model(mu, mu2, y, X)))
    return out


def updateChrStartEndIndexVec(chrVec):
    """Util.R:29-65 -- 0-based first/last post-QC marker index of chromosomes 1..22, -1 where absent."""
    chrVec = np.asarray(chrVec)
    start, end = [], []
    for c in range(1, 23):
        idx = np.nonzero(chrVec == c)[0]
        if len(idx):
            start.append(int(idx.min())); end.append(int(idx.max()))
        else:
            start.append(-1); end.append(-1)
    LOCO = sum(s != -1 for s in start) > 1
    return LOCO, np.array(start), np.array(end)


def extractVarianceRatio(g, model, family, marker_order, numMarkers=30, maxiterPCG=500, tolPCG=1e-5,
                         ratioCVcutoff=0.001, chr_of_marker=None):
    """FG.R:2152-2423, single (non-categorical) ratio, full-GRM path.  `marker_order` is the caller's
    permutation `sample(MACindex)` (FG.R:2242) over the VR store (or the GRM store when no hold-out)."""
    mu, eta, y, X = model["fitted_values"], model["linear_predictors"], model["y"], model["X"]
    tau = model["theta"]
    noK = model["obj_noK"]
    me = family.mu_eta(eta)
    W = (me / np.sqrt(family.variance(mu))) ** 2
    Sigma_iX = getSigma_X(g, W, tau, X, maxiterPCG, tolPCG)
    use_vr = g.isVarRatio
    N = g.N
    ratios = []
    pos = 0
    numMarkers0 = numMarkers
    while True:
        while len(ratios) < numMarkers0 and pos < len(marker_order):
            i = marker_order[pos]
            pos += 1
            G0 = g.Get_OneSNP_Geno(i, vr=use_vr).astype(np.float64)
            if G0.sum() / (2 * N) > 0.5:
                G0 = 2 - G0
            AC = G0.sum()
            Gt = G0 - noK["XXVX_inv"] @ (noK["XV"] @ G0)
            gn = Gt / np.sqrt(AC)
            Sigma_iG = getSigma_G(g, W, tau, Gt, maxiterPCG, tolPCG)
            var1a = Gt @ Sigma_iG - Gt @ Sigma_iX @ np.linalg.solve(X.T @ Sigma_iX, X.T @ Sigma_iG)
            var1 = var1a / AC
            var2null = float((mu * (1 - mu)) @ (gn * gn)) if model["traitType"] == "binary" else float(gn @ gn)
            ratios.append(var1 / var2null)
        cv = calCV(ratios)
        if cv > ratioCVcutoff and pos < len(marker_order):
            numMarkers0 += 10
        else:
            break
    return float(np.mean(ratios)), ratios
