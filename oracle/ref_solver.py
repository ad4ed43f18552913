"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libfg_ref{64,32}.so -- the reference's OWN solver-layer functions
(getDiagOfSigma, getCrossprod, getPCG1ofSigmaAndVector, getCoefficients, GetTrace[_q], getAIScore[_q], fitglmmaiRPCG[_q],
getSigma_X / _G, calCV and the _LOCO twins of /root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp:2322-3662), cut out of the
reference tree at build time and compiled unmodified against oracle/ref_fg/mini_arma.h (recipe: oracle/Makefile, `ref`).
The 64 variant reads the reference's `float` as `double`; the 32 variant is the reference as shipped.

The GRM product and the GRM diagonal come from an OracleGeno (oracle/oracle.py), which other artefacts pin; what this binding
pins is everything ABOVE the product: the PCG recurrence and its stopping rule, the covariance algebra, the trace estimator
with its CV retries, the AI score and the tau update.  Never imported by the product."""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)
DP = C.c_void_p


def available(bits=64):
    """bits: 64 / 32 = the solver layer over the oracle's product; "cpu" = the reference's whole CPU path (libfg_refcpu.so)"""
    return os.path.exists(os.path.join(_HERE, "_ref", "libfg_ref%s.so" % bits))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class RefSolver:
    """The reference solver on top of an OracleGeno's product / diagonal.  `U`: N x nmax matrix of +-1 probes; the reference
    draws rbinom(N, 1, 0.5) and maps 0/1 to -1/+1 (FG.cpp:3134-3137), so it is fed (U + 1) / 2 column by column."""

    def __init__(self, geno, bits=64, U=None):
        self.L = C.CDLL(os.path.join(_HERE, "_ref", "libfg_ref%d.so" % bits))
        L = self.L
        self._signatures()
        L.fgref_set_problem.argtypes = [C.c_int, C.c_int, DP, C.c_int, CB]
        L.fgref_set_loco.argtypes = [DP, C.c_int, C.c_int, CB]
        assert L.fgref_real_bytes() == bits // 8
        self.g, self.N = geno, geno.N
        self.products = 0

        def cb(bp, op, n):
            b = np.ctypeslib.as_array(bp, shape=(n,))
            np.ctypeslib.as_array(op, shape=(n,))[:] = geno.getCrossprodMatAndKin(b.copy())
            self.products += 1

        def cb_loco(bp, op, n):
            b = np.ctypeslib.as_array(bp, shape=(n,))
            np.ctypeslib.as_array(op, shape=(n,))[:] = geno.getCrossprodMatAndKin_LOCO(b.copy())
            self.products += 1
        self._cb, self._cb_loco = CB(cb), CB(cb_loco)            # keep the thunks alive
        self._diag = _f(geno.Get_Diagof_StdGeno())
        L.fgref_set_problem(geno.N, geno.M, _p(self._diag), int(bool(geno.setKinDiagtoOne)), self._cb)
        self._draws = None
        if U is not None:
            self.set_probes(U)

    def _signatures(self):
        L = self.L
        L.fgref_last_error.restype = C.c_char_p
        L.fgref_log.restype = C.c_char_p
        L.fgref_cal_cv.restype = C.c_double
        L.fgref_draws_used.restype = C.c_long
        L.fgref_cal_cv.argtypes = [DP, C.c_int]
        L.fgref_set_draws.argtypes = [DP, C.c_long]
        L.fgref_diag_of_sigma.argtypes = [DP, DP, C.c_int, DP]
        L.fgref_pcg.argtypes = [DP, DP, DP, C.c_int, C.c_double, C.c_int, DP]
        L.fgref_get_coefficients.argtypes = [DP, DP, C.c_int, DP, DP, C.c_int, C.c_double, C.c_int, DP, DP, DP, DP, DP]
        L.fgref_get_ai_score.argtypes = [C.c_int, DP, DP, C.c_int, DP, DP, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_double, DP, DP]
        L.fgref_fit_glmmai_rpcg.argtypes = [C.c_int, DP, DP, C.c_int, DP, DP, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_double,
                                            C.c_double]
        L.fgref_get_sigma_x.argtypes = [DP, DP, DP, C.c_int, C.c_int, C.c_double, DP]
        L.fgref_get_sigma_g.argtypes = [DP, DP, DP, C.c_int, C.c_double, DP]

    def set_probes(self, U):
        self._draws = _f(((np.asarray(U) + 1.0) / 2.0).T.reshape(-1))        # column after column
        self.L.fgref_set_draws(_p(self._draws), self._draws.size)

    def set_loco_chromosome(self, c):
        """The state setStartEndIndex(start_c, end_c, c) + set_Diagof_StdGeno_LOCO leave in the reference's genotype object."""
        g = self.g
        g.setStartEndIndex(g.startIndexVec[c], g.endIndexVec[c], c)
        self._dl = _f(g._diag_loco[:, c])
        self.L.fgref_set_loco(_p(self._dl), int(g.M), int(g.Msub_byChr[c]), self._cb_loco)

    def _ck(self, rc):
        if rc:
            raise RuntimeError(self.L.fgref_last_error().decode())

    def pcg_iterations(self):
        """iteration counts the reference printed since the last call ("iter from getPCG1ofSigmaAndVector <n>", FG.cpp:2798)"""
        text = self.L.fgref_log().decode()
        self.L.fgref_clear_log()
        return [int(m) for m in re.findall(r"iter from getPCG1ofSigmaAndVector (\d+)", text)]

    def calCV(self, x):
        x = _f(x)
        return self.L.fgref_cal_cv(_p(x), len(x))

    def getDiagOfSigma(self, w, tau, loco=False):
        w, tau, out = _f(w), _f(tau), np.zeros(self.N)
        self._ck(self.L.fgref_diag_of_sigma(_p(w), _p(tau), int(loco), _p(out)))
        return out

    def getPCG1ofSigmaAndVector(self, w, tau, b, maxiterPCG, tolPCG, loco=False, return_iter=False):
        w, tau, b, x = _f(w), _f(tau), _f(b), np.zeros(self.N)
        self.L.fgref_clear_log()
        self._ck(self.L.fgref_pcg(_p(w), _p(tau), _p(b), int(maxiterPCG), float(tolPCG), int(loco), _p(x)))
        return (x, self.pcg_iterations()[0]) if return_iter else x

    def getCoefficients(self, Y, X, w, tau, maxiterPCG, tolPCG, loco=False):
        Y, X, w, tau = _f(Y), np.asfortranarray(X, dtype=np.float64), _f(w), _f(tau)
        p = X.shape[1]
        SiY, SiX, cov = np.zeros(self.N), np.zeros((self.N, p), order="F"), np.zeros((p, p), order="F")
        alpha, eta = np.zeros(p), np.zeros(self.N)
        self._ck(self.L.fgref_get_coefficients(_p(Y), _p(X), p, _p(w), _p(tau), int(maxiterPCG), float(tolPCG), int(loco), _p(SiY), _p(SiX),
                                               _p(cov), _p(alpha), _p(eta)))
        return dict(Sigma_iY=SiY, Sigma_iX=SiX, cov=cov, alpha=alpha, eta=eta)

    def _ai(self, quant, Y, X, w, tau, SiY, SiX, cov, nrun, maxiterPCG, tolPCG, cvcut):
        Y, X, w, tau, SiY = _f(Y), np.asfortranarray(X, dtype=np.float64), _f(w), _f(tau), _f(SiY)
        SiX, cov = np.asfortranarray(SiX, dtype=np.float64), np.asfortranarray(cov, dtype=np.float64)
        o, PY = np.zeros(8), np.zeros(self.N)
        self._ck(self.L.fgref_get_ai_score(int(quant), _p(Y), _p(X), X.shape[1], _p(w), _p(tau), _p(SiY), _p(SiX), _p(cov), int(nrun),
                                           int(maxiterPCG), float(tolPCG), float(cvcut), _p(o), _p(PY)))
        return o, PY

    def getAIScore(self, Y, X, w, tau, SiY, SiX, cov, nrun, maxiterPCG, tolPCG, cvcut):
        o, PY = self._ai(False, Y, X, w, tau, SiY, SiX, cov, nrun, maxiterPCG, tolPCG, cvcut)
        # the trace probes are drawn first; getAIScore's own two solves draw nothing
        return dict(YPAPY=o[0], Trace=o[3], AI=o[6], PY=PY, nrun_used=int(round(o[7])))

    def getAIScore_q(self, Y, X, w, tau, SiY, SiX, cov, nrun, maxiterPCG, tolPCG, cvcut):
        o, PY = self._ai(True, Y, X, w, tau, SiY, SiX, cov, nrun, maxiterPCG, tolPCG, cvcut)
        return dict(YPAPY=o[0], YPA0PY=o[1], Trace=np.array([o[2], o[3]]), AI=np.array([[o[4], o[5]], [o[5], o[6]]]), PY=PY,
                    nrun_used=int(round(o[7])))

    def fitglmmaiRPCG(self, Y, X, w, tau, SiY, SiX, cov, nrun, maxiterPCG, tolPCG, tol, cvcut, quant=False):
        Y, X, w, SiY = _f(Y), np.asfortranarray(X, dtype=np.float64), _f(w), _f(SiY)
        SiX, cov = np.asfortranarray(SiX, dtype=np.float64), np.asfortranarray(cov, dtype=np.float64)
        t = _f(tau).copy()
        self._ck(self.L.fgref_fit_glmmai_rpcg(int(quant), _p(Y), _p(X), X.shape[1], _p(w), _p(t), _p(SiY), _p(SiX), _p(cov), int(nrun),
                                              int(maxiterPCG), float(tolPCG), float(tol), float(cvcut)))
        return t

    def getSigma_X(self, w, tau, X, maxiterPCG, tolPCG):
        w, tau, X = _f(w), _f(tau), np.asfortranarray(X, dtype=np.float64)
        out = np.zeros_like(X, order="F")
        self._ck(self.L.fgref_get_sigma_x(_p(w), _p(tau), _p(X), X.shape[1], int(maxiterPCG), float(tolPCG), _p(out)))
        return out

    def getSigma_G(self, w, tau, G, maxiterPCG, tolPCG):
        w, tau, G, out = _f(w), _f(tau), _f(G), np.zeros(self.N)
        self._ck(self.L.fgref_get_sigma_g(_p(w), _p(tau), _p(G), int(maxiterPCG), float(tolPCG), _p(out)))
        return out


class RefCPU(RefSolver):
    """oracle/_ref/libfg_refcpu.so: the reference's own CPU path END TO END, as shipped (fp32) -- genoClass (FG.cpp:37-1183: PLINK
    reader, QC, imputation, re-pack, standardised genotypes, diagonals), the OpenMP marker loop parallelCrossProd (FG.cpp:1576-1851)
    and the solver layer on top of it.  Nothing here comes from the oracle: files in, tau out, all reference text."""

    def __init__(self, U=None):
        # the reference keeps its genotype store in a file-global object that setgeno fills once per process (FG.cpp:1188); every
        # RefCPU therefore loads its own private copy of the library
        import shutil
        import tempfile
        self._tmp = tempfile.NamedTemporaryFile(prefix="libfg_refcpu_", suffix=".so", delete=False)
        self._tmp.close()
        shutil.copyfile(os.path.join(_HERE, "_ref", "libfg_refcpu.so"), self._tmp.name)
        self.L = C.CDLL(self._tmp.name)
        os.unlink(self._tmp.name)                      # the mapping stays valid; nothing is left behind
        L = self.L
        RefSolver._signatures(self)
        L.fgref_setgeno.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, DP, C.c_int, DP, C.c_int, C.c_double, C.c_double, C.c_int,
                                    C.c_double, C.c_double, C.c_int, DP, C.c_int]
        for f in ("fgref_N", "fgref_M_qc", "fgref_M_raw", "fgref_M_vr"):
            getattr(L, f).restype = C.c_long
        L.fgref_one_snp_geno.argtypes = [C.c_int, C.c_int, DP]
        L.fgref_one_snp_stdgeno.argtypes = [C.c_int, DP]
        L.fgref_crossprod.argtypes = [DP, C.c_int, DP]
        L.fgref_set_start_end_index_vec.argtypes = [DP, DP, C.c_int]
        L.fgref_set_start_end_index.argtypes = [C.c_int, C.c_int, C.c_int]
        assert L.fgref_real_bytes() == 4
        self._draws = None
        self.N = 0
        self._U = U

    def setgeno(self, bedfile, bimfile, famfile, subSampleInGeno, indicator, minMAF=0.0, maxMissing=1.0, isVarRatio=False,
                minMACvr=20.0, maxMACvr=-1.0, isDiagofKinSetAsOne=False, vr_rand_idx=None):
        sub = np.ascontiguousarray(subSampleInGeno, dtype=np.int32)
        ind = np.ascontiguousarray(indicator, dtype=np.uint8)
        vr = np.ascontiguousarray([] if vr_rand_idx is None else vr_rand_idx, dtype=np.int64)
        self._ck(self.L.fgref_setgeno(bedfile.encode(), bimfile.encode(), famfile.encode(), _p(sub), len(sub), _p(ind), len(ind),
                                      float(minMAF), float(maxMissing), int(isVarRatio), float(minMACvr), float(maxMACvr),
                                      int(isDiagofKinSetAsOne), _p(vr), len(vr)))
        self.N, self.M, self.M0, self.Mvr = self.L.fgref_N(), self.L.fgref_M_qc(), self.L.fgref_M_raw(), self.L.fgref_M_vr()
        self.L.fgref_clear_log()
        if self._U is not None:
            self.set_probes(self._U)

    def _vec(self, fn, n, dtype=np.float64):
        out = np.zeros(max(n, 1), dtype=dtype)
        fn(_p(out))
        return out[:n]

    def getAlleleFreqVec(self):
        return self._vec(self.L.fgref_allele_freq, self.M)

    def getInvStdVec(self):
        return self._vec(self.L.fgref_inv_std, self.M)

    def getMACVec(self):
        return self._vec(self.L.fgref_mac, self.M, np.int64)

    def getQCdMarkerIndex(self):
        return self._vec(self.L.fgref_qc_mask, self.M0, np.uint8).astype(bool)

    def getIndexVec_forVarRatio(self):
        return self._vec(self.L.fgref_vr_index, self.Mvr, np.int64)

    def getMACVec_forVarRatio(self):
        return self._vec(self.L.fgref_vr_mac, self.Mvr, np.int64)

    def Get_OneSNP_Geno(self, idx, vr=False):
        out = np.zeros(self.N, dtype=np.int64)
        self._ck(self.L.fgref_one_snp_geno(int(idx), int(vr), _p(out)))
        return out

    def Get_OneSNP_StdGeno(self, idx):
        out = np.zeros(self.N)
        self._ck(self.L.fgref_one_snp_stdgeno(int(idx), _p(out)))
        return out

    def Get_Diagof_StdGeno(self):
        out = np.zeros(self.N)
        self._ck(self.L.fgref_diag_stdgeno(_p(out)))
        return out

    def getCrossprodMatAndKin(self, b, loco=False):
        b, out = _f(b), np.zeros(self.N)
        self._ck(self.L.fgref_crossprod(_p(b), int(loco), _p(out)))
        return out

    def getCrossprodMatAndKin_LOCO(self, b):
        return self.getCrossprodMatAndKin(b, loco=True)

    def setStartEndIndexVec(self, start, end):
        s, e = np.ascontiguousarray(start, dtype=np.int64), np.ascontiguousarray(end, dtype=np.int64)
        self._ck(self.L.fgref_set_start_end_index_vec(_p(s), _p(e), len(s)))

    def setStartEndIndex(self, start, end, chromIndex):
        self._ck(self.L.fgref_set_start_end_index(int(start), int(end), int(chromIndex)))

    def set_Diagof_StdGeno_LOCO(self):
        self._ck(self.L.fgref_set_diag_loco())

    def set_loco_chromosome(self, c):
        raise NotImplementedError("use setStartEndIndex: this build holds the reference's own genotype object")


def fit_through_reference(o, r, fit0, U, trait, **kw):
    """The R-level loop of oracle.glmmkin_ai_PCG (FG.R:127-304 / 340-549) with every C++ export of that loop answered by the
    reference's compiled code (`r`, a RefSolver over the OracleGeno `o`) instead of by the oracle's restatement."""
    from . import oracle as O
    swap = dict(
        getCoefficients=lambda g, Y, X, w, tau, mi, tp, loco=False: r.getCoefficients(Y, X, w, tau, mi, tp, loco),
        getAIScore=lambda g, Y, X, w, tau, SiY, SiX, cov, nrun, mi, tp, cv, draw: r.getAIScore(Y, X, w, tau, SiY, SiX, cov, nrun, mi, tp, cv),
        getAIScore_q=lambda g, Y, X, w, tau, SiY, SiX, cov, nrun, mi, tp, cv, draw: r.getAIScore_q(Y, X, w, tau, SiY, SiX, cov, nrun, mi, tp, cv),
        fitglmmaiRPCG=lambda g, Y, X, w, tau, SiY, SiX, cov, nrun, mi, tp, tol, cv, draw: r.fitglmmaiRPCG(Y, X, w, tau, SiY, SiX, cov, nrun, mi, tp, tol, cv),
        fitglmmaiRPCG_q=lambda g, Y, X, w, tau, SiY, SiX, cov, nrun, mi, tp, tol, cv, draw: r.fitglmmaiRPCG(Y, X, w, tau, SiY, SiX, cov, nrun, mi, tp, tol, cv, quant=True))
    saved = {k: getattr(O, k) for k in swap}
    try:
        for k, v in swap.items():
            setattr(O, k, v)
        return O.glmmkin_ai_PCG(o, fit0, (0, 0), U, trait=trait, **kw)
    finally:
        for k, v in saved.items():
            setattr(O, k, v)
