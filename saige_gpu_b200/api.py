"""Host-side mirror of the reference's Rcpp export surface for the step-1 hot path.

Same names, argument order and meaning as the `[[Rcpp::export]]` functions of
/root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp that src/SAIGE/R/SAIGE_fitGLMM_fast.R calls (the list is in
SURVEY.md 8b), as methods of one object that owns the C-ABI context (the reference keeps a file-global `geno`).
Every call goes through libsaige_b200.so; nothing here computes on the CPU.

Differences from the reference, all forced by the boundary contract:
  * errors raise SaigeB200Error instead of printing and continuing (FG.cpp:762-765) or exit() (FG.cpp:1947);
  * random draws come from the caller: `vr_rand_idx` for setgeno (arma::randi, FG.cpp:866-868) and a `draw(n)`
    callable for the trace estimator (R's rbinom, FG.cpp:3134-3137);
  * vector arguments may also be N x k matrices where batching is possible (getCrossprodMatAndKin,
    getPCG1ofSigmaAndVector, getSigma_G).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import SaigeB200Error, PROBE_FN, CHROM_FN


def _f64(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class SaigeB200:
    def __init__(self, device=0, rank=0, world=1, nccl_id=None, engine="tensor"):
        self._L = _lib.lib()
        h = C.c_void_p()
        if world > 1:
            if nccl_id is None or len(nccl_id) != _lib.NCCL_ID_BYTES:
                raise SaigeB200Error("world > 1 needs the %d-byte NCCL id made by nccl_unique_id()" % _lib.NCCL_ID_BYTES)
            buf = (C.c_char * _lib.NCCL_ID_BYTES).from_buffer_copy(bytes(nccl_id))
            rc = self._L.sgb_create_dist(device, rank, world, C.cast(buf, C.c_void_p), C.byref(h))
        else:
            rc = self._L.sgb_create(device, C.byref(h))
        if rc:
            raise SaigeB200Error(self._L.sgb_last_error(None).decode())
        self._h = h
        self.rank, self.world = rank, world
        self._keep = []
        if engine != "tensor":
            self.set_engine(engine)

    @staticmethod
    def nccl_unique_id():
        L = _lib.lib()
        buf = (C.c_char * _lib.NCCL_ID_BYTES)()
        if L.sgb_nccl_unique_id(C.cast(buf, C.c_void_p)):
            raise SaigeB200Error(L.sgb_last_error(None).decode())
        return bytes(buf)

    def _ck(self, rc):
        if rc:
            raise SaigeB200Error(self._L.sgb_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.sgb_destroy(self._h)
            self._h = None

    closeGenoFile_plink = close          # FG.cpp:1191

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_engine(self, engine):
        self._ck(self._L.sgb_set_engine(self._h, {"tensor": 0, "f64": 1, "umma": 2, "imma": 3}[engine]))

    def setStep2Batched(self, on=True):
        """Step-2 score sums as one tensor-engine GEMM per chunk (default) or every variant through the per-variant kernel."""
        self._ck(self._L.sgb_step2_set_batched(self._h, 1 if on else 0))

    def setStep2ChunkBytes(self, nbytes):
        """Raw-row bytes per pipeline chunk of the step-2 marker loop (default 1 GB)."""
        self._ck(self._L.sgb_step2_set_chunk_bytes(self._h, int(nbytes)))

    def set_verbose(self, on=True):
        """Print the reference's PCG log lines (FG.cpp:2794-2798) from every solve."""
        self._ck(self._L.sgb_set_verbose(self._h, 1 if on else 0))

    def set_rhs_limbs(self, n):
        """Digits (5..7) per right-hand-side value of wide batches on the tcgen05 kernel; 7 = full 55-bit values."""
        self._ck(self._L.sgb_set_rhs_limbs(self._h, int(n)))

    def set_product_tolerance(self, rel_tol):
        self._ck(self._L.sgb_set_product_tolerance(self._h, float(rel_tol)))

    def sync(self):
        self._ck(self._L.sgb_device_sync(self._h))

    # ---- configuration exports ----
    def setminMAFforGRM(self, minMAFforGRM):
        self._ck(self._L.sgb_set_min_maf_for_grm(self._h, float(minMAFforGRM)))

    def setmaxMissingRateforGRM(self, maxMissingforGRM):
        self._ck(self._L.sgb_set_max_missing_rate_for_grm(self._h, float(maxMissingforGRM)))

    def setminMAC_VarianceRatio(self, t_minMACVarRatio, t_maxMACVarRatio, t_isVarianceRatioinGeno):
        self._ck(self._L.sgb_set_min_mac_variance_ratio(self._h, float(t_minMACVarRatio), float(t_maxMACVarRatio),
                                                        int(bool(t_isVarianceRatioinGeno))))

    # ---- genotype store ----
    def setgeno(self, bedfile, bimfile, famfile, subSampleInGeno, indicatorGenoSamplesWithPheno, memoryChunk=2.0,
                isDiagofKinSetAsOne=False, vr_rand_idx=None):
        sub = np.ascontiguousarray(subSampleInGeno, dtype=np.int32)
        ind = np.ascontiguousarray(indicatorGenoSamplesWithPheno, dtype=np.uint8)
        vr = np.ascontiguousarray([] if vr_rand_idx is None else vr_rand_idx, dtype=np.int32)
        self._ck(self._L.sgb_setgeno(self._h, bedfile.encode(), bimfile.encode(), famfile.encode(), _p(sub), len(sub),
                                     _p(ind), len(ind), int(bool(isDiagofKinSetAsOne)), _p(vr), len(vr)))
        self._dims()

    def setgeno_mem(self, bed_body, n_fam, n_bim, subSampleInGeno, indicatorGenoSamplesWithPheno,
                    isDiagofKinSetAsOne=False, vr_rand_idx=None):
        bed = np.ascontiguousarray(bed_body, dtype=np.uint8)
        if bed.size < ((n_fam + 3) // 4) * n_bim:
            raise SaigeB200Error("bed body shorter than n_bim * ceil(n_fam/4) bytes")
        sub = np.ascontiguousarray(subSampleInGeno, dtype=np.int32)
        ind = np.ascontiguousarray(indicatorGenoSamplesWithPheno, dtype=np.uint8)
        if len(ind) != n_fam:
            raise SaigeB200Error("indicator length != n_fam")
        vr = np.ascontiguousarray([] if vr_rand_idx is None else vr_rand_idx, dtype=np.int32)
        self._ck(self._L.sgb_setgeno_mem(self._h, _p(bed), n_fam, n_bim, _p(sub), len(sub), _p(ind),
                                         int(bool(isDiagofKinSetAsOne)), _p(vr), len(vr)))
        self._dims()

    def setgeno_synth(self, n_samples, n_markers, seed, t0, t1):
        t0 = np.ascontiguousarray(t0, dtype=np.uint32)
        t1 = np.ascontiguousarray(t1, dtype=np.uint32)
        assert len(t0) == n_markers and len(t1) == n_markers
        self._ck(self._L.sgb_setgeno_synth(self._h, n_samples, n_markers, seed, _p(t0), _p(t1)))
        self._dims()

    def synth_bed_rows(self, n_samples, m0, m1, seed, t0, t1, miss_rate=0.0, out=None):
        """Raw PLINK .bed rows of the synthetic markers [m0, m1) (t0 / t1: thresholds of ALL markers or of that slice) into
        `out` (a writable uint8 buffer, e.g. a slice of a shared memory map) or a new array.  Bench / test input."""
        t0 = np.ascontiguousarray(t0, dtype=np.uint32)
        t1 = np.ascontiguousarray(t1, dtype=np.uint32)
        if len(t0) != m1 - m0:
            t0, t1 = np.ascontiguousarray(t0[m0:m1]), np.ascontiguousarray(t1[m0:m1])
        nbytes = ((n_samples + 3) // 4) * (m1 - m0)
        if out is None:
            out = np.empty(nbytes, dtype=np.uint8)
        assert out.dtype == np.uint8 and out.size >= nbytes and out.flags["C_CONTIGUOUS"]
        self._ck(self._L.sgb_synth_bed_rows(self._h, n_samples, m0, m1, seed, _p(t0), _p(t1), float(miss_rate), _p(out)))
        return out

    def _dims(self):
        L, h = self._L, self._h
        self.N = L.sgb_get_nnomissing(h)
        self.M = L.sgb_get_num_qc_markers(h)
        self.M0 = L.sgb_get_total_marker(h)
        self.Mloc = L.sgb_get_num_local_markers(h)
        self.Mvr = L.sgb_get_num_vr_markers(h)

    def gettotalMarker(self):
        return self._L.sgb_get_total_marker(self._h)

    def getNnomissingOut(self):
        return self._L.sgb_get_nnomissing(self._h)

    def getMsub_MAFge_minMAFtoConstructGRM(self):
        return self._L.sgb_get_num_qc_markers(self._h)

    def _ivec(self, fn, n, dtype=np.int32):
        out = np.zeros(max(n, 1), dtype=dtype)
        self._ck(fn(self._h, _p(out)))
        return out[:n]

    def getAlleleFreqVec(self):
        return self._ivec(self._L.sgb_get_allele_freq_vec, self.M, np.float64)

    def getMACVec(self):
        return self._ivec(self._L.sgb_get_mac_vec, self.M)

    def getAlleleCountVec(self):
        return self._ivec(self._L.sgb_get_allele_count_vec, self.M)

    def getMACVec_forVarRatio(self):
        return self._ivec(self._L.sgb_get_mac_vec_for_var_ratio, self.Mvr)

    def getIndexVec_forVarRatio(self):
        return self._ivec(self._L.sgb_get_index_vec_for_var_ratio, self.Mvr)

    def getIsVarRatioGeno(self):
        return bool(self._L.sgb_get_is_var_ratio_geno(self._h))

    def getQCdMarkerIndex(self):
        return self._ivec(self._L.sgb_get_qcd_marker_index, self.M0, np.uint8).astype(bool)

    def Get_OneSNP_Geno(self, SNPIdx):
        out = np.zeros(self.N, dtype=np.int32)
        self._ck(self._L.sgb_get_one_snp_geno(self._h, int(SNPIdx), _p(out)))
        return out

    def Get_OneSNP_Geno_forVarRatio(self, SNPIdx):
        out = np.zeros(self.N, dtype=np.int32)
        self._ck(self._L.sgb_get_one_snp_geno_for_var_ratio(self._h, int(SNPIdx), _p(out)))
        return out

    def Get_OneSNP_StdGeno(self, SNPIdx):
        out = np.zeros(self.N, dtype=np.float64)
        self._ck(self._L.sgb_get_one_snp_stdgeno(self._h, int(SNPIdx), _p(out)))
        return out

    # ---- LOCO ----
    def setStartEndIndexVec(self, startIndex_vec, endIndex_vec):
        s = np.ascontiguousarray(startIndex_vec, dtype=np.int32)
        e = np.ascontiguousarray(endIndex_vec, dtype=np.int32)
        self._ck(self._L.sgb_set_start_end_index_vec(self._h, _p(s), _p(e), len(s)))
        self._loco_start, self._loco_end = [int(v) for v in s], [int(v) for v in e]

    def setStartEndIndex(self, startIndex, endIndex, chromIndex):
        self._ck(self._L.sgb_set_start_end_index(self._h, int(startIndex), int(endIndex), int(chromIndex)))

    def set_Diagof_StdGeno_LOCO(self):
        self._ck(self._L.sgb_set_diag_of_stdgeno_loco(self._h))

    # ---- GRM products ----
    def get_DiagofKin(self):
        out = np.zeros(self.N)
        self._ck(self._L.sgb_get_diag_of_kin(self._h, _p(out)))
        return out

    def _mat(self, b):
        b = np.asarray(b, dtype=np.float64)
        one = b.ndim == 1
        bm = _f64(b.reshape(self.N, -1))
        return bm, one

    def getCrossprodMatAndKin(self, bVec):
        b, one = self._mat(bVec)
        y = np.zeros_like(b, order="F")
        self._ck(self._L.sgb_get_crossprod_mat_and_kin(self._h, _p(b), b.shape[1], _p(y)))
        return y[:, 0].copy() if one else y

    def getCrossprodMatAndKin_LOCO(self, bVec):
        b, one = self._mat(bVec)
        y = np.zeros_like(b, order="F")
        self._ck(self._L.sgb_get_crossprod_mat_and_kin_loco(self._h, _p(b), b.shape[1], _p(y)))
        return y[:, 0].copy() if one else y

    def getDiagOfSigma(self, wVec, tauVec, loco=False):
        w, tau = _f64(wVec), _f64(tauVec)
        out = np.zeros(self.N)
        self._ck(self._L.sgb_get_diag_of_sigma(self._h, _p(w), _p(tau), int(loco), _p(out)))
        return out

    def getDiagOfSigma_LOCO(self, wVec, tauVec):
        return self.getDiagOfSigma(wVec, tauVec, True)

    def getCrossprod(self, bVec, wVec, tauVec, loco=False):
        b, one = self._mat(bVec)
        w, tau = _f64(wVec), _f64(tauVec)
        y = np.zeros_like(b, order="F")
        self._ck(self._L.sgb_get_crossprod(self._h, _p(b), b.shape[1], _p(w), _p(tau), int(loco), _p(y)))
        return y[:, 0].copy() if one else y

    def getCrossprod_LOCO(self, bVec, wVec, tauVec):
        return self.getCrossprod(bVec, wVec, tauVec, True)

    # ---- PCG ----
    def getPCG1ofSigmaAndVector(self, wVec, tauVec, bVec, maxiterPCG, tolPCG, loco=False, return_iter=False):
        b, one = self._mat(bVec)
        w, tau = _f64(wVec), _f64(tauVec)
        x = np.zeros_like(b, order="F")
        it = np.zeros(b.shape[1], dtype=np.int32)
        self._ck(self._L.sgb_get_pcg1_of_sigma_and_vector(self._h, _p(w), _p(tau), _p(b), b.shape[1], int(maxiterPCG),
                                                          float(tolPCG), int(loco), _p(x), _p(it)))
        xo = x[:, 0].copy() if one else x
        return (xo, (int(it[0]) if one else it)) if return_iter else xo

    def getPCG1ofSigmaAndVector_LOCO(self, wVec, tauVec, bVec, maxiterPCG, tolPCG):
        return self.getPCG1ofSigmaAndVector(wVec, tauVec, bVec, maxiterPCG, tolPCG, True)

    # ---- AI-REML ----
    def _probe_cb(self, draw, factory=None):
        """draw(n) -> the next n probe columns.  factory: makes a fresh draw(); used when the library announces a new trace
        estimate with count = 0 (sgb_glmmkin_ai_pcg), where the reference re-seeds its generator."""
        N = self.N
        state = [draw]

        def cb(user, n, count, out):
            try:
                if count == 0:
                    state[0] = factory()
                    return 0
                U = np.asarray(state[0](int(count)), dtype=np.float64).reshape(N, -1)
                if U.shape[1] != count:
                    return 1
                dst = np.ctypeslib.as_array(out, shape=(int(n) * int(count),))
                dst[:] = np.asfortranarray(U).ravel(order="F")
                return 0
            except Exception:      # never let an exception cross the C boundary
                return 1
        return PROBE_FN(cb)

    def getCoefficients(self, Yvec, Xmat, wVec, tauVec, maxiterPCG, tolPCG, loco=False):
        Y, X, w, tau = _f64(Yvec), _f64(np.asarray(Xmat).reshape(self.N, -1)), _f64(wVec), _f64(tauVec)
        p = X.shape[1]
        SiY, SiX = np.zeros(self.N), np.zeros((self.N, p), order="F")
        cov, alpha, eta = np.zeros((p, p), order="F"), np.zeros(p), np.zeros(self.N)
        self._ck(self._L.sgb_get_coefficients(self._h, _p(Y), _p(X), p, _p(w), _p(tau), int(maxiterPCG), float(tolPCG),
                                              int(loco), _p(SiY), _p(SiX), _p(cov), _p(alpha), _p(eta)))
        return dict(Sigma_iY=SiY, Sigma_iX=SiX, cov=cov, alpha=alpha, eta=eta)

    def getCoefficients_LOCO(self, Yvec, Xmat, wVec, tauVec, maxiterPCG, tolPCG):
        return self.getCoefficients(Yvec, Xmat, wVec, tauVec, maxiterPCG, tolPCG, True)

    # ---- driver loops of SAIGE_fitGLMM_fast.R as single calls (optional; see include/saige_b200.h) ----
    FAMILY = {"binomial": 0, "gaussian": 1}

    def Get_Coef(self, y, X, tau, family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter, loco=False):
        """Get_Coef / Get_Coef_LOCO (FG.R:2-35, 42-73) with the IRLS loop on the device; returns the R list."""
        X = _f64(np.asarray(X).reshape(self.N, -1))
        p = X.shape[1]
        y, off, tau, a0, e0 = _f64(y), _f64(offset), _f64(tau), _f64(alpha0), _f64(eta0)
        Y, eta, W, mu, SiY = (np.empty(self.N) for _ in range(5))
        SiX, cov, alpha = np.empty((self.N, p), order="F"), np.empty((p, p), order="F"), np.empty(p)
        nit = np.zeros(1, dtype=np.int32)
        self._ck(self._L.sgb_get_coef(self._h, self.FAMILY[getattr(family, "name", family)], _p(y), _p(X), p, _p(off), _p(tau),
                                      _p(a0), _p(e0), int(maxiter), int(maxiterPCG), float(tolPCG), int(loco), _p(Y), _p(alpha),
                                      _p(eta), _p(W), _p(cov), _p(SiY), _p(SiX), _p(mu), _p(nit)))
        return dict(Y=Y, alpha=alpha, eta=eta, W=W, cov=cov, sqrtW=np.sqrt(W), Sigma_iY=SiY, Sigma_iX=SiX, mu=mu, n_iter=int(nit[0]))

    @staticmethod
    def _chrom_cb(on_chrom, view):
        """ctypes callback that hands chromosome c's finished outputs (view(c) -> dict) to on_chrom(c, dict)."""
        if on_chrom is None:
            return C.cast(None, CHROM_FN)

        def cb(user, c):
            try:
                on_chrom(int(c), view(int(c)))
            except Exception:      # never let an exception cross the C boundary
                pass
        return CHROM_FN(cb)

    def Get_Coef_LOCO_all(self, y, X, tau, family, alpha0, eta0, offset, maxiterPCG, tolPCG, maxiter, on_chrom=None):
        """The leave-one-chromosome-out refit loop (FG.R:255-292) as one call; returns one dict per chromosome of
        setStartEndIndexVec (None where the chromosome has no range).  on_chrom(c, dict) is called as soon as chromosome c is
        done, while the next one runs on the GPU."""
        X = _f64(np.asarray(X).reshape(self.N, -1))
        p = X.shape[1]
        nchr = len(self._loco_start)
        y, off, tau, a0, e0 = _f64(y), _f64(offset), _f64(tau), _f64(alpha0), _f64(eta0)
        Y, eta, mu = (np.zeros((self.N, nchr), order="F") for _ in range(3))
        alpha, cov = np.zeros((p, nchr), order="F"), np.zeros((p * p, nchr), order="F")
        nit = np.zeros(nchr, dtype=np.int32)

        def view(c):
            return dict(Y=Y[:, c], alpha=alpha[:, c].copy(), eta=eta[:, c], mu=mu[:, c],
                        cov=cov[:, c].reshape(p, p, order="F").copy(), n_iter=int(nit[c]))
        cb = self._chrom_cb(on_chrom, view)
        self._ck(self._L.sgb_get_coef_loco_all(self._h, self.FAMILY[getattr(family, "name", family)], _p(y), _p(X), p, _p(off),
                                               _p(tau), _p(a0), _p(e0), int(maxiter), int(maxiterPCG), float(tolPCG), _p(Y),
                                               _p(alpha), _p(eta), _p(cov), _p(mu), _p(nit), cb, None))
        return [None if (s == -1 or e == -1) else view(c) for c, (s, e) in enumerate(zip(self._loco_start, self._loco_end))]

    def glmmkin_ai_PCG(self, trait, y, X, offset, alpha_fit0, eta_fit0, tauInit, maxiter, tol, nrun, tolPCG, maxiterPCG,
                       traceCVcutoff, LOCO, draw_factory, on_chrom=None):
        """glmmkin.ai_PCG_Rcpp_Binary / _Quantitative after setgeno (FG.R:127-304, 340-549) as one call (sgb_glmmkin_ai_pcg).
        draw_factory() -> draw(n): a fresh probe stream per trace estimate (GetTrace re-seeds, FG.cpp:3114).
        on_chrom(c, dict): called when the genome-wide fit (c = -1) / chromosome c's refit is complete, while the GPU goes on."""
        X = _f64(np.asarray(X).reshape(self.N, -1))
        p = X.shape[1]
        N = self.N
        y, off, a0, e0, ti = _f64(y), _f64(offset), _f64(alpha_fit0), _f64(eta_fit0), _f64(tauInit)
        tau, alpha, cov = np.zeros(2), np.zeros(p), np.zeros((p, p), order="F")
        eta, mu, Y = np.empty(N), np.empty(N), np.empty(N)
        conv, nout = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        nchr = len(self._loco_start) if LOCO else 0
        Yl, el, ml = (np.zeros((N, max(nchr, 1)), order="F") for _ in range(3))
        al, cl, nl = np.zeros((p, max(nchr, 1)), order="F"), np.zeros((p * p, max(nchr, 1)), order="F"), np.zeros(max(nchr, 1), dtype=np.int32)
        cb = self._probe_cb(None, draw_factory)

        def view(c):
            if c < 0:
                return dict(theta=tau, coefficients=alpha, linear_predictors=eta, fitted_values=mu, Y=Y, cov=cov,
                            converged=bool(conv[0]), n_outer=int(nout[0]))
            return dict(Y=Yl[:, c], alpha=al[:, c].copy(), eta=el[:, c], mu=ml[:, c], cov=cl[:, c].reshape(p, p, order="F").copy(),
                        n_iter=int(nl[c]))
        ccb = self._chrom_cb(on_chrom, view)
        self._ck(self._L.sgb_glmmkin_ai_pcg(self._h, int(trait == "quantitative"), _p(y), _p(X), p, _p(off), _p(a0), _p(e0), _p(ti),
                                            int(maxiter), float(tol), int(nrun), float(tolPCG), int(maxiterPCG), float(traceCVcutoff),
                                            int(bool(LOCO)), cb, None, _p(tau), _p(alpha), _p(eta), _p(mu), _p(Y), _p(cov), _p(conv),
                                            _p(nout), _p(Yl), _p(al), _p(el), _p(cl), _p(ml), _p(nl), ccb, None))
        out = dict(view(-1), loco=None)
        if LOCO:
            out["loco"] = [None if (s == -1 or e == -1) else view(c) for c, (s, e) in enumerate(zip(self._loco_start, self._loco_end))]
        return out

    def varianceRatioMarkers(self, marker_idx, from_vr_store, wVec, tauVec, Xmat, XV, XXVX_inv, Sigma_iX, mu2, maxiterPCG, tolPCG):
        """The marker loop of extractVarianceRatio (FG.R:2298-2378) for a batch of markers: (var1, var2null, AC)."""
        idx = np.ascontiguousarray(marker_idx, dtype=np.int64)
        X = _f64(np.asarray(Xmat).reshape(self.N, -1))
        p = X.shape[1]
        XV = _f64(np.asarray(XV).reshape(p, self.N))
        XX, SiX = _f64(np.asarray(XXVX_inv).reshape(self.N, p)), _f64(np.asarray(Sigma_iX).reshape(self.N, p))
        w, tau = _f64(wVec), _f64(tauVec)
        m2 = None if mu2 is None else _f64(mu2)
        v1, v2, ac = np.zeros(len(idx)), np.zeros(len(idx)), np.zeros(len(idx))
        self._ck(self._L.sgb_variance_ratio_markers(self._h, _p(idx), len(idx), int(bool(from_vr_store)), _p(w), _p(tau), _p(X), p,
                                                    _p(XV), _p(XX), _p(SiX), None if m2 is None else _p(m2), int(maxiterPCG),
                                                    float(tolPCG), _p(v1), _p(v2), _p(ac)))
        return v1, v2, ac

    def setProbeStreamFixed(self, on=True):
        """The first nrun probes are the same in every GetTrace call (set_seed(200), FG.cpp:3114): keep them on the device."""
        self._ck(self._L.sgb_set_probe_stream_fixed(self._h, 1 if on else 0))

    def _ai_args(self, Yvec, Xmat, wVec, tauVec, Sigma_iY, Sigma_iX, cov):
        X = _f64(np.asarray(Xmat).reshape(self.N, -1))
        p = X.shape[1]
        return (_f64(Yvec), X, p, _f64(wVec), _f64(tauVec).copy(), _f64(Sigma_iY),
                _f64(np.asarray(Sigma_iX).reshape(self.N, p)), _f64(np.asarray(cov).reshape(p, p)))

    def getAIScore(self, Yvec, Xmat, wVec, tauVec, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff,
                   draw):
        Y, X, p, w, tau, SiY, SiX, cv = self._ai_args(Yvec, Xmat, wVec, tauVec, Sigma_iY, Sigma_iX, cov)
        out4, PY = np.zeros(4), np.zeros(self.N)
        cb = self._probe_cb(draw)
        self._ck(self._L.sgb_get_ai_score(self._h, _p(Y), _p(X), p, _p(w), _p(tau), _p(SiY), _p(SiX), _p(cv), int(nrun),
                                          int(maxiterPCG), float(tolPCG), float(traceCVcutoff), cb, None, _p(out4), _p(PY)))
        return dict(YPAPY=out4[0], Trace=out4[1], PY=PY, AI=out4[2], nrun_used=int(out4[3]))

    def getAIScore_q(self, Yvec, Xmat, wVec, tauVec, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, traceCVcutoff,
                     draw):
        Y, X, p, w, tau, SiY, SiX, cv = self._ai_args(Yvec, Xmat, wVec, tauVec, Sigma_iY, Sigma_iX, cov)
        o, PY = np.zeros(8), np.zeros(self.N)
        cb = self._probe_cb(draw)
        self._ck(self._L.sgb_get_ai_score_q(self._h, _p(Y), _p(X), p, _p(w), _p(tau), _p(SiY), _p(SiX), _p(cv), int(nrun),
                                            int(maxiterPCG), float(tolPCG), float(traceCVcutoff), cb, None, _p(o), _p(PY)))
        return dict(YPAPY=o[0], YPA0PY=o[1], Trace=np.array([o[2], o[3]]), PY=PY,
                    AI=np.array([[o[4], o[5]], [o[5], o[6]]]), nrun_used=int(o[7]))

    def _fit(self, fn, Yvec, Xmat, wVec, tauVec, Sigma_iY, Sigma_iX, cov, nrun, maxiterPCG, tolPCG, tol, traceCVcutoff,
             draw):
        Y, X, p, w, tau, SiY, SiX, cv = self._ai_args(Yvec, Xmat, wVec, tauVec, Sigma_iY, Sigma_iX, cov)
        cb = self._probe_cb(draw)
        self._ck(fn(self._h, _p(Y), _p(X), p, _p(w), _p(tau), _p(SiY), _p(SiX), _p(cv), int(nrun), int(maxiterPCG),
                    float(tolPCG), float(tol), float(traceCVcutoff), cb, None))
        return dict(tau=tau)

    def fitglmmaiRPCG(self, *a, **kw):
        return self._fit(self._L.sgb_fit_glmmai_rpcg, *a, **kw)

    def fitglmmaiRPCG_q(self, *a, **kw):
        return self._fit(self._L.sgb_fit_glmmai_rpcg_q, *a, **kw)

    def getSigma_X(self, wVec, tauVec, Xmat, maxiterPCG, tolPCG, loco=False):
        X = _f64(np.asarray(Xmat).reshape(self.N, -1))
        w, tau = _f64(wVec), _f64(tauVec)
        out = np.zeros_like(X, order="F")
        self._ck(self._L.sgb_get_sigma_x(self._h, _p(w), _p(tau), _p(X), X.shape[1], int(maxiterPCG), float(tolPCG),
                                         int(loco), _p(out)))
        return out

    def getSigma_G(self, wVec, tauVec, Gvec, maxiterPCG, tolPCG, loco=False):
        G, one = self._mat(Gvec)
        w, tau = _f64(wVec), _f64(tauVec)
        out = np.zeros_like(G, order="F")
        self._ck(self._L.sgb_get_sigma_g(self._h, _p(w), _p(tau), _p(G), G.shape[1], int(maxiterPCG), float(tolPCG),
                                         int(loco), _p(out)))
        return out[:, 0].copy() if one else out

    def calCV(self, xVec):
        x = _f64(xVec)
        return self._L.sgb_cal_cv(_p(x), len(x))

    def innerProduct(self, x, y):
        x, y = _f64(x), _f64(y)
        return self._L.sgb_inner_product(_p(x), _p(y), len(x))

    # ---- step 2: setSAIGEobjInCPP + mainMarkerInCPP ----
    STEP2_COLUMNS = ("tested", "AC_Allele2", "AF_Allele2", "MissingRate", "BETA", "SE", "Tstat", "var", "p.value", "p.value.NA",
                     "Is.SPA", "AF_case", "AF_ctrl", "N_case", "N_ctrl", "N_case_hom", "N_case_het", "N_ctrl_hom", "N_ctrl_het",
                     "var2", "Is.Firth", "Firth.converged", "BETA_c", "SE_c", "Tstat_c", "var_c", "p.value_c", "p.value.NA_c",
                     "log.p.value", "log.p.value.NA", "log.p.value_c", "log.p.value.NA_c")

    def setSAIGEobjInCPP(self, model, varRatio, SPAcutoff, pos_in_fam):
        """model: dict with mu, res, mu2, y, X, XVX, XXVX_inv, XVX_inv_XV, S_a, tau, trait (readInGLMM.R:39-170)."""
        X = _f64(model["X"])
        N, p = X.shape
        arrs = [_f64(np.asarray(model[k], dtype=np.float64).reshape(-1) if k in ("mu", "res", "mu2", "y", "S_a", "tau") else model[k])
                for k in ("mu", "res", "mu2", "y", "X", "XVX", "XXVX_inv", "XVX_inv_XV", "S_a", "tau")]
        pos = np.ascontiguousarray(pos_in_fam, dtype=np.int32)
        if len(pos) != N:
            raise SaigeB200Error("pos_in_fam must have one entry per model sample")
        ratios = np.asarray(varRatio, dtype=np.float64).reshape(-1)
        self._ck(self._L.sgb_step2_set_model(self._h, N, p, int(model["trait"] == "binary"), *[_p(a) for a in arrs],
                                             float(ratios[0]), float(SPAcutoff), _p(pos)))
        self._step2_N = N
        if len(ratios) > 1:
            self.setVarianceRatios(ratios, model.get("cateVarRatioMinMACVecExclude", (10, 20.5)),
                                   model.get("cateVarRatioMaxMACVecInclude", (20.5,)))

    def setFirth(self, is_Firth_beta, pCutoffforFirth=0.01, offset=None, se_from_fit=True):
        """is_Firth_beta / pCutoffforFirth of SPAGMMATtest: Firth's bias-reduced BETA for binary-trait variants with
        p <= cutoff; offset = the null model's offset vector (None = zeros)."""
        off = None if offset is None else _f64(np.asarray(offset, dtype=np.float64).reshape(-1))
        self._ck(self._L.sgb_step2_set_firth(self._h, int(bool(is_Firth_beta)), float(pCutoffforFirth),
                                             None if off is None else _p(off), int(bool(se_from_fit))))

    def setVarianceRatios(self, ratioVec_null, cateVarRatioMinMACVecExclude=(10, 20.5), cateVarRatioMaxMACVecInclude=(20.5,)):
        """t_varRatio_null / t_cateVarRatio*Vec of setSAIGEobjInCPP: one ratio, or one per MAC category (assignVarianceRatio)."""
        r = _f64(np.asarray(ratioVec_null, dtype=np.float64).reshape(-1))
        lo = _f64(np.asarray(cateVarRatioMinMACVecExclude, dtype=np.float64).reshape(-1))
        hi = _f64(np.asarray(cateVarRatioMaxMACVecInclude, dtype=np.float64).reshape(-1))
        if len(r) > 1 and (len(lo) != len(r) or len(hi) != len(r) - 1):
            raise SaigeB200Error("ERROR! The number of variance ratios are different from the length of cateVarRatioMinMACVecExclude")
        self._ck(self._L.sgb_step2_set_variance_ratios(self._h, len(r), _p(r), _p(lo), _p(hi)))

    def setCondition(self, P2=None, XtP2=None, VarInv=None, Tstat_cond=None):
        """assignConditionFactors: P2 (N x q), XtP2 = XXVX_inv^T P2 (p x q), VarInv (q x q), Tstat_cond (q); no argument = off."""
        if P2 is None:
            self._ck(self._L.sgb_step2_set_condition(self._h, 0, None, None, None, None))
            return
        P2, XtP2, VarInv = _f64(np.asarray(P2, dtype=np.float64)), _f64(np.asarray(XtP2, dtype=np.float64)), _f64(np.asarray(VarInv, dtype=np.float64))
        T = _f64(np.asarray(Tstat_cond, dtype=np.float64).reshape(-1))
        q = len(T)
        if P2.shape != (self._step2_N, q) or VarInv.shape != (q, q) or XtP2.shape[1] != q:
            raise SaigeB200Error("condition factors have inconsistent shapes")
        self._ck(self._L.sgb_step2_set_condition(self._h, q, _p(P2), _p(XtP2), _p(VarInv), _p(T)))

    def setMaxMACforER(self, max_MAC_for_ER=4.0):
        """max_MAC_for_ER of SPAGMMATtest / setAssocTest_GlobalVarsInCPP: binary-trait variants with MAC <= this get the
        exact-test p-value (efficient resampling) when their score exceeds the SPA cutoff; negative = off."""
        self._ck(self._L.sgb_step2_set_er(self._h, float(max_MAC_for_ER)))

    def mainMarkerInCPP(self, bed_rows, n_fam, n_markers, min_MAF=0.0, min_MAC=0.5, max_missing=0.15, se_two_sided=True):
        bed = np.ascontiguousarray(bed_rows, dtype=np.uint8)
        if bed.size < ((n_fam + 3) // 4) * n_markers:
            raise SaigeB200Error("bed_rows shorter than n_markers * ceil(n_fam/4) bytes")
        out = np.zeros((n_markers, len(self.STEP2_COLUMNS)))
        self._ck(self._L.sgb_step2_test_markers(self._h, _p(bed), int(n_fam), int(n_markers), float(min_MAF), float(min_MAC),
                                                float(max_missing), int(bool(se_two_sided)), _p(out)))
        return out

    def mainMarkerInCPP_dosage(self, dosages, min_MAF=0.0, min_MAC=0.5, max_missing=0.15, se_two_sided=True, impute_method=1,
                               dosage_zerod_cutoff=0.2, dosage_zerod_MAC_cutoff=10.0):
        """The marker loop for dosage rows (VCF / BGEN input): dosages[n_markers x n_file_samples], negative or NaN = missing."""
        D = np.ascontiguousarray(dosages, dtype=np.float64)
        if D.ndim != 2:
            raise SaigeB200Error("dosages must be n_markers x n_file_samples")
        out = np.zeros((D.shape[0], len(self.STEP2_COLUMNS)))
        self._ck(self._L.sgb_step2_test_dosages(self._h, _p(D), D.shape[1], D.shape[0], float(min_MAF), float(min_MAC),
                                                float(max_missing), int(bool(se_two_sided)), int(impute_method),
                                                float(dosage_zerod_cutoff), float(dosage_zerod_MAC_cutoff), _p(out)))
        return out

    # ---- dense N x N GRM (BASELINE config 4): tcgen05 build, stored-GRM products ----
    def buildDenseGRM(self, weight_limbs=7):
        """K = Z Z^T / M on the tensor cores from the loaded 2-bit store (collective when distributed)."""
        self._ck(self._L.sgb_dense_grm_build(self._h, int(weight_limbs)))
        return self.denseGRMInfo()

    def freeDenseGRM(self):
        self._ck(self._L.sgb_dense_grm_free(self._h))

    def denseGRMInfo(self):
        o = np.zeros(6)
        self._ck(self._L.sgb_dense_grm_info(self._h, _p(o)))
        return {"weight_limbs": int(o[0]), "fixed_point_exponent": int(o[1]), "block_rows": int(o[2]),
                "stored_bytes": float(o[3]), "build_ms": float(o[4]), "int8_ops": float(o[5])}

    def getDenseGRMBlock(self, i0, ni, j0, nj):
        out = np.zeros((ni, nj), order="F")
        self._ck(self._L.sgb_dense_grm_get_block(self._h, int(i0), int(ni), int(j0), int(nj), _p(out)))
        return out

    def setGRMMode(self, mode):
        """'packed' (default: products from the 2-bit genotypes) or 'dense' (products / PCG from the stored matrix)."""
        self._ck(self._L.sgb_set_grm_mode(self._h, {"packed": 0, "dense": 1}[mode]))

    def writeDenseGRM(self, prefix, sample_ids=None, rows_per_read=2048):
        """GCTA-style <prefix>.grm.bin / .grm.N.bin / .grm.id (fp32 lower triangle incl. diagonal, row by row), the
        format of the reference's extdata/output/nfam_*_GRM.grm.bin."""
        N = self.N
        M = self.M
        with open(prefix + ".grm.bin", "wb") as fb, open(prefix + ".grm.N.bin", "wb") as fn:
            for i0 in range(0, N, rows_per_read):
                ni = min(rows_per_read, N - i0)
                blk = self.getDenseGRMBlock(i0, ni, 0, i0 + ni)
                for a in range(ni):
                    row = blk[a, :i0 + a + 1].astype(np.float32)
                    fb.write(row.tobytes())
                    fn.write(np.full(row.size, M, dtype=np.float32).tobytes())
        with open(prefix + ".grm.id", "w") as f:
            for i in range(N):
                sid = sample_ids[i] if sample_ids is not None else str(i + 1)
                f.write("%s\t%s\n" % (sid, sid))

    def bench_dense_build(self, weight_limbs, first_block_row, n_block_rows):
        self._ck(self._L.sgb_bench_dense_build(self._h, int(weight_limbs), int(first_block_row), int(n_block_rows)))
        return self.denseGRMInfo()

    # ---- bench hooks / counters ----
    def bench_crossprod_device(self, k, reps, seed=1):
        ms = np.zeros(reps, dtype=np.float32)
        mk = np.zeros(2 * reps, dtype=np.float32)
        self._ck(self._L.sgb_bench_crossprod_device(self._h, int(k), int(reps), int(seed), _p(ms), _p(mk)))
        return ms, mk.reshape(reps, 2)

    def bench_fetch_result(self, k):
        Y = np.zeros((self.N, k), order="F")
        B = np.zeros((self.N, k), order="F")
        self._ck(self._L.sgb_bench_fetch_result(self._h, int(k), _p(Y), _p(B)))
        return Y, B

    def counters(self):
        c = _lib.Counters()
        self._ck(self._L.sgb_get_counters(self._h, C.byref(c)))
        return {n: getattr(c, n) for n, _ in c._fields_}

    def reset_counters(self):
        self._ck(self._L.sgb_reset_counters(self._h))
