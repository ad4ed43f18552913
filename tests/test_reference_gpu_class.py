"""Runs the reference's OWN GPU matvec (src/SAIGE/src/gpuSymMatMult.cu, compiled unmodified into oracle/_ref by
oracle/Makefile) on the B200 and checks the oracle -- and through it the CUDA library -- against it.

gpuSymMatMult::sym_sgemv computes A (A^T x) on a dense fp32 standardised matrix with two cublasSgemv calls
(gpuSymMatMult.cu:246-287); gpuParallelCrossProd divides by the marker count (FG.cpp:1700-1706).  fp32 arithmetic
=> tolerance 2e-5 relative against the fp64 oracle (the oracle's own ref32 mode sits at ~3e-7)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgpusymmatmult_ref.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_reference_gpu_matvec_agrees_with_oracle_and_library(grm10k):
    from oracle import oracle as O
    from saige_gpu_b200 import SaigeB200
    ref = C.CDLL(REF_SO)
    ref.ref_set_matrix.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p]
    ref.ref_sym_sgemv.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p]
    ref.ref_sym_sgemv_range.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
    N0, M0 = grm10k["N0"], grm10k["M0"]
    o = O.OracleGeno(mode=O.REF32)
    o.minMAF, o.maxMissing = 0.01, 0.15
    o.setgeno(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    # the dense fp32 slab gpuDistributeSNPs uploads (FG.cpp:1915-1923): column j = Get_OneSNP_StdGeno(j) in float
    A = np.empty((N0, o.M), dtype=np.float32, order="F")
    for m in range(o.M):
        A[:, m] = o.Get_OneSNP_StdGeno(m).astype(np.float32)
    assert ref.ref_set_matrix(N0, o.M, A.ctypes.data) == 0
    x = (np.random.default_rng(0).integers(0, 2, N0) * 2.0 - 1.0).astype(np.float32)
    z = np.zeros(N0, dtype=np.float32)
    assert ref.ref_sym_sgemv(N0, x.ctypes.data, z.ctypes.data) == 0
    y_ref = z.astype(np.float64) / o.M
    o64 = O.OracleGeno()
    o64.minMAF, o64.maxMissing = 0.01, 0.15
    o64.setgeno(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    y_orc = o64.getCrossprodMatAndKin(x.astype(np.float64))
    scale = np.max(np.abs(y_orc))
    assert np.max(np.abs(y_ref - y_orc)) / scale < 2e-5
    # LOCO-style range product of the reference class (sym_sgemv_range, gpuSymMatMult.cu:178-241)
    zr = np.zeros(N0, dtype=np.float32)
    assert ref.ref_sym_sgemv_range(2000, 4000, N0, x.ctypes.data, zr.ctypes.data) == 0
    want = o64._crossprod_range(2000, 4000, x.astype(np.float64))
    assert np.max(np.abs(zr - want)) / np.max(np.abs(want)) < 2e-5
    g = SaigeB200(device=0)
    g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15)
    g.setgeno_mem(grm10k["bed"], N0, M0, np.arange(1, N0 + 1), np.ones(N0, np.uint8))
    y_lib = g.getCrossprodMatAndKin(x.astype(np.float64))
    assert np.max(np.abs(y_lib - y_ref)) / scale < 2e-5
    g.close()
    ref.ref_free()
