"""The C-ABI library loads and exports every symbol include/saige_b200.h declares; without a GPU the compute path
refuses to run (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "saige_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(sgb_[a-z0-9_]+)\s*\(", src))
    names -= {"sgb_probe_fn"}
    return sorted(names)


def test_header_declares_the_hot_path():
    names = declared_symbols()
    for must in ("sgb_setgeno", "sgb_get_crossprod_mat_and_kin", "sgb_get_crossprod_mat_and_kin_loco",
                 "sgb_get_diag_of_kin", "sgb_get_pcg1_of_sigma_and_vector", "sgb_get_coefficients", "sgb_get_ai_score",
                 "sgb_get_ai_score_q", "sgb_fit_glmmai_rpcg", "sgb_fit_glmmai_rpcg_q", "sgb_get_sigma_x", "sgb_get_sigma_g",
                 "sgb_set_diag_of_stdgeno_loco", "sgb_create_dist"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from saige_gpu_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build it first: python -c 'import __graft_entry__ as g; g.build()'"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), "libsaige_b200.so does not export %s" % name


def test_binding_covers_every_declared_symbol():
    from saige_gpu_b200 import _lib
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared_symbols())
    _lib.lib()


def test_library_has_no_torch_or_oracle_dependency():
    import subprocess
    from saige_gpu_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "libcudart" in out


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "saige_gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "saige_oracle" not in txt.replace(
                    "oracle/saige_oracle.c", ""), f


def test_no_cpu_fallback_without_gpu():
    from saige_gpu_b200 import SaigeB200, SaigeB200Error
    cudart = ctypes.CDLL("libcudart.so.12")
    n = ctypes.c_int(0)
    if cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(SaigeB200Error, match="no CPU fallback"):
        SaigeB200(device=0)


def test_host_helpers_without_device():
    """calCV / innerProduct are pure host functions of the ABI (FG.cpp:3104-3110)."""
    import numpy as np
    from saige_gpu_b200 import _lib
    L = _lib.lib()
    x = np.array([1.0, 2.0, 4.0, 8.0])
    cv = L.sgb_cal_cv(x.ctypes.data, 4)
    assert abs(cv - (np.std(x, ddof=1) / np.mean(x)) / 4) < 1e-15
    y = np.array([1.0, -1.0, 0.5, 2.0])
    assert L.sgb_inner_product(x.ctypes.data, y.ctypes.data, 4) == float(x @ y)


def test_raw_bed_generator_is_plink_shaped():
    """bench.py feeds the ingest and step-2 samples from saige_gpu_b200.synth.raw_bed (no oracle outside the CPU-baseline
    legs): right size, Hardy-Weinberg-like code frequencies, the requested missing rate, deterministic per seed."""
    import numpy as np
    from saige_gpu_b200 import synth
    n, m = 4001, 64
    bed = synth.raw_bed(n, m, seed=5, miss_rate=0.02)
    B0 = (n + 3) // 4
    assert bed.dtype == np.uint8 and bed.size == B0 * m
    assert np.array_equal(bed, synth.raw_bed(n, m, seed=5, miss_rate=0.02))
    rows = bed.reshape(m, B0)
    codes = np.stack([(rows >> (2 * j)) & 3 for j in range(4)], axis=2).reshape(m, -1)[:, :n]
    miss = (codes == 1).mean()
    assert 0.012 < miss < 0.028
    f = ((codes == 0) * 2 + (codes == 2)).sum(1) / (2.0 * (codes != 1).sum(1))      # A1 frequency per marker
    assert f.min() > 0.02 and f.max() < 0.56 and f.std() > 0.05


def test_r_probe_stream_matches_r():
    """ProbeStream(rng="R") reproduces R's set.seed + rbinom(N, 1, 0.5) stream.  Known-answer check on R's documented
    output  set.seed(1); runif(5) -> 0.2655087 0.3721239 0.5728534 0.9082078 0.2016819."""
    import numpy as np
    from saige_gpu_b200 import step1
    assert np.allclose(step1.r_unif_rand(1, 5), [0.2655087, 0.3721239, 0.5728534, 0.9082078, 0.2016819], atol=5e-8)
    u = step1.r_unif_rand(200, 12)
    ps = step1.ProbeStream(4, nmax=3, seed=200, rng="R")
    assert ps.U.shape == (4, 3) and set(np.unique(ps.U)) <= {-1.0, 1.0}
    assert np.array_equal(ps.U[:, 0], 2.0 * (u[:4] >= 0.5) - 1.0) and np.array_equal(ps.U[:, 2], 2.0 * (u[8:12] >= 0.5) - 1.0)
    d = ps.fresh()
    assert np.array_equal(d(2), ps.U[:, :2]) and np.array_equal(d(1), ps.U[:, 2:3])


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 (and as C++) with no project or CUDA include."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "saige_b200.h")
    subprocess.check_call(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", hdr])
    subprocess.check_call(["g++", "-fsyntax-only", "-x", "c++", "-Wall", "-Werror", hdr])


def test_rcpp_shim_type_checks_against_the_header():
    """R / Rcpp / Armadillo are not installed here, so the step-1 shim cannot be built for real; it is at least run through the
    compiler's front end with declaration-only stand-ins for their classes (tests/stubs/RcppArmadillo.h), which checks every
    sgb_* call of the shim -- name, arity, pointer and scalar types, callback signatures -- against include/saige_b200.h,
    in the single-process and in the pbdMPI configuration."""
    import subprocess
    shim = os.path.join(ROOT, "rcpp_shim", "SAIGE_fitGLMM_fast_b200.cpp")
    base = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-unused-function", "-Wno-unused-variable", "-Werror", "-DUSE_SAIGE_B200",
            "-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + os.path.join(ROOT, "include"), shim]
    subprocess.check_call(base)
    subprocess.check_call(base + ["-DUSE_pbdMPI"])
    # every export of the header that the step-1 R path needs is referenced by the shim
    src = open(shim).read()
    for sym in ("sgb_setgeno", "sgb_get_coefficients", "sgb_get_ai_score", "sgb_get_ai_score_q", "sgb_fit_glmmai_rpcg", "sgb_fit_glmmai_rpcg_q",
                "sgb_get_sigma_x", "sgb_get_sigma_g", "sgb_set_diag_of_stdgeno_loco", "sgb_get_coef", "sgb_glmmkin_ai_pcg",
                "sgb_variance_ratio_markers", "sgb_set_probe_stream_fixed"):
        assert sym in src, sym
