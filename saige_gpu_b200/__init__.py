"""saige_gpu_b200: B200-native (sm_100a) back end for the SAIGE step-1 null-GLMM hot path.

Layout (only what the path needs):
  csrc/            hand-written CUDA kernels + the C ABI  -> libsaige_b200.so (declared in include/saige_b200.h)
  _lib.py          ctypes binding of the C ABI (raises if the library or the GPU is missing -- no fallback)
  api.py           mirror of the reference's Rcpp export surface (same names / argument meaning)
  step1.py         mirror of the R driver between fitNULLGLMM and the exports (so step 1 can run without R)
  fitnull.py       fitNULLGLMM itself: phenotype file + PLINK files -> <prefix>.rda + <prefix>.varianceRatio.txt
  step2.py         SPAGMMATtest (single-variant tests from the step-1 files);  rdata.py: R save() reader / writer
  synth.py         synthetic workload of SURVEY.md 8(d)
"""
from ._lib import SaigeB200Error, LIB_PATH, EXPORTED_SYMBOLS  # noqa: F401
from .api import SaigeB200  # noqa: F401
from .fitnull import fitNULLGLMM  # noqa: F401
from .step2 import SPAGMMATtest  # noqa: F401
