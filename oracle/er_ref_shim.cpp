// C shim around the reference's exact-test core (src/SAIGE/src/Binary_global.hpp:17, Binary_global.cpp:69-80 SKAT_Exact,
// :113-119 GetProb) so the oracle harness can call the UNMODIFIED reference code through ctypes.  oracle/Makefile compiles
// Binary_ComputeExact.cpp, Binary_HyperGeo.cpp, Binary_global.cpp (+ the files they link against) from where they lie with
// -D_STAND_ALONE_ (the sources' own switch that drops <R.h>).  Test infrastructure only.
#include <cstddef>
#include "Binary_global.hpp"

void GetProb(int k, int ngroup, int ncase, int *group, double *weight, double *prob);

extern "C" {
void ref_skat_exact(int *resarray, int nres, int *nres_k, double *Z0, double *Z1, int k, int m, int total, int *total_k,
                    double *prob_k, double *odds, double *p1, int *IsExact, double *pval, double *pval_same, double *minP,
                    int test_type, double epsilon)
{
    SKAT_Exact(resarray, nres, nres_k, Z0, Z1, k, m, total, total_k, prob_k, odds, p1, IsExact, pval, pval_same, minP, test_type, epsilon);
}
void ref_get_prob(int k, int ngroup, int ncase, int *group, double *weight, double *prob)
{
    GetProb(k, ngroup, ncase, group, weight, prob);
}
}
