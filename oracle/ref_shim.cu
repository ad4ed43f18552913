// C shim around the reference's gpuSymMatMult class (src/SAIGE/src/gpuSymMatMult.hpp:13-36) so the
// oracle harness can call the UNMODIFIED reference GPU matvec through ctypes.  Test infrastructure only.
#include <cstddef>
#include "gpuSymMatMult.hpp"

static gpuSymMatMult *g_ref = nullptr;

extern "C" {
int ref_set_matrix(size_t n_rows, size_t n_cols, const float *A)
{
    delete g_ref;
    g_ref = new gpuSymMatMult();
    return g_ref->set_matrix(0, 0, n_rows, n_cols, A);
}
int ref_sym_sgemv(size_t n_elem, const float *x, float *ret) { return g_ref ? g_ref->sym_sgemv(0, n_elem, x, ret) : -1; }
int ref_sym_sgemv_range(size_t c0, size_t c1, size_t n_elem, const float *x, float *ret)
{
    return g_ref ? g_ref->sym_sgemv_range(0, c0, c1, n_elem, x, ret) : -1;
}
void ref_free(void) { delete g_ref; g_ref = nullptr; }
}
