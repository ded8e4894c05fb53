// TEST INFRASTRUCTURE: runs a generated process's Proc::matrix() on the CPU.
// The device functions are __host__ __device__, so the very code the GPU kernel executes can be
// checked against the oracle in the build container (no GPU there).  Never linked into the product.
#include MF_PROC_SOURCE

extern "C" int hostcheck_smatrix(const double* p, long long nevt, const double* par, const double* coup,
                                 long long coup_stride, double sqh, int only_comb, double* out) {
  for (long long ev = 0; ev < nevt; ++ev) {
    double m[Proc::NEXT][4];
    for (int i = 0; i < Proc::NEXT; ++i)
      for (int k = 0; k < 4; ++k) m[i][k] = p[(ev * Proc::NEXT + i) * 4 + k];
    cxd c[Proc::NCOUP > 0 ? Proc::NCOUP : 1];
    for (int j = 0; j < Proc::NCOUP; ++j) {
      const long long o = coup_stride ? ((long long)j * nevt + ev) : j;
      c[j] = mk(coup[2 * o], coup[2 * o + 1]);
    }
    if (only_comb >= 0)
      out[ev] = Proc::matrix(m, only_comb, par, c, sqh);
    else
      out[ev] = mf::smatrix_event<Proc>(m, par, c, sqh);
  }
  return 0;
}
