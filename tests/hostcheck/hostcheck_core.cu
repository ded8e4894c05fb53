// TEST INFRASTRUCTURE: CPU execution of the process-independent device functions
// (HELAS wavefunctions, ALOHA vertices, RAMBO, Philox, VEGAS map, PDF interpolation) for checks without a GPU.
#include <cstring>

#include "aloha_sm.cuh"
#include "helas.cuh"
#include "pdf.cuh"
#include "phasespace.cuh"
#include "philox.cuh"
#include "vegas.cuh"

using namespace mf;

template <int NEXT>
static void rf(const double* x, long long nevt, double sqrts, const double* masses, int massive, double shat_min,
               PSConst k, int lab, double* p, double* w, double* x1, double* x2) {
  constexpr int ND = 4 * (NEXT - 2) + 2;
  for (long long e = 0; e < nevt; ++e) {
    double m[NEXT][4];
    ramboflow<NEXT>(x + e * ND, sqrts, masses, massive != 0, shat_min, k, m, w[e], x1[e], x2[e]);
    if (lab) boost_to_lab<NEXT>(m, x1[e], x2[e]);
    for (int i = 0; i < NEXT; ++i)
      for (int c = 0; c < 4; ++c) p[(e * NEXT + i) * 4 + c] = m[i][c];
  }
}

extern "C" {

// kind: 0 ixxxxx, 1 oxxxxx, 2 vxxxxx.  p (nevt,4), out (6,nevt) complex interleaved
int hc_wavefunction(int kind, const double* p, long long nevt, double mass, int nhel, int nsf, double sqh,
                    double* out) {
  for (long long e = 0; e < nevt; ++e) {
    cxd w[6];
    if (kind == 0) ixxxxx(p + 4 * e, mass, nhel, nsf, w);
    if (kind == 1) oxxxxx(p + 4 * e, mass, nhel, nsf, w);
    if (kind == 2) vxxxxx(p + 4 * e, mass, nhel, nsf, sqh, w);
    for (int k = 0; k < 6; ++k) out[2 * (k * nevt + e)] = w[k].re, out[2 * (k * nevt + e) + 1] = w[k].im;
  }
  return 0;
}

static void get(const double* a, long long nevt, long long e, cxd w[6]) {
  for (int k = 0; k < 6; ++k) w[k] = mk(a[2 * (k * nevt + e)], a[2 * (k * nevt + e) + 1]);
}
static void put(double* a, long long nevt, long long e, const cxd w[6]) {
  for (int k = 0; k < 6; ++k) a[2 * (k * nevt + e)] = w[k].re, a[2 * (k * nevt + e) + 1] = w[k].im;
}

// routine ids follow madflow_b200/aloha_ids
int hc_aloha(int id, const double* A, const double* B, const double* C, const double* D, long long nevt,
             double cre, double cim, double M, double W, double* out) {
  const cxd coup = mk(cre, cim);
  for (long long e = 0; e < nevt; ++e) {
    cxd a[6], b[6], c[6], d[6], r[6];
    if (A) get(A, nevt, e, a);
    if (B) get(B, nevt, e, b);
    if (C) get(C, nevt, e, c);
    if (D) get(D, nevt, e, d);
    cxd amp = mk(0, 0);
    bool is_amp = false;
    switch (id) {
      case 0: amp = FFV1_0(a, b, c, coup); is_amp = true; break;
      case 1: FFV1_1(a, b, coup, M, W, r); break;
      case 2: FFV1_2(a, b, coup, M, W, r); break;
      case 3: VVV1P0_1(a, b, coup, M, W, r); break;
      case 4: amp = VVV1_0(a, b, c, coup); is_amp = true; break;
      case 5: FFV1P0_3(a, b, coup, M, W, r); break;
      case 6: amp = VVVV_0<1>(a, b, c, d, coup); is_amp = true; break;
      case 7: amp = VVVV_0<3>(a, b, c, d, coup); is_amp = true; break;
      case 8: amp = VVVV_0<4>(a, b, c, d, coup); is_amp = true; break;
      case 9: VVVVP0_1<1>(a, b, c, coup, M, W, r); break;
      case 10: VVVVP0_1<3>(a, b, c, coup, M, W, r); break;
      case 11: VVVVP0_1<4>(a, b, c, coup, M, W, r); break;
      default: return -1;
    }
    if (is_amp) out[2 * e] = amp.re, out[2 * e + 1] = amp.im;
    else put(out, nevt, e, r);
  }
  return 0;
}

int hc_ramboflow(int next, const double* x, long long nevt, double sqrts, const double* masses, double pi,
                 double acc, double gev2pb, int lab, double* p, double* w, double* x1, double* x2) {
  const int nout = next - 2;
  double msum = 0;
  for (int i = 0; i < nout; ++i) msum += masses[i];
  PSConst k;
  k.pi = pi, k.acc = acc, k.gev2pb = gev2pb;
  k.wt0 = std::log(pi / 2.0) * (nout - 1) - 2.0 * std::lgamma((double)(nout - 1)) - std::log((double)(nout - 1));
  k.inv_norm = 1.0 / std::pow(2 * pi, 3 * nout - 4);
  const int massive = msum != 0.0;
  switch (next) {
    case 4: rf<4>(x, nevt, sqrts, masses, massive, msum * msum, k, lab, p, w, x1, x2); break;
    case 5: rf<5>(x, nevt, sqrts, masses, massive, msum * msum, k, lab, p, w, x1, x2); break;
    case 6: rf<6>(x, nevt, sqrts, masses, massive, msum * msum, k, lab, p, w, x1, x2); break;
    case 7: rf<7>(x, nevt, sqrts, masses, massive, msum * msum, k, lab, p, w, x1, x2); break;
    default: return -1;
  }
  return 0;
}

int hc_philox(unsigned long long seed, unsigned iteration, unsigned long long first, long long nevt, int ndim,
              double* out) {
  for (long long e = 0; e < nevt; ++e)
    for (int j = 0; j < (ndim + 1) / 2; ++j) {
      double a, b;
      philox_pair(seed, iteration, first + e, j, a, b);
      out[e * ndim + 2 * j] = a;
      if (2 * j + 1 < ndim) out[e * ndim + 2 * j + 1] = b;
    }
  return 0;
}

int hc_vegas_map(const double* grid, const double* r, long long nevt, int ndim, double* x, int* bins, double* w) {
  for (long long e = 0; e < nevt; ++e) {
    double ww = 1.0;
    for (int d = 0; d < ndim; ++d) x[e * ndim + d] = vegas_map(grid + d * VEGAS_EDGES, r[e * ndim + d], bins[e * ndim + d], ww);
    w[e] = ww;
  }
  return 0;
}
// csrc/pdf.cuh on a packed table T: out (nevt, ncol) = x f(x, Q2) of the table columns `col`
int hc_pdf_xfx(const double* T, const int* col, int ncol, const double* x, const double* q2, long long nevt, double* out) {
  for (long long e = 0; e < nevt; ++e) {
    const PdfPoint c = pdf_locate(T, x[e], q2[e]);
    for (int k = 0; k < ncol; ++k) out[e * ncol + k] = pdf_eval(c, col[k]);
  }
  return 0;
}

int hc_pdf_alphas(const double* T, const double* q2, long long nevt, double* out) {
  for (long long e = 0; e < nevt; ++e) out[e] = pdf_alphas(T, q2[e]);
  return 0;
}

// event_scale<4> as the integrand kernels call it: alpha_s and luminosity of events given lab momenta (nevt,4,4)
int hc_event_scale(const double* T, int alpha_mode, double alpha_s, double mz2, double b0, double fixed_q2, int nch,
                   const signed char* fl1, const signed char* fl2, const double* p, const double* x1, const double* x2,
                   long long nevt, double* as, double* lumi) {
  mfp_integrand_args u;
  memset(&u, 0, sizeof(u));
  u.d_pdf = T, u.alpha_mode = alpha_mode, u.alpha_s = alpha_s, u.mz2 = mz2, u.b0 = b0, u.fixed_q2 = fixed_q2, u.nchannels = nch;
  for (int c = 0; c < nch; ++c) u.chan_fl1[c] = fl1[c], u.chan_fl2[c] = fl2[c];
  for (long long e = 0; e < nevt; ++e) {
    double m[4][4];
    for (int i = 0; i < 4; ++i)
      for (int c = 0; c < 4; ++c) m[i][c] = p[(e * 4 + i) * 4 + c];
    event_scale<4>(u, m, x1[e], x2[e], as[e], lumi[e]);
  }
  return 0;
}
}
