// PDF and alpha_s interpolation on an LHAPDF lhagrid1 table resident in HBM (L2-resident in practice:
// a full NNPDF member is ~0.5 MB).  Device restatement of what the reference gets from pdfflow
// (scripts/madflow_exec.py:412-413 `pdf.xfxQ2`, :431 `pdf.alphasQ2`; pdfflow itself is third party and
// absent: parity unpinned) = LHAPDF 6's LogBicubicInterpolator and AlphaS_Ipol: cubic Hermite splines in
// log x and log Q2, derivatives from finite differences (mean of the one-sided slopes; one-sided at the
// edges of a subgrid); subgrids split at the flavour thresholds.  Outside the grid x and Q2 are frozen at
// the edge.  oracle/pdf.py is the numpy statement of the same algorithm.
//
// Table (doubles), built by madflow_b200/pdf.py::pack_table:
//   [0] nsub  [1] nfl  [2] nas (alpha_s subgrids)  [3] Q2 of the first alpha_s knot  [4] its alpha_s
//   [5] log-log gradient below it  [6] Q2 of the last knot  [7] its alpha_s
//   PDF subgrid s at [8 + 8 s]:        nx, nq, off_x, off_logx, off_q2, off_logq2, off_xf, q2min
//   alpha_s subgrid a at [8 + 8 nsub + 4 a]:   n, off_q2, off_logq2, off_alphas
//   data: x[nx], log x[nx], Q2[nq], log Q2[nq], xf[nx][nq][nfl] per subgrid; Q2[n], log Q2[n], alpha_s[n]
// The knots are searched on x and Q2 themselves (not on their logarithms) so that a point on a knot lands
// in the same cell as in the oracle whatever the last bit of the device's log().
#pragma once
#include "../../include/madflow_b200_process.h"
#include "mf_complex.cuh"
#include "phasespace.cuh"

namespace mf {

constexpr int PDF_HEADER = 8;

MF_DEV double pdf_cubic(double t, double vl, double vdl, double vh, double vdh) {
  const double t2 = t * t, t3 = t2 * t;
  const double p0 = (2 * t3 - 3 * t2 + 1) * vl;
  const double m0 = (t3 - 2 * t2 + t) * vdl;
  const double p1 = (-2 * t3 + 3 * t2) * vh;
  const double m1 = (t3 - t2) * vdh;
  return p0 + m0 + p1 + m1;
}

// largest i in [0, n-2] with knots[i] <= v (v already clamped into the knot range)
MF_DEV int pdf_cell(const double* knots, int n, double v) {
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (knots[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// Everything about the point (x, Q2) that does not depend on the flavour
struct PdfPoint {
  const double* xf;
  int nx, nq, nfl, ix, iq;
  double tx, dlogx, dxl, dxr;   // dxl = log x[ix] - log x[ix-1], dxr = log x[ix+2] - log x[ix+1] (when they exist)
  double tq, dq0, dq1, dq2;
};

MF_DEV PdfPoint pdf_locate(const double* T, double x, double q2) {
  const int nsub = (int)T[0];
  int s = 0;
  while (s + 1 < nsub && q2 >= T[PDF_HEADER + 8 * (s + 1) + 7]) ++s;  // a threshold belongs to the upper subgrid
  const double* D = T + PDF_HEADER + 8 * s;
  PdfPoint c;
  c.nx = (int)D[0], c.nq = (int)D[1], c.nfl = (int)T[1];
  const double *xs = T + (long long)D[2], *lx = T + (long long)D[3], *qs = T + (long long)D[4], *lq = T + (long long)D[5];
  c.xf = T + (long long)D[6];
  x = x < xs[0] ? xs[0] : (x > xs[c.nx - 1] ? xs[c.nx - 1] : x);
  q2 = q2 < qs[0] ? qs[0] : (q2 > qs[c.nq - 1] ? qs[c.nq - 1] : q2);
  c.ix = pdf_cell(xs, c.nx, x);
  c.iq = pdf_cell(qs, c.nq, q2);
  c.dlogx = lx[c.ix + 1] - lx[c.ix];
  c.tx = (log(x) - lx[c.ix]) / c.dlogx;
  c.dxl = c.ix > 0 ? lx[c.ix] - lx[c.ix - 1] : 1.0;
  c.dxr = c.ix + 2 < c.nx ? lx[c.ix + 2] - lx[c.ix + 1] : 1.0;
  c.dq1 = lq[c.iq + 1] - lq[c.iq];
  c.tq = (log(q2) - lq[c.iq]) / c.dq1;
  c.dq0 = c.iq > 0 ? lq[c.iq] - lq[c.iq - 1] : 1.0;
  c.dq2 = c.iq + 2 < c.nq ? lq[c.iq + 2] - lq[c.iq + 1] : 1.0;
  return c;
}

// the x spline of flavour ifl on the Q2 knot j
MF_DEV double pdf_row(const PdfPoint& c, int j, int ifl) {
  const double* col = c.xf + (long long)j * c.nfl + ifl;
  const long long st = (long long)c.nq * c.nfl;
  const double f0 = col[c.ix * st], f1 = col[(c.ix + 1) * st];
  const double mid = (f1 - f0) / c.dlogx;
  double d0 = mid, d1 = mid;
  if (c.ix > 0) d0 = ((f0 - col[(c.ix - 1) * st]) / c.dxl + mid) / 2.0;
  if (c.ix + 2 < c.nx) d1 = (mid + (col[(c.ix + 2) * st] - f1) / c.dxr) / 2.0;
  return pdf_cubic(c.tx, f0, d0 * c.dlogx, f1, d1 * c.dlogx);
}

// x f(x, Q2) of the flavour with table index ifl (LogBicubicInterpolator::_interpolateXQ2)
MF_DEV double pdf_eval(const PdfPoint& c, int ifl) {
  const double vl = pdf_row(c, c.iq, ifl), vh = pdf_row(c, c.iq + 1, ifl);
  const double fwd = (vh - vl) / c.dq1;
  double vdl = fwd, vdh = fwd;
  if (c.iq > 0) vdl = (fwd + (vl - pdf_row(c, c.iq - 1, ifl)) / c.dq0) / 2.0;
  if (c.iq + 2 < c.nq) vdh = (fwd + (pdf_row(c, c.iq + 2, ifl) - vh) / c.dq2) / 2.0;
  return pdf_cubic(c.tq, vl, vdl * c.dq1, vh, vdh * c.dq1);
}

MF_DEV double pdf_xfxq2(const double* T, int ifl, double x, double q2) { return pdf_eval(pdf_locate(T, x, q2), ifl); }

// alpha_s(Q2) from the set's table (AlphaS_Ipol::calcAlphasQ2)
MF_DEV double pdf_alphas(const double* T, double q2) {
  if (q2 < T[3]) return T[4] * pow(q2 / T[3], T[5]);
  if (q2 > T[6]) return T[7];
  const int nsub = (int)T[0], nas = (int)T[2];
  const double* A = T + PDF_HEADER + 8 * nsub;
  int s = 0;
  while (s + 1 < nas && q2 >= T[(long long)A[4 * (s + 1) + 1]]) ++s;
  const int n = (int)A[4 * s];
  const double *qs = T + (long long)A[4 * s + 1], *lq = T + (long long)A[4 * s + 2], *as = T + (long long)A[4 * s + 3];
  const int i = pdf_cell(qs, n, q2);
  const double dl = lq[i + 1] - lq[i];
  const double mid = (as[i + 1] - as[i]) / dl;
  double d0 = mid, d1 = mid;
  if (i > 0) d0 = 0.5 * (mid + (as[i] - as[i - 1]) / (lq[i] - lq[i - 1]));
  if (i + 2 < n) d1 = 0.5 * ((as[i + 2] - as[i + 1]) / (lq[i + 2] - lq[i + 1]) + mid);
  return pdf_cubic((log(q2) - lq[i]) / dl, as[i], d0 * dl, as[i + 1], d1 * dl);
}

// Parton luminosity of one subprocess (madflow_exec.py:450-454):
//   sum_channels xf_a(x1, Q2) xf_b(x2, Q2) / x1 / x2,   channels = initial_states (+ mirrored)
MF_DEV double pdf_luminosity(const double* T, int nch, const signed char* fl1, const signed char* fl2, double x1, double x2,
                             double q2) {
  const PdfPoint c1 = pdf_locate(T, x1, q2), c2 = pdf_locate(T, x2, q2);
  double sum = 0.0;
  for (int c = 0; c < nch; ++c) sum += pdf_eval(c1, fl1[c]) * pdf_eval(c2, fl2[c]);
  return sum / x1 / x2;
}

// Scale, alpha_s and luminosity of one accepted event (madflow_exec.py:426-454): q2 = (sum_out mT / 2)^2 or the
// fixed scale; alpha_s frozen (mode 0), one-loop (mode 1) or from the set's table (mode 2); luminosity 1 without
// a table (--no_pdf, :437-438).  `m` = the momenta the matrix element sees.
template <int NEXT>
MF_DEV void event_scale(const mfp_integrand_args& u, const double m[NEXT][4], double x1, double x2, double& as, double& lumi) {
  double q2 = u.fixed_q2;
  if (!(q2 > 0.0) && (u.alpha_mode != 0 || u.d_pdf != nullptr)) {
    double smt = 0.0;
#pragma unroll
    for (int i = 2; i < NEXT; ++i) smt += cut_value(CUT_MT, m[i]);
    q2 = (smt / 2.0) * (smt / 2.0);
  }
  if (u.alpha_mode == 0) as = u.alpha_s;
  else if (u.alpha_mode == 1) as = u.alpha_s / (1.0 + u.alpha_s * u.b0 * log(q2 / u.mz2));
  else as = pdf_alphas(u.d_pdf, q2);
  lumi = u.d_pdf ? pdf_luminosity(u.d_pdf, u.nchannels, u.chan_fl1, u.chan_fl2, x1, x2, q2) : 1.0;
}

}  // namespace mf
