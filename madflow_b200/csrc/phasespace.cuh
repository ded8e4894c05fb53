// RAMBO phase space, x1/x2 sampling, flux factor, cuts and boost as FP64 device functions.
//
// Replaces python_package/madflow/phasespace.py (reference, TensorFlow):
//   _gen_unconstrained_momenta :121-142, _conformal_transformation :108-118, rambo :145-212,
//   _massive_xfactor :38-105, _get_x1x2 :215-233, _get_x1x2_onshell :236-254, ramboflow :257-319,
//   _boost_to_lab :322-356, PhaseSpaceGenerator cuts :405-478.
// One call = one event, everything in registers.  Differences from the reference, by design:
//   * the Newton iteration on the massive rescaling factor runs per event until f <= ACC
//     (<= 10 steps) instead of stopping for the whole batch when the first event converges
//     (phasespace.py:90-92), and the energies are those of the final factor;
//     this equals the reference evaluated with a batch of one event.
//   * constants arrive in PSConst: the caller chooses the reference's float32-rounded values
//     (PI, ACC, GeV->pb; see oracle/__init__.py) or exact doubles.
#pragma once
#include "mf_complex.cuh"

namespace mf {

constexpr int MF_MAX_OUT = 8;    // outgoing particles
constexpr int MF_MAX_CUTS = 16;

struct PSConst {
  double pi;       // phasespace.py:16
  double acc;      // phasespace.py:17
  double gev2pb;   // phasespace.py:316
  double wt0;      // (n-1) log(pi/2) - 2 lgamma(n-1) - log(n-1), n = number of outgoing (phasespace.py:185-187)
  double inv_norm; // 1/(2 pi)^(3n-4)                                     (phasespace.py:191)
};

// CUT_MIJ / CUT_DR act on a PAIR of particles (an extension: the reference only has single-particle cuts, which
// leave the final-state gluon-gluon collinear singularity of g g > t t~ g g (g) unregulated)
enum CutVar { CUT_PT = 0, CUT_MT = 1, CUT_MT2 = 2, CUT_MIJ = 3, CUT_DR = 4 };

struct Cut {
  int var;        // CutVar
  int particle;   // index into the nexternal momenta; pair variables: i + 256 * j
  int has_min, has_max;
  double vmin, vmax;
};

struct CutList {
  int n;
  Cut c[MF_MAX_CUTS];
};

// phasespace.py:405-422
MF_DEV double cut_value(int var, const double p[4]) {
  const double pt2 = p[1] * p[1] + p[2] * p[2];
  if (var == CUT_PT) return sqrt(pt2);
  const double m2 = p[0] * p[0] - (p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
  // the reference squares pt = sqrt(px^2+py^2) again: mt2 = m2 + pt**2
  const double pt = sqrt(pt2);
  const double v = m2 + pt * pt;
  return var == CUT_MT2 ? v : sqrt(v);
}

// invariant mass / Delta R = sqrt(d eta^2 + d phi^2) of a pair
MF_DEV double pair_cut_value(int var, const double a[4], const double b[4]) {
  if (var == CUT_MIJ) {
    const double e = a[0] + b[0], x = a[1] + b[1], y = a[2] + b[2], z = a[3] + b[3];
    const double m2 = e * e - (x * x + y * y + z * z);
    return m2 > 0.0 ? sqrt(m2) : 0.0;
  }
  const double pa = sqrt(a[1] * a[1] + a[2] * a[2] + a[3] * a[3]), pb = sqrt(b[1] * b[1] + b[2] * b[2] + b[3] * b[3]);
  const double deta = 0.5 * log((pa + a[3]) / (pa - a[3])) - 0.5 * log((pb + b[3]) / (pb - b[3]));
  double dphi = fabs(atan2(a[2], a[1]) - atan2(b[2], b[1]));
  if (dphi > M_PI) dphi = 2.0 * M_PI - dphi;
  return sqrt(deta * deta + dphi * dphi);
}

// phasespace.py:444-461: strict inequalities, all cuts must pass
template <int NEXT>
MF_DEV bool pass_cuts(const CutList& cuts, const double p[NEXT][4]) {
  bool ok = true;
  for (int i = 0; i < cuts.n; ++i) {
    const Cut& c = cuts.c[i];
    const double v = c.var >= CUT_MIJ ? pair_cut_value(c.var, p[c.particle & 0xff], p[(c.particle >> 8) & 0xff])
                                      : cut_value(c.var, p[c.particle]);
    if (c.has_min) ok = ok && (v > c.vmin);
    if (c.has_max) ok = ok && (v < c.vmax);
  }
  return ok;
}

// Massless RAMBO for NOUT particles at total energy sqrts; xr = 4*NOUT uniforms.
// Returns log-weight pieces in wt (the massless weight itself, phasespace.py:185-191).
template <int NOUT>
MF_DEV void rambo_massless(const double* xr, double sqrts, const PSConst& k, double p[NOUT][4], double& wt) {
  double q[NOUT][4];
  double Q[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const double costh = 2.0 * xr[4 * i + 0] - 1.0;
    const double sinth = sqrt(1.0 - costh * costh);
    const double phi = 2 * k.pi * xr[4 * i + 1];
    const double en = -1.0 * log(xr[4 * i + 2] * xr[4 * i + 3]);
    double sn, cs;
    sincos(phi, &sn, &cs);
    q[i][0] = en;
    q[i][1] = en * sinth * sn;
    q[i][2] = en * sinth * cs;
    q[i][3] = en * costh;
#pragma unroll
    for (int m = 0; m < 4; ++m) Q[m] += q[i][m];
  }
  const double qmass = sqrt(Q[0] * Q[0] - (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]));
  const double x = sqrts / qmass;
  const double b[4] = {-Q[0] / qmass, -Q[1] / qmass, -Q[2] / qmass, -Q[3] / qmass};
  const double gamma = -b[0];
  const double a = 1.0 / (1.0 + gamma);
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const double bq = q[i][1] * b[1] + q[i][2] * b[2] + q[i][3] * b[3];
    const double tmp = bq * a + q[i][0];
    p[i][0] = (q[i][0] * gamma + bq) * x;
    p[i][1] = (q[i][1] + b[1] * tmp) * x;
    p[i][2] = (q[i][2] + b[2] * tmp) * x;
    p[i][3] = (q[i][3] + b[3] * tmp) * x;
  }
  wt = k.wt0 + (2 * NOUT - 4) * log(sqrts);
}

// Full RAMBO: massless generation + (if any mass is non-zero) the massive rescaling.
// masses == nullptr or all zero -> massless.  On return wt is the phase-space weight.
template <int NOUT>
MF_DEV void rambo(const double* xr, double sqrts, const double* masses, bool massive, const PSConst& k,
                  double p[NOUT][4], double& wt) {
  double lw;
  rambo_massless<NOUT>(xr, sqrts, k, p, lw);
  if (!massive) {
    wt = exp(lw) * k.inv_norm;
    return;
  }
  double msum = 0.0;
#pragma unroll
  for (int i = 0; i < NOUT; ++i) msum += masses[i];
  const double r = msum / sqrts;
  double x = sqrt(1.0 - r * r);
  // Newton on f(x) = sum_i sqrt(m_i^2 + x^2 e_i^2) - sqrts, per event (phasespace.py:76-90).  The reference stops at
  // f <= ACC = 1e-14 GeV (absolute), which the round-off of f at sqrt(s) ~ 1e2..1e4 GeV never lets it reach: it would
  // always run all 10 steps.  Here the loop also stops once |f| is at the round-off level of sqrt(s); the steps that
  // are skipped would move x by < 1e-15 relative.
  const double f_noise = 8.0 * 2.220446049250313e-16 * sqrts;
  for (int it = 0; it < 10; ++it) {
    double f0 = -sqrts, g0 = 0.0;
#pragma unroll
    for (int i = 0; i < NOUT; ++i) {
      const double e2 = p[i][0] * p[i][0];
      const double ne = sqrt(masses[i] * masses[i] + e2 * (x * x));
      f0 += ne;
      g0 += e2 / ne;
    }
    if (!(f0 > k.acc)) break;
    x = x - f0 / (x * g0);
    if (fabs(f0) <= f_noise) break;   // this step was the last one that can change x
  }
  double wt2 = 1.0, wt3 = 0.0;
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const double e0 = p[i][0];
    const double ne = sqrt(masses[i] * masses[i] + (e0 * e0) * (x * x));
    const double v = e0 * x;
    wt2 *= v / ne;
    wt3 += v * v / ne;
    p[i][0] = ne;
    p[i][1] *= x;
    p[i][2] *= x;
    p[i][3] *= x;
  }
  lw += (2 * NOUT - 3) * log(x) + log(wt2 / wt3 * sqrts);
  wt = exp(lw) * k.inv_norm;
}

// phasespace.py:215-233
MF_DEV void get_x1x2(double u0, double u1, double shat_min, double s_in, double& shat, double& wgt, double& x1,
                     double& x2) {
  const double taumin = shat_min / s_in;
  const double delta = 1.0 - taumin;
  const double tau = u0 * delta + taumin;
  x1 = pow(tau, u1);
  x2 = tau / x1;
  wgt = delta * (-1.0 * log(tau));
  shat = x1 * x2 * s_in;
}

// ramboflow (phasespace.py:257-319) for NEXT external particles (2 incoming + NEXT-2 outgoing):
// xr holds 4*(NEXT-2)+2 numbers.  Momenta in the partonic centre-of-mass frame.
template <int NEXT>
MF_DEV void ramboflow(const double* xr, double com_sqrts, const double* masses, bool massive, double shat_min,
                      const PSConst& k, double p[NEXT][4], double& wgt, double& x1, double& x2) {
  constexpr int NOUT = NEXT - 2;
  double shat;
  get_x1x2(xr[0], xr[1], shat_min, com_sqrts * com_sqrts, shat, wgt, x1, x2);
  const double roots = sqrt(shat);
  double wtps;
  rambo<NOUT>(xr + 2, roots, masses, massive, k, &p[2], wtps);
  wgt *= wtps;
  const double ein = roots / 2.0;
  p[0][0] = ein, p[0][1] = 0.0, p[0][2] = 0.0, p[0][3] = ein;
  p[1][0] = ein, p[1][1] = 0.0, p[1][2] = 0.0, p[1][3] = -ein;
  wgt *= k.gev2pb;
  wgt /= 2 * shat;
}

// 2 -> 1 (phasespace.py:236-254, :293-297): xr[0] only
MF_DEV void ramboflow_2to1(double u, double com_sqrts, double mass, const PSConst& k, double p[3][4], double& wgt,
                           double& x1, double& x2) {
  const double s_in = com_sqrts * com_sqrts;
  const double ratio = mass / sqrt(s_in);
  const double tau_max = log(ratio);
  wgt = -2.0 * tau_max / s_in;
  const double tau = tau_max - 2.0 * u * tau_max;
  x1 = ratio * exp(tau);
  x2 = ratio * exp(-tau);
  const double shat = mass * mass;
  const double roots = sqrt(shat);
  const double ein = roots / 2.0;
  p[0][0] = ein, p[0][1] = 0.0, p[0][2] = 0.0, p[0][3] = ein;
  p[1][0] = ein, p[1][1] = 0.0, p[1][2] = 0.0, p[1][3] = -ein;
  p[2][0] = roots, p[2][1] = 0.0, p[2][2] = 0.0, p[2][3] = 0.0;
  wgt *= k.gev2pb;
  wgt /= 2 * shat;
}

// phasespace.py:322-356
template <int NEXT>
MF_DEV void boost_to_lab(double p[NEXT][4], double x1, double x2) {
  const double eta = -0.5 * log(x1 / x2);
  const double cth = cosh(eta), sth = sinh(eta);
#pragma unroll
  for (int i = 0; i < NEXT; ++i) {
    const double e = p[i][0], z = p[i][3];
    p[i][0] = e * cth + z * (-1.0 * sth);
    p[i][3] = e * (-1.0 * sth) + z * cth;
  }
}

}  // namespace mf
