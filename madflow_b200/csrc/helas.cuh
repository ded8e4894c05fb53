// HELAS external wavefunctions as FP64 device functions.
//
// Replaces python_package/madflow/wavefunctions_flow.py (reference, TensorFlow):
//   sxxxxx :33-51, ixxxxx :55-85 (+ :159-330), oxxxxx :88-116 (+ :334-462),
//   vxxxxx :119-154 (+ :467-747), sign :19-29.
// One call = one event.  `mass`, `nhel`, `nsf` are uniform over the grid (kernel parameters or
// loop counters), so the branches on them are warp-uniform; the per-event selects the reference
// expresses with tf.where (pp == 0, pp+pz == 0, pt == 0) are predicated selects here.
// A wavefunction is cxd w[6]: w[0], w[1] carry the momentum, w[2..5] the components.
#pragma once
#include "mf_complex.cuh"

namespace mf {

// wavefunctions_flow.py:19-29 -- x*sign(y) with sign(0) = 0 (TensorFlow semantics, not Fortran's)
MF_DEV double sign_tf(double x, double y) { return y > 0.0 ? x : (y < 0.0 ? -x : 0.0 * x); }

MF_DEV void sxxxxx(const double p[4], int nss, cxd w[3]) {
  w[0] = mk(p[0] * nss, p[3] * nss);
  w[1] = mk(p[1] * nss, p[2] * nss);
  w[2] = mk(1.0, 0.0);
}

// rest-frame spinor, wavefunctions_flow.py:372-387 (_ox_massive_pp_zero)
MF_DEV void pp_zero_spinor(double fmass, int nsf, int ip, int im, double v[4]) {
  const double sqm0 = sqrt(fabs(fmass));
  const double sqm1 = sign_tf(sqm0, fmass);
  const double s_im = (abs(im) == 0) ? sqm0 : sqm1;
  const double s_ip = (abs(ip) == 0) ? sqm0 : sqm1;
  v[0] = im * s_im;
  v[1] = ip * nsf * s_im;
  v[2] = im * nsf * s_ip;
  v[3] = ip * s_ip;
}

// Shared by ixxxxx / oxxxxx massive branches (wavefunctions_flow.py:204-229, :405-432).
// ysign = +1 (ixxxxx) or -1 (oxxxxx).  Returns sfomeg[2], chi[2] (chi[0] real).
MF_DEV void massive_blocks(const double p[4], double fmass, int nsf, int nh, double ysign, double& pp,
                           double sfomeg[2], cxd chi[2]) {
  pp = fmin(p[0], sqrt(mul_rn(p[1], p[1]) + mul_rn(p[2], p[2]) + mul_rn(p[3], p[3])));
  const double sf0 = (1 + nsf + (1 - nsf) * nh) * 0.5;
  const double sf1 = (1 + nsf - (1 - nsf) * nh) * 0.5;
  const double sq = sqrt(p[0] + pp);
  const double om0 = sq, om1 = fmass / sq;
  // ip = (1+nh)/2, im = (1-nh)/2 :  nh=+1 -> (1,0), nh=-1 -> (0,1)
  sfomeg[0] = sf0 * (nh == 1 ? om1 : om0);
  sfomeg[1] = sf1 * (nh == 1 ? om0 : om1);
  const double pp3 = fmax(pp + p[3], 0.0);
  const double den = sqrt(2.0 * pp * pp3);
  chi[1] = (pp3 == 0.0) ? mk(-nh, 0.0) : mk(nh * p[1] / den, ysign * p[2] / den);
  chi[0] = mk(sqrt(pp3 * 0.5 / pp), 0.0);
}

MF_DEV void ixxxxx(const double p[4], double fmass, int nhel, int nsf, cxd w[6]) {
  w[0] = mk(-p[0] * nsf, -p[3] * nsf);
  w[1] = mk(-p[1] * nsf, -p[2] * nsf);
  const int nh = nhel * nsf;
  if (fmass != 0.0) {
    double pp, sfomeg[2];
    cxd chi[2];
    massive_blocks(p, fmass, nsf, nh, 1.0, pp, sfomeg, chi);
    const int ip = (1 + nh) / 2, im = (1 - nh) / 2;
    if (pp == 0.0) {
      double v[4];
      pp_zero_spinor(fmass, nsf, im, ip, v);  // :181-183, (ip,im) exchanged
      for (int k = 0; k < 4; ++k) w[2 + k] = mk(v[k], 0.0);
    } else {
      const cxd c_im = chi[im], c_ip = chi[ip];
      w[2] = sfomeg[0] * c_im;
      w[3] = sfomeg[0] * c_ip;
      w[4] = sfomeg[1] * c_im;
      w[5] = sfomeg[1] * c_ip;
    }
  } else {
    const double sqp0p3 = sqrt(fmax(p[0] + p[3], 0.0)) * nsf;
    const cxd chi1 = (sqp0p3 == 0.0) ? mk(-nhel * sqrt(2.0 * p[0]), 0.0) : mk(nh * p[1] / sqp0p3, p[2] / sqp0p3);
    const cxd chi0 = mk(sqp0p3, 0.0), z = mk(0.0, 0.0);
    if (nh == 1) {
      w[2] = z, w[3] = z, w[4] = chi0, w[5] = chi1;
    } else {
      w[2] = chi1, w[3] = chi0, w[4] = z, w[5] = z;
    }
  }
}

MF_DEV void oxxxxx(const double p[4], double fmass, int nhel, int nsf, cxd w[6]) {
  w[0] = mk(p[0] * nsf, p[3] * nsf);
  w[1] = mk(p[1] * nsf, p[2] * nsf);
  const int nh = nhel * nsf;
  if (fmass != 0.0) {
    double pp, sfomeg[2];
    cxd chi[2];
    massive_blocks(p, fmass, nsf, nh, -1.0, pp, sfomeg, chi);
    if (pp == 0.0) {
      const int ip = -((1 - nh) / 2) * nhel;  // :353-354
      const int im = ((1 + nh) / 2) * nhel;
      double v[4];
      pp_zero_spinor(fmass, nsf, ip, im, v);
      for (int k = 0; k < 4; ++k) w[2 + k] = mk(v[k], 0.0);
    } else {
      const int ip = (1 + nh) / 2, im = (1 - nh) / 2;
      const cxd c_im = chi[im], c_ip = chi[ip];
      w[2] = sfomeg[1] * c_im;
      w[3] = sfomeg[1] * c_ip;
      w[4] = sfomeg[0] * c_im;
      w[5] = sfomeg[0] * c_ip;
    }
  } else {
    const double sqp0p3 = sqrt(fmax(p[0] + p[3], 0.0)) * nsf;
    const cxd chi0 = (sqp0p3 == 0.0) ? mk(-nhel * sqrt(2.0 * p[0]), 0.0) : mk(nh * p[1] / sqp0p3, -p[2] / sqp0p3);
    const cxd chi1 = mk(sqp0p3, 0.0), z = mk(0.0, 0.0);
    if (nh == 1) {
      w[2] = chi1, w[3] = chi0, w[4] = z, w[5] = z;
    } else {
      w[2] = z, w[3] = z, w[4] = chi0, w[5] = chi1;
    }
  }
}

// sqh: sqrt(1/2) as the caller wants it (the reference's is float32-rounded, wavefunctions_flow.py:10)
MF_DEV void vxxxxx(const double p[4], double vmass, int nhel, int nsv, double sqh, cxd w[6]) {
  w[0] = mk(p[0] * nsv, p[3] * nsv);
  w[1] = mk(p[1] * nsv, p[2] * nsv);
  if (nhel == 4) {  // BRST check, :467-515
    const double d = (vmass == 0.0) ? p[0] : vmass;
    for (int k = 0; k < 4; ++k) w[2 + k] = mk(p[k] / d, 0.0);
    return;
  }
  const double pt2 = mul_rn(p[1], p[1]) + mul_rn(p[2], p[2]);
  const int ahel = abs(nhel);
  const double hel0 = 1.0 - ahel;
  const double nsvahl = nsv * ahel;
  if (vmass != 0.0) {  // :548-678
    const double pp = fmin(p[0], sqrt(pt2 + mul_rn(p[3], p[3])));
    const double pt = fmin(pp, sqrt(pt2));
    if (pp == 0.0) {
      w[2] = mk(1.0, 0.0);  // sic: the reference leaves v[0] = 1 in the rest frame (:588)
      w[3] = mk(-nhel * sqh, 0.0);
      w[4] = mk(0.0, nsvahl * sqh);
      w[5] = mk(hel0, 0.0);
      return;
    }
    const double emp = p[0] / (vmass * pp);
    w[2] = mk(hel0 * pp / vmass, 0.0);
    w[5] = mk(mul_rn(hel0 * p[3], emp) + mul_rn(nhel * pt / pp, sqh), 0.0);
    if (pt != 0.0) {
      const double pzpt = p[3] / (pp * pt) * sqh * nhel;
      w[3] = mk(mul_rn(hel0 * p[1], emp) - mul_rn(p[1], pzpt), -nsvahl * p[2] / pt * sqh);
      w[4] = mk(mul_rn(hel0 * p[2], emp) - mul_rn(p[2], pzpt), nsvahl * p[1] / pt * sqh);
    } else {
      w[3] = mk(-nhel * sqh, 0.0);
      w[4] = mk(0.0, nsvahl * sign_tf(sqh, p[3]));
    }
  } else {  // :683-747
    const double pp = p[0];
    const double pt = sqrt(pt2);
    w[2] = mk(0.0, 0.0);
    w[5] = mk(nhel * pt / pp * sqh, 0.0);
    if (pt != 0.0) {
      const double pzpt = p[3] / (pp * pt) * sqh * nhel;
      w[3] = mk(-p[1] * pzpt, -nsv * p[2] / pt * sqh);
      w[4] = mk(-p[2] * pzpt, nsv * p[1] / pt * sqh);
    } else {
      w[3] = mk(-nhel * sqh, 0.0);
      w[4] = mk(0.0, nsv * sign_tf(sqh, p[3]));
    }
  }
}

}  // namespace mf
