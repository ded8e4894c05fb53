// ALOHA vertex routines of the SM QCD sector as FP64 device functions.
//
// Replaces the TensorFlow routines madflow's ALOHA writer emits
// (madgraph_plugin/PyOut_create_aloha.py:124-197).  Pinned by the frozen MG5 2.9.2 output in
// python_package/madflow/tests/mockup_debug_me.py: FFV1_0 :42-54, FFV1_1 :66-194,
// FFV1_2 :206-328, VVV1P0_1 :340-397.  The others (VVV1_0, FFV1P0_3, VVVV{1,3,4}_0,
// VVVV{1,3,4}P0_1) follow from the UFO Lorentz structures with the same conventions:
//   amplitude  = -i * COUP * L(contracted)
//   off-shell  = -i * COUP * L^mu / (P^2 - M(M - iW)),  P = -(Re w0, Re w1, Im w1, Im w0)
// The spinor routines are written with the slashed objects factored out:
//   X = F2bar * Vslash  (row),   Y = Vslash * F1  (column),   F1.X = F2.Y = J.V
// which is algebraically identical to the generated expressions (verified to 1e-15 by
// tests/test_aloha*.py) at ~40% of their operation count; momenta, masses, widths are real.
#pragma once
#include "mf_complex.cuh"

namespace mf {

struct Mom {
  double e, x, y, z;
};

MF_DEV Mom mom_of(const cxd w[6], double s) { return Mom{s * w[0].re, s * w[1].re, s * w[1].im, s * w[0].im}; }

// COUP / (P^2 - M(M - iW))  (mockup_debug_me.py:87)
MF_DEV cxd propagator(cxd coup, const Mom& P, double M, double W) {
  // rounded like the reference's P0**2 - P1**2 - P2**2 - P3**2 - M*(M - iW): no FMA contraction,
  // the cancellation in P^2 amplifies any differently rounded product by E^2/P^2
  const double p2 = ((mul_rn(P.e, P.e) - mul_rn(P.x, P.x)) - mul_rn(P.y, P.y)) - mul_rn(P.z, P.z);
  return cdiv(coup, mk(p2 - mul_rn(M, M), M * W));
}

// X = F2bar * Vslash : X[0..3] <-> spinor slots 2..5
MF_DEV void slash_row(const cxd F2[6], const cxd V[6], cxd X[4]) {
  const cxd vp = V[2] + V[5], vm = V[2] - V[5];
  const cxd a = mk(V[3].re - V[4].im, V[3].im + V[4].re);  // V3 + i V4
  const cxd b = mk(V[3].re + V[4].im, V[3].im - V[4].re);  // V3 - i V4
  X[0] = F2[4] * vp + F2[5] * a;
  X[1] = F2[4] * b + F2[5] * vm;
  X[2] = F2[2] * vm - F2[3] * a;
  X[3] = F2[3] * vp - F2[2] * b;
}

// Y = Vslash * F1 : Y[0..3] <-> spinor slots 2..5
MF_DEV void slash_col(const cxd F1[6], const cxd V[6], cxd Y[4]) {
  const cxd vp = V[2] + V[5], vm = V[2] - V[5];
  const cxd a = mk(V[3].re - V[4].im, V[3].im + V[4].re);
  const cxd b = mk(V[3].re + V[4].im, V[3].im - V[4].re);
  Y[0] = F1[4] * vm - F1[5] * b;
  Y[1] = F1[5] * vp - F1[4] * a;
  Y[2] = F1[2] * vp + F1[3] * b;
  Y[3] = F1[2] * a + F1[3] * vm;
}

MF_DEV cxd FFV1_0(const cxd F1[6], const cxd F2[6], const cxd V3[6], cxd COUP) {
  cxd X[4];
  slash_row(F2, V3, X);
  const cxd t = F1[2] * X[0] + F1[3] * X[1] + F1[4] * X[2] + F1[5] * X[3];
  return COUP * mul_mi(t);
}

// off-shell outgoing-flow fermion (result goes where an oxxxxx wavefunction goes)
MF_DEV void FFV1_1(const cxd F2[6], const cxd V3[6], cxd COUP, double M1, double W1, cxd F1[6]) {
  F1[0] = F2[0] + V3[0];
  F1[1] = F2[1] + V3[1];
  const Mom P = mom_of(F1, -1.0);
  const cxd iD = mul_i(propagator(COUP, P, M1, W1));
  cxd X[4];
  slash_row(F2, V3, X);
  const double Pp = P.e + P.z, Pm = P.e - P.z;
  const cxd Pa = mk(P.x, P.y), Pb = mk(P.x, -P.y);
  F1[2] = iD * (M1 * X[0] - Pp * X[2] - Pa * X[3]);
  F1[3] = iD * (M1 * X[1] - Pb * X[2] - Pm * X[3]);
  F1[4] = iD * (M1 * X[2] - Pm * X[0] + Pa * X[1]);
  F1[5] = iD * (M1 * X[3] + Pb * X[0] - Pp * X[1]);
}

// off-shell incoming-flow fermion (result goes where an ixxxxx wavefunction goes)
MF_DEV void FFV1_2(const cxd F1[6], const cxd V3[6], cxd COUP, double M2, double W2, cxd F2[6]) {
  F2[0] = F1[0] + V3[0];
  F2[1] = F1[1] + V3[1];
  const Mom P = mom_of(F2, -1.0);
  const cxd iD = mul_i(propagator(COUP, P, M2, W2));
  cxd Y[4];
  slash_col(F1, V3, Y);
  const double Pp = P.e + P.z, Pm = P.e - P.z;
  const cxd Pa = mk(P.x, P.y), Pb = mk(P.x, -P.y);
  F2[2] = iD * (M2 * Y[0] + Pm * Y[2] - Pb * Y[3]);
  F2[3] = iD * (M2 * Y[1] + Pp * Y[3] - Pa * Y[2]);
  F2[4] = iD * (M2 * Y[2] + Pp * Y[0] + Pb * Y[1]);
  F2[5] = iD * (M2 * Y[3] + Pa * Y[0] + Pm * Y[1]);
}

// Minkowski products on the polarisation slots
MF_DEV cxd vdot(const cxd A[6], const cxd B[6]) { return A[2] * B[2] - A[3] * B[3] - A[4] * B[4] - A[5] * B[5]; }
MF_DEV cxd pdot(const Mom& P, const cxd V[6]) { return P.e * V[2] - P.x * V[3] - P.y * V[4] - P.z * V[5]; }

// gluon current from the quark line: V3^mu = -i COUP J^mu / P3^2,  FFV1_0 = -i COUP (J.V)
MF_DEV void FFV1P0_3(const cxd F1[6], const cxd F2[6], cxd COUP, double M3, double W3, cxd V3[6]) {
  V3[0] = F1[0] + F2[0];
  V3[1] = F1[1] + F2[1];
  const Mom P = mom_of(V3, -1.0);
  const cxd D = mul_mi(propagator(COUP, P, M3, W3));
  const cxd a = F1[2] * F2[4], b = F1[3] * F2[5], c = F1[4] * F2[2], d = F1[5] * F2[3];
  const cxd e = F1[2] * F2[5], f = F1[3] * F2[4], g = F1[4] * F2[3], h = F1[5] * F2[2];
  V3[2] = D * (a + b + c + d);
  V3[3] = D * (g + h - e - f);
  V3[4] = D * mul_mi(e - f - g + h);
  V3[5] = D * (b + c - a - d);
}

MF_DEV void VVV1P0_1(const cxd V2[6], const cxd V3[6], cxd COUP, double M1, double W1, cxd V1[6]) {
  const Mom P2 = mom_of(V2, 1.0), P3 = mom_of(V3, 1.0);
  V1[0] = V2[0] + V3[0];
  V1[1] = V2[1] + V3[1];
  const Mom P1 = mom_of(V1, -1.0);
  const Mom d12 = Mom{P1.e - P2.e, P1.x - P2.x, P1.y - P2.y, P1.z - P2.z};  // V3.(P1-P2) = TMP1-TMP2
  const Mom d13 = Mom{P1.e - P3.e, P1.x - P3.x, P1.y - P3.y, P1.z - P3.z};  // V2.(P1-P3) = TMP3-TMP4
  const cxd t12 = pdot(d12, V3), t34 = pdot(d13, V2), t5 = vdot(V3, V2);
  const cxd D = mul_mi(propagator(COUP, P1, M1, W1));  // -i * denom
  // V1^mu = denom*( TMP5*(-i)(P2-P3)^mu + V2^mu*(-i)(TMP1-TMP2) + V3^mu*(+i)(TMP3-TMP4) )
  const double q[4] = {P2.e - P3.e, P2.x - P3.x, P2.y - P3.y, P2.z - P3.z};
#pragma unroll
  for (int k = 0; k < 4; ++k) V1[2 + k] = D * (q[k] * t5 + V2[2 + k] * t12 - V3[2 + k] * t34);
}

MF_DEV cxd VVV1_0(const cxd V1[6], const cxd V2[6], const cxd V3[6], cxd COUP) {
  const Mom P1 = mom_of(V1, 1.0), P2 = mom_of(V2, 1.0), P3 = mom_of(V3, 1.0);
  const Mom d12 = Mom{P1.e - P2.e, P1.x - P2.x, P1.y - P2.y, P1.z - P2.z};
  const Mom d31 = Mom{P3.e - P1.e, P3.x - P1.x, P3.y - P1.y, P3.z - P1.z};
  const Mom d23 = Mom{P2.e - P3.e, P2.x - P3.x, P2.y - P3.y, P2.z - P3.z};
  const cxd L = vdot(V1, V2) * pdot(d12, V3) + vdot(V1, V3) * pdot(d31, V2) + vdot(V2, V3) * pdot(d23, V1);
  return COUP * mul_mi(L);
}

// four-gluon contact terms; KIND in {1,3,4} names the UFO structure VVVV<KIND>
template <int KIND>
MF_DEV cxd VVVV_0(const cxd V1[6], const cxd V2[6], const cxd V3[6], const cxd V4[6], cxd COUP) {
  cxd L;
  if (KIND == 1) L = vdot(V1, V4) * vdot(V2, V3) - vdot(V1, V3) * vdot(V2, V4);
  if (KIND == 3) L = vdot(V1, V4) * vdot(V2, V3) - vdot(V1, V2) * vdot(V3, V4);
  if (KIND == 4) L = vdot(V1, V3) * vdot(V2, V4) - vdot(V1, V2) * vdot(V3, V4);
  return COUP * mul_mi(L);
}

template <int KIND>
MF_DEV void VVVVP0_1(const cxd V2[6], const cxd V3[6], const cxd V4[6], cxd COUP, double M1, double W1, cxd V1[6]) {
  V1[0] = V2[0] + V3[0] + V4[0];
  V1[1] = V2[1] + V3[1] + V4[1];
  const Mom P1 = mom_of(V1, -1.0);
  const cxd D = mul_mi(propagator(COUP, P1, M1, W1));
  cxd s, t;  // L^mu = A^mu * s - B^mu * t
  const cxd* A;
  const cxd* B;
  if (KIND == 1) { A = V4; s = vdot(V2, V3); B = V3; t = vdot(V2, V4); }
  if (KIND == 3) { A = V4; s = vdot(V2, V3); B = V2; t = vdot(V3, V4); }
  if (KIND == 4) { A = V3; s = vdot(V2, V4); B = V2; t = vdot(V3, V4); }
#pragma unroll
  for (int k = 2; k < 6; ++k) V1[k] = D * (A[k] * s - B[k] * t);
}

}  // namespace mf
