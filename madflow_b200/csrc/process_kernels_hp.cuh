// Helicity-parallel matrix-element kernels ("hp"): a thread block evaluates E events at a time and
// its threads are the (event, helicity combination) pairs.
//
// The reference evaluates the full HELAS call list once per helicity combination
// (matrix_method_python.inc:99-102: `for hel in self.helicities: ans += self.matrix(...)`), so a
// wavefunction over the leg subset S is recomputed 2^n times although it only depends on the
// 2^|S| helicities of its own legs.  Here every wavefunction is computed ONCE per distinct
// helicity assignment of its legs and kept in shared memory:
//
//   phase 1   external wavefunctions: n legs x 2 helicities x E events
//   phase 2   off-shell currents level by level (level = number of legs); the work items of a
//             level are (current, event, helicity variant), spread over all threads of the block;
//             table driven, one copy of each vertex routine in the instruction stream: hp_units
//             (term rows + work items), or -- g g > t t~ g g (g) -- the packed units slu_units
//             (warp trips of one class of objects, one 64-bit word per (unit, term), see "SLU")
//   phase 3   amplitudes, batch by batch: (a) the pair objects of the batch (see HpPair), (b) the
//             amplitudes of ALL helicity combinations on the FP64 tensor cores: amplitude[vq][vx] =
//             sum_k Q_k[vq] x_k[vx] is a (variants of Q x 4)(4 x variants of x) complex matrix
//             product, evaluated in 8x8 tiles by four mma.sync.m8n8k4.f64 per tile and written to
//             the amplitude buffer [row][helicity combination], (c) the JAMP sums: thread
//             (e, h, colour group) adds the buffer rows into its JAMP registers through generated
//             code (static offsets, coefficients +-1, +-i folded into the adds)
//   phase 4   colour contraction per thread (colour groups exchange JAMPs through shared memory),
//             then the sum over helicities and colour groups is a warp-shuffle + shared-memory
//             reduction per event.
//
// Shared memory: every event of the block owns HP_EVSTRIDE cxd (= 16 bytes): HP_WFSIZE of
// wavefunctions, HP_SCRATCH of pair objects, HP_NB * HP_NHP of amplitudes.  Wavefunction w lives at
// offset P::wf(w).off of its event's area:
//   [ 0..1 ]              momentum slots w0, w1  (helicity independent)
//   [ 2 + hp_slot(k,nv,v) ]  component k = 0..3 (HELAS slots 2..5), helicity variant v < nv = 2^|S|
// Variant bits follow the ascending leg order of S; bit = 1 means helicity +1.  hp_slot = k*nv + (v ^ 2k):
// the XOR makes BOTH access patterns bank-conflict free for 16-byte elements -- 8 consecutive variants
// of one component (current / pair-object phases) and the tensor-core fragment (2 variants x 4 components
// per quarter warp).
#pragma once
#include <type_traits>

#include "pipeline_kernels.cuh"
#include "process_kernels.cuh"

namespace mf {

enum HpType : unsigned char {
  HP_VXXXXX = 0, HP_OXXXXX = 1, HP_IXXXXX = 2,
  HP_FFV1_1 = 3, HP_FFV1_2 = 4, HP_FFV1P0_3 = 5, HP_VVV1P0_1 = 6,
  HP_VVVV1P0_1 = 7, HP_VVVV3P0_1 = 8, HP_VVVV4P0_1 = 9
};

struct HpWf {
  unsigned int off;      // start of the block, in cxd per event
  unsigned short nv;     // helicity variants = 2^(number of legs)
  unsigned short legs;   // bit mask of the external legs below this wavefunction
};

struct HpExt {
  unsigned char type, leg;
  signed char nsf, mass_idx;   // mass_idx < 0: massless
  unsigned short out;
};

// Vertices.  Every vertex is evaluated as a NUMERATOR in dual form: with one of its lines left open,
//   Q[k] = -i COUP (vertex contracted with its other lines)_k,   metric signs folded in,
// so that closing the open line with a wavefunction x gives the amplitude  amp = sum_k x[2+k] Q[k]  and a current is
// the same numerator times its propagator (hp_unit below):
//   ROW  (O, G):    Obar Gslash            FFV1_1(O, G)   | FFV1_0 closed with x = I
//   COL  (I, G):    Gslash I               FFV1_2(I, G)   | FFV1_0 closed with x = O
//   CUR  (I, O):    Obar gamma^mu I        FFV1P0_3(I, O) | FFV1_0 closed with x = G
//   VVV  (V2, V3):  three-gluon vertex     VVV1P0_1       | VVV1_0, the other two lines in cyclic order
//   VVVV (3 lines): contact term           VVVVkP0_1      | VVVVk_0
// The reference evaluates the whole vertex for each of the 2^n helicity combinations
// (matrix_method_python.inc:99-102); here a numerator costs 2^|its legs| evaluations and each of the 2^n
// combinations of an amplitude only a 4-term complex dot product (the tensor-core tiles below).
enum HpVertex : unsigned char { HP_Q_ROW = 0, HP_Q_COL = 1, HP_Q_CUR = 2, HP_Q_VVV = 3, HP_Q_VVVV = 4 };
// + HP_F_NOMOM: a current that nothing else is built from stores no momentum slots
enum HpFinish : unsigned char { HP_F_NONE = 0, HP_F_G = 1, HP_F_O = 2, HP_F_I = 3, HP_F_NOMOM = 4 };

// One term of an object (current or pair object = vertex numerator).  An object is the sum of its terms,
//   sum_t phase_t * numerator_t,   phase in {1, -1, i, -i}
// -- sub-diagrams over the same legs with linearly dependent colour factors (madflow_b200/recursion.py), all sharing
// one propagator -- and is stored once per helicity variant of its legs.  (32 bytes, fetched with two 16-byte loads.)
struct alignas(16) HpTerm {
  unsigned char type, nin, coup, phase;   // phase code 0..3 = 1, -1, i, -i (the sign of the coupling folded in)
  unsigned char term[2];                  // HP_Q_VVVV: sign<<6 | vector<<4 | dotA<<2 | dotB  (indices into the inputs)
  unsigned short out_nv, out_off;         // the object: helicity variants, offset in the event area (cxd)
  unsigned short in_off[3], in_nv[3];     // inputs: offset of the wavefunction block in the event area, variants
  unsigned char vmask[3];                 // bits of the object's variant index that form the input's variant index
  unsigned char finish;                   // HpFinish: none (numerator), or the propagator of a g / o / i current
  signed char mass_idx, width_idx;        // < 0: ZERO
  unsigned char pad[4];
};
static_assert(sizeof(HpTerm) == 32, "HpTerm is fetched as two 16-byte words");

struct alignas(8) HpWorkItem {  // (term, helicity variant of the object) and, precomputed, the variants of the term's
  unsigned short term;          // inputs that belong to it.  A UNIT = the work items of one (object, variant), stored
  unsigned char v, iv[3];       // consecutively; unit descriptor = first item | number of items << 24.
  unsigned short pad;
};
static_assert(sizeof(HpWorkItem) == 8, "HpWorkItem is fetched with one 8-byte load");

// One tensor-core work item of the amplitude phase: rows = 8 helicity variants of the pair object Q
// (from variant q0), columns = 8 variants of the wavefunction x (from x0); element (r, c) is the
// amplitude of helicity combination rowh[r] | colh[c] and goes to row `slot` of the amplitude buffer.
struct alignas(16) HpTile {
  unsigned short q, x;          // component 0 of Q and of x in the event area (cxd)
  unsigned short qnv, xnv;      // component strides = helicity variants of the objects
  unsigned char q0, x0;         // first variant of the tile
  unsigned char qvalid, xvalid; // rows / columns in use (the rest is padding)
  unsigned short slot, flags;   // row of the amplitude buffer; flags unused
  unsigned char rowh[8], colh[8];
};

struct HpBatch {   // one (helicity pass, batch): the units of its pair objects and its tiles
  int unit_begin, unit_end, tile_begin, tile_end;
};

// a 32-byte table row with two 16-byte loads instead of one load per field
template <class T>
MF_DEV T hp_fetch32(const T* p) {
  static_assert(sizeof(T) == 32, "two 16-byte words");
  union {
    uint4 w[2];
    T t;
  } u;
  u.w[0] = reinterpret_cast<const uint4*>(p)[0];
  u.w[1] = reinterpret_cast<const uint4*>(p)[1];
  return u.t;
}

// phase 1: work item `it` in [0, NEXT*E*2) = (leg, event, helicity)
template <class P>
MF_DEV void hp_externals(int it, int E, const double* mom /*[E][NEXT][4]*/, const double* par, double sqh, cxd* wf) {
  const int leg = it / (2 * E);
  const int r = it - leg * 2 * E;
  const int e = r >> 1, bit = r & 1;
  const HpExt x = P::ext(leg);
  const double* p = mom + (e * P::NEXT + x.leg) * 4;
  const double mass = x.mass_idx < 0 ? 0.0 : par[x.mass_idx];
  cxd w[6];
  const int hel = 2 * bit - 1;
  if (x.type == HP_VXXXXX) vxxxxx(p, mass, hel, x.nsf, sqh, w);
  else if (x.type == HP_OXXXXX) oxxxxx(p, mass, hel, x.nsf, w);
  else ixxxxx(p, mass, hel, x.nsf, w);
  cxd* o = wf + e * P::HP_EVSTRIDE + P::wf(x.out).off;
  if (bit == 0) o[0] = w[0], o[1] = w[1];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[2 + k * 2 + bit] = w[2 + k];
}

// position of helicity combination h inside a row of the amplitude buffer: folding bits 3..5 onto bits 0..2
// keeps the JAMP threads' reads (consecutive h) conflict free and spreads the tensor-core tiles' stores
// (8 lanes = 3 arbitrary helicity bits) over the banks
MF_DEV int hp_abuf_pos(int h) { return h ^ ((h >> 3) & 7); }

// position of (component k, helicity variant v) inside an object of nv variants (see the header)
MF_DEV int hp_slot(int k, int nv, int v) { return k * nv + (v ^ ((2 * k) & (nv - 1))); }

// the four components of one helicity variant of a wavefunction block (slots 2..5), and all 6 slots
MF_DEV void hp_load_comp(const cxd* blk, int nv, int v, cxd out[6]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) out[2 + k] = blk[2 + hp_slot(k, nv, v)];
}
MF_DEV void hp_load(const cxd* blk, int nv, int v, cxd out[6]) {
  out[0] = blk[0], out[1] = blk[1];
  hp_load_comp(blk, nv, v, out);
}

// K += sign * in[vi] * (in[da] . in[db]) for one term of a four-gluon structure (all indices warp-uniform)
MF_DEV void hp_quartic_term(unsigned char d, const cxd a[6], const cxd b[6], const cxd c[6], cxd d01, cxd d02, cxd d12,
                            cxd K[4]) {
  const int vi = (d >> 4) & 3, key = ((d >> 2) & 3) + (d & 3);  // dot: {0,1} -> 1, {0,2} -> 2, {1,2} -> 3
  cxd dot = key == 1 ? d01 : (key == 2 ? d02 : d12);
  if (d & 0x40) dot = -dot;
  switch (vi) {
    case 0:
#pragma unroll
      for (int k = 0; k < 4; ++k) K[k] = fma_c(a[2 + k], dot, K[k]);
      break;
    case 1:
#pragma unroll
      for (int k = 0; k < 4; ++k) K[k] = fma_c(b[2 + k], dot, K[k]);
      break;
    default:
#pragma unroll
      for (int k = 0; k < 4; ++k) K[k] = fma_c(c[2 + k], dot, K[k]);
      break;
  }
}

// The five vertex numerators (dual form, see HpVertex): Q[k] += f * (vertex with one line open)_k, f = -i COUP phase.
// Shared by the table-driven routine (hp_unit_terms) and the class-specialised straight-line units (slu_*).
MF_DEV void hp_q_row(const cxd a[6], const cxd b[6], cxd f, cxd Q[4]) {  // a = O (F2), b = G
  cxd X[4];
  slash_row(a, b, X);
#pragma unroll
  for (int k = 0; k < 4; ++k) Q[k] = fma_c(f, X[k], Q[k]);
}
MF_DEV void hp_q_col(const cxd a[6], const cxd b[6], cxd f, cxd Q[4]) {  // a = I (F1), b = G
  cxd Y[4];
  slash_col(a, b, Y);
#pragma unroll
  for (int k = 0; k < 4; ++k) Q[k] = fma_c(f, Y[k], Q[k]);
}
MF_DEV void hp_q_cur(const cxd a[6], const cxd b[6], cxd f, cxd Q[4]) {  // a = I (F1), b = O (F2): J^mu with FFV1_0 = -i COUP (J.V)
  const cxd tp = fma_c(a[5], b[3], a[2] * b[4]), tm = fma_c(a[4], b[2], a[3] * b[5]);  // t0 + t3, t1 + t2
  const cxd u0 = a[2] * b[5], u1 = a[3] * b[4], u2 = a[4] * b[3], u3 = a[5] * b[2];
  Q[0] = fma_c(f, tp + tm, Q[0]);
  Q[1] = fma_c(f, (u0 - u3) + (u1 - u2), Q[1]);          // -J1
  Q[2] = fma_c(f, mul_i((u0 + u3) - (u1 + u2)), Q[2]);   // -J2
  Q[3] = fma_c(f, tp - tm, Q[3]);                        // -J3
}
// a = V2, b = V3 of VVV1_0(V1,V2,V3) with their momentum slots; P1 = -(P2+P3); the only kind that needs the momenta
MF_DEV void hp_q_vvv(const cxd a[6], const cxd b[6], cxd f, cxd Q[4]) {
  const Mom P2 = mom_of(a, 1.0), P3 = mom_of(b, 1.0);
  const Mom d12 = Mom{-2.0 * P2.e - P3.e, -2.0 * P2.x - P3.x, -2.0 * P2.y - P3.y, -2.0 * P2.z - P3.z};  // P1-P2
  const Mom d31 = Mom{2.0 * P3.e + P2.e, 2.0 * P3.x + P2.x, 2.0 * P3.y + P2.y, 2.0 * P3.z + P2.z};      // P3-P1
  const cxd s3 = pdot(d12, b), s2 = pdot(d31, a), s23 = vdot(a, b);
  const double q[4] = {P2.e - P3.e, P2.x - P3.x, P2.y - P3.y, P2.z - P3.z};
  const cxd fm = -f;
  Q[0] = fma_c(f, fma_c(a[2], s3, fma_c(b[2], s2, q[0] * s23)), Q[0]);
#pragma unroll
  for (int k = 1; k < 4; ++k) Q[k] = fma_c(fm, fma_c(a[2 + k], s3, fma_c(b[2 + k], s2, q[k] * s23)), Q[k]);
}
// all four-gluon structures over the same three lines at once: K = ca a (b.c) + cb b (a.c) + cc c (a.b) with small
// integer weights (the UFO structures VVVV1/3/4 and the relative signs of the merged terms, codegen.py)
MF_DEV void hp_q_vvvv_merged(const cxd a[6], const cxd b[6], const cxd c[6], double ca, double cb, double cc, cxd f, cxd Q[4]) {
  const cxd da = ca * vdot(b, c), db = cb * vdot(a, c), dc = cc * vdot(a, b);
  const cxd fm = -f;
  Q[0] = fma_c(f, fma_c(a[2], da, fma_c(b[2], db, c[2] * dc)), Q[0]);
#pragma unroll
  for (int k = 1; k < 4; ++k) Q[k] = fma_c(fm, fma_c(a[2 + k], da, fma_c(b[2 + k], db, c[2 + k] * dc)), Q[k]);
}

// One unit of the current / pair-object phases: (object, helicity variant) of the event whose area is ev_e.  The
// terms of the object are evaluated in turn and added up (hp_unit_terms); the propagator is applied (currents) and the
// result stored by hp_unit_finish.  An object with many terms is split over 2^k neighbouring lanes (k = bits 28-29 of
// the unit descriptor), each evaluating a few terms; hp_units adds the partial sums with warp shuffles in a fixed order.
template <class P>
MF_DEV void hp_unit_terms(const unsigned unit, const cxd* coup_e, const cxd* ev_e, cxd Q[4], HpTerm& t, HpWorkItem& it) {
  const int begin = (int)(unit & 0xffffffu), count = (int)((unit >> 24) & 0xfu);
#pragma unroll 1
  for (int j = 0; j < count; ++j) {
    it = P::work_item(begin + j);
    t = P::term(it.term);
    const cxd* ab = ev_e + t.in_off[0];
    const cxd* bb = ev_e + t.in_off[1];
    cxd a[6], b[6];
    hp_load_comp(ab, t.in_nv[0], it.iv[0], a);
    hp_load_comp(bb, t.in_nv[1], it.iv[1], b);
    cxd f = mul_mi(coup_e[t.coup]);  // -i * COUP * phase
    if (t.phase & 2) f = mul_i(f);
    if (t.phase & 1) f = -f;
    switch (t.type) {
      case HP_Q_ROW: hp_q_row(a, b, f, Q); break;
      case HP_Q_COL: hp_q_col(a, b, f, Q); break;
      case HP_Q_CUR: hp_q_cur(a, b, f, Q); break;
      case HP_Q_VVV:
        a[0] = ab[0], a[1] = ab[1], b[0] = bb[0], b[1] = bb[1];
        hp_q_vvv(a, b, f, Q);
        break;
      default: {  // HP_Q_VVVV: K = sum_t sign_t * in[vec_t] * (in[dotA_t] . in[dotB_t])
        cxd c[6];
        hp_load_comp(ev_e + t.in_off[2], t.in_nv[2], it.iv[2], c);
        // the three Minkowski products once; each term picks one of them and one vector (warp-uniform)
        const cxd d01 = vdot(a, b), d02 = vdot(a, c), d12 = vdot(b, c);
        cxd K[4] = {mk(0, 0), mk(0, 0), mk(0, 0), mk(0, 0)};
        hp_quartic_term(t.term[0], a, b, c, d01, d02, d12, K);
        hp_quartic_term(t.term[1], a, b, c, d01, d02, d12, K);
        const cxd fm = -f;
        Q[0] = fma_c(f, K[0], Q[0]);
#pragma unroll
        for (int k = 1; k < 4; ++k) Q[k] = fma_c(fm, K[k], Q[k]);
      } break;
    }
  }
}

// Propagator + store of one (object, variant): Q = the summed numerators, w = the momentum slots of the object (the
// sum of its inputs'), o = its block in the event area.  fin = HpFinish (& 3), nomom: no momentum slots are stored.
MF_DEV void hp_finish_store(int fin, bool nomom, const cxd Q[4], const cxd w[2], double M, double W, cxd* o, int nv, int v) {
  const Mom Pm = Mom{-w[0].re, -w[1].re, -w[1].im, -w[0].im};
  const cxd inv = propagator(mk(1.0, 0.0), Pm, M, W);
  cxd r[4];
  if (fin == HP_F_G) {  // V^mu = numerator^mu / (P^2 - ..): undo the metric signs of the dual form
    r[0] = inv * Q[0];
    const cxd ninv = -inv;
#pragma unroll
    for (int k = 1; k < 4; ++k) r[k] = ninv * Q[k];
  } else {
    // fermion propagators (FFV1_1 / FFV1_2 of aloha_sm.cuh with i COUP X / den = -Q / den)
    const cxd ninv = -inv;
    const double Pp = Pm.e + Pm.z, Pn = Pm.e - Pm.z;
    const cxd Pa = mk(Pm.x, Pm.y), Pb = mk(Pm.x, -Pm.y);
    if (fin == HP_F_O) {
      r[0] = ninv * (M * Q[0] - Pp * Q[2] - Pa * Q[3]);
      r[1] = ninv * (M * Q[1] - Pb * Q[2] - Pn * Q[3]);
      r[2] = ninv * (M * Q[2] - Pn * Q[0] + Pa * Q[1]);
      r[3] = ninv * (M * Q[3] + Pb * Q[0] - Pp * Q[1]);
    } else {
      r[0] = ninv * (M * Q[0] + Pn * Q[2] - Pb * Q[3]);
      r[1] = ninv * (M * Q[1] + Pp * Q[3] - Pa * Q[2]);
      r[2] = ninv * (M * Q[2] + Pp * Q[0] + Pb * Q[1]);
      r[3] = ninv * (M * Q[3] + Pa * Q[0] + Pn * Q[1]);
    }
  }
  if (v == 0 && !nomom) o[0] = w[0], o[1] = w[1];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[2 + hp_slot(k, nv, v)] = r[k];
}

template <class P>
MF_DEV void hp_unit_finish(const HpTerm& t, const HpWorkItem& it, const cxd Q[4], const double* par, cxd* ev_e) {
  cxd* o = ev_e + t.out_off;
  const int nv = t.out_nv, v = it.v;
  if (t.finish == HP_F_NONE) {  // a vertex numerator (pair object): components only
#pragma unroll
    for (int k = 0; k < 4; ++k) o[hp_slot(k, nv, v)] = Q[k];
    return;
  }
  // a current: momentum slots = the sum of its inputs', propagator 1 / (P^2 - M(M - iW)) with P = -(sum)
  cxd w[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    w[k] = ev_e[t.in_off[0] + k] + ev_e[t.in_off[1] + k];
    if (t.nin > 2) w[k] += ev_e[t.in_off[2] + k];
  }
  const double M = t.mass_idx < 0 ? 0.0 : par[t.mass_idx];
  const double W = t.width_idx < 0 ? 0.0 : par[t.width_idx];
  hp_finish_store(t.finish & 3, (t.finish & HP_F_NOMOM) != 0, Q, w, M, W, o, nv, v);
}

// The units [begin, begin + n) of a phase for the E events of the block; thread w = tid, tid + nthreads, .. takes unit
// w / E for event w % E.  SPLIT: the phase holds units split over neighbouring lanes (stride E); all lanes of a warp
// then run the same number of trips and take part in the shuffles.  On the host the parts of a group are simply
// evaluated one after the other.
template <class P, bool SPLIT>
MF_DEV void hp_units(int begin, int n, int tid, int nthreads, const double* par, const cxd* coup, cxd* ev) {
  constexpr int E = P::HP_E, EVS = P::HP_EVSTRIDE;
  const int total = n * E;
#ifdef __CUDA_ARCH__
  const int bound = SPLIT ? ((total + 31) & ~31) : total;
#pragma unroll 1
  for (int w = tid; w < bound; w += nthreads) {
    const int ii = w / E, e = w - ii * E;
    const unsigned unit = w < total ? P::unit(begin + ii) : 0u;
    cxd Q[4] = {mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0)};
    HpTerm t;
    HpWorkItem it;
    hp_unit_terms<P>(unit, coup + e * P::NCOUP, ev + e * EVS, Q, t, it);
    if constexpr (SPLIT) {
      // partial sums of a group of 2^glog lanes (stride E) -> its first lane, always in the order of a binary tree
      const int glog = (int)((unit >> 28) & 3u), g = 1 << glog, part = ii & (g - 1);
      const int gmax = __reduce_max_sync(0xffffffffu, glog);
#pragma unroll 1
      for (int r = 0; r < gmax; ++r) {
        const int o = 1 << r;
        const bool take = r < glog && (part & ((o << 1) - 1)) == 0;   // this lane adds the lane `o` parts further on
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double re = __shfl_down_sync(0xffffffffu, Q[k].re, o * E);
          const double im = __shfl_down_sync(0xffffffffu, Q[k].im, o * E);
          if (take) Q[k].re += re, Q[k].im += im;
        }
      }
      if (part != 0) continue;
    }
    if ((unit >> 24) & 0xfu) hp_unit_finish<P>(t, it, Q, par, ev + e * EVS);
  }
#else
  for (int w = tid; w < total; w += nthreads) {
    const int ii = w / E, e = w - ii * E;
    const unsigned unit = P::unit(begin + ii);
    const int g = 1 << ((unit >> 28) & 3u);
    if (ii & (g - 1)) continue;   // a part of a group: evaluated with its first unit
    cxd Q[4] = {mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0)};
    HpTerm t;
    HpWorkItem it;
    for (int pp = g - 1; pp >= 0; --pp) hp_unit_terms<P>(P::unit(begin + ii + pp), coup + e * P::NCOUP, ev + e * EVS, Q, t, it);
    hp_unit_finish<P>(t, it, Q, par, ev + e * EVS);
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// Packed units ("SLU", codegen.hp_config SLU).  The table-driven routine above spends two thirds of its ~270 warp
// instructions per term on fetching and decoding table rows, on address arithmetic and on per-lane type switches.
// Here the objects of a phase are sorted into CLASSES = (sequence of vertex kinds, kind of propagator); the units of
// a phase are packed into warp TRIPS of one class each, statically assigned to the warps of the block (longest
// processing time first), so that everything about the kind of work is warp-uniform and comes from one descriptor
// per trip in constant memory.  What is left per lane is one 64-bit word per (unit, term) and one per unit, read
// coalesced by the lanes of the trip:
//   input descriptor (22 bits) = block offset in the event area (14) | helicity variant (5) << 14 | log2 variants (3) << 19
//   term word   .x = input A | f-index << 22      .y = input B            f-index = coupling * 4 + phase code (31: null term)
//   4-gluon 2nd .x = input C                      .y = (ca + 2) | (cb + 2) << 3 | (cc + 2) << 6
//   unit word   .x = output descriptor (offset = the block start, or start - 2 without momentum slots)
//   trip        .x = first word   .y = units in use | HpFinish << 8 | (mass + 1) << 12 | (width + 1) << 16 | terms << 20
//               .z = vertex kinds, 3 bits per term     .w = log2(parts per unit)
// f = -i COUP phase is looked up in a per-event table of the 4 phases of every coupling (shared memory).
// The code stays a LOOP over the terms with a warp-uniform switch over the five vertex kinds: one function per
// class (the first version, profiles/r02w_*) saved the same instructions but ran out of the instruction cache --
// 14 400 instructions, 24 % of the stall samples "no instruction" -- and gained nothing.
struct SluIn {
  const cxd* blk;   // the block's momentum slots; components from blk + 2
  int idx[4];       // position of the variant's component k behind blk + 2
};
MF_DEV SluIn slu_in(unsigned d, const cxd* ev_e) {
  const int off = (int)(d & 0x3fffu), v = (int)((d >> 14) & 31u), lg = (int)((d >> 19) & 7u), mask = (1 << lg) - 1;
  SluIn r;
  r.blk = ev_e + off;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.idx[k] = (k << lg) + (v ^ ((2 * k) & mask));
  return r;
}
MF_DEV void slu_load(const SluIn& in, cxd a[6]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) a[2 + k] = in.blk[2 + in.idx[k]];
}
MF_DEV void slu_load_mom(const SluIn& in, cxd a[6]) { a[0] = in.blk[0], a[1] = in.blk[1]; }


// one term of kind KIND (HpVertex) of one unit: Q += f * numerator; with want_mom also the sum of the inputs'
// momentum slots (a current's own momentum).  PF (tables too large for L1, g g > t t~ g g g: +2.6 %; g g > t t~ g g:
// -3 %): tw / tw2 = the next two words of the unit, already in registers; the words after them are fetched here, before
// the arithmetic, so that they arrive from L2 while it runs.
template <int KIND, bool PF>
MF_DEV void slu_term(const uint2* w, int stride, int& j, uint2& tw, uint2& tw2, const cxd* ev_e, const cxd* ftab_e, cxd Q[4], cxd mw[2],
                     bool want_mom) {
  uint2 t, t2 = make_uint2(0u, 0u);
  if (PF) {
    t = tw, t2 = tw2;
    if (KIND == HP_Q_VVVV) {
      tw = w[(j + 2) * stride], tw2 = w[(j + 3) * stride];
      j += 2;
    } else {
      tw = tw2, tw2 = w[(j + 2) * stride];
      ++j;
    }
  } else {
    t = w[j * stride];
    ++j;
    if (KIND == HP_Q_VVVV) {
      t2 = w[j * stride];
      ++j;
    }
  }
  const SluIn ia = slu_in(t.x & 0x3fffffu, ev_e), ib = slu_in(t.y & 0x3fffffu, ev_e);
  const cxd f = ftab_e[(t.x >> 22) & 31u];
  cxd a[6], b[6];
  slu_load(ia, a);
  slu_load(ib, b);
  if (KIND == HP_Q_VVV || want_mom) {
    slu_load_mom(ia, a);
    slu_load_mom(ib, b);
    mw[0] = a[0] + b[0], mw[1] = a[1] + b[1];
  }
  if (KIND == HP_Q_ROW) hp_q_row(a, b, f, Q);
  if (KIND == HP_Q_COL) hp_q_col(a, b, f, Q);
  if (KIND == HP_Q_CUR) hp_q_cur(a, b, f, Q);
  if (KIND == HP_Q_VVV) hp_q_vvv(a, b, f, Q);
  if (KIND == HP_Q_VVVV) {
    const SluIn ic = slu_in(t2.x & 0x3fffffu, ev_e);
    cxd c[6];
    slu_load(ic, c);
    if (want_mom) {
      slu_load_mom(ic, c);
      mw[0] += c[0], mw[1] += c[1];
    }
    const double ca = (double)((int)(t2.y & 7u) - 2), cb = (double)((int)((t2.y >> 3) & 7u) - 2), cc = (double)((int)((t2.y >> 6) & 7u) - 2);
    hp_q_vvvv_merged(a, b, c, ca, cb, cc, f, Q);
  }
}

// f-table of an event: SLU_NF entries, [coupling * 4 + phase code] = -i COUP x {1, -1, i, -i}; the rest, in particular
// the last one (the f-index of a null term), is zero
constexpr int SLU_NF = 32;
MF_DEV void hp_fill_ftab(cxd coup, cxd* f4) {
  const cxd f = mul_mi(coup);
  f4[0] = f, f4[1] = -f, f4[2] = mul_i(f), f4[3] = -mul_i(f);
}

// the terms of one lane's share of a unit: `kinds` = 3 bits per term (warp-uniform), words from w[stride], w[2 stride], ..
template <bool PF>
MF_DEV void slu_terms(const uint2* w, int stride, unsigned kinds, int nterms, bool want_mom, const cxd* ev_e, const cxd* ftab_e,
                      cxd Q[4], cxd mw[2]) {
  int j = 1;   // PF: tw = word j, tw2 = word j + 1 (the table ends with padding: reading ahead never leaves it)
  uint2 tw = make_uint2(0u, 0u), tw2 = tw;
  if (PF) tw = w[stride], tw2 = w[2 * stride];
#pragma unroll 1
  for (int q = 0; q < nterms; ++q, kinds >>= 3) {
    const bool wm = q == 0 && want_mom;   // the momentum slots come with the first term
    switch (kinds & 7u) {
      case HP_Q_ROW: slu_term<HP_Q_ROW, PF>(w, stride, j, tw, tw2, ev_e, ftab_e, Q, mw, wm); break;
      case HP_Q_COL: slu_term<HP_Q_COL, PF>(w, stride, j, tw, tw2, ev_e, ftab_e, Q, mw, wm); break;
      case HP_Q_CUR: slu_term<HP_Q_CUR, PF>(w, stride, j, tw, tw2, ev_e, ftab_e, Q, mw, wm); break;
      case HP_Q_VVV: slu_term<HP_Q_VVV, PF>(w, stride, j, tw, tw2, ev_e, ftab_e, Q, mw, wm); break;
      default: slu_term<HP_Q_VVVV, PF>(w, stride, j, tw, tw2, ev_e, ftab_e, Q, mw, wm); break;
    }
  }
}

// The trips of phase `ph` that belong to `warp`.  Proc::slu_range(ph * warps + warp) = first and last + 1 trip.  A
// unit with many terms is split into g = 2^k PARTS evaluated by neighbouring lanes (every part the same sequence of
// vertex kinds, padded with null terms whose f-index points at a zero) and added up with warp shuffles in the fixed
// order of a binary tree; part 0 applies the propagator and stores.  lane = ((unit * g) + part) * E + event; word j of
// (unit, part) at [first word + j * (32/E) + unit * g + part]; lanes beyond the trip's units repeat its first unit
// (they take part in the shuffles) and store nothing.
template <class P>
MF_DEV void slu_units(int ph, int warp, int lane, const cxd* ftab, const double* par, cxd* ev) {
  constexpr int E = P::HP_E, LPU = 32 / E, NW = P::HP_THREADS / 32;
  static_assert(32 % E == 0, "events per block divide the warp");
  const int lu = lane / E, e = lane - lu * E;
  const cxd* ftab_e = ftab + e * SLU_NF;
  cxd* ev_e = ev + e * P::HP_EVSTRIDE;
  const int2 r = P::slu_range(ph * NW + warp);
#pragma unroll 1
  for (int t = r.x; t < r.y; ++t) {
    const uint4 d = P::slu_trip(t);
    const int glog = P::HP_SLU_SPLIT ? (int)d.w : 0, part = lu & ((1 << glog) - 1);   // HP_SLU_SPLIT false: no trip is split
    const bool active = (lu >> glog) < (int)(d.y & 0xffu);
    const int fin = (int)((d.y >> 8) & 7u), nterms = (int)(d.y >> 20);
    cxd Q[4] = {mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0), mk(0.0, 0.0)};
    cxd mw[2] = {mk(0.0, 0.0), mk(0.0, 0.0)};
#ifdef __CUDA_ARCH__
    const uint2* w = P::slu_words() + d.x + lu;
    // the unit's header word: with tables beyond L1 fetched before the terms (needed right after them; g g > t t~ g g g
    // +1.1 %), else after them (g g > t t~ g g: one more live register through the terms costs 2 %)
    const unsigned od_early = P::HP_SLU_PREFETCH ? w[0].x : 0u;
    slu_terms<P::HP_SLU_PREFETCH>(w, LPU, d.z, nterms, fin != HP_F_NONE && part == 0, ev_e, ftab_e, Q, mw);
    if constexpr (P::HP_SLU_SPLIT) {
#pragma unroll 1
      for (int o = (1 << glog) >> 1; o > 0; o >>= 1) {   // lane p adds lane p + o: the sum arrives in part 0
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          Q[k].re += __shfl_down_sync(0xffffffffu, Q[k].re, o * E);
          Q[k].im += __shfl_down_sync(0xffffffffu, Q[k].im, o * E);
        }
      }
    }
    if (!active || part != 0) continue;
#else
    // on the host the lane of part 0 evaluates all parts in turn and adds them in the order of the device's tree
    if (!active || part != 0) continue;
    const uint2* w = P::slu_words() + d.x + lu;
    // the unit's header word: with tables beyond L1 fetched before the terms (needed right after them; g g > t t~ g g g
    // +1.1 %), else after them (g g > t t~ g g: one more live register through the terms costs 2 %)
    const unsigned od_early = P::HP_SLU_PREFETCH ? w[0].x : 0u;
    cxd Qp[8][4];
    for (int pp = 0; pp < (1 << glog); ++pp) {
      for (int k = 0; k < 4; ++k) Qp[pp][k] = mk(0.0, 0.0);
      cxd mwp[2];
      slu_terms<P::HP_SLU_PREFETCH>(w + pp, LPU, d.z, nterms, fin != HP_F_NONE && pp == 0, ev_e, ftab_e, Qp[pp], pp == 0 ? mw : mwp);
    }
    for (int o = (1 << glog) >> 1; o > 0; o >>= 1)
      for (int pp = 0; pp < o; ++pp)
        for (int k = 0; k < 4; ++k) Qp[pp][k] += Qp[pp + o][k];
    for (int k = 0; k < 4; ++k) Q[k] = Qp[0][k];
#endif
    const unsigned od = P::HP_SLU_PREFETCH ? od_early : w[0].x;
    cxd* o = ev_e + (od & 0x3fffu);
    const int v = (int)((od >> 14) & 31u), nv = 1 << ((od >> 19) & 7u);
    if (fin == HP_F_NONE) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[hp_slot(k, nv, v)] = Q[k];
    } else {
      const int mi = (int)((d.y >> 12) & 15u) - 1, wi = (int)((d.y >> 16) & 15u) - 1;
      hp_finish_store(fin & 3, (fin & HP_F_NOMOM) != 0, Q, mw, mi < 0 ? 0.0 : par[mi], wi < 0 ? 0.0 : par[wi], o, nv, v);
    }
  }
}

// vtab[mask * NCOMB + h]: the helicity variant, of an object over the leg set `mask`, that belongs to
// helicity combination h (= the bits of h at the positions in mask, packed).  Filled once per block.
template <class P>
MF_DEV void hp_fill_vtab(int idx, unsigned char* vtab) {
  const int mask = idx / P::NCOMB, h = idx - mask * P::NCOMB;
  int out = 0, pos = 0;
  for (int b = 0; b < P::NEXT; ++b)
    if (mask & (1 << b)) {
      out |= ((h >> b) & 1) << pos;
      ++pos;
    }
  vtab[idx] = (unsigned char)out;
}

// wavefunction w as helicity combination h sees it (used by the straight-line flavour)
template <class P>
MF_DEV void hp_load_amp(const cxd* wf_e, const unsigned char* vtab, int h, int w, cxd out[6]) {
  const HpWf d = P::wf(w);
  hp_load(wf_e + d.off, d.nv, vtab[d.legs * P::NCOMB + h], out);
}

// D += A(8x4) B(4x8) on the FP64 tensor cores.  Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// {d0, d1} = D[lane/4][2*(lane%4) + {0, 1}]  (layout checked on the device by tools/ubench.cu)
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
#ifdef __CUDA_ARCH__
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
#endif
}

// explicit shared-space accesses for the amplitude phase (byte address in the shared window)
__device__ __forceinline__ cxd lds_cxd(unsigned addr) {
  cxd v = {0.0, 0.0};
#ifdef __CUDA_ARCH__
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(addr));
#endif
  return v;
}
__device__ __forceinline__ void sts_cxd(unsigned addr, double re, double im) {
#ifdef __CUDA_ARCH__
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(re), "d"(im) : "memory");
#endif
}

// NT tiles of the amplitude phase for the E events of the block, executed by a full warp: the
// descriptors are read first (warp-uniform), then all fragment loads are issued, then the NT*E*4
// tensor-core instructions, then the stores -- NT*E independent tiles in flight per warp.
// `evs` = shared-window byte address of the block's event areas.
struct HpTileWords {
  uint4 w0;  // q, x, qnv, xnv | q0, x0, qvalid, xvalid | slot
  uint4 w1;  // rowh[8] | colh[8]
};
__device__ __forceinline__ HpTileWords hp_tile_words(const HpTile* t) {
  const uint4* tp = reinterpret_cast<const uint4*>(t);
  return HpTileWords{tp[0], tp[1]};
}

template <class P, int NT>
__device__ __forceinline__ void hp_mma_tiles(const HpTileWords (&tile)[NT], int lane, unsigned evs) {
  constexpr int E = P::HP_E;
  constexpr unsigned EVB = P::HP_EVSTRIDE * 16u;
  const int r = lane >> 2, k = lane & 3;
  unsigned qi[NT], xi[NT], d0[NT], d1[NT];
  bool s0[NT], s1[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const uint4 w0 = tile[t].w0, w1 = tile[t].w1;
    const int q = w0.x & 0xffff, x = w0.x >> 16, qnv = w0.y & 0xffff, xnv = w0.y >> 16;
    const int q0 = w0.z & 0xff, x0 = (w0.z >> 8) & 0xff, qvalid = (w0.z >> 16) & 0xff, xvalid = w0.z >> 24;
    const int slot = w0.w & 0xffff;
    qi[t] = evs + 16u * (q + hp_slot(k, qnv, (q0 + r) & (qnv - 1)));
    xi[t] = evs + 16u * (x + hp_slot(k, xnv, (x0 + r) & (xnv - 1)));
    const int hq = (((r & 4) ? w1.y : w1.x) >> (8 * (r & 3))) & 0xff;
    const unsigned hc = (((k & 2) ? w1.w : w1.z) >> (16 * (k & 1))) & 0xffff;
    d0[t] = evs + 16u * (P::HP_WFSIZE + P::HP_SCRATCH + slot * P::HP_NHP + hp_abuf_pos(hq | (hc & 0xff)));
    d1[t] = evs + 16u * (P::HP_WFSIZE + P::HP_SCRATCH + slot * P::HP_NHP + hp_abuf_pos(hq | (hc >> 8)));
    s0[t] = r < qvalid && 2 * k < xvalid, s1[t] = r < qvalid && 2 * k + 1 < xvalid;
  }
  cxd qa[NT][E], xb[NT][E];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int e = 0; e < E; ++e) qa[t][e] = lds_cxd(qi[t] + e * EVB), xb[t][e] = lds_cxd(xi[t] + e * EVB);
  double cr0[NT][E], cr1[NT][E], ci0[NT][E], ci1[NT][E];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      cr0[t][e] = cr1[t][e] = ci0[t][e] = ci1[t][e] = 0.0;
      dmma_m8n8k4(cr0[t][e], cr1[t][e], qa[t][e].re, xb[t][e].re);
      dmma_m8n8k4(ci0[t][e], ci1[t][e], qa[t][e].re, xb[t][e].im);
    }
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      dmma_m8n8k4(cr0[t][e], cr1[t][e], -qa[t][e].im, xb[t][e].im);
      dmma_m8n8k4(ci0[t][e], ci1[t][e], qa[t][e].im, xb[t][e].re);
    }
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (s0[t]) sts_cxd(d0[t] + e * EVB, cr0[t][e], ci0[t][e]);
      if (s1[t]) sts_cxd(d1[t] + e * EVB, cr1[t][e], ci1[t][e]);
    }
}

// the same tile on the CPU (tests/hostcheck): plain loops over its rows and columns
template <class P>
inline void hp_mma_tile_host(const HpTile* tp, const cxd* ev_e, cxd* abuf_e) {
  for (int r = 0; r < tp->qvalid; ++r)
    for (int c = 0; c < tp->xvalid; ++c) {
      cxd amp = mk(0.0, 0.0);
      for (int k = 0; k < 4; ++k)
        amp = fma_c(ev_e[tp->q + hp_slot(k, tp->qnv, tp->q0 + r)], ev_e[tp->x + hp_slot(k, tp->xnv, tp->x0 + c)], amp);
      abuf_e[tp->slot * P::HP_NHP + hp_abuf_pos(tp->rowh[r] | tp->colh[c])] = amp;
    }
}

// ------------------------------------------------------------------------------------------------
// JAMP accumulators in Tensor Memory.  A thread's JAMPs (HP_NJ complex numbers) are live from the first batch of a
// helicity pass to the colour contraction, but only touched in the short JAMP phase of every batch; kept in registers
// they take 60-96 of a thread's registers away from the current / pair / tile phases (g g > t t~ g g g, 128 registers
// per thread: 5 KB of spills per thread, all of them in the pair phase).  Blackwell's TMEM (256 KB per SM, 128 lanes x
// 512 columns of 32 bits) is not used otherwise -- there is no FP64 tcgen05.mma -- so the accumulators are parked there
// between JAMP phases: tcgen05.st after a batch, tcgen05.ld before the next.  Shape 32x32b: lane i of a warp owns TMEM
// lane 32 * (warp % 4) + i, its words are consecutive columns; warps that share a lane quarter use different columns.
template <int NWORDS>
__device__ __forceinline__ void tmem_store_words(unsigned taddr, const unsigned (&r)[NWORDS]) {
  static_assert(NWORDS % 16 == 0, "multiples of 16 words");
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int c = 0; c < NWORDS; c += 16)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr + c), "r"(r[c]), "r"(r[c + 1]), "r"(r[c + 2]), "r"(r[c + 3]), "r"(r[c + 4]), "r"(r[c + 5]),
                 "r"(r[c + 6]), "r"(r[c + 7]), "r"(r[c + 8]), "r"(r[c + 9]), "r"(r[c + 10]), "r"(r[c + 11]), "r"(r[c + 12]),
                 "r"(r[c + 13]), "r"(r[c + 14]), "r"(r[c + 15])
                 : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
}
template <int NWORDS>
__device__ __forceinline__ void tmem_load_words(unsigned taddr, unsigned (&r)[NWORDS]) {
  static_assert(NWORDS % 16 == 0, "multiples of 16 words");
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int c = 0; c < NWORDS; c += 16)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[c]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]), "=r"(r[c + 6]),
                   "=r"(r[c + 7]), "=r"(r[c + 8]), "=r"(r[c + 9]), "=r"(r[c + 10]), "=r"(r[c + 11]), "=r"(r[c + 12]),
                   "=r"(r[c + 13]), "=r"(r[c + 14]), "=r"(r[c + 15])
                 : "r"(taddr + c)
                 : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#endif
}
// words per thread (padded to a multiple of 16) and columns per block (a power of two >= 32) of the JAMP store
template <class P>
struct HpTmem {
  static constexpr int NWORDS = ((4 * P::HP_NJ + 15) / 16) * 16;
  static constexpr int WARPS_PER_QUARTER = (P::HP_THREADS / 32 + 3) / 4;
  static constexpr int NEED = NWORDS * WARPS_PER_QUARTER;
  static constexpr int NCOLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
  static_assert(NEED <= 512, "the JAMPs of a block do not fit in Tensor Memory");
};
template <class P>
__device__ __forceinline__ void hp_jamp_park(unsigned tmem_base, const cxd (&J)[P::HP_NJ]) {
#ifdef __CUDA_ARCH__
  constexpr int NW = HpTmem<P>::NWORDS;
  const int warp = threadIdx.x >> 5;
  unsigned r[NW];
#pragma unroll
  for (int j = 0; j < NW / 4; ++j) {
    const cxd v = j < P::HP_NJ ? J[j] : mk(0.0, 0.0);
    r[4 * j] = (unsigned)__double2loint(v.re), r[4 * j + 1] = (unsigned)__double2hiint(v.re);
    r[4 * j + 2] = (unsigned)__double2loint(v.im), r[4 * j + 3] = (unsigned)__double2hiint(v.im);
  }
  tmem_store_words<NW>(tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * NW), r);
#endif
}
template <class P>
__device__ __forceinline__ void hp_jamp_fetch(unsigned tmem_base, cxd (&J)[P::HP_NJ]) {
#ifdef __CUDA_ARCH__
  constexpr int NW = HpTmem<P>::NWORDS;
  const int warp = threadIdx.x >> 5;
  unsigned r[NW];
  tmem_load_words<NW>(tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * NW), r);
#pragma unroll
  for (int j = 0; j < P::HP_NJ; ++j)
    J[j] = mk(__hiloint2double((int)r[4 * j + 1], (int)r[4 * j]), __hiloint2double((int)r[4 * j + 3], (int)r[4 * j + 2]));
#endif
}

// Optional phase timers (-DMF_HP_PROFILE, tools/profile_phases.py): SM cycles spent by each block in
// externals / currents / pair objects / amplitudes / JAMP / colour+reduction, summed over blocks.
#ifdef MF_HP_PROFILE
__device__ unsigned long long g_hp_prof[8];
#define MF_PROF_DECL long long prof_t = clock64();
#define MF_PROF(slot)                                                   \
  do {                                                                  \
    if (threadIdx.x == 0) {                                             \
      const long long now_ = clock64();                                 \
      atomicAdd(&g_hp_prof[slot], (unsigned long long)(now_ - prof_t)); \
      prof_t = now_;                                                    \
    }                                                                   \
  } while (0)
#else
#define MF_PROF_DECL
#define MF_PROF(slot)
#endif

// ------------------------------------------------------------------------------------------------
// Table-driven colour contraction  sum_ab J_a* cf_ab J_b  of one helicity pass.  The JAMPs of the pass sit in
// shared memory as two planes (Re, Im) of HP_NCP rows (colours, padded to a multiple of 8) x HP_PLANE doubles
// (helicity combinations + 4: the row stride = 4 mod 16 keeps the tensor-core fragment loads conflict free).
// cfsym is the colour matrix symmetrised in 8x8 blocks (0 below the block diagonal, cf on it, 2 cf above), so
// that  sum_ab J_a cf_ab J_b = sum_a J_a (cfsym J)_a  with about half the products.
//
// Tensor-core flavour: a warp owns (plane, 8 helicity combinations) = one column tile of the HP_NCP x (2 HP_NHP)
// matrix Z = cfsym [Re J | Im J]; its B fragments (all colours of its columns) stay in registers, the A
// fragments (cfsym, the same for every block and event) come through L1, each 8x8 tile of Z costs up to
// HP_NCP/4 mma.sync.m8n8k4.f64 and is contracted with J on the spot.
template <class P>
__device__ __forceinline__ double hp_colour_mma(const double* planes /* shared: [2][NCP][PLANE] */, int tid, int only_hl) {
  constexpr int NCP = P::HP_NCP, PL = P::HP_PLANE, NHP = P::HP_NHP, T = P::HP_THREADS;
  constexpr int NKK = NCP / 4, NU = 2 * (NHP / 8);
  const int warp = tid >> 5, lane = tid & 31, r = lane >> 2, k = lane & 3;
  const double* cf = P::cfsym();
  double me = 0.0;
#pragma unroll 1
  for (int u = warp; u < NU; u += T / 32) {
    const int pl = u / (NHP / 8), j = u - pl * (NHP / 8);
    const double* plane = planes + pl * NCP * PL;
    double B[NKK];
#pragma unroll
    for (int kk = 0; kk < NKK; ++kk) B[kk] = plane[(4 * kk + k) * PL + 8 * j + r];
    const int h0 = 8 * j + 2 * k;
    const bool keep0 = only_hl < 0 || only_hl == h0, keep1 = only_hl < 0 || only_hl == h0 + 1;
#pragma unroll 1
    for (int i = 0; i < NCP / 8; ++i) {
      double c0 = 0.0, c1 = 0.0;
      if constexpr (P::HP_CF_FRAG) {   // A fragments of two k-blocks with one coalesced 16-byte load per lane
        const double2* af = reinterpret_cast<const double2*>(P::cfsym_frag()) + (size_t)i * (NKK / 2) * 32 + lane;
#pragma unroll
        for (int kk2 = 0; kk2 < NKK / 2; ++kk2)
          if (kk2 >= i) {
            const double2 a2 = __ldg(af + kk2 * 32);
            dmma_m8n8k4(c0, c1, a2.x, B[2 * kk2]);
            dmma_m8n8k4(c0, c1, a2.y, B[2 * kk2 + 1]);
          }
      } else {
        const double* arow = cf + (8 * i + r) * NCP + k;
#pragma unroll
        for (int kk = 0; kk < NKK; ++kk)
          if (kk >= 2 * i) dmma_m8n8k4(c0, c1, __ldg(arow + 4 * kk), B[kk]);
      }
      const double* jr = plane + (8 * i + r) * PL + h0;
      if (keep0) me += jr[0] * c0;
      if (keep1) me += jr[1] * c1;
    }
  }
  return me / P::HP_COLOUR_DENOM;
}

// CUDA-core flavour of the same contraction (the A/B partner of hp_colour_mma): thread (helicity combination,
// colour group) takes the rows of its own colours
template <class P>
MF_DEV double hp_colour_loop(const double* planes, int hl, int cg, const double* cf) {
  constexpr int NCP = P::HP_NCP, PL = P::HP_PLANE;
  double me = 0.0;
  const int lo = cg * P::HP_NJ, hi = (lo + P::HP_NJ < P::NCOLOR) ? lo + P::HP_NJ : P::NCOLOR;
  for (int a = lo; a < hi; ++a) {
    double tr = 0.0, ti = 0.0;
    for (int b = 8 * (a / 8); b < NCP; ++b) {
      const double w = cf[a * NCP + b];
      tr += w * planes[b * PL + hl];
      ti += w * planes[(NCP + b) * PL + hl];
    }
    me += planes[a * PL + hl] * tr + planes[(NCP + a) * PL + hl] * ti;
  }
  return me / P::HP_COLOUR_DENOM;
}

// ------------------------------------------------------------------------------------------------
// The E-event matrix-element evaluation used by both kernels.  `mom` [E][NEXT][4], `coup` [E][NCOUP]
// and the event areas `ev` [E][HP_EVSTRIDE] live in shared memory; returns, in the first thread of
// every event, the event's |M|^2 summed over helicities and colours and averaged (other threads
// return garbage).  Thread tid = (e * HP_NCG + colour group) * HP_NHP + helicity combination of the pass.
template <class P>
__device__ __forceinline__ double hp_smatrix_block(int nev /* <= E valid events */, const double* mom,
                                                   const cxd* coup, const cxd* ftab /* [E][NCOUP][4 phases], SLU */,
                                                   const double* par, double sqh, cxd* ev,
                                                   const unsigned char* vtab, double* red /* [T/32] */, int only_h,
                                                   unsigned tmem_base = 0u) {
  constexpr int E = P::HP_E, NHP = P::HP_NHP, NCG = P::HP_NCG, TE = NHP * NCG, T = E * TE, EVS = P::HP_EVSTRIDE;
  const int tid = threadIdx.x;
  MF_PROF_DECL
  if constexpr (2 * E <= 32 && P::NEXT <= 32) {
    // leg l runs in warp l % NW, lanes [l / NW * 2E, +2E): the legs sharing a warp are mostly of one
    // kind, so the vector / spinor routines of different legs run side by side instead of in turn
    constexpr int NW = T / 32, PER = 32 / (2 * E);
    const int warp = tid >> 5, lane = tid & 31;
    for (int l0 = 0; l0 < P::NEXT; l0 += NW * PER) {
      const int leg = l0 + warp + NW * (lane / (2 * E));
      if (leg < P::NEXT && lane < PER * 2 * E) hp_externals<P>(leg * 2 * E + lane % (2 * E), E, mom, par, sqh, ev);
    }
  } else {
    for (int it = tid; it < P::NEXT * E * 2; it += T) hp_externals<P>(it, E, mom, par, sqh, ev);
  }
  __syncthreads();
  MF_PROF(0);
#pragma unroll 1
  for (int L = 2; L <= P::HP_MAXLEVEL; ++L) {
    // units (object, variant) of this level: the currents and the pair objects that stay in shared memory
    if constexpr (P::HP_SLU)
      slu_units<P>(L - 2, tid >> 5, tid & 31, ftab, par, ev);
    else
      hp_units<P, P::HP_SPLIT>(P::level_begin(L), P::level_begin(L + 1) - P::level_begin(L), tid, T, par, coup, ev);
    __syncthreads();
  }
  MF_PROF(1);
  const int e = tid / TE, rr = tid - e * TE, cg = rr / NHP, h = rr - cg * NHP;
  const cxd* wf_e = ev + e * EVS;
  double me = 0.0;
  if constexpr (P::HP_UNROLL) {
    // short amplitude lists: straight-line code, whole vertices per helicity combination
    me = P::hp_amps_unrolled(wf_e, vtab, h, coup + e * P::NCOUP);
    if (only_h >= 0) me = (h == only_h) ? me : 0.0;
  } else {
    cxd* abuf_h = ev + e * EVS + P::HP_WFSIZE + P::HP_SCRATCH + hp_abuf_pos(h);
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll 1
    for (int pass = 0; pass < P::HP_NPASS; ++pass) {
      cxd J[P::HP_NJ];
#pragma unroll
      for (int j = 0; j < P::HP_NJ; ++j) J[j] = mk(0.0, 0.0);
#pragma unroll 1
      for (int bi = 0; bi < P::HP_NBATCH; ++bi) {
        const HpBatch bt = P::batch(pass * P::HP_NBATCH + bi);
#ifdef MF_HP_EXPERIMENT_NOJ   // measurement only (wrong results): the JAMPs are not kept across the pair phase
#pragma unroll
        for (int j = 0; j < P::HP_NJ; ++j) J[j] = mk(0.0, 0.0);
#endif
        if constexpr (P::HP_SLU)
          slu_units<P>(P::HP_MAXLEVEL - 1 + pass * P::HP_NBATCH + bi, warp, lane, ftab, par, ev);
        else
          hp_units<P, P::HP_SPLIT>(bt.unit_begin, bt.unit_end - bt.unit_begin, tid, T, par, coup, ev);
        __syncthreads();   // pair objects complete; the JAMP reads of the batch before are done (amplitude buffer reused)
        MF_PROF(2);
        {
          // every warp takes MT tiles per trip; an index past the end repeats the batch's last tile
          // (same values stored twice) so that the trips stay straight-line
          constexpr int NW = T / 32, MT = P::HP_TILES_IN_FLIGHT;
          const unsigned evs = (unsigned)__cvta_generic_to_shared(ev);
          const int last = bt.tile_end - 1;
          // the descriptors of the next trip are fetched while the current trip computes
          HpTileWords cur[MT], nxt[MT];
          int w = bt.tile_begin + warp;
#pragma unroll
          for (int t = 0; t < MT; ++t) cur[t] = hp_tile_words(P::tile(w + t * NW < last ? w + t * NW : last));
#pragma unroll 1
          for (; w < bt.tile_end; w += MT * NW) {
            const int wn = w + MT * NW;
#pragma unroll
            for (int t = 0; t < MT; ++t) nxt[t] = hp_tile_words(P::tile(wn + t * NW < last ? wn + t * NW : last));
            hp_mma_tiles<P, MT>(cur, lane, evs);
#pragma unroll
            for (int t = 0; t < MT; ++t) cur[t] = nxt[t];
          }
        }
        __syncthreads();
        MF_PROF(3);
#ifdef __CUDA_ARCH__
#ifndef MF_HP_EXPERIMENT_TMEM_ALLOC_ONLY
        if constexpr (P::HP_TMEM_J) {   // the JAMPs were parked in Tensor Memory during the pair and tile phases
          if (bi > 0) {
            hp_jamp_fetch<P>(tmem_base, J);
          } else {   // defined on every path: nothing of J is live before this point
#pragma unroll
            for (int j = 0; j < P::HP_NJ; ++j) J[j] = mk(0.0, 0.0);
          }
        }
#endif
#endif
        P::jamp_batch(bi, cg, abuf_h, J);
#ifdef __CUDA_ARCH__
#ifndef MF_HP_EXPERIMENT_TMEM_ALLOC_ONLY
        if constexpr (P::HP_TMEM_J) {
          if (bi + 1 < P::HP_NBATCH) hp_jamp_park<P>(tmem_base, J);
        }
#endif
#endif
        MF_PROF(4);
      }
      // the helicity combination of this thread within the whole table, and the row asked for (if any)
      const int hfull = pass * NHP + h;
      const int only_hl = only_h < 0 ? -1 : (only_h / NHP == pass ? only_h % NHP : NHP);  // NHP: none in this pass
      if constexpr (P::HP_COLOUR == 0) {
        const double m = P::colour_sum(0, J, nullptr);
        me += (only_h < 0 || hfull == only_h) ? m : 0.0;
      } else if constexpr (P::HP_COLOUR == 1) {
        // exchange the JAMPs of the colour groups through the (now free) scratch + amplitude buffer area
        static_assert(P::HP_EVSTRIDE - P::HP_WFSIZE >= P::NCOLOR * NHP, "JAMP exchange area too small");
        cxd* jb = ev + e * EVS + P::HP_WFSIZE + h;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < P::HP_NJ; ++j)
          if (cg * P::HP_NJ + j < P::NCOLOR) jb[(cg * P::HP_NJ + j) * NHP] = J[j];
        __syncthreads();
        const double m = P::colour_sum(cg, J, jb);
        me += (only_h < 0 || hfull == only_h) ? m : 0.0;
        if (pass + 1 < P::HP_NPASS) __syncthreads();  // the next pass overwrites the exchange area
      } else {
        static_assert(P::HP_EVSTRIDE - P::HP_WFSIZE >= P::HP_NCP * P::HP_PLANE, "JAMP exchange area too small");
        static_assert(E == 1 || P::HP_COLOUR == 3, "the tensor-core colour contraction handles one event per block");
        double* planes = reinterpret_cast<double*>(ev + e * EVS + P::HP_WFSIZE);
        constexpr int NCP = P::HP_NCP, PL = P::HP_PLANE;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < P::HP_NJ; ++j)
          if (cg * P::HP_NJ + j < P::NCOLOR) {
            planes[(cg * P::HP_NJ + j) * PL + h] = J[j].re;
            planes[(NCP + cg * P::HP_NJ + j) * PL + h] = J[j].im;
          }
        for (int i = P::NCOLOR * NHP + rr; i < NCP * NHP; i += TE) {  // zero the padding colours
          planes[(i / NHP) * PL + i % NHP] = 0.0;
          planes[(NCP + i / NHP) * PL + i % NHP] = 0.0;
        }
        __syncthreads();
        if constexpr (P::HP_COLOUR == 2) {
          me += hp_colour_mma<P>(planes, tid, only_hl);
        } else {
          const double m = hp_colour_loop<P>(planes, h, cg, P::cfsym());
          me += (only_h < 0 || hfull == only_h) ? m : 0.0;
        }
        if (pass + 1 < P::HP_NPASS) __syncthreads();
      }
    }
  }
  if (e >= nev) me = 0.0;
  // sum over the helicities (and colour groups) of one event
  constexpr int W = TE < 32 ? TE : 32;
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) me += __shfl_down_sync(0xffffffffu, me, o, W);
  if (TE > 32) {
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) red[warp] = me;
    __syncthreads();
    if (rr == 0) {
      me = 0.0;
#pragma unroll
      for (int k = 0; k < TE / 32; ++k) me += red[e * (TE / 32) + k];
    }
  }
  MF_PROF(5);
  return only_h >= 0 ? me : me / P::DENOM;
}

template <class P>
struct HpSmatrixSmem {
  static constexpr int E = P::HP_E, T = P::HP_THREADS;
  double mom[E * P::NEXT * 4];
  cxd coup[E * (P::NCOUP > 0 ? P::NCOUP : 1)];
  cxd ftab[P::HP_SLU ? E * SLU_NF : 1];   // -i COUP x {1, -1, i, -i} per event and coupling, zero beyond (SLU)
  double red[T / 32 + 1];
  unsigned tmem_addr;   // base of the block's Tensor Memory allocation (HP_TMEM_J)
  unsigned char vtab[P::HP_UNROLL ? (1 << P::NEXT) * P::NCOMB : 16];  // straight-line flavour only
  // followed by the event areas: cxd ev[E][HP_EVSTRIDE] = wavefunctions | pair objects | amplitude buffer
};

template <class P>
__global__ void __launch_bounds__(P::HP_THREADS, P::HP_MINBLOCKS) smatrix_kernel_hp(const SmatrixArgs a) {
  constexpr int E = P::HP_E, TE = P::HP_NHP * P::HP_NCG, T = E * TE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HpSmatrixSmem<P>& s = *reinterpret_cast<HpSmatrixSmem<P>*>(smem_raw);
  cxd* evarea = reinterpret_cast<cxd*>(smem_raw + ((sizeof(HpSmatrixSmem<P>) + 15) / 16) * 16);
  const int tid = threadIdx.x;
  if constexpr (P::HP_UNROLL)
    for (int i = tid; i < (1 << P::NEXT) * P::NCOMB; i += T) hp_fill_vtab<P>(i, s.vtab);
  unsigned tmem_base = 0u;
  if constexpr (P::HP_TMEM_J) {   // Tensor Memory for the JAMP accumulators: warp 0 allocates, everybody reads the address
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s.tmem_addr)),
                   "n"(HpTmem<P>::NCOLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = s.tmem_addr;
  }
  int only_h = -1;
  if (a.only_comb >= 0) {
    only_h = 0;
    for (int j = 0; j < P::NEXT; ++j) only_h |= ((P::hel(a.only_comb, j) + 1) >> 1) << j;
  }
  // plain mode: one "segment" holding all nevt events, blocks stride over its groups of E events;
  // segmented mode: block b owns segments b, b + gridDim.x, ... and walks their valid events
  const bool segmented = a.seg_count != nullptr;
  const int nseg = segmented ? a.nseg : 1;
  // The inputs of the NEXT group of events (one momentum component and one coupling per thread) are fetched into
  // registers as soon as the current group's sit in shared memory, so that their DRAM latency runs under the whole
  // evaluation of the current group instead of in front of its first barrier.  Measured: g g > t t~ g g g (one event per
  // block of 16 warps) +9 %, g g > t t~ g g (two blocks per SM hide the latency already; 6 more registers) -1.3 %,
  // g g > t t~ g -2.5 %: per process (codegen.hp_config PREFIN).
  constexpr bool PREF = P::HP_PREFETCH_INPUTS && E * P::NEXT * 4 <= T && E * (P::NCOUP > 0 ? P::NCOUP : 1) <= T;
  auto load_mom = [&](long long ev0, int nev, int i) {
    const int e = i / (P::NEXT * 4), r = i - e * (P::NEXT * 4);
    const long long ev = ev0 + (e < nev ? e : 0);  // pad a partial group with a valid event
    return a.layout == MFP_LAYOUT_AOS ? a.p[ev * (P::NEXT * 4) + r] : a.p[(long long)r * a.nevt + ev];
  };
  auto load_coup = [&](long long ev0, int nev, int i) {   // alpha_s of the event (x), or its coupling (x, y)
    const int e = i / P::NCOUP, c = i - e * P::NCOUP;
    const long long ev = ev0 + (e < nev ? e : 0);
    if (a.alpha_s) return make_double2(a.alpha_s[ev], 0.0);
    return reinterpret_cast<const double2*>(a.coup)[a.coup_stride ? (long long)c * a.nevt + ev : c];
  };
  double mom_next = 0.0;
  double2 coup_next = make_double2(0.0, 0.0);
  bool have_next = false;
  for (int sg = segmented ? blockIdx.x : 0; sg < nseg; sg += segmented ? gridDim.x : 1) {
    const long long base = segmented ? (long long)sg * a.seg_size : 0;
    const long long cnt = segmented ? a.seg_count[sg] : a.nevt;
    const long long ngroups = (cnt + E - 1) / E;
    const long long gstep = segmented ? 1 : gridDim.x;
    for (long long g = segmented ? 0 : blockIdx.x; g < ngroups; g += gstep) {
      const long long ev0 = base + g * E;
      const int nev = (int)((base + cnt - ev0) < E ? (base + cnt - ev0) : E);
      if constexpr (PREF) {
        if (tid < E * P::NEXT * 4) s.mom[tid] = have_next ? mom_next : load_mom(ev0, nev, tid);
      } else {
        for (int i = tid; i < E * P::NEXT * 4; i += T) s.mom[i] = load_mom(ev0, nev, i);
      }
      for (int i = tid; i < E * P::NCOUP; i += T) {
        const int e = i / P::NCOUP, c = i - e * P::NCOUP;
        const double2 v = (PREF && have_next) ? coup_next : load_coup(ev0, nev, i);
        if (a.alpha_s) {  // couplings from alpha_s: G = 2 sqrt(pi alpha_s) (parameters.py:13-15)
          const double G = 2.0 * sqrt(M_PI * v.x);
          double gp = 1.0;
          for (int q = 0; q < P::coup_power(c); ++q) gp *= G;
          s.coup[i] = mk(P::coup_re(c) * gp, P::coup_im(c) * gp);
        } else {
          s.coup[i] = mk(v.x, v.y);
        }
        if constexpr (P::HP_SLU) hp_fill_ftab(s.coup[i], s.ftab + e * SLU_NF + 4 * c);
      }
      if constexpr (P::HP_SLU)
        for (int i = tid; i < E * SLU_NF; i += T)
          if (i % SLU_NF >= 4 * P::NCOUP) s.ftab[i] = mk(0.0, 0.0);
      if constexpr (PREF) {
        have_next = g + gstep < ngroups;   // within the segment; a new segment starts with a plain load
        if (have_next) {
          const long long ev0n = base + (g + gstep) * E;
          const int nevn = (int)((base + cnt - ev0n) < E ? (base + cnt - ev0n) : E);
          if (tid < E * P::NEXT * 4) mom_next = load_mom(ev0n, nevn, tid);
          if (tid < E * P::NCOUP) coup_next = load_coup(ev0n, nevn, tid);
        }
      }
      __syncthreads();
      const double me = hp_smatrix_block<P>(nev, s.mom, s.coup, s.ftab, a.par, a.sqh, evarea, s.vtab, s.red, only_h, tmem_base);
      const int e = tid / TE;
      if (tid - e * TE == 0 && e < nev) a.out[ev0 + e] = me;
      __syncthreads();
    }
  }
  if constexpr (P::HP_TMEM_J) {
    __syncthreads();
    if (tid < 32)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(HpTmem<P>::NCOLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
template <class P>
size_t hp_smatrix_smem() {
  return ((sizeof(HpSmatrixSmem<P>) + 15) / 16) * 16 + sizeof(cxd) * (size_t)P::HP_EVSTRIDE * P::HP_E;
}

template <class P>
int launch_smatrix_hp(const double* d_p, int layout, long long nevt, const double* par, const double* d_coup,
                      long long coup_stride, double sqh, double* d_out, int only_comb, cudaStream_t st) {
  if (nevt <= 0) return 0;
  if (layout != MFP_LAYOUT_AOS && layout != MFP_LAYOUT_SOA) return fail_msg("mfp_smatrix: unknown layout");
  if (P::NCOUP > 0 && d_coup == nullptr) return fail_msg("mfp_smatrix: couplings missing");
  SmatrixArgs a;
  a.p = d_p, a.layout = layout, a.nevt = nevt, a.coup = d_coup, a.coup_stride = coup_stride, a.sqh = sqh;
  a.out = d_out, a.only_comb = only_comb;
  a.alpha_s = nullptr, a.seg_count = nullptr, a.seg_size = 0, a.nseg = 0;
  for (int i = 0; i < MFP_MAX_PARAMS; ++i) a.par[i] = i < P::NPAR ? par[i] : 0.0;
  const size_t smem = hp_smatrix_smem<P>();
  cudaError_t e = cudaFuncSetAttribute(smatrix_kernel_hp<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail("smatrix_kernel_hp smem attribute", e);
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smatrix_kernel_hp<P>, P::HP_THREADS, smem);
  if (per_sm < 1) per_sm = 1;
  long long blocks = (nevt + P::HP_E - 1) / P::HP_E;
  const long long cap = (long long)sms * per_sm;
  if (blocks > cap) blocks = cap;
  smatrix_kernel_hp<P><<<(unsigned)blocks, P::HP_THREADS, smem, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("smatrix_kernel_hp launch", e);
  return 0;
}

// grid size override (mfp_set_integrand_blocks): libraries that work on the same events must cut their event
// buffers into the same segments
static int g_blocks_override = 0;

template <class P>
int integrand_blocks_hp() {
  if (g_blocks_override > 0) return g_blocks_override;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = hp_smatrix_smem<P>();
  cudaFuncSetAttribute(smatrix_kernel_hp<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smatrix_kernel_hp<P>, P::HP_THREADS, smem);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

// segments of the event buffer: a few per matrix-element block so that the blocks stay balanced
template <class P>
int hp_segments(int nblocks) { return nblocks * 4; }

template <class P>
long long integrand_workspace_hp(long long nevents) {
  constexpr int NDIM = 4 * (P::NEXT - 2) + 2;
  return (long long)event_buffer_layout(nevents, hp_segments<P>(integrand_blocks_hp<P>()), P::NEXT, NDIM, nullptr, nullptr);
}

// the segmentation the last launch cut this workspace into: the event view must be laid out the same way even if the
// caller's grid size differed from integrand_blocks_hp() or the override changed in between
static int g_last_nseg = 0;
static long long g_last_nevents = -1;
static const void* g_last_workspace = nullptr;

template <class P>
int integrand_events_hp(void* d_workspace, long long nevents, mfp_event_view* out) {
  constexpr int NDIM = 4 * (P::NEXT - 2) + 2;
  EventBuffer b;
  const bool same = g_last_nseg > 0 && g_last_nevents == nevents && g_last_workspace == d_workspace;
  event_buffer_layout(nevents, same ? g_last_nseg : hp_segments<P>(integrand_blocks_hp<P>()), P::NEXT, NDIM, d_workspace, &b);
  out->d_mom = b.mom, out->d_weight = b.w, out->d_me = b.me, out->d_alpha_s = b.as, out->capacity = b.cap;
  out->d_bins = b.bins;
  return 0;
}

// One pass of the integrand = three launches on the caller's stream (see pipeline_kernels.cuh)
template <class P>
int launch_integrand_hp(const mfp_integrand_args* u, cudaStream_t st) {
  constexpr int NDIM = 4 * (P::NEXT - 2) + 2;
  IntegrandArgs ia;
  if (int rc = prepare_integrand_args<P>(u, ia)) return rc;
  const int nseg = hp_segments<P>(u->nblocks);
  GenArgs g;
  g.u = ia.u, g.massive = ia.massive, g.shat_min = ia.shat_min, g.ps = ia.ps, g.cuts = ia.cuts;
  const size_t need = event_buffer_layout(u->nevents, nseg, P::NEXT, NDIM, u->d_workspace, &g.buf);
  if (!u->d_workspace || (size_t)u->workspace_bytes < need)
    return fail_msg("mfp_integrand: workspace missing or smaller than mfp_integrand_workspace(nevents)");
  g_last_nseg = nseg, g_last_nevents = u->nevents, g_last_workspace = u->d_workspace;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ggrid = nseg < sms * 8 ? nseg : sms * 8;
  ps_generate_kernel<P::NEXT><<<ggrid, GEN_BLOCK, 0, st>>>(g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("ps_generate_kernel launch", e);

  SmatrixArgs a;
  a.p = g.buf.mom, a.layout = MFP_LAYOUT_AOS, a.nevt = g.buf.cap, a.coup = nullptr, a.coup_stride = 0;
  a.sqh = u->sqh, a.out = g.buf.me, a.only_comb = -1;
  a.alpha_s = g.buf.as, a.seg_count = g.buf.count, a.seg_size = g.buf.seg, a.nseg = nseg;
  for (int i = 0; i < MFP_MAX_PARAMS; ++i) a.par[i] = i < P::NPAR ? u->par[i] : 0.0;
  const size_t smem = hp_smatrix_smem<P>();
  e = cudaFuncSetAttribute(smatrix_kernel_hp<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail("smatrix_kernel_hp smem attribute", e);
  smatrix_kernel_hp<P><<<u->nblocks, P::HP_THREADS, smem, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("smatrix_kernel_hp launch", e);

  if (u->skip_accumulate) return 0;  // the caller sums several subprocesses (mf_vegas_accumulate_sum)
  accumulate_kernel<<<u->nblocks, ACC_BLOCK, accumulate_smem(NDIM), st>>>(
      g.buf.me, g.buf.w, g.buf.bins, g.buf.cap, NDIM, u->accumulate_hist, u->d_partial);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("accumulate_kernel launch", e);
  return 0;
}

// kernel flavour: 0 = the process's default (P::USE_HP), 1 = one event per thread, 2 = helicity-parallel
static int g_variant = 0;
template <class P>
bool use_hp() { return P::HP_AVAILABLE && (!P::HAS_THREAD || (g_variant == 0 ? P::USE_HP : g_variant == 2)); }

template <class P>
int dispatch_smatrix(const double* d_p, int layout, long long nevt, const double* par, const double* d_coup,
                     long long cs, double sqh, double* d_out, int only_comb, cudaStream_t st) {
  if constexpr (P::HAS_THREAD) {
    if (!use_hp<P>()) return launch_smatrix<P>(d_p, layout, nevt, par, d_coup, cs, sqh, d_out, only_comb, st);
  }
  return launch_smatrix_hp<P>(d_p, layout, nevt, par, d_coup, cs, sqh, d_out, only_comb, st);
}

}  // namespace mf

#ifdef MF_HP_PROFILE
#define MF_DEFINE_PROFILE_HOOK                                                          \
  int mfp_profile_read(unsigned long long* out8, int reset) {                           \
    cudaDeviceSynchronize();                                                            \
    if (out8) cudaMemcpyFromSymbol(out8, mf::g_hp_prof, 8 * sizeof(unsigned long long)); \
    if (reset) {                                                                        \
      unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};                               \
      cudaMemcpyToSymbol(mf::g_hp_prof, z, sizeof(z));                                  \
    }                                                                                   \
    return 0;                                                                           \
  }
#else
#define MF_DEFINE_PROFILE_HOOK
#endif

#define MF_DEFINE_PROCESS(P)                                                                                   \
  extern "C" {                                                                                                 \
  int mfp_get_info(mfp_info* o) {                                                                              \
    if (!o) return mf::fail_msg("mfp_get_info: null pointer");                                                 \
    memset(o, 0, sizeof(*o));                                                                                  \
    strncpy(o->name, P::name(), sizeof(o->name) - 1);                                                          \
    o->nexternal = P::NEXT, o->ninitial = P::NINIT, o->ncomb = P::NCOMB, o->ncolor = P::NCOLOR;                \
    o->ndiags = P::NDIAGS, o->namps = P::NAMPS, o->nwavefuncs = P::NWF, o->nparams = P::NPAR;                  \
    o->ncouplings = P::NCOUP, o->ndim = 4 * (P::NEXT - 2) + 2;                                                 \
    o->block_threads = mf::use_hp<P>() ? P::HP_THREADS : P::BLOCK;                                        \
    o->denominator = P::DENOM, o->flops_per_event = P::FLOPS;                                                  \
    return 0;                                                                                                  \
  }                                                                                                            \
  const char* mfp_param_name(int i) { return (i >= 0 && i < P::NPAR) ? P::param_name(i) : ""; }                \
  const char* mfp_coupling_name(int i) { return (i >= 0 && i < P::NCOUP) ? P::coupling_name(i) : ""; }         \
  int mfp_coupling_def(int i, double* re, double* im, int* power) {                                            \
    if (i < 0 || i >= P::NCOUP) return mf::fail_msg("mfp_coupling_def: index out of range");                   \
    *re = P::coup_re(i), *im = P::coup_im(i), *power = P::coup_power(i);                                       \
    return 0;                                                                                                  \
  }                                                                                                            \
  int mfp_helicity(int ic, int leg) {                                                                          \
    return (ic >= 0 && ic < P::NCOMB && leg >= 0 && leg < P::NEXT) ? P::hel(ic, leg) : 0;                      \
  }                                                                                                            \
  int mfp_set_variant(int v) {                                                                                 \
    if (v < 0 || v > 2) return mf::fail_msg("mfp_set_variant: 0 default, 1 thread-per-event, 2 helicity-parallel"); \
    if (v == 2 && !P::HP_AVAILABLE)                                                                            \
      return mf::fail_msg("mfp_set_variant: the helicity-parallel flavour does not know this process's vertices"); \
    if (v == 1 && !P::HAS_THREAD)                                                                              \
      return mf::fail_msg("mfp_set_variant: the one-event-per-thread flavour is not compiled for this process"); \
    mf::g_variant = v;                                                                                         \
    return 0;                                                                                                  \
  }                                                                                                            \
  int mfp_get_variant(void) { return mf::use_hp<P>() ? 2 : 1; }                                                \
  int mfp_set_integrand_blocks(int nblocks) {                                                                  \
    if (nblocks < 0) return mf::fail_msg("mfp_set_integrand_blocks: 0 = automatic, > 0 = grid size");          \
    mf::g_blocks_override = nblocks;                                                                           \
    return 0;                                                                                                  \
  }                                                                                                            \
  int mfp_smatrix(const double* d_p, int layout, int64_t nevt, const double* par, const double* d_coup,        \
                  int64_t cs, double sqh, double* d_out, void* st) {                                           \
    return mf::dispatch_smatrix<P>(d_p, layout, nevt, par, d_coup, cs, sqh, d_out, -1, (cudaStream_t)st);      \
  }                                                                                                            \
  int mfp_matrix_hel(const double* d_p, int layout, int64_t nevt, int ic, const double* par,                   \
                     const double* d_coup, int64_t cs, double sqh, double* d_out, void* st) {                  \
    if (ic < 0 || ic >= P::NCOMB) return mf::fail_msg("mfp_matrix_hel: helicity row out of range");            \
    return mf::dispatch_smatrix<P>(d_p, layout, nevt, par, d_coup, cs, sqh, d_out, ic, (cudaStream_t)st);      \
  }                                                                                                            \
  int mfp_smatrix_host(const double* h_p, int layout, int64_t nevt, const double* par, const double* h_coup,   \
                       int64_t cs, double sqh, double* h_out) {                                                \
    return mf::smatrix_host<P>(mf::dispatch_smatrix<P>, h_p, layout, nevt, par, h_coup, cs, sqh, h_out);       \
  }                                                                                                            \
  int mfp_integrand_blocks(void) {                                                                             \
    if constexpr (P::HAS_THREAD) {                                                                             \
      if (!mf::use_hp<P>()) return mf::integrand_blocks<P>();                                                  \
    }                                                                                                          \
    return mf::integrand_blocks_hp<P>();                                                                       \
  }                                                                                                            \
  int64_t mfp_integrand_workspace(int64_t nevents) {                                                           \
    return mf::use_hp<P>() ? mf::integrand_workspace_hp<P>(nevents) : 0;                                       \
  }                                                                                                            \
  int mfp_integrand(const mfp_integrand_args* a, void* st) {                                                   \
    if (!a) return mf::fail_msg("mfp_integrand: null args");                                                   \
    if constexpr (P::HAS_THREAD) {                                                                             \
      if (!mf::use_hp<P>()) return mf::launch_integrand<P>(a, (cudaStream_t)st);                               \
    }                                                                                                          \
    return mf::launch_integrand_hp<P>(a, (cudaStream_t)st);                                                    \
  }                                                                                                            \
  int mfp_integrand_events(void* d_workspace, int64_t nevents, mfp_event_view* out) {                          \
    if (!out || !d_workspace) return mf::fail_msg("mfp_integrand_events: null pointer");                       \
    if (!mf::use_hp<P>()) return mf::fail_msg("mfp_integrand_events: the one-event-per-thread flavour has no event buffer"); \
    return mf::integrand_events_hp<P>(d_workspace, nevents, out);                                              \
  }                                                                                                            \
  const char* mfp_last_error(void) { return mf::g_err; }                                                       \
  MF_DEFINE_PROFILE_HOOK                                                                                       \
  }
