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
//             table driven (HpItem), one copy of each ALOHA routine in the instruction stream
//   phase 3   amplitudes + JAMP sums: thread (e, h) loops over the amplitude table, reading each
//             input current's variant for ITS helicity h (a warp's 32 helicities touch <= 8
//             variants of a current: one shared-memory wavefront per load); the JAMP updates are
//             a generated switch over the amplitude index, so the JAMPs stay in registers
//   phase 4   colour contraction per thread, then the sum over helicities is a warp-shuffle +
//             shared-memory reduction per event.
//
// Shared-memory layout of wavefunction w (sizes in cxd = 16 bytes), for the E events of the block:
//   [e][ 0..1 ]                momentum slots w0, w1  (helicity independent)
//   [e][ 2 + k*nv + v ]        component k = 0..3 (HELAS slots 2..5), helicity variant v < nv = 2^|S|
// starting at  P::wf_off(w)*E + e*(2 + 4*nv).  Variant bits follow the ascending leg order of S;
// bit = 1 means helicity +1.
#pragma once
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

enum HpAmpType : unsigned char { HP_FFV1_0 = 0, HP_VVV1_0 = 1, HP_VVVV1_0 = 2, HP_VVVV3_0 = 3, HP_VVVV4_0 = 4 };

struct HpAmp {
  unsigned char type, nin, coup, coup_neg;
  unsigned short in[4];
};

struct HpItem {
  unsigned char type, nin;
  signed char mass_idx, width_idx;   // < 0: ZERO
  unsigned char coup, coup_neg;
  unsigned short out;
  unsigned short in[3];
  unsigned long long vmap[3];        // 4 bits per output variant: the input's variant index
};

// number of cxd needed for the wavefunctions of E events
template <class P>
MF_HD constexpr int hp_wf_cxd(int E) { return P::HP_WFSIZE * E; }

template <class P>
MF_DEV cxd* hp_wf(cxd* wf, int E, int w, int e) {
  const HpWf d = P::wf(w);
  return wf + (size_t)d.off * E + (size_t)e * (2 + 4 * d.nv);
}

// phase 1: work item `it` in [0, NEXT*E*2)
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
  cxd* o = hp_wf<P>(wf, E, x.out, e);
  if (bit == 0) o[0] = w[0], o[1] = w[1];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[2 + k * 2 + bit] = w[2 + k];
}

template <class P>
MF_DEV void hp_load(const cxd* wf, int E, int w, int e, int v, cxd out[6]) {
  const HpWf d = P::wf(w);
  const cxd* s = wf + (size_t)d.off * E + (size_t)e * (2 + 4 * d.nv);
  out[0] = s[0], out[1] = s[1];
#pragma unroll
  for (int k = 0; k < 4; ++k) out[2 + k] = s[2 + k * d.nv + v];
}

// phase 2: work item (item index `idx`, event e, output variant v)
template <class P>
MF_DEV void hp_current(int idx, int e, int v, int E, const double* par, const cxd* coup_e, cxd* wf) {
  const HpItem it = P::item(idx);
  cxd a[6], b[6], c[6], r[6];
  hp_load<P>(wf, E, it.in[0], e, (int)((it.vmap[0] >> (4 * v)) & 15ull), a);
  hp_load<P>(wf, E, it.in[1], e, (int)((it.vmap[1] >> (4 * v)) & 15ull), b);
  if (it.nin > 2) hp_load<P>(wf, E, it.in[2], e, (int)((it.vmap[2] >> (4 * v)) & 15ull), c);
  cxd cp = coup_e[it.coup];
  if (it.coup_neg) cp = -cp;
  const double M = it.mass_idx < 0 ? 0.0 : par[it.mass_idx];
  const double W = it.width_idx < 0 ? 0.0 : par[it.width_idx];
  switch (it.type) {
    case HP_FFV1_1: FFV1_1(a, b, cp, M, W, r); break;
    case HP_FFV1_2: FFV1_2(a, b, cp, M, W, r); break;
    case HP_FFV1P0_3: FFV1P0_3(a, b, cp, M, W, r); break;
    case HP_VVV1P0_1: VVV1P0_1(a, b, cp, M, W, r); break;
    case HP_VVVV1P0_1: VVVVP0_1<1>(a, b, c, cp, M, W, r); break;
    case HP_VVVV3P0_1: VVVVP0_1<3>(a, b, c, cp, M, W, r); break;
    default: VVVVP0_1<4>(a, b, c, cp, M, W, r); break;
  }
  const HpWf d = P::wf(it.out);
  cxd* o = wf + (size_t)d.off * E + (size_t)e * (2 + 4 * d.nv);
  if (v == 0) o[0] = r[0], o[1] = r[1];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[2 + k * d.nv + v] = r[2 + k];
}

// phase 3 for thread (event e, helicity bits h): loop over the amplitude table.  The loop body is a
// handful of routines selected by a block-uniform switch, and the JAMP updates of amplitude `ai` are a
// generated `switch (ai)` whose cases address the JAMP registers statically -- the whole phase is a few
// tens of KB of instructions and stays in the instruction cache (a fully unrolled amplitude list is
// ~300 KB for g g > t t~ g g and stalls on instruction fetch: profiles/r01_ttxgg_hp_v1.summary.txt).
// vtab[w*NCOMB + h] is the helicity variant of wavefunction w that belongs to helicity combination h.
template <class P>
MF_DEV void hp_load_amp(const cxd* wf, const unsigned char* vtab, int E, int e, int h, int w, cxd out[6]) {
  const HpWf d = P::wf(w);
  const cxd* s = wf + (size_t)d.off * E + (size_t)e * (2 + 4 * d.nv);
  const int v = vtab[w * P::NCOMB + h];
  out[0] = s[0], out[1] = s[1];
#pragma unroll
  for (int k = 0; k < 4; ++k) out[2 + k] = s[2 + k * d.nv + v];
}

template <class P>
MF_DEV double hp_amplitudes(const cxd* wf, const unsigned char* vtab, int E, int e, int h, const cxd* coup) {
  // short amplitude lists are emitted as straight-line code (it fits the instruction cache and has no
  // loop overhead); long ones run the table-driven loop below
  if (P::HP_UNROLL) return P::hp_amps_unrolled(wf, vtab, E, e, h, coup);
  cxd J[P::NCOLOR];
#pragma unroll
  for (int j = 0; j < P::NCOLOR; ++j) J[j] = mk(0.0, 0.0);
#pragma unroll 1
  for (int ai = 0; ai < P::HP_NAMPS; ++ai) {
    const HpAmp it = P::amp(ai);
    cxd a[6], b[6], c[6], d[6];
    hp_load_amp<P>(wf, vtab, E, e, h, it.in[0], a);
    hp_load_amp<P>(wf, vtab, E, e, h, it.in[1], b);
    hp_load_amp<P>(wf, vtab, E, e, h, it.in[2], c);
    cxd cp = coup[it.coup];
    if (it.coup_neg) cp = -cp;
    cxd amp;
    switch (it.type) {
      case HP_FFV1_0: amp = FFV1_0(a, b, c, cp); break;
      case HP_VVV1_0: amp = VVV1_0(a, b, c, cp); break;
      default:
        hp_load_amp<P>(wf, vtab, E, e, h, it.in[3], d);
        if (it.type == HP_VVVV1_0) amp = VVVV_0<1>(a, b, c, d, cp);
        else if (it.type == HP_VVVV3_0) amp = VVVV_0<3>(a, b, c, d, cp);
        else amp = VVVV_0<4>(a, b, c, d, cp);
        break;
    }
    P::jamp_accumulate(ai, amp, J);
  }
  return P::colour_sum(J);
}

// vtab: helicity variant of every wavefunction for every helicity combination (block-wide, once)
template <class P>
MF_DEV void hp_fill_vtab(int idx, unsigned char* vtab) {
  const int w = idx / P::NCOMB, h = idx - w * P::NCOMB;
  const unsigned legs = P::wf(w).legs;
  int out = 0, pos = 0;
  for (int b = 0; b < P::NEXT; ++b)
    if (legs & (1u << b)) {
      out |= ((h >> b) & 1) << pos;
      ++pos;
    }
  vtab[idx] = (unsigned char)out;
}

// ------------------------------------------------------------------------------------------------
// The E-event matrix-element evaluation used by both kernels.  `mom` [E][NEXT][4], `coup` [E][NCOUP]
// and `wf` live in shared memory; returns, in thread (e, h = 0), the event's |M|^2 summed over
// helicities and colours and averaged (other threads return garbage).
template <class P>
__device__ __forceinline__ double hp_smatrix_block(int nev /* <= E valid events */, const double* mom,
                                                   const cxd* coup, const double* par, double sqh, cxd* wf,
                                                   const unsigned char* vtab, double* red /* [T/32] */,
                                                   int only_h) {
  constexpr int E = P::HP_E, NH = P::NCOMB, T = E * NH;
  const int tid = threadIdx.x;
  for (int it = tid; it < P::NEXT * E * 2; it += T) hp_externals<P>(it, E, mom, par, sqh, wf);
  __syncthreads();
#pragma unroll 1
  for (int L = 2; L <= P::HP_MAXLEVEL; ++L) {
    const int begin = P::level_begin(L), cnt = P::level_begin(L + 1) - begin;
    const int nv = 1 << L;
    const int total = cnt * E * nv;
#pragma unroll 1
    for (int w = tid; w < total; w += T) {
      const int ci = w / (E * nv);
      const int r = w - ci * (E * nv);
      const int e = r / nv, v = r - e * nv;
      hp_current<P>(begin + ci, e, v, E, par, coup + e * P::NCOUP, wf);
    }
    __syncthreads();
  }
  const int e = tid / NH, h = tid - e * NH;
  double me = hp_amplitudes<P>(wf, vtab, E, e, h, coup + e * P::NCOUP);
  if (only_h >= 0) me = (h == only_h) ? me : 0.0;
  if (e >= nev) me = 0.0;
  // sum over helicities of one event
  constexpr int W = NH < 32 ? NH : 32;
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) me += __shfl_down_sync(0xffffffffu, me, o, W);
  if (NH > 32) {
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) red[warp] = me;
    __syncthreads();
    if (h == 0) {
      me = 0.0;
#pragma unroll
      for (int k = 0; k < NH / 32; ++k) me += red[e * (NH / 32) + k];
    }
  }
  return only_h >= 0 ? me : me / P::DENOM;
}

template <class P>
struct HpSmatrixSmem {
  static constexpr int E = P::HP_E, T = E * P::NCOMB;
  double mom[E * P::NEXT * 4];
  cxd coup[E * (P::NCOUP > 0 ? P::NCOUP : 1)];
  double red[T / 32 + 1];
  unsigned char vtab[P::HP_NWF * P::NCOMB];
  // followed by cxd wf[HP_WFSIZE * E]
};

template <class P>
__global__ void __launch_bounds__(P::HP_E* P::NCOMB, P::HP_MINBLOCKS) smatrix_kernel_hp(const SmatrixArgs a) {
  constexpr int E = P::HP_E, NH = P::NCOMB, T = E * NH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HpSmatrixSmem<P>& s = *reinterpret_cast<HpSmatrixSmem<P>*>(smem_raw);
  cxd* wf = reinterpret_cast<cxd*>(smem_raw + ((sizeof(HpSmatrixSmem<P>) + 15) / 16) * 16);
  const int tid = threadIdx.x;
  for (int i = tid; i < P::HP_NWF * P::NCOMB; i += T) hp_fill_vtab<P>(i, s.vtab);
  const long long ngroups = (a.nevt + E - 1) / E;
  for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const long long ev0 = g * E;
    const int nev = (int)((a.nevt - ev0) < E ? (a.nevt - ev0) : E);
    for (int i = tid; i < E * P::NEXT * 4; i += T) {
      const int e = i / (P::NEXT * 4), r = i - e * (P::NEXT * 4);
      const long long ev = ev0 + (e < nev ? e : 0);  // pad with a valid event
      s.mom[i] = a.layout == MFP_LAYOUT_AOS ? a.p[ev * (P::NEXT * 4) + r] : a.p[(long long)r * a.nevt + ev];
    }
    for (int i = tid; i < E * P::NCOUP; i += T) {
      const int e = i / P::NCOUP, c = i - e * P::NCOUP;
      const long long ev = ev0 + (e < nev ? e : 0);
      const double2 v = reinterpret_cast<const double2*>(a.coup)[a.coup_stride ? (long long)c * a.nevt + ev : c];
      s.coup[i] = mk(v.x, v.y);
    }
    __syncthreads();
    int only_h = -1;
    if (a.only_comb >= 0) {
      only_h = 0;
      for (int j = 0; j < P::NEXT; ++j) only_h |= ((P::hel(a.only_comb, j) + 1) >> 1) << j;
    }
    const double me = hp_smatrix_block<P>(nev, s.mom, s.coup, a.par, a.sqh, wf, s.vtab, s.red, only_h);
    const int e = tid / NH, h = tid - e * NH;
    if (h == 0 && e < nev) a.out[ev0 + e] = me;
    __syncthreads();
  }
}

// fused integrand, hp flavour: phase-space generation one event per thread, accepted events
// queued in shared memory, matrix elements E events at a time
template <class P>
struct HpIntegrandSmem {
  static constexpr int NDIM = 4 * (P::NEXT - 2) + 2;
  static constexpr int E = P::HP_E, T = E * P::NCOMB;
  static constexpr int QCAP = T + E;
  double grid[NDIM * VEGAS_EDGES];
  double hist[NDIM * VEGAS_BINS];
  double qmom[QCAP][P::NEXT * 4];   // event-major: the ME phase copies whole events
  double qw[QCAP];
  double qas[QCAP];
  unsigned char qbin[QCAP][NDIM];
  int warp_count[32];
  double red3[3][32];
  cxd coup[E * (P::NCOUP > 0 ? P::NCOUP : 1)];
  double red[T / 32 + 1];
  unsigned char vtab[P::HP_NWF * P::NCOMB];
};

template <class P>
__global__ void __launch_bounds__(P::HP_E* P::NCOMB, P::HP_MINBLOCKS) integrand_kernel_hp(const IntegrandArgs a) {
  using S = HpIntegrandSmem<P>;
  constexpr int NDIM = S::NDIM, E = P::HP_E, NH = P::NCOMB, T = E * NH, NWARP = T / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& s = *reinterpret_cast<S*>(smem_raw);
  cxd* wf = reinterpret_cast<cxd*>(smem_raw + ((sizeof(S) + 15) / 16) * 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < NDIM * VEGAS_EDGES; i += T) s.grid[i] = a.u.d_grid[i];
  for (int i = tid; i < NDIM * VEGAS_BINS; i += T) s.hist[i] = 0.0;
  for (int i = tid; i < P::HP_NWF * P::NCOMB; i += T) hp_fill_vtab<P>(i, s.vtab);
  __syncthreads();

  double s1 = 0.0, s2 = 0.0, cnt = 0.0;
  int qcount = 0;
  const long long ntiles = (a.u.nevents + T - 1) / T;
  const long long my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  for (long long k = 0; k <= my_tiles; ++k) {
    const bool flush = (k == my_tiles);
    if (!flush) {
      const long long tile = blockIdx.x + k * gridDim.x;
      const long long local = tile * T + tid;
      bool ok = false;
      double m[P::NEXT][4];
      double wgt = 0.0, as = 0.0;
      unsigned char bins[NDIM];
      if (local < a.u.nevents) {
        const unsigned long long ev = a.u.first_event + (unsigned long long)local;
        double xr[NDIM];
        double w = 1.0;
#pragma unroll
        for (int j = 0; j < (NDIM + 1) / 2; ++j) {
          double u0, u1;
          philox_pair(a.u.seed, a.u.iteration, ev, j, u0, u1);
          int b;
          xr[2 * j] = vegas_map(&s.grid[(2 * j) * VEGAS_EDGES], vegas_confine(u0), b, w);
          bins[2 * j] = (unsigned char)b;
          if (2 * j + 1 < NDIM) {
            xr[2 * j + 1] = vegas_map(&s.grid[(2 * j + 1) * VEGAS_EDGES], vegas_confine(u1), b, w);
            bins[2 * j + 1] = (unsigned char)b;
          }
        }
        double x1, x2;
        ramboflow<P::NEXT>(xr, a.u.com_sqrts, a.u.masses, a.massive != 0, a.shat_min, a.ps, m, wgt, x1, x2);
        ok = pass_cuts<P::NEXT>(a.cuts, m);
        ok = ok && (wgt == wgt) && (wgt != 0.0);
        if (ok) {
          if (a.u.lab_frame) boost_to_lab<P::NEXT>(m, x1, x2);
          double q2 = 0.0;
          if (a.u.alpha_mode != 0) {
            double smt = 0.0;
#pragma unroll
            for (int i = 2; i < P::NEXT; ++i) smt += cut_value(CUT_MT, m[i]);
            q2 = (smt / 2.0) * (smt / 2.0);
          }
          as = alpha_s_of(a.u, q2);
          wgt *= w * a.u.inv_total_events;
        }
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s.warp_count[warp] = __popc(ballot);
      __syncthreads();
      int base = qcount, total = 0;
#pragma unroll
      for (int wv = 0; wv < NWARP; ++wv) {
        const int c = s.warp_count[wv];
        if (wv < warp) base += c;
        total += c;
      }
      if (ok) {
        const int slot = base + __popc(ballot & ((1u << lane) - 1u));
#pragma unroll
        for (int i = 0; i < P::NEXT; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) s.qmom[slot][i * 4 + c] = m[i][c];
        s.qw[slot] = wgt;
        s.qas[slot] = as;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) s.qbin[slot][d] = bins[d];
      }
      qcount += total;
      __syncthreads();
    }
    // matrix elements, E queued events at a time (all of them when flushing)
    while (qcount >= E || (flush && qcount > 0)) {
      const int nev = qcount < E ? qcount : E;
      const int first = qcount - nev;
      // couplings of these events
      for (int i = tid; i < nev * P::NCOUP; i += T) {
        const int e = i / P::NCOUP, c = i - e * P::NCOUP;
        const double G = 2.0 * sqrt(M_PI * s.qas[first + e]);
        double g = 1.0;
        for (int q = 0; q < P::coup_power(c); ++q) g *= G;
        s.coup[i] = mk(P::coup_re(c) * g, P::coup_im(c) * g);
      }
      // pad missing events of a partial group with copies of the first one (results discarded)
      if (nev < E) {
        for (int i = tid; i < (E - nev) * P::NEXT * 4; i += T) {
          const int e = nev + i / (P::NEXT * 4), r = i % (P::NEXT * 4);
          s.qmom[first + e][r] = s.qmom[first][r];
        }
        for (int i = tid; i < (E - nev) * P::NCOUP; i += T) s.coup[nev * P::NCOUP + i] = s.coup[i % P::NCOUP];
      }
      __syncthreads();
      const double me = hp_smatrix_block<P>(nev, &s.qmom[first][0], s.coup, a.u.par, a.u.sqh, wf, s.vtab, s.red, -1);
      const int e = tid / NH, h = tid - e * NH;
      if (h == 0 && e < nev) {
        const int slot = first + e;
        const double t = me * s.qw[slot];
        const double t2 = t * t;
        s1 += t, s2 += t2, cnt += 1.0;
        if (a.u.accumulate_hist) {
#pragma unroll 1
          for (int d = 0; d < NDIM; ++d) atomicAdd(&s.hist[d * VEGAS_BINS + s.qbin[slot][d]], t2);
        }
      }
      qcount -= nev;
      __syncthreads();
    }
  }

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) s.red3[0][warp] = s1, s.red3[1][warp] = s2, s.red3[2][warp] = cnt;
  __syncthreads();
  double* out = a.u.d_partial + (long long)blockIdx.x * (VEGAS_HEADER + NDIM * VEGAS_BINS);
  if (tid == 0) {
    double t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int wv = 0; wv < NWARP; ++wv) t1 += s.red3[0][wv], t2 += s.red3[1][wv], t3 += s.red3[2][wv];
    out[0] = t1, out[1] = t2, out[2] = t3, out[3] = 0.0;
  }
  for (int i = tid; i < NDIM * VEGAS_BINS; i += T) out[VEGAS_HEADER + i] = s.hist[i];
}

// ------------------------------------------------------------------------------------------------
template <class P>
size_t hp_smatrix_smem() { return ((sizeof(HpSmatrixSmem<P>) + 15) / 16) * 16 + sizeof(cxd) * P::HP_WFSIZE * P::HP_E; }
template <class P>
size_t hp_integrand_smem() { return ((sizeof(HpIntegrandSmem<P>) + 15) / 16) * 16 + sizeof(cxd) * P::HP_WFSIZE * P::HP_E; }

template <class P>
int launch_smatrix_hp(const double* d_p, int layout, long long nevt, const double* par, const double* d_coup,
                      long long coup_stride, double sqh, double* d_out, int only_comb, cudaStream_t st) {
  if (nevt <= 0) return 0;
  if (layout != MFP_LAYOUT_AOS && layout != MFP_LAYOUT_SOA) return fail_msg("mfp_smatrix: unknown layout");
  if (P::NCOUP > 0 && d_coup == nullptr) return fail_msg("mfp_smatrix: couplings missing");
  SmatrixArgs a;
  a.p = d_p, a.layout = layout, a.nevt = nevt, a.coup = d_coup, a.coup_stride = coup_stride, a.sqh = sqh;
  a.out = d_out, a.only_comb = only_comb;
  for (int i = 0; i < MFP_MAX_PARAMS; ++i) a.par[i] = i < P::NPAR ? par[i] : 0.0;
  const size_t smem = hp_smatrix_smem<P>();
  cudaError_t e = cudaFuncSetAttribute(smatrix_kernel_hp<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail("smatrix_kernel_hp smem attribute", e);
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smatrix_kernel_hp<P>, P::HP_E * P::NCOMB, smem);
  if (per_sm < 1) per_sm = 1;
  long long blocks = (nevt + P::HP_E - 1) / P::HP_E;
  const long long cap = (long long)sms * per_sm;
  if (blocks > cap) blocks = cap;
  smatrix_kernel_hp<P><<<(unsigned)blocks, P::HP_E * P::NCOMB, smem, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("smatrix_kernel_hp launch", e);
  return 0;
}

template <class P>
int integrand_blocks_hp() {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = hp_integrand_smem<P>();
  cudaFuncSetAttribute(integrand_kernel_hp<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, integrand_kernel_hp<P>, P::HP_E * P::NCOMB, smem);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

template <class P>
int launch_integrand_hp(const mfp_integrand_args* u, cudaStream_t st) {
  IntegrandArgs a;
  if (int rc = prepare_integrand_args<P>(u, a)) return rc;
  const size_t smem = hp_integrand_smem<P>();
  cudaError_t e = cudaFuncSetAttribute(integrand_kernel_hp<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail("integrand_kernel_hp smem attribute", e);
  integrand_kernel_hp<P><<<u->nblocks, P::HP_E * P::NCOMB, smem, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("integrand_kernel_hp launch", e);
  return 0;
}

// kernel flavour: 0 = the process's default (P::USE_HP), 1 = one event per thread, 2 = helicity-parallel
static int g_variant = 0;
template <class P>
bool use_hp() { return g_variant == 0 ? P::USE_HP : g_variant == 2; }

template <class P>
int dispatch_smatrix(const double* d_p, int layout, long long nevt, const double* par, const double* d_coup,
                     long long cs, double sqh, double* d_out, int only_comb, cudaStream_t st) {
  return use_hp<P>() ? launch_smatrix_hp<P>(d_p, layout, nevt, par, d_coup, cs, sqh, d_out, only_comb, st)
                     : launch_smatrix<P>(d_p, layout, nevt, par, d_coup, cs, sqh, d_out, only_comb, st);
}

}  // namespace mf

#define MF_DEFINE_PROCESS(P)                                                                                   \
  extern "C" {                                                                                                 \
  int mfp_get_info(mfp_info* o) {                                                                              \
    if (!o) return mf::fail_msg("mfp_get_info: null pointer");                                                 \
    memset(o, 0, sizeof(*o));                                                                                  \
    strncpy(o->name, P::name(), sizeof(o->name) - 1);                                                          \
    o->nexternal = P::NEXT, o->ninitial = P::NINIT, o->ncomb = P::NCOMB, o->ncolor = P::NCOLOR;                \
    o->ndiags = P::NDIAGS, o->namps = P::NAMPS, o->nwavefuncs = P::NWF, o->nparams = P::NPAR;                  \
    o->ncouplings = P::NCOUP, o->ndim = 4 * (P::NEXT - 2) + 2;                                                 \
    o->block_threads = mf::use_hp<P>() ? P::HP_E * P::NCOMB : P::BLOCK;                                        \
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
    mf::g_variant = v;                                                                                         \
    return 0;                                                                                                  \
  }                                                                                                            \
  int mfp_get_variant(void) { return mf::use_hp<P>() ? 2 : 1; }                                                \
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
    return mf::use_hp<P>() ? mf::integrand_blocks_hp<P>() : mf::integrand_blocks<P>();                         \
  }                                                                                                            \
  int mfp_integrand(const mfp_integrand_args* a, void* st) {                                                   \
    if (!a) return mf::fail_msg("mfp_integrand: null args");                                                   \
    return mf::use_hp<P>() ? mf::launch_integrand_hp<P>(a, (cudaStream_t)st)                                   \
                           : mf::launch_integrand<P>(a, (cudaStream_t)st);                                     \
  }                                                                                                            \
  const char* mfp_last_error(void) { return mf::g_err; }                                                       \
  }
