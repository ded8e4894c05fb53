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
  int only_h = -1;
  if (a.only_comb >= 0) {
    only_h = 0;
    for (int j = 0; j < P::NEXT; ++j) only_h |= ((P::hel(a.only_comb, j) + 1) >> 1) << j;
  }
  // plain mode: one "segment" holding all nevt events, blocks stride over its groups of E events;
  // segmented mode: block b owns segments b, b + gridDim.x, ... and walks their valid events
  const bool segmented = a.seg_count != nullptr;
  const int nseg = segmented ? a.nseg : 1;
  for (int sg = segmented ? blockIdx.x : 0; sg < nseg; sg += segmented ? gridDim.x : 1) {
    const long long base = segmented ? (long long)sg * a.seg_size : 0;
    const long long cnt = segmented ? a.seg_count[sg] : a.nevt;
    const long long ngroups = (cnt + E - 1) / E;
    for (long long g = segmented ? 0 : blockIdx.x; g < ngroups; g += segmented ? 1 : gridDim.x) {
      const long long ev0 = base + g * E;
      const int nev = (int)((base + cnt - ev0) < E ? (base + cnt - ev0) : E);
      for (int i = tid; i < E * P::NEXT * 4; i += T) {
        const int e = i / (P::NEXT * 4), r = i - e * (P::NEXT * 4);
        const long long ev = ev0 + (e < nev ? e : 0);  // pad a partial group with a valid event
        s.mom[i] = a.layout == MFP_LAYOUT_AOS ? a.p[ev * (P::NEXT * 4) + r] : a.p[(long long)r * a.nevt + ev];
      }
      for (int i = tid; i < E * P::NCOUP; i += T) {
        const int e = i / P::NCOUP, c = i - e * P::NCOUP;
        const long long ev = ev0 + (e < nev ? e : 0);
        if (a.alpha_s) {  // couplings from alpha_s: G = 2 sqrt(pi alpha_s) (parameters.py:13-15)
          const double G = 2.0 * sqrt(M_PI * a.alpha_s[ev]);
          double gp = 1.0;
          for (int q = 0; q < P::coup_power(c); ++q) gp *= G;
          s.coup[i] = mk(P::coup_re(c) * gp, P::coup_im(c) * gp);
        } else {
          const double2 v = reinterpret_cast<const double2*>(a.coup)[a.coup_stride ? (long long)c * a.nevt + ev : c];
          s.coup[i] = mk(v.x, v.y);
        }
      }
      __syncthreads();
      const double me = hp_smatrix_block<P>(nev, s.mom, s.coup, a.par, a.sqh, wf, s.vtab, s.red, only_h);
      const int e = tid / NH, h = tid - e * NH;
      if (h == 0 && e < nev) a.out[ev0 + e] = me;
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <class P>
size_t hp_smatrix_smem() { return ((sizeof(HpSmatrixSmem<P>) + 15) / 16) * 16 + sizeof(cxd) * P::HP_WFSIZE * P::HP_E; }

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
  const size_t smem = hp_smatrix_smem<P>();
  cudaFuncSetAttribute(smatrix_kernel_hp<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smatrix_kernel_hp<P>, P::HP_E * P::NCOMB, smem);
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
  smatrix_kernel_hp<P><<<u->nblocks, P::HP_E * P::NCOMB, smem, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("smatrix_kernel_hp launch", e);

  accumulate_kernel<<<u->nblocks, ACC_BLOCK, (NDIM * VEGAS_BINS + 24) * sizeof(double), st>>>(
      g.buf.me, g.buf.w, g.buf.bins, g.buf.cap, NDIM, u->accumulate_hist, u->d_partial);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("accumulate_kernel launch", e);
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
  int64_t mfp_integrand_workspace(int64_t nevents) {                                                           \
    return mf::use_hp<P>() ? mf::integrand_workspace_hp<P>(nevents) : 0;                                       \
  }                                                                                                            \
  int mfp_integrand(const mfp_integrand_args* a, void* st) {                                                   \
    if (!a) return mf::fail_msg("mfp_integrand: null args");                                                   \
    return mf::use_hp<P>() ? mf::launch_integrand_hp<P>(a, (cudaStream_t)st)                                   \
                           : mf::launch_integrand<P>(a, (cudaStream_t)st);                                     \
  }                                                                                                            \
  const char* mfp_last_error(void) { return mf::g_err; }                                                       \
  }
