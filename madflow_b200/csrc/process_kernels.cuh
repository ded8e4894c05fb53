// Kernels and C-ABI glue shared by every generated process translation unit.
//
// A generated file defines `struct Proc` (constants, helicity table, `matrix()` = the ordered HELAS
// call list + JAMP + colour contraction of ONE helicity, i.e. the body of the reference's
// Matrix_<proc>.matrix, madgraph_plugin/template_files/matrix_method_python.inc:106-138) and then
// expands MF_DEFINE_PROCESS(Proc), which instantiates:
//
//   smatrix_kernel<Proc>    Matrix_<proc>.smatrix (inc:80-104): one event per thread, runtime
//                           (warp-uniform) loop over all helicity rows, result / denominator.
//   integrand_kernel<Proc>  the whole per-event integrand of scripts/madflow_exec.py:422-470 fused
//                           in one persistent kernel: Philox -> VEGAS map (grid in shared memory)
//                           -> x1,x2 -> RAMBO -> cuts -> boost -> alpha_s/couplings -> smatrix ->
//                           weight -> per-block sums + shared-memory histogram.  Events that fail
//                           the cuts are dropped before the matrix element: accepted events are
//                           queued in shared memory and the matrix element always runs on full
//                           blocks (the reference compacts with tf.boolean_mask, phasespace.py:506-515).
#pragma once
#include <cstring>
#include <mutex>
#include <cstdio>
#include <cstring>

#include "../../include/madflow_b200_process.h"
#include "aloha_sm.cuh"
#include "helas.cuh"
#include "pdf.cuh"
#include "phasespace.cuh"
#include "philox.cuh"
#include "vegas.cuh"

namespace mf {

static thread_local char g_err[512] = "";
static inline int fail(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -1;
}
static inline int fail_msg(const char* what) {
  snprintf(g_err, sizeof(g_err), "%s", what);
  return -2;
}

struct SmatrixArgs {
  const double* p;
  int layout;
  long long nevt;
  double par[MFP_MAX_PARAMS];
  const double* coup;
  long long coup_stride;
  double sqh;
  double* out;
  int only_comb;  // >= 0: evaluate this helicity row only and do not average (test hook)
  // segmented input (integrand pipeline): events live in `nseg` segments of `seg_size` slots of which
  // the first seg_count[s] are valid; couplings then come from alpha_s[slot]
  const double* alpha_s;
  const int* seg_count;
  long long seg_size;
  int nseg;
};

template <class P>
MF_DEV void load_momenta(const double* p, int layout, long long nevt, long long ev, double m[P::NEXT][4]) {
  if (layout == MFP_LAYOUT_AOS) {
    const double4* q = reinterpret_cast<const double4*>(p) + ev * P::NEXT;
#pragma unroll
    for (int i = 0; i < P::NEXT; ++i) {
      const double4 v = q[i];
      m[i][0] = v.x, m[i][1] = v.y, m[i][2] = v.z, m[i][3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < P::NEXT; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) m[i][k] = p[(long long)(i * 4 + k) * nevt + ev];
  }
}

template <class P>
MF_DEV double smatrix_event(const double m[P::NEXT][4], const double* par, const cxd* coup, double sqh) {
  double acc = 0.0;
#pragma unroll 1
  for (int ic = 0; ic < P::NCOMB; ++ic) acc += P::matrix(m, ic, par, coup, sqh);
  return acc / P::DENOM;
}

template <class P>
__global__ void __launch_bounds__(P::BLOCK, P::MINBLOCKS) smatrix_kernel(const SmatrixArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long ev = (long long)blockIdx.x * blockDim.x + threadIdx.x; ev < a.nevt; ev += stride) {
    double m[P::NEXT][4];
    load_momenta<P>(a.p, a.layout, a.nevt, ev, m);
    cxd coup[P::NCOUP > 0 ? P::NCOUP : 1];
    const double2* c2 = reinterpret_cast<const double2*>(a.coup);
#pragma unroll
    for (int c = 0; c < P::NCOUP; ++c) {
      const double2 v = a.coup_stride ? c2[(long long)c * a.nevt + ev] : c2[c];
      coup[c] = mk(v.x, v.y);
    }
    if (a.only_comb >= 0)
      a.out[ev] = P::matrix(m, a.only_comb, a.par, coup, a.sqh);
    else
      a.out[ev] = smatrix_event<P>(m, a.par, coup, a.sqh);
  }
}

// ------------------------------------------------------------------------------------------------
// fused integrand
struct IntegrandArgs {
  mfp_integrand_args u;  // the caller's description
  // derived on the host
  int massive;
  double shat_min;
  PSConst ps;
  CutList cuts;
};

template <class P>
struct IntegrandSmem {
  static constexpr int NDIM = 4 * (P::NEXT - 2) + 2;
  static constexpr int QCAP = 2 * P::BLOCK;
  double grid[NDIM * VEGAS_EDGES];
  double hist[P::BLOCK / 32][NDIM * VEGAS_BINS];   // one per warp: deterministic accumulation (vegas.cuh::warp_hist_add)
  double qmom[P::NEXT * 4][QCAP];
  double qw[QCAP];      // xjac * phase-space weight
  double qas[QCAP];     // alpha_s
  unsigned char qbin[NDIM][QCAP];
  int warp_count[32];
  double red[3][32];
};

// `valid`: this thread has an event in `slot` (all threads of a warp must call: the histogram update is a warp
// collective with a fixed summation order)
template <class P>
__device__ __forceinline__ void integrand_process_entry(const IntegrandArgs& a, IntegrandSmem<P>& s, int slot, bool valid,
                                                        double& s1, double& s2, double& cnt) {
  constexpr int NDIM = IntegrandSmem<P>::NDIM;
  if (!valid) slot = 0;
  double m[P::NEXT][4];
#pragma unroll
  for (int i = 0; i < P::NEXT; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) m[i][k] = s.qmom[i * 4 + k][slot];
  // couplings from alpha_s: G = 2 sqrt(pi alpha_s) (parameters.py:13-15), c = (re + i im) G^power
  const double G = 2.0 * sqrt(M_PI * s.qas[slot]);
  cxd coup[P::NCOUP > 0 ? P::NCOUP : 1];
#pragma unroll
  for (int c = 0; c < P::NCOUP; ++c) {
    double g = 1.0;
    for (int k = 0; k < P::coup_power(c); ++k) g *= G;
    coup[c] = mk(P::coup_re(c) * g, P::coup_im(c) * g);
  }
  const double me = valid ? smatrix_event<P>(m, a.u.par, coup, a.u.sqh) : 0.0;
  const double t = valid ? me * s.qw[slot] : 0.0;
  const double t2 = t * t;
  s1 += t;
  s2 += t2;
  cnt += valid ? 1.0 : 0.0;
  if (a.u.accumulate_hist) {
    double* whist = s.hist[threadIdx.x >> 5];
#pragma unroll 1
    for (int d = 0; d < NDIM; ++d) warp_hist_add(whist + d * VEGAS_BINS, s.qbin[d][slot], t2, valid);
  }
}

template <class P>
__global__ void __launch_bounds__(P::BLOCK, P::MINBLOCKS) integrand_kernel(const IntegrandArgs a) {
  constexpr int NDIM = IntegrandSmem<P>::NDIM;
  constexpr int B = P::BLOCK;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  IntegrandSmem<P>& s = *reinterpret_cast<IntegrandSmem<P>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = B / 32;

  for (int i = tid; i < NDIM * VEGAS_EDGES; i += B) s.grid[i] = a.u.d_grid[i];
  for (int i = tid; i < NWARP * NDIM * VEGAS_BINS; i += B) (&s.hist[0][0])[i] = 0.0;
  __syncthreads();

  double s1 = 0.0, s2 = 0.0, cnt = 0.0;
  int qcount = 0;
  const long long ntiles = (a.u.nevents + B - 1) / B;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long local = tile * B + tid;
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
      ok = pass_cuts<P::NEXT>(a.cuts, m);  // on centre-of-mass momenta (phasespace.py:506-508)
      // a vanishing or non-finite weight cannot contribute; drop it like a cut event
      ok = ok && (wgt == wgt) && (wgt != 0.0);
      if (ok) {
        if (a.u.lab_frame) boost_to_lab<P::NEXT>(m, x1, x2);
        double lumi;
        event_scale<P::NEXT>(a.u, m, x1, x2, as, lumi);  // madflow_exec.py:426-454
        wgt *= lumi;
        wgt *= w * a.u.inv_total_events;
      }
    }
    // block-wide stable compaction of the accepted events into the shared queue
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
        for (int k = 0; k < 4; ++k) s.qmom[i * 4 + k][slot] = m[i][k];
      s.qw[slot] = wgt;
      s.qas[slot] = as;
#pragma unroll
      for (int d = 0; d < NDIM; ++d) s.qbin[d][slot] = bins[d];
    }
    qcount += total;
    __syncthreads();
    if (qcount >= B) {  // at most once per tile: qcount < 2B always
      integrand_process_entry<P>(a, s, qcount - B + tid, true, s1, s2, cnt);
      qcount -= B;
      __syncthreads();
    }
  }
  if (__any_sync(0xffffffffu, tid < qcount)) integrand_process_entry<P>(a, s, tid, tid < qcount, s1, s2, cnt);

  // deterministic block reduction of the two sums
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) s.red[0][warp] = s1, s.red[1][warp] = s2, s.red[2][warp] = cnt;
  __syncthreads();
  double* out = a.u.d_partial + (long long)blockIdx.x * (VEGAS_HEADER + NDIM * VEGAS_BINS);
  if (tid == 0) {
    double t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int wv = 0; wv < NWARP; ++wv) t1 += s.red[0][wv], t2 += s.red[1][wv], t3 += s.red[2][wv];
    out[0] = t1, out[1] = t2, out[2] = t3, out[3] = 0.0;
  }
  for (int i = tid; i < NDIM * VEGAS_BINS; i += B) {
    double h = 0.0;
#pragma unroll
    for (int wv = 0; wv < NWARP; ++wv) h += s.hist[wv][i];   // warp order
    out[VEGAS_HEADER + i] = h;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
template <class P>
int launch_smatrix(const double* d_p, int layout, long long nevt, const double* par, const double* d_coup,
                   long long coup_stride, double sqh, double* d_out, int only_comb, cudaStream_t st) {
  if (nevt <= 0) return 0;
  if (layout != MFP_LAYOUT_AOS && layout != MFP_LAYOUT_SOA) return fail_msg("mfp_smatrix: unknown layout");
  if (P::NCOUP > 0 && d_coup == nullptr) return fail_msg("mfp_smatrix: couplings missing");
  SmatrixArgs a;
  a.p = d_p, a.layout = layout, a.nevt = nevt, a.coup = d_coup, a.coup_stride = coup_stride, a.sqh = sqh;
  a.out = d_out, a.only_comb = only_comb;
  a.alpha_s = nullptr, a.seg_count = nullptr, a.seg_size = 0, a.nseg = 0;
  for (int i = 0; i < MFP_MAX_PARAMS; ++i) a.par[i] = i < P::NPAR ? par[i] : 0.0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (nevt + P::BLOCK - 1) / P::BLOCK;
  const long long cap = (long long)sms * P::MINBLOCKS * 8;
  if (blocks > cap) blocks = cap;  // grid-stride; a whole multiple of the resident set
  smatrix_kernel<P><<<(unsigned)blocks, P::BLOCK, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("smatrix_kernel launch", e);
  return 0;
}

template <class P>
int integrand_blocks() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = 1;
  cudaFuncSetAttribute(integrand_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)sizeof(IntegrandSmem<P>));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, integrand_kernel<P>, P::BLOCK, sizeof(IntegrandSmem<P>));
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

template <class P>
int prepare_integrand_args(const mfp_integrand_args* u, IntegrandArgs& a) {
  constexpr int NOUT = P::NEXT - 2;
  if (u->nevents <= 0) return fail_msg("mfp_integrand: nevents must be positive");
  if (u->ncuts < 0 || u->ncuts > MFP_MAX_CUTS) return fail_msg("mfp_integrand: too many cuts");
  if (u->nblocks <= 0) return fail_msg("mfp_integrand: nblocks must come from mfp_integrand_blocks()");
  if (!u->d_grid || !u->d_partial) return fail_msg("mfp_integrand: null grid or partial buffer");
  if (u->skip_accumulate && !u->d_workspace)
    return fail_msg("mfp_integrand: skip_accumulate needs the event buffer of the helicity-parallel flavour");
  if (u->alpha_mode < 0 || u->alpha_mode > 2) return fail_msg("mfp_integrand: alpha_mode must be 0, 1 or 2");
  if (u->alpha_mode == 2 && !u->d_pdf) return fail_msg("mfp_integrand: alpha_mode 2 needs the PDF table (d_pdf)");
  if (u->d_pdf && (u->nchannels <= 0 || u->nchannels > MFP_MAX_CHANNELS))
    return fail_msg("mfp_integrand: a PDF table needs 1..MFP_MAX_CHANNELS initial-state channels");
  a.u = *u;
  double msum = 0.0;
  for (int i = 0; i < NOUT; ++i) msum += u->masses[i];
  a.massive = msum != 0.0;
  a.shat_min = msum * msum;
  a.ps.pi = u->pi, a.ps.acc = u->acc, a.ps.gev2pb = u->gev2pb;
  a.ps.wt0 = std::log(u->pi / 2.0) * (NOUT - 1) - 2.0 * std::lgamma((double)(NOUT - 1)) - std::log((double)(NOUT - 1));
  a.ps.inv_norm = 1.0 / std::pow(2 * u->pi, 3 * NOUT - 4);
  a.cuts.n = u->ncuts;
  for (int i = 0; i < u->ncuts; ++i) {
    const mfp_cut& c = u->cuts[i];
    const int pi_ = c.particle & 0xff, pj_ = c.var >= CUT_MIJ ? (c.particle >> 8) & 0xff : 0;
    if (c.particle < 0 || pi_ >= P::NEXT || pj_ >= P::NEXT || (c.var < CUT_MIJ && c.particle >= P::NEXT))
      return fail_msg("mfp_integrand: cut on a non-existent particle");
    if (c.var < 0 || c.var > CUT_DR) return fail_msg("mfp_integrand: unknown cut variable");
    a.cuts.c[i] = Cut{c.var, c.particle, c.has_min, c.has_max, c.vmin, c.vmax};
  }
  return 0;
}

template <class P>
int launch_integrand(const mfp_integrand_args* u, cudaStream_t st) {
  if (u->skip_accumulate) return fail_msg("mfp_integrand: skip_accumulate needs the helicity-parallel flavour");
  IntegrandArgs a;
  if (int rc = prepare_integrand_args<P>(u, a)) return rc;
  cudaError_t e = cudaFuncSetAttribute(integrand_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(IntegrandSmem<P>));
  if (e != cudaSuccess) return fail("integrand_kernel smem attribute", e);
  integrand_kernel<P><<<u->nblocks, P::BLOCK, sizeof(IntegrandSmem<P>), st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("integrand_kernel launch", e);
  return 0;
}

// Host-buffer entry point (mfp_smatrix_host): pageable host arrays in, pageable host array out.  The events go
// through the device in chunks over two streams with PINNED staging buffers that the library keeps between calls
// (no cudaMalloc / cudaFree and no pageable cudaMemcpy per call): while chunk i computes, chunk i+1 is staged and
// copied.  One pipeline per process library, calls are serialised.
struct HostPipe {
  static constexpr long long CHUNK = 1 << 17;
  double *d_p[2] = {nullptr, nullptr}, *d_c[2] = {nullptr, nullptr}, *d_o[2] = {nullptr, nullptr};
  double *h_p[2] = {nullptr, nullptr}, *h_c[2] = {nullptr, nullptr}, *h_o[2] = {nullptr, nullptr};
  cudaStream_t st[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  bool ready = false;
  int init(size_t pb, size_t cb) {
    if (ready) return 0;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      if (e == cudaSuccess) e = cudaMalloc(&d_p[i], pb);
      if (e == cudaSuccess) e = cudaMalloc(&d_o[i], CHUNK * sizeof(double));
      if (e == cudaSuccess && cb) e = cudaMalloc(&d_c[i], cb);
      if (e == cudaSuccess) e = cudaMallocHost(&h_p[i], pb);
      if (e == cudaSuccess) e = cudaMallocHost(&h_o[i], CHUNK * sizeof(double));
      if (e == cudaSuccess && cb) e = cudaMallocHost(&h_c[i], cb);
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return fail("mfp_smatrix_host: staging buffers", e);
    ready = true;
    return 0;
  }
};

template <class P, class Launch>
int smatrix_host(Launch launch, const double* h_p, int layout, long long nevt, const double* par,
                 const double* h_coup, long long coup_stride, double sqh, double* h_out) {
  if (nevt <= 0) return 0;
  if (layout != MFP_LAYOUT_AOS && layout != MFP_LAYOUT_SOA) return fail_msg("mfp_smatrix_host: unknown layout");
  static HostPipe pipe;
  static std::mutex lock;
  std::lock_guard<std::mutex> guard(lock);
  constexpr long long CH = HostPipe::CHUNK;
  constexpr int ROWS = P::NEXT * 4;
  const size_t pb = (size_t)CH * ROWS * sizeof(double);
  const size_t cb = (size_t)CH * P::NCOUP * 2 * sizeof(double);   // per-event couplings; frozen ones use the first slot
  if (int rc = pipe.init(pb, cb)) return rc;
  cudaError_t e;
  const long long nchunks = (nevt + CH - 1) / CH;
  auto drain = [&](long long c) -> int {   // results of chunk c: wait, then pinned -> caller's array
    const int s = (int)(c & 1);
    if ((e = cudaEventSynchronize(pipe.done[s])) != cudaSuccess) return fail("mfp_smatrix_host: chunk", e);
    const long long off = c * CH, n = (nevt - off) < CH ? (nevt - off) : CH;
    memcpy(h_out + off, pipe.h_o[s], (size_t)n * sizeof(double));
    return 0;
  };
  for (long long c = 0; c < nchunks; ++c) {
    const int s = (int)(c & 1);
    if (c >= 2)
      if (int rc = drain(c - 2)) return rc;   // frees staging set s
    const long long off = c * CH, n = (nevt - off) < CH ? (nevt - off) : CH;
    if (layout == MFP_LAYOUT_AOS) {
      memcpy(pipe.h_p[s], h_p + off * ROWS, (size_t)n * ROWS * sizeof(double));
    } else {
      for (int r = 0; r < ROWS; ++r) memcpy(pipe.h_p[s] + (long long)r * n, h_p + (long long)r * nevt + off, (size_t)n * sizeof(double));
    }
    cudaMemcpyAsync(pipe.d_p[s], pipe.h_p[s], (size_t)n * ROWS * sizeof(double), cudaMemcpyHostToDevice, pipe.st[s]);
    if (P::NCOUP > 0) {
      if (coup_stride) {
        for (int k = 0; k < P::NCOUP; ++k)
          memcpy(pipe.h_c[s] + 2 * (long long)k * n, h_coup + 2 * ((long long)k * nevt + off), (size_t)n * 2 * sizeof(double));
        cudaMemcpyAsync(pipe.d_c[s], pipe.h_c[s], (size_t)n * P::NCOUP * 2 * sizeof(double), cudaMemcpyHostToDevice, pipe.st[s]);
      } else {
        memcpy(pipe.h_c[s], h_coup, (size_t)P::NCOUP * 2 * sizeof(double));
        cudaMemcpyAsync(pipe.d_c[s], pipe.h_c[s], (size_t)P::NCOUP * 2 * sizeof(double), cudaMemcpyHostToDevice, pipe.st[s]);
      }
    }
    if (int rc = launch(pipe.d_p[s], layout, n, par, pipe.d_c[s], coup_stride, sqh, pipe.d_o[s], -1, pipe.st[s])) return rc;
    cudaMemcpyAsync(pipe.h_o[s], pipe.d_o[s], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, pipe.st[s]);
    if ((e = cudaEventRecord(pipe.done[s], pipe.st[s])) != cudaSuccess) return fail("mfp_smatrix_host: record", e);
  }
  for (long long c = nchunks >= 2 ? nchunks - 2 : 0; c < nchunks; ++c)
    if (int rc = drain(c)) return rc;
  return 0;
}

}  // namespace mf
