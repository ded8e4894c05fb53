// Event-generation and accumulation kernels of the three-stage integrand pipeline used by the
// helicity-parallel flavour:
//
//   ps_generate_kernel<NEXT>   Philox -> VEGAS map -> x1,x2 -> RAMBO -> cuts -> boost -> scale, alpha_s, luminosity;
//                              one event per thread; every block appends the events that pass the
//                              cuts to ITS OWN segment of the event buffer (stable order, so the
//                              buffer contents do not depend on scheduling) and zero-fills the rest
//   smatrix_kernel_hp<P>       matrix elements of the accepted events, segment by segment
//   accumulate_kernel          t = |M|^2 * w; per-block sums and histogram (fixed order reduction
//                              afterwards by vegas_reduce_kernel)
//
// The reference does the same three steps as separate TensorFlow graphs with a boolean_mask
// compaction in between (phasespace.py:506-515, madflow_exec.py:444-468).  Buffers (HBM), for a
// capacity of `cap` event slots split into `nseg` segments of `seg` slots:
//   mom    (cap, NEXT, 4) f64     w (cap) f64 = xjac * ps weight * lumi      as (cap) f64 alpha_s
//   bins   (ndim, cap) u8         me (cap) f64                                count (nseg) i32
#pragma once
#include "pdf.cuh"
#include "phasespace.cuh"
#include "philox.cuh"
#include "vegas.cuh"
#include "../../include/madflow_b200_process.h"

namespace mf {

struct EventBuffer {
  double* mom;
  double* w;
  double* as;
  double* me;
  unsigned char* bins;
  int* count;
  long long cap;   // total slots = nseg * seg
  long long seg;   // slots per segment (multiple of the generator's block size)
  long long per;   // generated events per segment = ceil(nevents / nseg): the segments are filled evenly
  int nseg;
};

struct GenArgs {
  mfp_integrand_args u;
  int massive;
  double shat_min;
  PSConst ps;
  CutList cuts;
  EventBuffer buf;
};

constexpr int GEN_BLOCK = 128;

template <int NEXT>
__global__ void __launch_bounds__(GEN_BLOCK) ps_generate_kernel(const GenArgs a) {
  constexpr int NDIM = 4 * (NEXT - 2) + 2, B = GEN_BLOCK, NWARP = B / 32;
  __shared__ double sgrid[NDIM * VEGAS_EDGES];
  __shared__ int warp_count[NWARP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NDIM * VEGAS_EDGES; i += B) sgrid[i] = a.u.d_grid[i];
  __syncthreads();
  const EventBuffer& q = a.buf;
  for (int sgi = blockIdx.x; sgi < q.nseg; sgi += gridDim.x) {
    const long long ev_begin = (long long)sgi * q.per;  // segment sgi generates events [ev_begin, ev_begin + per)
    const long long base = (long long)sgi * q.seg;      // ... and owns slots [base, base + seg)
    int filled = 0;
    for (long long off = 0; off < q.per; off += B) {
      const long long local = ev_begin + off + tid;
      bool ok = false;
      double m[NEXT][4];
      double wgt = 0.0, as = 0.0;
      unsigned char bins[NDIM];
      if (off + tid < q.per && local < a.u.nevents) {
        const unsigned long long ev = a.u.first_event + (unsigned long long)local;
        double xr[NDIM];
        double w = 1.0;
#pragma unroll
        for (int j = 0; j < (NDIM + 1) / 2; ++j) {
          double u0, u1;
          philox_pair(a.u.seed, a.u.iteration, ev, j, u0, u1);
          int b;
          xr[2 * j] = vegas_map(&sgrid[(2 * j) * VEGAS_EDGES], vegas_confine(u0), b, w);
          bins[2 * j] = (unsigned char)b;
          if (2 * j + 1 < NDIM) {
            xr[2 * j + 1] = vegas_map(&sgrid[(2 * j + 1) * VEGAS_EDGES], vegas_confine(u1), b, w);
            bins[2 * j + 1] = (unsigned char)b;
          }
        }
        double x1, x2;
        ramboflow<NEXT>(xr, a.u.com_sqrts, a.u.masses, a.massive != 0, a.shat_min, a.ps, m, wgt, x1, x2);
        ok = pass_cuts<NEXT>(a.cuts, m);  // on centre-of-mass momenta (phasespace.py:506-508)
        ok = ok && (wgt == wgt) && (wgt != 0.0);
        if (ok) {
          if (a.u.lab_frame) boost_to_lab<NEXT>(m, x1, x2);
          double lumi;
          event_scale<NEXT>(a.u, m, x1, x2, as, lumi);  // madflow_exec.py:426-454
          wgt *= lumi;
          wgt *= w * a.u.inv_total_events;
        }
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) warp_count[warp] = __popc(ballot);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int wv = 0; wv < NWARP; ++wv) {
        const int c = warp_count[wv];
        if (wv < warp) before += c;
        total += c;
      }
      if (ok) {
        const long long slot = base + filled + before + __popc(ballot & ((1u << lane) - 1u));
#pragma unroll
        for (int i = 0; i < NEXT; ++i)
          reinterpret_cast<double4*>(q.mom)[slot * NEXT + i] = make_double4(m[i][0], m[i][1], m[i][2], m[i][3]);
        q.w[slot] = wgt;
        q.as[slot] = as;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) q.bins[(long long)d * q.cap + slot] = bins[d];
      }
      filled += total;
      __syncthreads();
    }
    // unused tail of the segment: zero weight and zero matrix element
    for (long long s = base + filled + tid; s < base + q.seg; s += B) {
      q.w[s] = 0.0;
      q.me[s] = 0.0;
#pragma unroll
      for (int d = 0; d < NDIM; ++d) q.bins[(long long)d * q.cap + s] = 0;
    }
    if (tid == 0) q.count[sgi] = filled;
  }
}

constexpr int ACC_BLOCK = 128;
// shared memory of the accumulation kernels: one histogram per warp + the reduction scratch
inline size_t accumulate_smem(int ndim) { return ((size_t)(ACC_BLOCK / 32) * ndim * VEGAS_BINS + 24) * sizeof(double); }
// t = value(e) = f * xjac; per block: [sum t, sum t^2, #(t != 0), 0] + histogram of t^2 over (dimension, bin).
// Every sum is formed in an order that depends on the launch shape only (warp_hist_add), not on scheduling.
template <class Value>
__device__ __forceinline__ void accumulate_block(Value value, const unsigned char* bins, long long nevt, int ndim,
                                                 int with_hist, double* partial) {
  extern __shared__ double shist[];  // [warp][ndim*50], then 3*8 for the reduction
  constexpr int NW = ACC_BLOCK / 32;
  const int nh = ndim * VEGAS_BINS;
  double* red = shist + NW * nh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* whist = shist + warp * nh;
  for (int i = threadIdx.x; i < NW * nh; i += blockDim.x) shist[i] = 0.0;
  __syncthreads();
  double s1 = 0.0, s2 = 0.0, cnt = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long rounds = (nevt + stride - 1) / stride;   // the same trip count for every lane of the warp
  for (long long r = 0; r < rounds; ++r) {
    const long long e = r * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double t = e < nevt ? value(e) : 0.0;
    const double t2 = t * t;
    s1 += t;
    s2 += t2;
    cnt += (t != 0.0) ? 1.0 : 0.0;
    if (with_hist && __any_sync(0xffffffffu, t2 != 0.0))
      for (int d = 0; d < ndim; ++d)
        warp_hist_add(whist + d * VEGAS_BINS, t2 != 0.0 ? bins[(long long)d * nevt + e] : 0, t2, t2 != 0.0);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) red[warp] = s1, red[8 + warp] = s2, red[16 + warp] = cnt;
  __syncthreads();
  double* out = partial + (long long)blockIdx.x * (VEGAS_HEADER + ndim * VEGAS_BINS);
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int w = 0; w < NW; ++w) a += red[w], b += red[8 + w], c += red[16 + w];
    out[0] = a, out[1] = b, out[2] = c, out[3] = 0.0;
  }
  for (int i = threadIdx.x; i < nh; i += blockDim.x) {
    double h = 0.0;
    for (int w = 0; w < NW; ++w) h += shist[w * nh + i];   // warp order
    out[VEGAS_HEADER + i] = h;
  }
}

__global__ void __launch_bounds__(ACC_BLOCK) accumulate_kernel(const double* f, const double* xjac,
                                                               const unsigned char* bins, long long nevt, int ndim,
                                                               int with_hist, double* partial) {
  accumulate_block([=](long long e) { return f[e] * xjac[e]; }, bins, nevt, ndim, with_hist, partial);
}

// several subprocesses on the same events (madflow_exec.py:444-455: ret += luminosity_i * smatrix_i):
// t = sum_i f_i * w_i with w_i = xjac * phase-space weight * luminosity_i
constexpr int ACC_MAX_TERMS = 16;   // p p > t t~ j j has 12 subprocesses
struct AccTerms {
  int n;
  const double* f[ACC_MAX_TERMS];
  const double* w[ACC_MAX_TERMS];
};
__global__ void __launch_bounds__(ACC_BLOCK) accumulate_sum_kernel(const AccTerms terms, const unsigned char* bins,
                                                                   long long nevt, int ndim, int with_hist,
                                                                   double* partial) {
  accumulate_block(
      [=](long long e) {
        double t = 0.0;
        for (int i = 0; i < terms.n; ++i) t += terms.f[i][e] * terms.w[i][e];
        return t;
      },
      bins, nevt, ndim, with_hist, partial);
}

// carve the caller's workspace into the event buffer; returns bytes needed (buf may be null)
inline size_t event_buffer_layout(long long nevents, int nseg, int next, int ndim, void* workspace, EventBuffer* buf) {
  const long long per = (nevents + nseg - 1) / nseg;
  const long long seg = ((per + GEN_BLOCK - 1) / GEN_BLOCK) * GEN_BLOCK;
  const long long cap = seg * nseg;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  const size_t o_mom = take((size_t)cap * next * 4 * sizeof(double));
  const size_t o_w = take((size_t)cap * sizeof(double));
  const size_t o_as = take((size_t)cap * sizeof(double));
  const size_t o_me = take((size_t)cap * sizeof(double));
  const size_t o_bins = take((size_t)cap * ndim);
  const size_t o_cnt = take((size_t)nseg * sizeof(int));
  if (buf && workspace) {
    char* p = static_cast<char*>(workspace);
    buf->mom = reinterpret_cast<double*>(p + o_mom);
    buf->w = reinterpret_cast<double*>(p + o_w);
    buf->as = reinterpret_cast<double*>(p + o_as);
    buf->me = reinterpret_cast<double*>(p + o_me);
    buf->bins = reinterpret_cast<unsigned char*>(p + o_bins);
    buf->count = reinterpret_cast<int*>(p + o_cnt);
    buf->cap = cap, buf->seg = seg, buf->per = per, buf->nseg = nseg;
  }
  return off;
}

}  // namespace mf
