// Event output of the integrand pipeline: weighted histograms of kinematic observables and
// unweighting + compaction of the generated events, both on the device.
//
// The reference hands every event of every chunk to Python (tf.py_function -> per-event dicts -> MG5's
// lhe_parser -> gzip, python_package/madflow/lhe_writer.py:151-239) and histograms the LHE file afterwards
// (example/compare_mg5_hists.py:16-57).  Here the events of the pipeline's HBM buffer (momenta, weight
// factors) are histogrammed where they are, and only the events that survive the unweighting
// (probability |w| / w_max, Philox counter = global slot index) travel to the host writer.
#pragma once
#include "philox.cuh"

namespace mf {

enum Observable { OBS_PT = 0, OBS_ETA = 1, OBS_RAPIDITY = 2, OBS_ENERGY = 3, OBS_MASS = 4 };

// p = (E, px, py, pz); FourMomentum of MG5's lhe_parser (lhe_writer.py:385-433): pt, pseudorapidity, rapidity
MF_DEV double observable_value(int obs, const double p[4]) {
  const double pt2 = p[1] * p[1] + p[2] * p[2];
  switch (obs) {
    case OBS_PT: return sqrt(pt2);
    case OBS_ETA: {
      const double pabs = sqrt(pt2 + p[3] * p[3]);
      return 0.5 * log((pabs + p[3]) / (pabs - p[3]));
    }
    case OBS_RAPIDITY: return 0.5 * log((p[0] + p[3]) / (p[0] - p[3]));
    case OBS_ENERGY: return p[0];
    default: {
      const double m2 = p[0] * p[0] - pt2 - p[3] * p[3];
      return m2 > 0.0 ? sqrt(m2) : 0.0;
    }
  }
}

// bin of value v in [lo, hi) with nbins bins: 0 = underflow, 1..nbins, nbins+1 = overflow (NaN -> overflow)
MF_DEV int histogram_bin(double v, double lo, double inv_width, int nbins) {
  const double t = (v - lo) * inv_width;
  if (!(t >= 0.0)) return t < 0.0 ? 0 : nbins + 1;
  return t >= (double)nbins ? nbins + 1 : 1 + (int)t;
}

constexpr int EVH_BLOCK = 256;
constexpr int EVH_HIST_BLOCK = 128;
constexpr int EVH_MAX_BINS = 1022;

// partial[block][nbins + 2] = weights of the block's slots per bin; weight of slot i = w1[i] * (w2 ? w2[i] : 1); slots with
// weight 0 are skipped.  One histogram per warp in shared memory, filled in a fixed order (vegas.cuh::warp_hist_add)
// and merged in warp order; event_histogram_reduce_kernel then adds the blocks in block order: the result does not
// depend on scheduling (floating-point atomics did).
__global__ void __launch_bounds__(EVH_HIST_BLOCK) event_histogram_kernel(const double* mom, const double* w1, const double* w2,
                                                                        long long nevt, int next, int particle, int obs,
                                                                        double lo, double inv_width, int nbins, double* partial) {
  extern __shared__ double sh[];   // [warp][nbins + 2]
  constexpr int NW = EVH_HIST_BLOCK / 32;
  const int nb = nbins + 2;
  for (int i = threadIdx.x; i < NW * nb; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  double* whist = sh + (threadIdx.x >> 5) * nb;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long rounds = (nevt + stride - 1) / stride;
  for (long long r = 0; r < rounds; ++r) {
    const long long e = r * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double w = 0.0;
    int bin = 0;
    if (e < nevt) w = w2 ? w1[e] * w2[e] : w1[e];
    if (w != 0.0) {
      const double4 v = reinterpret_cast<const double4*>(mom)[e * next + particle];
      const double p[4] = {v.x, v.y, v.z, v.w};
      bin = histogram_bin(observable_value(obs, p), lo, inv_width, nbins);
    }
    if (__any_sync(0xffffffffu, w != 0.0)) warp_hist_add(whist, bin, w, w != 0.0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    double h = 0.0;
    for (int wv = 0; wv < NW; ++wv) h += sh[wv * nb + i];
    partial[(long long)blockIdx.x * nb + i] = h;
  }
}

__global__ void event_histogram_reduce_kernel(const double* partial, int nblocks, int nb, double* hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  double h = 0.0;
  for (int b = 0; b < nblocks; ++b) h += partial[(long long)b * nb + i];
  hist[i] += h;
}

// per block: partial[block] = {max |w|, sum |w|, sum w^2} over its slots (the caller reduces the blocks in a fixed
// order, so the statistics -- and the unweighting threshold derived from them -- do not depend on scheduling)
__global__ void __launch_bounds__(EVH_BLOCK) weight_stats_kernel(const double* w1, const double* w2, long long nevt, double* partial) {
  __shared__ double red[3][EVH_BLOCK / 32];
  double m = 0.0, s1 = 0.0, s2 = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nevt; e += stride) {
    const double w = fabs(w2 ? w1[e] * w2[e] : w1[e]);
    if (w == w) {
      m = w > m ? w : m;
      s1 += w;
      s2 += w * w;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_down_sync(0xffffffffu, m, o);
    m = other > m ? other : m;
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[0][warp] = m, red[1][warp] = s1, red[2][warp] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < EVH_BLOCK / 32; ++w) {
      m = red[0][w] > m ? red[0][w] : m;
      s1 += red[1][w];
      s2 += red[2][w];
    }
    partial[blockIdx.x * 3 + 0] = m, partial[blockIdx.x * 3 + 1] = s1, partial[blockIdx.x * 3 + 2] = s2;
  }
}

// Unweighting: slot i survives with probability |w_i| / wmax (always, if |w_i| >= wmax) and is appended to the
// output with weight sign(w_i) * max(|w_i|, wmax).  The random number depends only on (seed, first_index + i).
// out_index receives the global slot index so that the host can restore a scheduling-independent order.
__global__ void __launch_bounds__(EVH_BLOCK) select_events_kernel(const double* mom, const double* w1, const double* w2,
                                                                 long long nevt, int next, double wmax, unsigned long long seed,
                                                                 unsigned long long first_index, double* out_mom, double* out_w,
                                                                 long long* out_index, int* count, long long capacity) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nevt; e += stride) {
    const double w = w2 ? w1[e] * w2[e] : w1[e];
    if (w == 0.0 || w != w) continue;
    double u0, u1;
    philox_pair(seed, 0xFFFFFFFFu, first_index + (unsigned long long)e, 0xFFFFFFFFu, u0, u1);
    const double a = fabs(w);
    if (a < u0 * wmax) continue;
    const int slot = atomicAdd(count, 1);
    if (slot >= capacity) continue;  // the counter keeps counting: the host sees the overflow
    for (int i = 0; i < next; ++i)
      reinterpret_cast<double4*>(out_mom)[(long long)slot * next + i] = reinterpret_cast<const double4*>(mom)[e * next + i];
    out_w[slot] = (w < 0.0 ? -1.0 : 1.0) * (a > wmax ? a : wmax);
    out_index[slot] = (long long)(first_index + (unsigned long long)e);
  }
}

}  // namespace mf
