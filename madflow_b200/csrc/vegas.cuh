// VEGAS grid mapping (device side).  Replaces vegasflow's `_generate_random_array`
// ([EXT] vegasflow 1.x, not under /root/reference; call sites scripts/madflow_exec.py:487-525).
//   u in (1e-8, 1-1e-8);  xn = BINS*(1-u);  k = floor(xn);  x = lo_k + (hi_k - lo_k)*(xn - k);
//   weight *= BINS*(hi_k - lo_k).      Grid: edges[ndim][BINS+1], edges[d][0]=0, edges[d][BINS]=1.
#pragma once
#include "mf_complex.cuh"

namespace mf {

constexpr int VEGAS_BINS = 50;
constexpr int VEGAS_EDGES = VEGAS_BINS + 1;
constexpr double VEGAS_TECH_CUT = 1e-8;
// accumulator layout: [sum t, sum t^2, events that reached the matrix element, reserved] + histogram
constexpr int VEGAS_HEADER = 4;

MF_DEV double vegas_confine(double u) { return VEGAS_TECH_CUT + u * (1.0 - 2.0 * VEGAS_TECH_CUT); }

// edges: this dimension's VEGAS_EDGES numbers (shared or global memory)
MF_DEV double vegas_map(const double* edges, double r, int& bin, double& w) {
  const double xn = VEGAS_BINS * (1.0 - r);
  int k = (int)xn;
  k = k < 0 ? 0 : (k > VEGAS_BINS - 1 ? VEGAS_BINS - 1 : k);
  const double lo = edges[k], hi = edges[k + 1];
  const double delta = hi - lo;
  bin = k;
  w *= delta * VEGAS_BINS;
  return lo + delta * (xn - k);
}

// Deterministic histogram accumulation.  hist[bin] += v for the lanes of one warp, in a FIXED order: the lanes that
// hit the same bin are summed by the lowest of them in ascending lane order and the sum is added with a plain store
// to a histogram that only THIS WARP writes (no floating-point atomics: their order -- and with it the last bits of
// the VEGAS grid, which the refinement amplifies on a spiky integrand -- changed from run to run).  The warp-private
// histograms of a block are merged in warp order, the blocks in block order (vegas_reduce_kernel).
// All 32 lanes must call; `active` = this lane has something to add.
#ifdef __CUDACC__
__device__ __forceinline__ void warp_hist_add(double* whist, int bin, double v, bool active) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(full, active ? bin : -1);
  const int n = __popc(peers);
  const int nmax = __reduce_max_sync(full, active ? n : 0);
  double sum = 0.0;
  unsigned rem = peers;
  for (int k = 0; k < nmax; ++k) {
    const int src = rem ? __ffs(rem) - 1 : lane;  // the k-th lane of my group
    rem &= rem - 1;
    const double x = __shfl_sync(full, v, src);
    if (k < n) sum += x;
  }
  if (active && lane == __ffs(peers) - 1) whist[bin] += sum;
  __syncwarp();
}
#endif

}  // namespace mf
