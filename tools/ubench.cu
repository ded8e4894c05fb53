// Micro-benchmarks behind the kernel design decisions in DESIGN.md (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench
// 1. FP64 DFMA rate vs. resident warps per SM and independent chains per thread (latency hiding)
// 2. FP64 tensor-core rate: mma.sync.aligned.m8n8k4.f64 (DMMA), same sweep
// 3. shared-memory LDS.128 rate for broadcast-heavy access (8 distinct addresses per warp)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c0[i] = threadIdx.x * 1e-3 + i, c1[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}

// DMMA correctness + fragment layout check: C = A(8x4) * B(4x8)
__global__ void dmma_check(double* out) {
  const int lane = threadIdx.x;
  const int row = lane >> 2, k = lane & 3;  // A[row][k]; B[k][col = lane>>2]
  double a = 1.0 + row * 10 + k;            // A[r][k] = 1 + 10 r + k
  double b = (k + 1) * ((lane >> 2) + 1) * 0.5;  // B[k][c] = (k+1)(c+1)/2
  double c0 = 0, c1 = 0;
  dmma(c0, c1, a, b);
  out[lane * 2] = c0;      // expected C[row][2*(lane&3)]
  out[lane * 2 + 1] = c1;  // expected C[row][2*(lane&3)+1]
}

// half of the warps run DFMA chains, the other half DMMA chains: do the two share one pipe?
__global__ void mixed_kernel(double* out, int iters, double a, double b, int mode /*0 both, 1 dfma only, 2 dmma only*/) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (((warp >> 2) & 1) == 0) {  // warps 0-3, 8-11: one DFMA and one DMMA warp group on every SM sub-partition
    if (mode == 2) return;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < 4 * iters; ++it) {  // x4: about the same pipe time as the DMMA warps
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
  } else {
    if (mode == 1) return;
    double c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) c0[i] = threadIdx.x * 1e-3 + i, c1[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dmma(c0[i], c1[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
  }
  if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void lds_kernel(double* out, int iters, int stride_mask) {
  __shared__ double2 buf[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(i, -i);
  __syncthreads();
  double s = 0;
  int idx = (threadIdx.x & stride_mask);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      double2 v = buf[(idx + i * 37 + it) & 2047];
      s += v.x;
    }
  }
  if (s == 12345.678) out[0] = s;
}

template <class F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %.0f MHz\n", prop.name, sms, prop.clockRate / 1e3);
  double* out;
  CK(cudaMalloc(&out, 1024));
  {
    double h[64];
    dmma_check<<<1, 32>>>(out);
    CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int lane = 0; lane < 32; ++lane)
      for (int j = 0; j < 2; ++j) {
        const int r = lane >> 2, c = 2 * (lane & 3) + j;
        double ref = 0;
        for (int k = 0; k < 4; ++k) ref += (1.0 + r * 10 + k) * ((k + 1) * (c + 1) * 0.5);
        const double d = h[lane * 2 + j] - ref;
        maxerr = d * d > maxerr ? d * d : maxerr;
      }
    printf("dmma fragment layout check: max sq err %.3g (A[lane/4][lane%%4], B[lane%%4][lane/4], C[lane/4][2*(lane%%4)+{0,1}])\n", maxerr);
  }
  const int iters = 20000;
  printf("\nDFMA: TFLOP/s (2 flop per FMA) vs warps/SM (rows) and independent chains per thread (cols 1,2,4,8)\n");
  for (int warps : {4, 8, 12, 16, 32, 64}) {
    printf("  warps/SM %2d:", warps);
    const int threads = warps >= 32 ? 1024 : warps * 32, blocks = sms * (warps >= 32 ? warps / 32 : 1);
    float ms;
    ms = time_it([&] { dfma_kernel<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(" %7.2f", 2.0 * 1 * iters * (double)blocks * threads / ms / 1e9);
    ms = time_it([&] { dfma_kernel<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(" %7.2f", 2.0 * 2 * iters * (double)blocks * threads / ms / 1e9);
    ms = time_it([&] { dfma_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(" %7.2f", 2.0 * 4 * iters * (double)blocks * threads / ms / 1e9);
    ms = time_it([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(" %7.2f\n", 2.0 * 8 * iters * (double)blocks * threads / ms / 1e9);
  }
  printf("\nDMMA m8n8k4: TFLOP/s (512 flop per warp instruction) vs warps/SM and independent accumulators (1,2,4,8)\n");
  for (int warps : {4, 8, 12, 16, 32, 64}) {
    printf("  warps/SM %2d:", warps);
    const int threads = warps >= 32 ? 1024 : warps * 32, blocks = sms * (warps >= 32 ? warps / 32 : 1);
    const double nw = (double)blocks * threads / 32;
    float ms;
    ms = time_it([&] { dmma_kernel<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(" %7.2f", 512.0 * 1 * iters * nw / ms / 1e9);
    ms = time_it([&] { dmma_kernel<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(" %7.2f", 512.0 * 2 * iters * nw / ms / 1e9);
    ms = time_it([&] { dmma_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(" %7.2f", 512.0 * 4 * iters * nw / ms / 1e9);
    ms = time_it([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf(" %7.2f\n", 512.0 * 8 * iters * nw / ms / 1e9);
  }
  printf("\nDFMA and DMMA side by side (16 warps/SM, half the warps of every sub-partition DFMA x8 chains, the others DMMA x4 chains): ms for the same iteration count\n");
  {
    const int threads = 512, blocks = sms;
    for (int mode : {1, 2, 0}) {
      float ms = time_it([&] { mixed_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, mode); });
      const double nw = (double)blocks * threads / 32 / 2;
      const double dfma_tf = mode != 2 ? 2.0 * 8 * 4 * iters * nw * 32 / ms / 1e9 : 0.0;
      const double dmma_tf = mode != 1 ? 512.0 * 4 * iters * nw / ms / 1e9 : 0.0;
      printf("  mode %s: %.3f ms  DFMA %.2f TFLOP/s + DMMA %.2f TFLOP/s = %.2f\n", mode == 0 ? "both" : (mode == 1 ? "DFMA only" : "DMMA only"), ms,
             dfma_tf, dmma_tf, dfma_tf + dmma_tf);
    }
  }
  printf("\nLDS.128: bytes/clk/SM delivered to lanes (512 B per warp instruction), 8 warps/SM, ILP 8; distinct addresses per warp 1/8/32\n");
  for (int mask : {0, 7, 31}) {
    const int threads = 256, blocks = sms;
    float ms = time_it([&] { lds_kernel<8><<<blocks, threads>>>(out, iters, mask); });
    const double instr = 8.0 * iters * blocks * threads / 32;
    printf("  mask %2d: %.1f warp-instr/us/SM -> %.2f instr/clk/SM at %.0f MHz\n", mask, instr / ms / 1e3 / sms,
           instr / ms / 1e3 / sms / (prop.clockRate / 1e3), prop.clockRate / 1e3);
  }
  return 0;
}
