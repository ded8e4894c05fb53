// Minimal FP64 complex type for the matrix-element device code.
// A plain {re, im} pair: every product below is written so that nvcc contracts it into DFMA.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>

// every device function is also host-callable so tests can run the very same code on the CPU
#define MF_DEV __host__ __device__ __forceinline__
#define MF_HD __host__ __device__ __forceinline__

// Product that is never contracted into an FMA.  Used where the reference's expression is
// ill-conditioned (external wavefunctions of forward/backward particles, P^2 of nearly on-shell
// propagators): there a differently rounded intermediate is amplified by E/(E+pz) or E^2/P^2, so
// these few operations are rounded exactly like the reference's separate multiply and add.
MF_HD double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  double r = a * b;
#ifdef __FMA__
  asm volatile("" : "+x"(r));
#endif
  return r;
#endif
}

struct cxd {
  double re, im;
};

MF_HD cxd mk(double re, double im) { return cxd{re, im}; }
MF_HD cxd operator+(cxd a, cxd b) { return cxd{a.re + b.re, a.im + b.im}; }
MF_HD cxd operator-(cxd a, cxd b) { return cxd{a.re - b.re, a.im - b.im}; }
MF_HD cxd operator-(cxd a) { return cxd{-a.re, -a.im}; }
MF_HD cxd operator*(cxd a, cxd b) { return cxd{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
MF_HD cxd operator*(double a, cxd b) { return cxd{a * b.re, a * b.im}; }
MF_HD cxd operator*(cxd a, double b) { return cxd{a.re * b, a.im * b}; }
MF_HD cxd& operator+=(cxd& a, cxd b) {
  a.re += b.re;
  a.im += b.im;
  return a;
}
MF_HD cxd& operator-=(cxd& a, cxd b) {
  a.re -= b.re;
  a.im -= b.im;
  return a;
}
MF_HD cxd conj(cxd a) { return cxd{a.re, -a.im}; }
MF_HD cxd mul_i(cxd a) { return cxd{-a.im, a.re}; }    // i*a, exact
MF_HD cxd mul_mi(cxd a) { return cxd{a.im, -a.re}; }   // -i*a, exact
MF_HD double norm2(cxd a) { return a.re * a.re + a.im * a.im; }
// a + b*c and a - b*c with the four DFMAs explicit
MF_HD cxd fma_c(cxd b, cxd c, cxd a) {
  return cxd{a.re + b.re * c.re - b.im * c.im, a.im + b.re * c.im + b.im * c.re};
}
MF_HD cxd cdiv(cxd a, cxd b) {
  const double inv = 1.0 / (b.re * b.re + b.im * b.im);
  return cxd{(a.re * b.re + a.im * b.im) * inv, (a.im * b.re - a.re * b.im) * inv};
}
