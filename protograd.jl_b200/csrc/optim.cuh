// optim.cuh — per-element optimizer arithmetic shared by the arena kernels (arena_kernels.cu) and the fused
// gradient-exchange + update kernel over NVLink peer memory (dp.cu).  src/optimizers/adam.jl:32-41.
#pragma once
#include "common.cuh"

namespace {

template <typename P> struct Ar;
template <> struct Ar<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Ar<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

// adam: in = {p, m, v, g}; out = {p, m, v, m_hat, v_hat}
template <typename T, typename P> struct AdamF {
  P s, b1, omb1, b2, omb2, c1, c2, eps;  // c1 = 1 - b1^it, c2 = 1 - b2^it
  const double* dev;                     // non-null: c1, c2 come from device scalars dev[1], dev[2]
  __device__ __forceinline__ void operator()(const T* in, T* out) const {
    const P c1 = dev ? (P)dev[1] : this->c1, c2 = dev ? (P)dev[2] : this->c2;
    T g = in[3];
    T m = (T)Ar<P>::add(Ar<P>::mul(b1, (P)in[1]), Ar<P>::mul(omb1, (P)g));
    T g2 = Ar<T>::mul(g, g);
    T v = (T)Ar<P>::add(Ar<P>::mul(b2, (P)in[2]), Ar<P>::mul(omb2, (P)g2));
    T mh = (T)Ar<P>::div((P)m, c1);
    T vh = (T)Ar<P>::div((P)v, c2);
    P den = Ar<P>::add((P)sqrt_rn(vh), eps);
    P q = Ar<P>::div((P)mh, den);
    out[0] = (T)Ar<P>::sub((P)in[0], Ar<P>::mul(s, q));
    out[1] = m; out[2] = v; out[3] = mh; out[4] = vh;
  }
};
// fast f32 Adam: reciprocal bias corrections, approximate division (normwise ~1e-7 of exact)
struct AdamFastF {
  float s, b1, omb1, b2, omb2, rc1, rc2, eps;
  const double* dev;
  __device__ __forceinline__ void operator()(const float* in, float* out) const {
    const float rc1 = dev ? (float)(1.0 / dev[1]) : this->rc1, rc2 = dev ? (float)(1.0 / dev[2]) : this->rc2;
    float g = in[3];
    float m = fmaf(b1, in[1], omb1 * g);
    float v = fmaf(b2, in[2], omb2 * g * g);
    float mh = m * rc1, vh = v * rc2;
    out[0] = in[0] - s * __fdividef(mh, sqrtf(vh) + eps);
    out[1] = m; out[2] = v; out[3] = mh; out[4] = vh;
  }
};


}  // namespace
