// nnpm_butterfly.cuh -- the register-resident radix-R (R = 2 .. 32) sub-transforms every own FFT pass is built from
// (x passes nnpm_rowfft.cuh, y passes and the fused z pass nnpm_colfft.cuh): compile-time twiddles exp(+-2 pi i m / 32),
// decimation-in-frequency (natural in, bit-reversed out) and decimation-in-time (bit-reversed in, natural out) networks
// unrolled by template recursion.  They implement the rfft / irfft of compute_potential (src/ParticleMesh.jl:481, :502)
// together with the inter-pass twiddles of the kernels that use them.
//
// Pure arithmetic on double2 with no CUDA intrinsics, so the same text also compiles as host C++: tests/test_butterfly_host.py
// builds it with g++ (shim: __device__ / __host__ / __forceinline__ defined away, a plain double2) and checks every radix
// and both signs against the DFT definition without a GPU.
#pragma once

__host__ __device__ constexpr double zf_cos32(int m) {  // cos(2 pi m / 32), any m
  constexpr double ZF_COS32[9] = {1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025452, 0.7071067811865476,
                                  0.5555702330196023, 0.38268343236508984, 0.19509032201612833, 0.0};
  m = ((m % 32) + 32) % 32;
  if (m > 16) m = 32 - m;                 // cos(-x) = cos(x)
  return m <= 8 ? ZF_COS32[m] : -ZF_COS32[16 - m];
}
__host__ __device__ constexpr double zf_sin32(int m) { return zf_cos32(m - 8); }  // sin(x) = cos(x - pi/2)
__host__ __device__ constexpr int zf_log2(int r) { return r <= 1 ? 0 : 1 + zf_log2(r / 2); }
__host__ __device__ constexpr int zf_brev(int q, int bits) {
  int o = 0;
  for (int b = 0; b < bits; b++) o |= ((q >> b) & 1) << (bits - 1 - b);
  return o;
}

__device__ __forceinline__ double2 zadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 zsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 zmul(double2 a, double c, double s) {  // a * (c + i s)
  return make_double2(fma(a.x, c, -a.y * s), fma(a.x, s, a.y * c));
}
// a * exp(SIGN * 2 pi i m32 / 32) with the trivial cases folded at compile time
template <int M32, int SIGN>
__device__ __forceinline__ double2 ztw(double2 a) {
  constexpr int m = ((M32 % 32) + 32) % 32;
  if (m == 0) return a;
  if (m == 16) return make_double2(-a.x, -a.y);
  if (m == 8) return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);   // * (+-i)
  if (m == 24) return SIGN > 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
  constexpr double c = zf_cos32(m), s = (SIGN > 0 ? 1.0 : -1.0) * zf_sin32(m);
  return zmul(a, c, s);
}

// decimation in frequency: natural-order input, bit-reversed output; a[BASE .. BASE+R)
template <int R, int BASE, int SIGN>
struct ZDif {
  template <int I>
  static __device__ __forceinline__ void bf(double2* a) {
    if constexpr (I < R / 2) {
      const double2 u = a[BASE + I], v = a[BASE + I + R / 2];
      a[BASE + I] = zadd(u, v);
      a[BASE + I + R / 2] = ztw<I*(32 / R), SIGN>(zsub(u, v));
      bf<I + 1>(a);
    }
  }
  static __device__ __forceinline__ void run(double2* a) {
    if constexpr (R >= 2) {
      bf<0>(a);
      ZDif<R / 2, BASE, SIGN>::run(a);
      ZDif<R / 2, BASE + R / 2, SIGN>::run(a);
    }
  }
};
// decimation in time: bit-reversed input, natural-order output
template <int R, int BASE, int SIGN>
struct ZDit {
  template <int I>
  static __device__ __forceinline__ void bf(double2* a) {
    if constexpr (I < R / 2) {
      const double2 e = a[BASE + I], o = ztw<I*(32 / R), SIGN>(a[BASE + I + R / 2]);
      a[BASE + I] = zadd(e, o);
      a[BASE + I + R / 2] = zsub(e, o);
      bf<I + 1>(a);
    }
  }
  static __device__ __forceinline__ void run(double2* a) {
    if constexpr (R >= 2) {
      ZDit<R / 2, BASE, SIGN>::run(a);
      ZDit<R / 2, BASE + R / 2, SIGN>::run(a);
      bf<0>(a);
    }
  }
};
