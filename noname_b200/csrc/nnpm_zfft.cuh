// nnpm_zfft.cuh -- K3: fused [z-FFT -> Green's function multiply -> z-iFFT] in ONE pass over HBM.
//
// compute_potential (src/ParticleMesh.jl:473-505) is rfft -> divide by Ginv -> irfft.  The x/y
// parts of the 3-D transform are done by cuFFT (batched 2-D r2c / c2r over the slab's planes);
// this kernel does everything along z on the (slab-transposed) half spectrum T[k][col], col =
// (jl, ic) flattened: it reads C adjacent columns, transforms them along k with a two-pass
// Stockham FFT of length N = R1*R2 (radix-R sub-transforms held entirely in registers, one
// shared-memory exchange between the passes), multiplies by scale/Ginv(i,j,k), transforms back
// and writes the columns in place -- 32 B/element of HBM traffic instead of the 96 B/element of
// [cuFFT z forward, Green kernel, cuFFT z inverse].
//
//   forward : pass A  radix R1, input rows j + r*R2 straight from global memory (DIF: natural in,
//                     bit-reversed out) -> smem[(j*R1 + q)]
//             pass B  radix R2 on smem[j + r*R1] * W_N^(r*j) (DIF) -> thread j holds kz = j + q*R1
//   Green   : in registers, on the bit-reversed layout
//   inverse : pass A' radix R2 (DIT: bit-reversed in, natural out) -> smem[(j*R2 + r)]
//             pass B' radix R1 on smem[j + r*R2] * conj(W_N^(r*j)) (DIF inverse) -> rows j + q*R2
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

// 1/x for normal, finite x: hardware seed (MUFU.RCP64H) + two Newton steps, error < 1 ulp-ish;
// skips the special-case handling of a full IEEE division (|Ginv| lies in [4 sin^2(pi/N), 12]).
__device__ __forceinline__ double zf_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

static bool zfused_factor(i64 N, int* r1, int* r2) {
  switch (N) {
    case 64: *r1 = 8; *r2 = 8; return true;
    case 128: *r1 = 16; *r2 = 8; return true;
    case 256: *r1 = 16; *r2 = 16; return true;
    case 512: *r1 = 32; *r2 = 16; return true;
    case 1024: *r1 = 32; *r2 = 32; return true;
    default: return false;
  }
}
static bool zfused_supported(nnpm_ctx* ctx) {
  int a, b;
  if (ctx->R != 3) return false;
  if (zfused_factor(ctx->g.N3, &a, &b)) return true;
  // N3 = 2 x a two-pass length (2048): the radix-2 wrapper of the persistent kernel (k_zfused_p<.., TWO>)
  return ctx->g.N3 % 2 == 0 && zfused_factor(ctx->g.N3 / 2, &a, &b);
}
