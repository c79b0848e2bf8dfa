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

#include "nnpm_butterfly.cuh"   // zf_cos32 / zf_brev, ztw, ZDif / ZDit: the register-resident radix-R sub-transforms

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
