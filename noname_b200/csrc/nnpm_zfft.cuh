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

template <int R1, int R2, int C>
__global__ void __launch_bounds__(C*(R1 > R2 ? R1 : R2), 1)
k_zfused(double2* __restrict__ T, i64 ncol, i64 colbeg, i64 colend, i64 icn, i64 j0, const double* __restrict__ s1, const double* __restrict__ s2,
         const double* __restrict__ s3, const double2* __restrict__ tw, int rank3, double scale) {
  constexpr int RM = R1 > R2 ? R1 : R2, B1 = zf_log2(R1), B2 = zf_log2(R2);
  extern __shared__ double2 zsm[];  // [N][C] exchange buffer, then tw[N] (double2), then s3[N] (double)
  constexpr int N = R1 * R2;
  double2* stw = zsm + N * C;
  double* ss3 = (double*)(stw + N);
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    stw[t] = __ldg(tw + t);
    ss3[t] = __ldg(s3 + t);
  }
  const int c = threadIdx.x % C, j = threadIdx.x / C;
  // columns [colbeg, colend) of the ncol-wide matrix (a sub-range when the solve is pipelined with the
  // NVLink transposes)
  const i64 col = colbeg + blockIdx.x * (i64)C + c;
  const bool colok = col < colend;
  double2* Tc = T + col;
  double2 v[RM];

  // ---- forward pass A: radix R1 over rows j + r*R2, j < R2
  if (j < R2) {
#pragma unroll
    for (int r = 0; r < R1; r++) v[r] = colok ? Tc[(i64)(j + r * R2) * ncol] : make_double2(0.0, 0.0);
    ZDif<R1, 0, -1>::run(v);
#pragma unroll
    for (int q = 0; q < R1; q++) zsm[(j * R1 + q) * C + c] = v[zf_brev(q, B1)];
  }
  __syncthreads();
  // ---- forward pass B: radix R2 over smem[j + r*R1], j < R1
  if (j < R1) {
#pragma unroll
    for (int r = 0; r < R2; r++) v[r] = zsm[(j + r * R1) * C + c];
  }
  __syncthreads();  // every read of the exchange buffer is done before pass A' overwrites it
  if (j < R1) {
#pragma unroll
    for (int r = 1; r < R2; r++) {
      const double2 w = stw[r * j];
      v[r] = zmul(v[r], w.x, w.y);
    }
    ZDif<R2, 0, -1>::run(v);
    // ---- Green's function: X[kz] *= scale / Ginv(i, j, kz), kz = j + q*R1 held at v[brev(q)]
    //      Ginv = -4 (sin^2(kx/2) + sin^2(ky/2) + sin^2(kz/2))   ParticleMesh.jl:494-501
    double sxy = 0.0;
    bool dc_col = false;
    if (colok) {
      const i64 ic = col % icn, jl = col / icn;
      sxy = s1[ic] + s2[j0 + jl];
      dc_col = (ic == 0 && j0 + jl == 0);
    }
#pragma unroll
    for (int q = 0; q < R2; q++) {
      const int kz = j + q * R1;
      const double ginv = rank3 ? -4.0 * (sxy + ss3[kz]) : -4.0 * sxy;
      double f = (ginv != 0.0) ? scale * zf_rcp(ginv) : scale;
      if (dc_col && kz == 0) f = 0.0;  // rt[1,1,1] = 0  (:501)
      double2& e = v[zf_brev(q, B2)];
      e.x *= f;
      e.y *= f;
    }
    // ---- inverse pass A': radix R2, bit-reversed registers -> natural order
    ZDit<R2, 0, +1>::run(v);
#pragma unroll
    for (int r = 0; r < R2; r++) zsm[(j * R2 + r) * C + c] = v[r];
  }
  __syncthreads();
  // ---- inverse pass B': radix R1 over smem[j + r*R2], j < R2 -> rows j + q*R2
  if (j < R2) {
#pragma unroll
    for (int r = 0; r < R1; r++) v[r] = zsm[(j + r * R2) * C + c];
#pragma unroll
    for (int r = 1; r < R1; r++) {
      const double2 w = stw[r * j];
      v[r] = zmul(v[r], w.x, -w.y);
    }
    ZDif<R1, 0, +1>::run(v);
    if (colok) {
#pragma unroll
      for (int q = 0; q < R1; q++) Tc[(i64)(j + q * R2) * ncol] = v[zf_brev(q, B1)];
    }
  }
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
  // N3 = 2 x a two-pass length (2048): only the persistent kernel has the radix-2 wrapper (k_zfused_p<.., TWO>)
  return ctx->opt_zfused == 2 && !ctx->opt_colfft_z && ctx->g.N3 % 2 == 0 && zfused_factor(ctx->g.N3 / 2, &a, &b);
}

template <int R1, int R2, int C>
static int zfused_go(nnpm_ctx* ctx, double2* T, double scale, i64 colbeg, i64 colend) {
  const i64 ncol = ctx->njl * ctx->icn;
  if (colend < 0) { colbeg = 0; colend = ncol; }
  const size_t smem = (size_t)R1 * R2 * (C * sizeof(double2) + sizeof(double2) + sizeof(double));
  CK(cudaFuncSetAttribute(k_zfused<R1, R2, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((colend - colbeg + C - 1) / C);
  k_zfused<R1, R2, C><<<grid, C*(R1 > R2 ? R1 : R2), smem, ctx->st>>>(T, ncol, colbeg, colend, ctx->icn, ctx->njl * ctx->r, ctx->s1, ctx->s2,
                                                                      ctx->s3, ctx->ztw, 1, scale);
  LAUNCHED(ctx);
  KCHECK();
  return NNPM_OK;
}

static int zfused_launch(nnpm_ctx* ctx, double2* T, double scale, i64 colbeg = 0, i64 colend = -1) {
  const i64 N = ctx->g.N3;
  if (!ctx->ztw) {  // W_N^m = exp(-2 pi i m / N), m < N, in extended precision on the host
    std::vector<double2> h((size_t)N);
    for (i64 m = 0; m < N; m++) {
      const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)N;
      h[m] = make_double2((double)cosl(a), (double)sinl(a));
    }
    CKR(dalloc(ctx, &ctx->ztw, (size_t)N));
    CK(cudaMemcpyAsync(ctx->ztw, h.data(), (size_t)N * sizeof(double2), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
  }
  const bool c4 = ctx->opt_zcols == 4;
  switch (N) {
    case 64: return zfused_go<8, 8, 8>(ctx, T, scale, colbeg, colend);
    case 128: return zfused_go<16, 8, 8>(ctx, T, scale, colbeg, colend);
    case 256: return zfused_go<16, 16, 8>(ctx, T, scale, colbeg, colend);
    case 512: return c4 ? zfused_go<32, 16, 4>(ctx, T, scale, colbeg, colend) : zfused_go<32, 16, 8>(ctx, T, scale, colbeg, colend);
    case 1024: return c4 ? zfused_go<32, 32, 4>(ctx, T, scale, colbeg, colend) : zfused_go<32, 32, 8>(ctx, T, scale, colbeg, colend);
  }
  return fail(ctx, NNPM_ERR_STATE, "fused z pass does not support N3 = %lld", (long long)N);
}
