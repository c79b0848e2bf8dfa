// nnpm_pk.cuh -- device-side powerSpectrum (src/ParticleMesh.jl:65-100, fftfreq :202-206).
#pragma once

// sum of the owned (unpadded) cells, one double atomicAdd per block
__global__ void k_mesh_sum(const double* __restrict__ own, Geom g, double* out) {
  __shared__ double s[32];
  const i64 ncell = g.N1 * g.N2 * g.nzl;
  double acc = 0.0;
  for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < ncell; t += (i64)gridDim.x * blockDim.x) {
    i64 i = t % g.N1, r = t / g.N1;
    acc += own[i + g.rowp * r];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) atomicAdd(out, acc);
  }
}

// delta = rho/rho_mean - 1   (:75)
__global__ void k_delta(double* __restrict__ own, Geom g, double mean) {
  const i64 ncell = g.N1 * g.N2 * g.nzl;
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= ncell) return;
  i64 i = t % g.N1, r = t / g.N1;
  own[i + g.rowp * r] = __dsub_rn(__ddiv_rn(own[i + g.rowp * r], mean), 1.0);
}

__device__ __forceinline__ double fftfreq_dev(i64 i, i64 N) {  // :202-206, 0-based i
  const i64 s = N / 2;
  const i64 m = i - ((i > s) ? N : 0);
  return __ddiv_rn(__dmul_rn(6.283185307179586, (double)m), (double)N);
}

// shell binning of |delta_k|^2 over the half spectrum T[k][jl][ic]; the Hermitian mirror of
// column ic (0 < ic < N1-ic) has the same |k| and the same power, so it is added with weight 2.
__global__ void k_pk_bin(const double2* __restrict__ T, i64 icn, i64 njl, i64 j0, Geom g, double dk, int lenk,
                         double fac, double* __restrict__ power, double* __restrict__ counts) {
  extern __shared__ double sb[];  // [2*lenk]
  for (int t = threadIdx.x; t < 2 * lenk; t += blockDim.x) sb[t] = 0.0;
  __syncthreads();
  const i64 total = icn * njl * g.N3;
  for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    i64 i = t % icn, r = t / icn;
    i64 jl = r % njl, k = r / njl;
    double kx = fftfreq_dev(i, g.N1), ky = fftfreq_dev(j0 + jl, g.N2), kz = fftfreq_dev(k, g.N3);
    double ka = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)), __dmul_rn(kz, kz)));
    i64 n = (i64)rint(__ddiv_rn(ka, dk));
    if (n >= 1 && n <= lenk) {
      double2 c = T[t];
      double a = sqrt(__dadd_rn(__dmul_rn(c.x, c.x), __dmul_rn(c.y, c.y)));  // abs(fft(delta))
      double p = __dmul_rn(__dmul_rn(a, a), fac);
      const i64 im = (i == 0) ? 0 : g.N1 - i;
      const double w = (im != i && im >= icn) ? 2.0 : 1.0;
      atomicAdd(&sb[n - 1], w * p);
      atomicAdd(&sb[lenk + n - 1], w);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < lenk; t += blockDim.x) {
    if (sb[lenk + t] != 0.0) {
      atomicAdd(&power[t], sb[t]);
      atomicAdd(&counts[t], sb[lenk + t]);
    }
  }
}

extern "C" int nnpm_power_spectrum(nnpm_ctx* ctx, double* karray_out, double* power_out, int32_t lenk) {
  if (!ctx || !karray_out || !power_out) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  const Geom& g = ctx->g;
  const i64 nmax = g.N1 > g.N2 ? (g.N1 > g.N3 ? g.N1 : g.N3) : (g.N2 > g.N3 ? g.N2 : g.N3);
  const int want = (int)nearbyint((double)nmax / 2.0);  // round(Int, maximum(dims)/2)  :81
  if (lenk != want) return fail(ctx, NNPM_ERR_ARG, "lenk must be round(max(dims)/2) = %d", want);
  const double dk = 2 * M_PI / (double)nmax;  // :79
  if (!ctx->pk_buf) CKR(dalloc(ctx, &ctx->pk_buf, (size_t)2 * lenk + 2));
  double* d_power = ctx->pk_buf;
  double* d_counts = ctx->pk_buf + lenk;
  double* d_sum = ctx->pk_buf + 2 * lenk;
  CK(cudaMemsetAsync(ctx->pk_buf, 0, (size_t)(2 * lenk + 2) * 8, ctx->st));
  CK(cudaMemsetAsync(ctx->d_red + 3, 0, 8, ctx->st));
  CKR(do_deposit(ctx, false, 0.0));
  double* own = ctx->mesh + g.plane * g.glo;
  const i64 ncl = g.N1 * g.N2 * g.nzl;
  k_mesh_sum<<<ctx->nsm * 4, 256, 0, ctx->st>>>(own, g, d_sum);
  LAUNCHED(ctx);
  KCHECK();
  double sum = 0.0;
  CK(cudaMemcpyAsync(&sum, d_sum, 8, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  if (ctx->P > 1) CKR(comm_allreduce_sum(ctx, &sum, 1));
  const double mean = sum / ((double)g.N1 * (double)g.N2 * (double)g.N3);
  k_delta<<<nblk(ncl, 256), 256, 0, ctx->st>>>(own, g, mean);
  LAUNCHED(ctx);
  double2* T = nullptr;
  CKR(fft_forward(ctx, &T, true, nullptr, false));
  const double ncell = (double)g.N1 * (double)g.N2 * (double)g.N3;
  const double fac = 1.0 / (ncell * ncell);  // prod(box)/prod(dims)^2 with box = 1  :77
  k_pk_bin<<<ctx->nsm * 2, 256, (size_t)2 * lenk * 8, ctx->st>>>(T, ctx->icn, ctx->njl, ctx->njl * ctx->r, g, dk, lenk, fac,
                                                            d_power, d_counts);
  LAUNCHED(ctx);
  KCHECK();
  std::vector<double> h((size_t)2 * lenk);
  CK(cudaMemcpyAsync(h.data(), ctx->pk_buf, (size_t)2 * lenk * 8, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  if (ctx->P > 1) CKR(comm_allreduce_sum(ctx, h.data(), 2 * lenk));
  for (int n = 0; n < lenk; n++) {
    karray_out[n] = (double)(n + 1) * dk;
    power_out[n] = h[n] / h[lenk + n];  // power ./ counts (NaN for empty bins, as the reference)
  }
  ctx->field_state = 0;
  ctx->acc_valid = false;
  return NNPM_OK;
}
