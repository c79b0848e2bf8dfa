// nnpm_ic.cuh -- device-side initial conditions that need the FFT machinery or the cosmology scalars:
//   nnpm_init_grf      PowerSpectrumParticles (src/ParticleMesh.jl:209-254) + scaleFreePS (src/InputPowerSpectra.jl:26):
//                      seeded Gaussian-random-field Zel'dovich displacements on the particle lattice
//   nnpm_init_grf_peaks  the same with multiplePeaksPS (src/InputPowerSpectra.jl:6-24, :28-36), the spectrum the shipped
//                      run/PM/particle_mesh.conf names
//   nnpm_init_pancake  ZeldovichPancakeParticles (src/ParticleMesh.jl:177-198)
// Included by nnpm.cu after the transform code (uses fft_inverse and the slab transposes).
#pragma once

// fftfreq (ParticleMesh.jl:202-206) for a 0-based index: 2 pi (idx - N [idx > N/2]) / N, evaluated left to right
__device__ __forceinline__ double ic_kf(int idx, int N) {
  const int w = idx > N / 2 ? idx - N : idx;
  return __ddiv_rn(__dmul_rn(6.283185307179586, (double)w), (double)N);
}

// multiplePeaksPS: P(k) = k^n where |k - k_p| / k < lnwidth_p / 2 for some peak p, 0 elsewhere (InputPowerSpectra.jl:28-36)
struct PeakTab {
  int n;
  double k[NNPM_MAX_PEAKS], w[NNPM_MAX_PEAKS];
};

// complex amplitude (ak - i bk)/2 of mode (i, j, k), ak = sqrt(P(|k|)) g1 / k^2, bk likewise with g2 (:230-236);
// (g1, g2) = Box-Muller of two uniforms drawn from mix64 keyed by (seed, linear index of the mode in the half spectrum)
template <bool PEAKS>
__device__ __forceinline__ double2 ic_mode_amp(int i, int j, int k, int N1, int N2, int N3, u64 seed, double ps_index,
                                               const PeakTab& pk, double& kx, double& ky, double& kz) {
  kx = ic_kf(i, N1);
  ky = ic_kf(j, N2);
  kz = N3 > 1 ? ic_kf(k, N3) : 0.0;
  const double k2 = kx * kx + ky * ky + kz * kz;
  if (k2 == 0.0) return make_double2(0.0, 0.0);
  const double ka = sqrt(k2);
  if (PEAKS) {
    bool inside = false;
    for (int q = 0; q < pk.n; q++)
      if (__ddiv_rn(fabs(__dsub_rn(ka, pk.k[q])), ka) < __ddiv_rn(pk.w[q], 2.0)) inside = true;
    if (!inside) return make_double2(0.0, 0.0);
  }
  const double psv = sqrt(pow(ka, ps_index));
  const u64 lin = (u64)i + (u64)(N1 / 2 + 1) * ((u64)j + (u64)N2 * (u64)k);
  const u64 h = mix64(seed ^ mix64(lin));
  const double u1 = u01(mix64(h + 1)), u2 = u01(mix64(h + 2));
  const double r = sqrt(-2.0 * log(u1));
  const double g1 = r * cospi(2.0 * u2), g2 = r * sinpi(2.0 * u2);
  return make_double2(psv * g1 / k2 / 2.0, -(psv * g2 / k2) / 2.0);
}

// component D of the displacement spectrum at (i, j, k): amp * k_D, with the planes i = 0 and i = N1/2 made Hermitian
// ((j, k) whose mirror (-j, -k) has the smaller linear index takes the conjugate of the mirror; self-mirrored bins are real)
template <bool PEAKS>
__device__ __forceinline__ double2 ic_mode_value(int D, int i, int j, int k, int N1, int N2, int N3, u64 seed, double ps_index,
                                                 const PeakTab& pk) {
  bool conj = false, real_only = false;
  if (i == 0 || 2 * i == N1) {
    const int jm = (N2 - j) % N2, km = (N3 - k) % N3;
    const int a = j + N2 * k, b = jm + N2 * km;
    if (b < a) { j = jm; k = km; conj = true; }
    else if (b == a) real_only = true;
  }
  double kf[3];
  const double2 amp = ic_mode_amp<PEAKS>(i, j, k, N1, N2, N3, seed, ps_index, pk, kf[0], kf[1], kf[2]);
  double2 c = make_double2(amp.x * kf[D], amp.y * kf[D]);
  if (conj) c.y = -c.y;
  if (real_only) c.y = 0.0;
  return c;
}

// fill the (slab-transposed) half spectrum T[k][jl][ic], jl = j - j0
template <bool PEAKS>
__global__ void k_ic_grf_fill(double2* __restrict__ T, int D, int icn, int njl, int j0, int N1, int N2, int N3, u64 seed,
                              double ps_index, const __grid_constant__ PeakTab pk) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = blockIdx.y;
  if (ic >= icn) return;
  for (int k = blockIdx.z; k < N3; k += gridDim.z)
    T[ic + (i64)icn * (jl + (i64)njl * k)] = ic_mode_value<PEAKS>(D, ic, j0 + jl, k, N1, N2, N3, seed, ps_index, pk);
}

// psi_D at this rank's lattice sites (the transformed mesh, unnormalised) -> dst[t]; the squares are summed in a fixed
// order: thread-sequential over a grid stride, block tree, one partial per block
__global__ void __launch_bounds__(256)
k_ic_grf_store(double* __restrict__ dst, const double* __restrict__ own, i64 nloc, int N1, int N2, i64 rowp, double* __restrict__ partial) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < nloc; t += (i64)gridDim.x * blockDim.x) {
    const i64 i = t % N1, jk = t / N1;   // jk = j + N2 * kl: rows of the slab in order
    const double p = own[i + rowp * jk];
    dst[t] = p;
    acc += p * p;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// x = wrap(q + s psi), v = vfac s psi  (psi held in the velocity arrays), ids = lattice index
template <int RANK>
__global__ void k_ic_grf_apply(PartPtrs p, uint32_t* ids, i64 nloc, i64 first, i64 P1, i64 P2, i64 P3, double s, double vfac) {
  const i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= nloc) return;
  const i64 gidx = first + t;
  const i64 i = gidx % P1, j = (gidx / P1) % P2, k = gidx / (P1 * P2);
  double q[3];
  q[0] = __ddiv_rn(__dsub_rn((double)(i + 1), 0.5), (double)P1);
  q[1] = __ddiv_rn(__dsub_rn((double)(j + 1), 0.5), (double)P2);
  q[2] = (RANK == 3) ? __ddiv_rn(__dsub_rn((double)(k + 1), 0.5), (double)P3) : 0.0;
#pragma unroll
  for (int d = 0; d < RANK; d++) {
    const double psi = __dmul_rn(s, p.v[d][t]);
    p.x[d][t] = wrap01(__dadd_rn(q[d], psi));
    p.v[d][t] = __dmul_rn(vfac, psi);
  }
  ids[t] = (uint32_t)gidx;
}

template <int RANK>
__global__ void k_init_pancake(PartPtrs p, uint32_t* ids, i64 nloc, i64 first, i64 P1, i64 P2, i64 P3, double xamp, double vamp) {
  const i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= nloc) return;
  const i64 gidx = first + t;
  const i64 i = gidx % P1, j = (gidx / P1) % P2, k = gidx / (P1 * P2);
  double q[3];
  q[0] = __ddiv_rn(__dsub_rn((double)(i + 1), 0.5), (double)P1);
  q[1] = __ddiv_rn(__dsub_rn((double)(j + 1), 0.5), (double)P2);
  q[2] = (RANK == 3) ? __ddiv_rn(__dsub_rn((double)(k + 1), 0.5), (double)P3) : 0.0;
  const double sn = sin(__dmul_rn(6.283185307179586, q[0]));   // sin(2pi * kZ * x[1,:]), kZ = 1  (:193-194)
#pragma unroll
  for (int d = 0; d < RANK; d++) {
    p.x[d][t] = wrap01(d == 0 ? __dadd_rn(q[0], __dmul_rn(xamp, sn)) : q[d]);
    p.v[d][t] = d == 0 ? __dmul_rn(vamp, sn) : 0.0;
  }
  ids[t] = (uint32_t)gidx;
}

static int ic_finish(nnpm_ctx* ctx, i64 per, double m) {
  ctx->np = per;
  ctx->m = m;
  ctx->first_id = 0;
  ctx->sorted_np = 0;
  ctx->ids_dirty = ctx->P > 1;
  ctx->acc_valid = ctx->pa_valid = false;
  ctx->rho_keep_valid = false;
  ctx->field_state = 0;
  CK(cudaStreamSynchronize(ctx->st));
  if (ctx->P > 1) return nnpm_redistribute(ctx);
  return NNPM_OK;
}

extern "C" int nnpm_init_pancake(nnpm_ctx* ctx, const int64_t pdims[3], double xamp, double vamp, double m) {
  if (!ctx || !pdims) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  const i64 P1 = pdims[0], P2 = pdims[1], P3 = ctx->R == 3 ? pdims[2] : 1;
  if (P1 * P2 * P3 != ctx->cfg.npart_total) return fail(ctx, NNPM_ERR_ARG, "prod(pdims) != npart_total");
  if (ctx->P > 1 && P3 % ctx->P) return fail(ctx, NNPM_ERR_ARG, "pdims[2] must be divisible by nranks");
  const i64 per = P3 / ctx->P * P1 * P2, first = per * ctx->r;
  if (per > ctx->cap) return fail(ctx, NNPM_ERR_CAPACITY, "pancake ICs need capacity %lld", (long long)per);
  if (per) {
    if (ctx->R == 3) k_init_pancake<3><<<nblk(per, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, per, first, P1, P2, P3, xamp, vamp);
    else k_init_pancake<2><<<nblk(per, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, per, first, P1, P2, P3, xamp, vamp);
    LAUNCHED(ctx);
    KCHECK();
  }
  return ic_finish(ctx, per, m);
}

static int init_grf_impl(nnpm_ctx* ctx, const int64_t pdims[3], uint64_t seed, double ps_index, const PeakTab& pk,
                         double disp_rms, double vfac, double m) {
  CK(cudaSetDevice(ctx->cfg.device));
  const i64 P1 = pdims[0], P2 = pdims[1], P3 = ctx->R == 3 ? pdims[2] : 1;
  if (P1 * P2 * P3 != ctx->cfg.npart_total) return fail(ctx, NNPM_ERR_ARG, "prod(pdims) != npart_total");
  if (ctx->P > 1 && (P3 % ctx->P || P2 % ctx->P || P3 / ctx->P < 4))
    return fail(ctx, NNPM_ERR_ARG, "nranks > 1 needs pdims[1], pdims[2] divisible by nranks and pdims[2] / nranks >= 4");
  const i64 per = P3 / ctx->P * P1 * P2, first = per * ctx->r;
  if (per > ctx->cap) return fail(ctx, NNPM_ERR_CAPACITY, "random-field ICs need capacity %lld", (long long)per);
  // the inverse transforms run on a mesh of the particle lattice's shape: this context's own when the shapes agree,
  // otherwise a temporary context that borrows this one's stream and communicator
  nnpm_ctx* ic = ctx;
  if (P1 != ctx->g.N1 || P2 != ctx->g.N2 || P3 != ctx->g.N3) {
    nnpm_config c2 = ctx->cfg;
    c2.ngrid[0] = P1; c2.ngrid[1] = P2; c2.ngrid[2] = P3;
    c2.npart_total = 1; c2.part_capacity = 1; c2.sort_every = 0; c2.timing = 0;
    c2.stream = (void*)ctx->st;
    int rc = nnpm_create(&ic, &c2);
    if (rc != NNPM_OK) return fail(ctx, rc, "random-field ICs: lattice context: %s", nnpm_last_error(nullptr));
    ic->nccl = ctx->nccl;
    ic->comm = ctx->comm;
    ic->comm_borrowed = true;   // peer_ok stays false: transposes through grouped ncclSend / ncclRecv
  }
  auto done = [&](int rc) {
    if (ic != ctx) {
      if (rc != NNPM_OK) ctx->err = ic->err;
      nnpm_destroy(ic);
    }
    return rc;
  };
#define CKI(call) do { int rc3_ = (call); if (rc3_ != NNPM_OK) return done(rc3_); } while (0)
  const Geom& g = ic->g;
  const int NB = 1024;
  if (!ctx->ic_partial) CKI(dalloc(ctx, &ctx->ic_partial, (size_t)NB * 3));
  double* partial = ctx->ic_partial;
  double2* T = ic->P > 1 ? (double2*)ic->meshT : (double2*)(ic->mesh + g.plane * g.glo);
  for (int d = 0; d < ctx->R; d++) {
    dim3 grid(nblk(ic->icn, 128), (unsigned)ic->njl, (unsigned)(P3 < 64 ? P3 : 64));
    if (pk.n > 0)
      k_ic_grf_fill<true><<<grid, 128, 0, ctx->st>>>(T, d, (int)ic->icn, (int)ic->njl, (int)(ic->njl * ic->r), (int)P1, (int)P2,
                                                     (int)P3, (u64)seed, ps_index, pk);
    else
      k_ic_grf_fill<false><<<grid, 128, 0, ctx->st>>>(T, d, (int)ic->icn, (int)ic->njl, (int)(ic->njl * ic->r), (int)P1, (int)P2,
                                                      (int)P3, (u64)seed, ps_index, pk);
    LAUNCHED(ctx);
    if (cudaGetLastError() != cudaSuccess) return done(fail(ctx, NNPM_ERR_CUDA, "k_ic_grf_fill launch failed"));
    CKI(fft_inverse(ic, true, nullptr, false));
    if (per) {
      k_ic_grf_store<<<NB, 256, 0, ctx->st>>>(ctx->pp.v[d], ic->mesh + g.plane * g.glo, per, (int)P1, (int)P2, g.rowp, partial + NB * d);
      LAUNCHED(ctx);
    } else {
      if (cudaMemsetAsync(partial + NB * d, 0, NB * 8, ctx->st) != cudaSuccess) return done(fail(ctx, NNPM_ERR_CUDA, "memset failed"));
    }
  }
  ic->field_state = 0;
  std::vector<double> hp((size_t)NB * 3, 0.0);
  if (cudaMemcpyAsync(hp.data(), partial, (size_t)NB * ctx->R * 8, cudaMemcpyDeviceToHost, ctx->st) != cudaSuccess ||
      cudaStreamSynchronize(ctx->st) != cudaSuccess)
    return done(fail(ctx, NNPM_ERR_CUDA, "random-field ICs: reading the displacement norm failed: %s", cudaGetErrorString(cudaGetLastError())));
  double sum = 0.0;
  for (int i = 0; i < NB * ctx->R; i++) sum += hp[i];
  if (ctx->P > 1) CKI(comm_allreduce_sum(ctx, &sum, 1));
  const double rms = sqrt(sum / ((double)ctx->cfg.npart_total * ctx->R));
  const double s = rms > 0.0 ? disp_rms / rms : 0.0;
  if (per) {
    if (ctx->R == 3) k_ic_grf_apply<3><<<nblk(per, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, per, first, P1, P2, P3, s, vfac);
    else k_ic_grf_apply<2><<<nblk(per, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, per, first, P1, P2, P3, s, vfac);
    LAUNCHED(ctx);
  }
#undef CKI
  int rc = done(cudaGetLastError() == cudaSuccess ? NNPM_OK : fail(ctx, NNPM_ERR_CUDA, "random-field ICs: apply kernel failed"));
  if (rc != NNPM_OK) return rc;
  return ic_finish(ctx, per, m);
}

extern "C" int nnpm_init_grf(nnpm_ctx* ctx, const int64_t pdims[3], uint64_t seed, double ps_index, double disp_rms,
                             double vfac, double m) {
  if (!ctx || !pdims) return NNPM_ERR_ARG;
  PeakTab pk;
  memset(&pk, 0, sizeof pk);
  return init_grf_impl(ctx, pdims, seed, ps_index, pk, disp_rms, vfac, m);
}

extern "C" int nnpm_init_grf_peaks(nnpm_ctx* ctx, const int64_t pdims[3], uint64_t seed, double ps_index, int32_t npeaks,
                                   const double* kpeak, const double* lnwidth, double disp_rms, double vfac, double m) {
  if (!ctx || !pdims || !kpeak || !lnwidth) return NNPM_ERR_ARG;
  if (npeaks < 1 || npeaks > NNPM_MAX_PEAKS) return fail(ctx, NNPM_ERR_ARG, "npeaks must be 1 .. %d", NNPM_MAX_PEAKS);
  PeakTab pk;
  memset(&pk, 0, sizeof pk);
  pk.n = npeaks;
  for (int q = 0; q < npeaks; q++) {
    if (!(kpeak[q] > 0.0) || !(lnwidth[q] > 0.0)) return fail(ctx, NNPM_ERR_ARG, "peak %d: k and lnwidth must be positive", q);
    pk.k[q] = kpeak[q];
    pk.w[q] = lnwidth[q];
  }
  return init_grf_impl(ctx, pdims, seed, ps_index, pk, disp_rms, vfac, m);
}
