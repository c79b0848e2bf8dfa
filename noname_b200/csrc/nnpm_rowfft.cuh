// nnpm_rowfft.cuh -- x passes of the 3-D transform: batched real-to-complex / complex-to-real FFTs along the
// contiguous mesh rows, in place in the padded r2c layout (row = N1 doubles + 2 of padding = N1/2+1 complex).
// Replaces cuFFT's batched 1-D D2Z / Z2D plans in compute_potential (src/ParticleMesh.jl:481, :502): the library
// kernels reach 0.72-0.75 of the measured HBM copy peak at 1024^3 (profiles/r01_ncu_summary.txt).
//
// A real transform of length N = 2M is one complex transform of length M on z[n] = x[2n] + i x[2n+1] plus an
// O(M) butterfly between bins k and M-k:
//   forward   X[k]   = E[k] + W_N^k O[k],  E[k] = (Z[k] + conj Z[M-k]) / 2,  O[k] = -i (Z[k] - conj Z[M-k]) / 2,  k = 0 .. M
//   inverse   Z'[k]  = (X[k] + conj X[M-k]) + i conj(W_N^k) (X[k] - conj X[M-k]),   z' = IDFT_M(Z') = N x   (unnormalised,
//             like cuFFT's Z2D: the 1/(N1 N2 N3) sits in the Green's-function factor of the z pass)
//
// Data movement: every CTA owns a shared-memory slot per row (C = 4 rows at a time); the rows arrive by ONE bulk
// asynchronous copy each (cp.async.bulk, 8 KB at N1 = 1024, completion on an mbarrier) and leave by one bulk copy
// each -- the threads never touch global memory.  A slot holds the row, then (aliased) the padded exchange buffer of
// the two register-resident radix passes (ZDif of nnpm_zfft.cuh), then the spectrum.  CTAs are small (4 warps, 51 KB
// at N1 = 1024) and persistent, four per SM, so the load / store phases of one CTA run under the arithmetic of its
// siblings: ~128 KB per SM in flight without any intra-CTA pipelining.
#pragma once

__device__ __forceinline__ void bulk_load_1d(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

// L2 eviction-priority hints for the fused 2-D transform: data read or written for the last time goes in as
// evict_first, the intermediate between its two passes as evict_last
__device__ __forceinline__ u64 l2_policy_evict_first() {
  u64 p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ u64 l2_policy_evict_last() {
  u64 p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_load_1d_hint(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar, u64 pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d_hint(void* gdst, const void* ssrc, unsigned bytes, u64 pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned long long* bar, u64 pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, u64 pol) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(tm), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2), "l"(pol)
               : "memory");
}

struct RowFftArgs {
  double* mesh;        // first row
  i64 rowp;            // row pitch in doubles (2 * (N1/2 + 1))
  i64 nrows;           // rows to transform (rows are contiguous: row r at mesh + r * rowp)
  const double2* twM;  // W_M^m, m < M
  const double2* twN;  // W_N^k, k < M
};

// Complex FFT of length M = R1 * R2 (R1 >= R2) on the C rows held in the slots, natural order in and out, in place.
// T = C * R2 threads, every lane busy in both passes (FP64 issue slots are the co-limit of this kernel: a half-idle
// warp costs as many as a full one):
//   pass A  C * R2 sub-transforms of radix R1, one per thread: row c = tid / R2, j = tid % R2: elements j + r*R2 ->
//           DIF -> exchange E[q*(R2+1) + j]   (padded: both sides of the exchange are bank-conflict free)
//   pass B  C * R1 sub-transforms of radix R2, KB = R1 / R2 per thread (held together in the same R1 registers):
//           E[j*(R2+1) + r] * W_M^(r j) -> DIF -> element j + q*R1
template <int R1, int R2, int C, int SLOT, int SIGN>
__device__ __forceinline__ void rowfft_core(double2* slots, const double2* stwM, int tid) {
  static_assert(R1 >= R2 && R1 % R2 == 0, "factor order");
  constexpr int T = C * R2, KB = R1 / R2, B1 = zf_log2(R1), B2 = zf_log2(R2), EST = R2 + 1;
  double2 v[R1];
  {
    double2* slot = slots + (tid / R2) * SLOT;
    const int j = tid % R2;
#pragma unroll
    for (int r = 0; r < R1; r++) v[r] = slot[j + r * R2];
    ZDif<R1, 0, SIGN>::run(v);
    __syncthreads();  // every row is in registers: the slots become the exchange buffers
#pragma unroll
    for (int q = 0; q < R1; q++) slot[q * EST + j] = v[zf_brev(q, B1)];
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < KB; it++) {
    const int t = tid + it * T;
    const double2* slot = slots + (t / R1) * SLOT;
    const int j = t % R1;
#pragma unroll
    for (int r = 0; r < R2; r++) v[it * R2 + r] = slot[j * EST + r];
#pragma unroll
    for (int r = 1; r < R2; r++) {
      const double2 w = stwM[r * j];
      v[it * R2 + r] = zmul(v[it * R2 + r], w.x, SIGN < 0 ? w.y : -w.y);
    }
  }
  ZDif<R2, 0, SIGN>::run(v);
  if (KB == 2) ZDif<R2, R2, SIGN>::run(v);
  __syncthreads();  // the exchange is consumed: the slots take the result in natural order
#pragma unroll
  for (int it = 0; it < KB; it++) {
    const int t = tid + it * T;
    double2* slot = slots + (t / R1) * SLOT;
    const int j = t % R1;
#pragma unroll
    for (int q = 0; q < R2; q++) slot[j + q * R1] = v[it * R2 + zf_brev(q, B2)];
  }
  __syncthreads();
}

// INVERSE = false: real rows -> half spectrum (r2c, forward sign);  true: half spectrum -> real rows (c2r, unnormalised)
template <int R1, int R2, bool INVERSE>
__global__ void __launch_bounds__(4 * R2, (R1 * R2 > 512 ? 2 : 4))
k_rowfft(RowFftArgs a) {
  constexpr int C = 4, M = R1 * R2, T = C * R2;
  constexpr int SLOT = R1 * (R2 + 1) > M + 1 ? R1 * (R2 + 1) : M + 1;   // complex elements per row slot
  extern __shared__ __align__(128) unsigned char rf_smem[];
  double2* slots = (double2*)rf_smem;       // [C][SLOT]
  double2* stwM = slots + C * SLOT;         // [M]      W_M^m
  double2* stwH = stwM + M;                 // [M/2+1]  forward: -i/2 W_N^k;  inverse: i conj(W_N^k)
  __shared__ __align__(8) unsigned long long full;
  const int tid = threadIdx.x;
  for (int t = tid; t < M; t += T) stwM[t] = __ldg(a.twM + t);
  for (int t = tid; t <= M / 2; t += T) {
    const double2 w = __ldg(a.twN + t);
    stwH[t] = INVERSE ? make_double2(w.y, w.x) : make_double2(0.5 * w.y, -0.5 * w.x);
  }
  if (tid == 0) mbar_init(&full, 1);
  __syncthreads();
  constexpr unsigned IN_BYTES = INVERSE ? (M + 1) * 16u : M * 16u;
  constexpr unsigned OUT_BYTES = INVERSE ? M * 16u : (M + 1) * 16u;
  constexpr int NPAIR = M / 2 + 1;          // bins k = 0 .. M/2, each paired with M - k
  const i64 ngroups = (a.nrows + C - 1) / C;
  unsigned phase = 0u;
  for (i64 g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const i64 row0 = g * C;
    const int nr = (int)((a.nrows - row0) < C ? (a.nrows - row0) : C);
    if (tid == 0) {
      // the previous group's bulk stores must have finished READING the slots before they are refilled
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      mbar_expect_tx(&full, (unsigned)nr * IN_BYTES);
      for (int r = 0; r < nr; r++) bulk_load_1d(slots + r * SLOT, a.mesh + (row0 + r) * a.rowp, IN_BYTES, &full);
    }
    mbar_wait(&full, phase);
    phase ^= 1u;
    if (INVERSE) {
      // Z'[k] = (X[k] + conj X[M-k]) + i conj(W_N^k) (X[k] - conj X[M-k]), pairs (k, M-k) in place; k = 0 also consumes X[M]
      for (int t = tid; t < C * NPAIR; t += T) {
        double2* slot = slots + (t / NPAIR) * SLOT;
        const int k = t % NPAIR;
        double2 xa = slot[k], xb = slot[M - k];
        if (k == 0) { xa.y = 0.0; xb.y = 0.0; }                      // the DC and Nyquist bins of a real signal are real
        const double2 h = stwH[k];                                   // i conj(W_N^k)
        const double2 e = make_double2(xa.x + xb.x, xa.y - xb.y);    // X[k] + conj X[M-k]
        const double2 d = make_double2(xa.x - xb.x, xa.y + xb.y);    // X[k] - conj X[M-k]
        const double2 u = make_double2(h.x * d.x - h.y * d.y, h.x * d.y + h.y * d.x);   // i conj(W_N^k) d
        slot[k] = make_double2(e.x + u.x, e.y + u.y);
        // bin M-k:  conj(e) + i conj(W_N^(M-k)) (-conj d) = conj(e) - conj(u)
        if (k != 0 && k != M - k) slot[M - k] = make_double2(e.x - u.x, -e.y + u.y);
      }
      __syncthreads();
      rowfft_core<R1, R2, C, SLOT, +1>(slots, stwM, tid);
    } else {
      rowfft_core<R1, R2, C, SLOT, -1>(slots, stwM, tid);
      // X[k] = s/2 + h d,  X[M-k] = conj(s/2) - conj(h d),  s = Z[k] + conj Z[M-k], d = Z[k] - conj Z[M-k], h = -i/2 W_N^k
      for (int t = tid; t < C * NPAIR; t += T) {
        double2* slot = slots + (t / NPAIR) * SLOT;
        const int k = t % NPAIR;
        const double2 za = slot[k], zb = slot[k == 0 ? 0 : M - k];
        const double2 h = stwH[k];
        const double2 sh = make_double2(0.5 * (za.x + zb.x), 0.5 * (za.y - zb.y));
        const double2 d = make_double2(za.x - zb.x, za.y + zb.y);
        const double2 u = make_double2(h.x * d.x - h.y * d.y, h.x * d.y + h.y * d.x);
        slot[k] = make_double2(sh.x + u.x, sh.y + u.y);
        slot[M - k] = make_double2(sh.x - u.x, -sh.y + u.y);
      }
      __syncthreads();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> bulk-copy reads
    __syncthreads();
    if (tid == 0) {
      for (int r = 0; r < nr; r++) bulk_store_1d(a.mesh + (row0 + r) * a.rowp, slots + r * SLOT, OUT_BYTES);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- host side -------------------------------------------------------------------------------
static bool rowfft_supported_n(i64 N1) {
  int r1, r2;
  return N1 % 2 == 0 && zfused_factor(N1 / 2, &r1, &r2);
}

static int rowfft_tables(nnpm_ctx* ctx) {
  if (ctx->xtwM) return NNPM_OK;
  const i64 N = ctx->g.N1, M = N / 2;
  std::vector<double2> hM((size_t)M), hN((size_t)M);
  const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
  for (i64 m = 0; m < M; m++) {
    const long double aM = -tau * (long double)m / (long double)M, aN = -tau * (long double)m / (long double)N;
    hM[m] = make_double2((double)cosl(aM), (double)sinl(aM));
    hN[m] = make_double2((double)cosl(aN), (double)sinl(aN));
  }
  CKR(dalloc(ctx, &ctx->xtwM, (size_t)M));
  CKR(dalloc(ctx, &ctx->xtwN, (size_t)M));
  CK(cudaMemcpyAsync(ctx->xtwM, hM.data(), (size_t)M * sizeof(double2), cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(ctx->xtwN, hN.data(), (size_t)M * sizeof(double2), cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return NNPM_OK;
}

template <int R1, int R2, bool INVERSE>
static int rowfft_go(nnpm_ctx* ctx, const RowFftArgs& a) {
  constexpr int M = R1 * R2;
  constexpr int SLOT = R1 * (R2 + 1) > M + 1 ? R1 * (R2 + 1) : M + 1;
  const size_t smem = (size_t)4 * SLOT * 16 + (size_t)M * 16 + (size_t)(M / 2 + 1) * 16;
  CK(cudaFuncSetAttribute(k_rowfft<R1, R2, INVERSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const i64 ngroups = (a.nrows + 3) / 4;
  const i64 maxg = (i64)ctx->nsm * (M > 512 ? 2 : 4);
  const unsigned grid = (unsigned)(ngroups < maxg ? ngroups : maxg);
  k_rowfft<R1, R2, INVERSE><<<grid, 4 * R2, smem, ctx->st>>>(a);
  LAUNCHED(ctx);
  KCHECK();
  return NNPM_OK;
}

// x pass over planes [plane0, plane0 + nplanes) of the owned slab, in place
static int rowfft_x(nnpm_ctx* ctx, bool inverse, i64 plane0 = 0, i64 nplanes = -1) {
  const Geom& g = ctx->g;
  CKR(rowfft_tables(ctx));
  if (nplanes < 0) nplanes = g.nzl;
  RowFftArgs a;
  a.mesh = ctx->mesh + g.plane * (g.glo + plane0);
  a.rowp = g.rowp;
  a.nrows = g.N2 * nplanes;
  a.twM = ctx->xtwM;
  a.twN = ctx->xtwN;
  if (a.nrows == 0) return NNPM_OK;
#define RF(R1_, R2_) return inverse ? rowfft_go<R1_, R2_, true>(ctx, a) : rowfft_go<R1_, R2_, false>(ctx, a)
  switch (g.N1 / 2) {
    case 64: RF(8, 8);
    case 128: RF(16, 8);
    case 256: RF(16, 16);
    case 512: RF(32, 16);
    case 1024: RF(32, 32);
  }
#undef RF
  return fail(ctx, NNPM_ERR_STATE, "row FFT does not support N1 = %lld", (long long)g.N1);
}

// ------------------------------------------------------------------------------------------------------------------
#ifndef NNPM_EXPERIMENTS
// The fused x + y transform (k_fft2d, option "fft2d") is bit-identical to the separate passes but measured slower
// (DESIGN.md section 3): only built with -DNNPM_EXPERIMENTS (make EXPERIMENTS=1).
static bool fft2d_supported(const nnpm_ctx*) { return false; }
static int fft2d_xy(nnpm_ctx* ctx, bool, i64 = 0, i64 = -1) { return fail(ctx, NNPM_ERR_STATE, "fft2d: built without -DNNPM_EXPERIMENTS"); }
#else
// Fused 2-D transform of the owned planes with the intermediate held in L2 (k_fft2d).
//
// Run as separate kernels, the x pass and the y pass each read and write the whole mesh from / to HBM: 4 x 8.6 GB per
// direction at 1024^3.  A plane is only 8.4 MB, though, and the 126 MB L2 holds a dozen of them.  Here ONE persistent
// kernel does both passes, plane block by plane block: its CTAs draw work items from a global queue ordered
//     X(block 0) | X(block 1) Y(block 0) | X(block 2) Y(block 1) | ... | Y(last block)          (forward; B planes per block)
// where X(p, g) transforms 8 rows of plane p (bulk-copy staged, as k_rowfft) and Y(p, c) one group of 4 columns of plane
// p (TMA tensor copies, as k_colfft).  A Y item needs every row of its plane: each finished X item bumps a per-plane
// counter (release), a Y item waits until the counter is complete (acquire).  Because items are started in queue order
// and X items never wait, the oldest unfinished item can always run: no deadlock, and since a plane's Y items sit a
// whole block behind its X items the wait is almost never taken.  The x output of a block (B x 8.4 MB) is still in L2
// when the y pass reads it, and is overwritten there before it is ever evicted: HBM sees one read and one write of the
// mesh per direction instead of two each.  Inverse direction: Y items first, X items (c2r) one block behind.
// ------------------------------------------------------------------------------------------------------------------
struct Fft2dArgs {
  double* mesh;          // first owned plane
  i64 rowp, plane;       // row / plane pitch in doubles
  int nplanes, bplanes;  // planes to transform, planes per block (nplanes % bplanes == 0)
  int icn, nicg;         // complex columns, column groups of 4
  const double2* twM;    // x pass: W_M^m
  const double2* twN;    // x pass: W_N^k
  const double2* twY;    // y pass: W_N2^m
  unsigned* queue;       // [0] next item; zeroed before the launch
  unsigned* done;        // [nplanes] finished first-pass items per plane; zeroed before the launch
  unsigned* fault;       // set when a dependency wait gave up (never in a correct run)
};

template <int R1, int R2, int R1Y, int R2Y, bool INVERSE>
__global__ void __launch_bounds__(128, 2)
k_fft2d(const __grid_constant__ CUtensorMap tmap, Fft2dArgs a) {
  constexpr int CX = 128 / R2;                       // rows per X item: every lane busy in pass A of the row FFT
  constexpr int M = R1 * R2, NY = R1Y * R2Y;
  constexpr int SLOT = R1 * (R2 + 1) > M + 1 ? R1 * (R2 + 1) : M + 1;
  constexpr int CY = 4, ESTY = R2Y + 1, B1Y = zf_log2(R1Y), B2Y = zf_log2(R2Y), RMY = R1Y > R2Y ? R1Y : R2Y;
  constexpr int YBUF = R1Y * ESTY * CY > NY * CY ? R1Y * ESTY * CY : NY * CY;   // landing zone / exchange / output, aliased
  constexpr int BUF = CX * SLOT > YBUF ? CX * SLOT : YBUF;
  constexpr int RB = NY < 256 ? NY : 256, NOPS = NY / RB;
  static_assert(4 * RMY <= 128 && CX * R2 == 128, "thread mapping");
  extern __shared__ __align__(128) unsigned char f2_smem[];
  double2* buf = (double2*)f2_smem;                  // [BUF]
  double2* stwM = buf + BUF;                         // [M]
  double2* stwH = stwM + M;                          // [M/2+1]
  double2* stwY = stwH + (M / 2 + 1);                // [NY]
  __shared__ __align__(8) unsigned long long full;
  __shared__ unsigned s_item;
  const int tid = threadIdx.x;
  for (int t = tid; t < M; t += 128) stwM[t] = __ldg(a.twM + t);
  for (int t = tid; t <= M / 2; t += 128) {
    const double2 w = __ldg(a.twN + t);
    stwH[t] = INVERSE ? make_double2(w.y, w.x) : make_double2(0.5 * w.y, -0.5 * w.x);
  }
  for (int t = tid; t < NY; t += 128) stwY[t] = __ldg(a.twY + t);
  if (tid == 0) mbar_init(&full, 1);
  __syncthreads();
  // first-pass items read their input for the last time and write the intermediate; second-pass items read the
  // intermediate for the last time and write the final result
  const u64 pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
  const int N2 = NY;
  const unsigned xpp = (unsigned)(N2 / CX), ypp = (unsigned)a.nicg;          // items per plane
  const unsigned B = (unsigned)a.bplanes, nb = (unsigned)a.nplanes / B;
  // first pass = X (forward) or Y (inverse); second pass one block behind
  const unsigned f_pp = INVERSE ? ypp : xpp, s_pp = INVERSE ? xpp : ypp;
  const unsigned S = B * (f_pp + s_pp);
  const unsigned total = (unsigned)a.nplanes * (xpp + ypp);
  constexpr unsigned IN_BYTES = INVERSE ? (M + 1) * 16u : M * 16u;
  constexpr unsigned OUT_BYTES = INVERSE ? M * 16u : (M + 1) * 16u;
  constexpr int NPAIR = M / 2 + 1;
  constexpr int SIGNY = INVERSE ? +1 : -1;
  unsigned phase = 0u;
  for (;;) {
    if (tid == 0) s_item = atomicAdd(a.queue, 1u);
    __syncthreads();
    const unsigned item = s_item;
    __syncthreads();
    if (item >= total) break;
    // decode: (first | second pass, plane, sub-item)
    bool second;
    unsigned pl, sub;
    if (item < B * f_pp) { second = false; pl = item / f_pp; sub = item % f_pp; }
    else {
      const unsigned i2 = item - B * f_pp, s = 1u + i2 / S, r = i2 % S;
      if (s < nb && r < B * f_pp) { second = false; pl = s * B + r / f_pp; sub = r % f_pp; }
      else {
        const unsigned r2 = s < nb ? r - B * f_pp : r;
        second = true; pl = (s - 1u) * B + r2 / s_pp; sub = r2 % s_pp;
      }
    }
    const bool is_x = (second == INVERSE);
    if (second) {
      // dependency: every first-pass item of this plane has finished (and its stores are visible)
      if (tid == 0) {
        unsigned spins = 0;
        while (atomicAdd(a.done + pl, 0u) < f_pp) {
          __nanosleep(200);
          if (++spins > (1u << 24)) { atomicExch(a.fault, 1u); break; }
        }
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");   // the data was written, and is about to be read, through the async proxy
      }
    }
    double* pbase = a.mesh + a.plane * (i64)pl;
    if (is_x) {
      const i64 row0 = (i64)sub * CX;
      if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_expect_tx(&full, (unsigned)CX * IN_BYTES);
        for (int r = 0; r < CX; r++) bulk_load_1d_hint(buf + r * SLOT, pbase + (row0 + r) * a.rowp, IN_BYTES, &full, pol_first);
      }
      mbar_wait(&full, phase);
      phase ^= 1u;
      if (INVERSE) {
        for (int t = tid; t < CX * NPAIR; t += 128) {
          double2* slot = buf + (t / NPAIR) * SLOT;
          const int k = t % NPAIR;
          double2 xa = slot[k], xb = slot[M - k];
          if (k == 0) { xa.y = 0.0; xb.y = 0.0; }
          const double2 h = stwH[k];
          const double2 e = make_double2(xa.x + xb.x, xa.y - xb.y);
          const double2 d = make_double2(xa.x - xb.x, xa.y + xb.y);
          const double2 u = make_double2(h.x * d.x - h.y * d.y, h.x * d.y + h.y * d.x);
          slot[k] = make_double2(e.x + u.x, e.y + u.y);
          if (k != 0 && k != M - k) slot[M - k] = make_double2(e.x - u.x, -e.y + u.y);
        }
        __syncthreads();
        rowfft_core<R1, R2, CX, SLOT, +1>(buf, stwM, tid);
      } else {
        rowfft_core<R1, R2, CX, SLOT, -1>(buf, stwM, tid);
        for (int t = tid; t < CX * NPAIR; t += 128) {
          double2* slot = buf + (t / NPAIR) * SLOT;
          const int k = t % NPAIR;
          const double2 za = slot[k], zb = slot[k == 0 ? 0 : M - k];
          const double2 h = stwH[k];
          const double2 sh = make_double2(0.5 * (za.x + zb.x), 0.5 * (za.y - zb.y));
          const double2 d = make_double2(za.x - zb.x, za.y + zb.y);
          const double2 u = make_double2(h.x * d.x - h.y * d.y, h.x * d.y + h.y * d.x);
          slot[k] = make_double2(sh.x + u.x, sh.y + u.y);
          slot[M - k] = make_double2(sh.x - u.x, -sh.y + u.y);
        }
        __syncthreads();
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        for (int r = 0; r < CX; r++) bulk_store_1d_hint(pbase + (row0 + r) * a.rowp, buf + r * SLOT, OUT_BYTES, second ? pol_first : pol_last);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      // ---- Y item: column group `sub` (4 complex columns) of plane pl, all N2 rows, single buffer
      const int c = tid % CY, j = tid / CY;
      if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_expect_tx(&full, (unsigned)(NY * CY * sizeof(double2)));
#pragma unroll
        for (int o = 0; o < NOPS; o++) tma_load_3d_hint(buf + o * RB * CY, &tmap, (int)sub * 8, o * RB, (int)pl, &full, pol_first);
      }
      mbar_wait(&full, phase);
      phase ^= 1u;
      double2 v[RMY];
      if (j < R2Y) {
#pragma unroll
        for (int r = 0; r < R1Y; r++) v[r] = buf[(j + r * R2Y) * CY + c];
        ZDif<R1Y, 0, SIGNY>::run(v);
      }
      __syncthreads();   // the group is in registers: the buffer becomes the padded exchange
      if (j < R2Y) {
#pragma unroll
        for (int q = 0; q < R1Y; q++) buf[(q * ESTY + j) * CY + c] = v[zf_brev(q, B1Y)];
      }
      __syncthreads();
      if (j < R1Y) {
#pragma unroll
        for (int r = 0; r < R2Y; r++) v[r] = buf[(j * ESTY + r) * CY + c];
#pragma unroll
        for (int r = 1; r < R2Y; r++) {
          const double2 w = stwY[r * j];
          v[r] = zmul(v[r], w.x, SIGNY < 0 ? w.y : -w.y);
        }
        ZDif<R2Y, 0, SIGNY>::run(v);
      }
      __syncthreads();   // the exchange is consumed: natural order for the tensor store
      if (j < R1Y) {
#pragma unroll
        for (int q = 0; q < R2Y; q++) buf[(j + q * R1Y) * CY + c] = v[zf_brev(q, B2Y)];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int o = 0; o < NOPS; o++) tma_store_3d_hint(&tmap, buf + o * RB * CY, (int)sub * 8, o * RB, (int)pl, second ? pol_first : pol_last);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (!second && tid == 0) {
      // publish: the stores must be COMPLETE (not merely read out of shared memory) before the counter moves
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
      __threadfence();
      atomicAdd(a.done + pl, 1u);
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int R1, int R2, int R1Y, int R2Y, bool INVERSE>
static int fft2d_go(nnpm_ctx* ctx, const Fft2dArgs& a) {
  constexpr int CX = 128 / R2, M = R1 * R2, NY = R1Y * R2Y;
  constexpr int SLOT = R1 * (R2 + 1) > M + 1 ? R1 * (R2 + 1) : M + 1;
  constexpr int YBUF = R1Y * (R2Y + 1) * 4 > NY * 4 ? R1Y * (R2Y + 1) * 4 : NY * 4;
  constexpr int BUF = CX * SLOT > YBUF ? CX * SLOT : YBUF;
  const size_t smem = ((size_t)BUF + M + (M / 2 + 1) + NY) * 16;
  CK(cudaFuncSetAttribute(k_fft2d<R1, R2, R1Y, R2Y, INVERSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const i64 total = (i64)a.nplanes * (NY / CX + a.nicg);
  const i64 maxg = (i64)ctx->nsm * 2;
  const unsigned grid = (unsigned)(total < maxg ? total : maxg);
  k_fft2d<R1, R2, R1Y, R2Y, INVERSE><<<grid, 128, smem, ctx->st>>>(ctx->tm_y, a);
  LAUNCHED(ctx);
  KCHECK();
  return NNPM_OK;
}

static bool fft2d_supported(const nnpm_ctx* ctx) {
  const Geom& g = ctx->g;
  return g.N1 == g.N2 && (g.N1 == 128 || g.N1 == 256 || g.N1 == 512 || g.N1 == 1024);
}

// fused x + y passes over planes [plane0, plane0 + nplanes) of the owned slab, in place
static int fft2d_xy(nnpm_ctx* ctx, bool inverse, i64 plane0 = 0, i64 nplanes = -1) {
  const Geom& g = ctx->g;
  CKR(rowfft_tables(ctx));
  CKR(colfft_twiddles(ctx, &ctx->ytw, g.N2));
  if (!ctx->have_tm_y) {
    CKR(colfft_tensor_map(ctx, &ctx->tm_y, ctx->mesh + g.plane * g.glo, g.N2, g.nzl, ctx->icn * 16, g.plane * 8, true, g.N2));
    ctx->have_tm_y = true;
  }
  if (nplanes < 0) nplanes = g.nzl;
  if (nplanes == 0) return NNPM_OK;
  if (!ctx->f2_sync) CKR(dalloc(ctx, &ctx->f2_sync, (size_t)g.nzl + 8));
  CK(cudaMemsetAsync(ctx->f2_sync, 0, (size_t)(g.nzl + 8) * 4, ctx->st));
  Fft2dArgs a;
  a.mesh = ctx->mesh + g.plane * g.glo;   // plane indices are relative to the tensor map's origin: the first owned plane
  a.rowp = g.rowp;
  a.plane = g.plane;
  a.nplanes = (int)nplanes;
  const int want = ctx->opt_fft2d_block > 0 ? ctx->opt_fft2d_block : 4;
  int b = want;
  while (b > 1 && nplanes % b) b--;
  a.bplanes = b;
  a.icn = (int)ctx->icn;
  a.nicg = (int)((ctx->icn + 3) / 4);
  a.twM = ctx->xtwM; a.twN = ctx->xtwN; a.twY = ctx->ytw;
  a.queue = ctx->f2_sync;
  a.fault = ctx->f2_sync + 1;
  a.done = ctx->f2_sync + 8;
  if (plane0 != 0) return fail(ctx, NNPM_ERR_STATE, "fused 2-D transform: plane sub-ranges are not supported");
#define F2(R1_, R2_, RY1_, RY2_) return inverse ? fft2d_go<R1_, R2_, RY1_, RY2_, true>(ctx, a) : fft2d_go<R1_, R2_, RY1_, RY2_, false>(ctx, a)
  switch (g.N1) {
    case 128: F2(8, 8, 16, 8);
    case 256: F2(16, 8, 16, 16);
    case 512: F2(16, 16, 32, 16);
    case 1024: F2(32, 16, 32, 32);
  }
#undef F2
  return fail(ctx, NNPM_ERR_STATE, "fused 2-D transform does not support N1 = N2 = %lld", (long long)g.N1);
}
#endif  // NNPM_EXPERIMENTS
