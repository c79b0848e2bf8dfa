// nnpm_colfft.cuh -- strided-column FFT engine with TMA staging (sm_100a), used for
//   * the y pass of the 2-D transforms (forward after cuFFT's x r2c, inverse before its x c2r), and
//   * the fused z pass [z-FFT -> Green's function multiply (ParticleMesh.jl:482-502) -> z-iFFT].
//
// Why: a column of the half spectrum is N elements 8 KB .. 8 MB apart.  Load/store instructions
// issued from registers (k_zfused, cuFFT's regular_fft) keep too few bytes in flight per SM for that
// access pattern (ncu: 25-31 % of DRAM peak, lg_throttle / long_scoreboard stalls).  Here the TMA
// engine moves whole column groups: one persistent CTA per SM cycles two shared-memory buffers of
// [N][4] double2 (a group of 4 adjacent columns = 64 B per row); while the CTA transforms group g in
// registers + a padded exchange buffer, cp.async.bulk.tensor loads of group g+1 and the tensor store
// of group g-1 are in flight.  The threads never touch global memory.
//
//   tensor map   3-D over doubles: (2*icn, n_mid, n_outer) = (ic pairs, j, k); box (8, 1|RB, RB|1)
//   y pass       column = (plane, ic), rows = j:  coords (8*icg, row0, plane)
//   z pass       column = (jl, ic),   rows = k:  coords (8*icg, jl, row0)
//   columns past icn (513 = 128*4 + 1) are out of bounds: TMA zero-fills them on load and drops
//   them on store.
//
// Arithmetic: N = R1*R2, two register-resident radix passes (ZDif/ZDit of nnpm_zfft.cuh).
//   pass A  thread n' (< R2): rows n' + n1*R2 -> radix-R1 DIF -> y1[k1][n']  -> exchange E[k1*(R2+1) + n']
//   pass B  thread k1 (< R1): E[k1*(R2+1) + n'] * W_N^(n' k1) -> radix-R2 DIF -> X[k1 + R1*k2]
// E is padded by one element per k1 row so that both sides of the exchange are bank-conflict free
// with 4 columns per row (quarter-warps touch two rows whose indices differ by an odd number).
#pragma once

enum { CF_FWD = 0, CF_INV = 1 };

struct ColFftArgs {
  int ngroups;        // column groups = n_outer_cols * nicg
  int nicg;           // groups per outer index = ceil(icn / columns per group)
  int icn;
  int run;            // adjacent groups a CTA processes back to back (1 or 2)
  int outer0;         // first outer index (plane / local row) of this launch: a pass can run on a sub-range
  i64 j0;             // global j of local row-block start (Green's function, z pass)
  const double* s1;   // sin^2 tables (z pass)
  const double* s2;
  const double* s3;
  const double2* tw;  // W_N^m, m < N
  double scale;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// y pass: rows run along tensor dimension 1 (j), the outer column index along dimension 2 (plane).
// PEERS > 1 (forward pass of the slab-decomposed solve): the transformed group is not stored back in place but
// straight into the slab-transposed spectrum T_d[k][jl][ic] of the rank d = j / njl that owns row block d -- one
// TMA tensor store per destination through that peer's tensor map (NVLink peer memory, or local memory for
// d == this rank), i.e. the y pass and the forward all-to-all are ONE kernel and the NVLink transfer of group g
// overlaps the arithmetic of group g + 1.
struct PeerMaps { CUtensorMap m[8]; };
// The transposes fused into the transform kernels' stores ("fuse_transpose" = 1) are bit-identical to the separate
// transposes but measured slower: their PEER = true instantiations are only built with
// -DNNPM_EXPERIMENTS (make EXPERIMENTS=1); the default library refuses the option.
#ifdef NNPM_EXPERIMENTS
constexpr bool NNPM_EXP_PEER = true;
#else
constexpr bool NNPM_EXP_PEER = false;
#endif

template <int R1, int R2, int MODE, bool PEER>
__global__ void __launch_bounds__(4 * (R1 > R2 ? R1 : R2), 1)
k_colfft(const __grid_constant__ CUtensorMap tmap, ColFftArgs a, const __grid_constant__ PeerMaps pm, int npeers, int njl, int kz0) {
  constexpr int C = 4, N = R1 * R2, RM = R1 > R2 ? R1 : R2, B1 = zf_log2(R1), B2 = zf_log2(R2);
  constexpr int RB = N < 256 ? N : 256, NOPS = N / RB;         // rows per TMA operation (box limit 256)
  constexpr int EST = R2 + 1;                                   // padded exchange row
  constexpr int SIGN = (MODE == CF_INV) ? +1 : -1;              // sign of the transform
  extern __shared__ __align__(128) unsigned char cf_smem[];
  double2* buf0 = (double2*)cf_smem;                            // [2][N][C]
  double2* E = buf0 + 2 * N * C;                                // [R1][EST][C]
  double2* stw = E + R1 * EST * C;                              // [N]
  __shared__ __align__(8) unsigned long long full[2];
  const int tid = threadIdx.x, c = tid % C, j = tid / C;
  for (int t = tid; t < N; t += blockDim.x) stw[t] = __ldg(a.tw + t);
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
  }
  __syncthreads();
  // Group schedule: CTA b works on runs of `a.run` adjacent groups, runs dealt round-robin to the CTAs:
  //   g(b, it) = ((it / run) * gridDim + b) * run + it % run.
  // At any time the CTAs together cover one contiguous stretch of every row (DRAM page locality).
  const int run = a.run;
  auto group_of = [&](int it) { return ((it / run) * (int)gridDim.x + (int)blockIdx.x) * run + it % run; };
  auto issue_load = [&](int g, int b) {  // thread 0 only
    const int og = g / a.nicg, icg = g - og * a.nicg, outer = og + a.outer0;
    mbar_expect_tx(&full[b], (unsigned)(N * C * sizeof(double2)));
#pragma unroll
    for (int o = 0; o < NOPS; o++) tma_load_3d(buf0 + (b * N + o * RB) * C, &tmap, icg * 8, o * RB, outer, &full[b]);
  };
  int g = group_of(0);
  if (tid == 0 && g < a.ngroups) issue_load(g, 0);
  unsigned phase[2] = {0u, 0u};
  double2 v[RM];
  for (int it = 0; g < a.ngroups; it++) {
    const int b = it & 1;
    const int gnext = group_of(it + 1);
    const int og_ = g / a.nicg, icg = g - og_ * a.nicg, outer = og_ + a.outer0;
    // prefetch the next group into the other buffer: its previous contents were stored one iteration
    // ago; the store must have finished READING shared memory before the load overwrites it
    if (tid == 0 && gnext < a.ngroups) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue_load(gnext, b ^ 1);
    }
    double2* B = buf0 + b * N * C;
    mbar_wait(&full[b], phase[b]);
    phase[b] ^= 1u;
    // ---- pass A: radix R1 over rows j + r*R2
    if (j < R2) {
#pragma unroll
      for (int r = 0; r < R1; r++) v[r] = B[(j + r * R2) * C + c];
      ZDif<R1, 0, SIGN>::run(v);
#pragma unroll
      for (int q = 0; q < R1; q++) E[(q * EST + j) * C + c] = v[zf_brev(q, B1)];
    }
    __syncthreads();
    // ---- pass B: radix R2 over n' for k1 = j
    if (j < R1) {
#pragma unroll
      for (int r = 0; r < R2; r++) v[r] = E[(j * EST + r) * C + c];
#pragma unroll
      for (int r = 1; r < R2; r++) {
        const double2 w = stw[r * j];
        v[r] = zmul(v[r], w.x, SIGN < 0 ? w.y : -w.y);
      }
      ZDif<R2, 0, SIGN>::run(v);  // v[brev(q)] = X[j + q*R1]
#pragma unroll
      for (int q = 0; q < R2; q++) B[(j + q * R1) * C + c] = v[zf_brev(q, B2)];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> TMA store reads
    __syncthreads();
    if (tid == 0) {
      if (PEER) {
        // rows [d*njl, (d+1)*njl) of the transformed group belong to rank d: T_d[kz0 + outer][jl][ic]
        const int rbp = njl < 256 ? njl : 256;
        for (int d = 0; d < npeers; d++)
          for (int o = 0; o < njl; o += rbp) tma_store_3d(&pm.m[d], B + (d * njl + o) * C, icg * 8, o, kz0 + outer);
      } else {
#pragma unroll
        for (int o = 0; o < NOPS; o++) tma_store_3d(&tmap, B + o * RB * C, icg * 8, o * RB, outer);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    g = gnext;
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete (also the remote ones) before exit
}

// Destination of the fused z pass' output rows.  Row n (global plane k = n) of column (jl, ic) is written to
//   base[n >> shift] + (n & mask) * plane_c + (row0 + jl) * icn + ic                       (complex elements)
// * in place (single GPU, or the un-fused multi-GPU path on meshT): shift = 31, base[0] = T, plane_c = njl * icn,
//   row0 = 0 -- the same element the column was read from;
// * fused backward slab transpose (nranks > 1, peer memory): base[d] = rank d's mesh slab mapped over NVLink (or
//   the local one), shift = log2(nzl), plane_c = plane / 2, row0 = njl * rank: the z pass stores every row
//   straight into the [kl][j][ic] layout of the rank that owns plane k, so the inverse all-to-all costs no extra
//   pass over HBM and its NVLink transfer runs under the arithmetic of the next column group.
struct ZDst {
  double2* base[8];
  int shift, mask;
  i64 plane_c, row0;
};

// Persistent form of k_zfused (nnpm_zfft.cuh) with the NEXT column group prefetched by TMA: 8 adjacent
// columns per group (128-byte rows: exact DRAM traffic), 8 warps, the [N][8] shared-memory buffer is
// both the landing zone of the tensor load and the exchange buffer of the two radix passes.  As soon
// as the last exchange read of group g is done (start of the final inverse pass), one thread issues the
// tensor load of group g + gridDim.x, so the load latency and transfer run under the final pass'
// arithmetic and its stores instead of in front of the next group's first pass.
//
// TWO: transform length NT = 2 * R1 * R2 (N3 = 2048 = 2 x 32 x 32, the fine mesh of BASELINE config 5) by one
// extra radix-2 step around the two-pass engine, without a third pass over shared memory.  A group is 4 real
// columns (64-byte rows, [NT][4] in shared memory = the same 128 KB); the 8 "virtual columns" the 8 thread
// columns work on are (real column, parity): virtual column (cc, 0) transforms e[n] = a[n] + a[n + N] and
// yields the even bins X[2q], (cc, 1) transforms o[n] = (a[n] - a[n + N]) W_NT^n and yields the odd bins
// X[2q + 1] (decimation in frequency).  Pass A forms e / o on the fly from the landing zone (two reads instead
// of one); the Green factor uses the true bin kz = 2q + parity; after the inverse passes the two parities of a
// column sit in lanes l and l ^ 4 of one warp and are combined by one shuffle exchange:
// x[n] = E[n] + W_NT^-n O[n],  x[n + N] = E[n] - W_NT^-n O[n].
template <int R1, int R2, bool TWO, bool PEER>
__global__ void __launch_bounds__(8 * (R1 > R2 ? R1 : R2), 1)
k_zfused_p(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ZDst dst, ColFftArgs a) {
  constexpr int C = 8, N = R1 * R2, RM = R1 > R2 ? R1 : R2, B1 = zf_log2(R1), B2 = zf_log2(R2);
  constexpr int NT = TWO ? 2 * N : N;   // transform length
  constexpr int CR = TWO ? 4 : 8;       // real columns per group
  constexpr int TS = TWO ? 2 : 1;       // W_N^m = stw[TS * m]  (stw holds W_NT^m)
  constexpr int RB = NT < 256 ? NT : 256, NOPS = NT / RB;
  extern __shared__ __align__(128) unsigned char cf_smem[];
  double2* zsm = (double2*)cf_smem;  // [NT][CR]
  double2* stw = zsm + NT * CR;
  double* ss3 = (double*)(stw + NT);
  __shared__ __align__(8) unsigned long long full;
  const int tid = threadIdx.x, c = tid % C, j = tid / C;
  const int cc = TWO ? (c & 3) : c, par = TWO ? (c >> 2) : 0;
  // element (row r of the length-N sub-transform, virtual column c) of the exchange buffer
  auto Z = [&](int r) { return TWO ? (r + N * par) * CR + cc : r * C + c; };
  for (int t = tid; t < NT; t += blockDim.x) {
    stw[t] = __ldg(a.tw + t);
    ss3[t] = __ldg(a.s3 + t);
  }
  if (tid == 0) mbar_init(&full, 1);
  __syncthreads();
  auto issue_load = [&](int g) {  // thread 0 only
    const int og = g / a.nicg, icg = g - og * a.nicg;
    mbar_expect_tx(&full, (unsigned)(NT * CR * sizeof(double2)));
#pragma unroll
    for (int o = 0; o < NOPS; o++) tma_load_3d(zsm + o * RB * CR, &tmap, icg * (2 * CR), og + a.outer0, o * RB, &full);
  };
  int g = blockIdx.x;
  if (tid == 0 && g < a.ngroups) issue_load(g);
  unsigned phase = 0u;
  double2 v[RM];
  for (; g < a.ngroups; g += gridDim.x) {
    const int og = g / a.nicg, icg = g - og * a.nicg, jl = og + a.outer0;
    const int ic = icg * CR + cc;
    const bool colok = ic < a.icn;
    double sxy = 1.0;
    if (colok) sxy = __ldg(a.s1 + ic) + __ldg(a.s2 + a.j0 + jl);
    const bool dc_col = (ic == 0 && a.j0 + jl == 0);
    mbar_wait(&full, phase);
    phase ^= 1u;
    // ---- forward pass A: radix R1 over rows j + r*R2 (from the landing zone)
    if (j < R2) {
#pragma unroll
      for (int r = 0; r < R1; r++) {
        const int rr = j + r * R2;
        if (!TWO) {
          v[r] = zsm[rr * C + c];
        } else {
          const double2 a0 = zsm[rr * CR + cc], a1 = zsm[(rr + N) * CR + cc];
          if (par) {
            const double2 w = stw[rr];
            v[r] = zmul(zsub(a0, a1), w.x, w.y);
          } else {
            v[r] = zadd(a0, a1);
          }
        }
      }
      ZDif<R1, 0, -1>::run(v);
    }
    __syncthreads();  // inputs are in registers: the buffer becomes the exchange
    if (j < R2) {
#pragma unroll
      for (int q = 0; q < R1; q++) zsm[Z(j * R1 + q)] = v[zf_brev(q, B1)];
    }
    __syncthreads();
    // ---- forward pass B, Green's function, inverse pass A'
    if (j < R1) {
#pragma unroll
      for (int r = 0; r < R2; r++) v[r] = zsm[Z(j + r * R1)];
    }
    __syncthreads();
    if (j < R1) {
#pragma unroll
      for (int r = 1; r < R2; r++) {
        const double2 w = stw[TS * r * j];
        v[r] = zmul(v[r], w.x, w.y);
      }
      ZDif<R2, 0, -1>::run(v);
#pragma unroll
      for (int q = 0; q < R2; q++) {  // X[kz] *= scale / Ginv   (ParticleMesh.jl:494-501)
        const int kz = TWO ? 2 * (j + q * R1) + par : j + q * R1;
        const double ginv = -4.0 * (sxy + ss3[kz]);
        double f = (ginv != 0.0) ? a.scale * zf_rcp(ginv) : a.scale;
        if (dc_col && kz == 0) f = 0.0;
        double2& e = v[zf_brev(q, B2)];
        e.x *= f;
        e.y *= f;
      }
      ZDit<R2, 0, +1>::run(v);
#pragma unroll
      for (int r = 0; r < R2; r++) zsm[Z(j * R2 + r)] = v[r];
    }
    __syncthreads();
    // ---- inverse pass B': radix R1 -> rows j + q*R2, stored straight from registers
    if (j < R2) {
#pragma unroll
      for (int r = 0; r < R1; r++) v[r] = zsm[Z(j + r * R2)];
    }
    __syncthreads();  // the exchange is consumed: the buffer is free for the next group's tensor load
    if (tid == 0 && g + (int)gridDim.x < a.ngroups) issue_load(g + gridDim.x);
    if (j < R2) {  // warp-uniform: R2 is a multiple of the 4 values of j a warp holds
#pragma unroll
      for (int r = 1; r < R1; r++) {
        const double2 w = stw[TS * r * j];
        v[r] = zmul(v[r], w.x, -w.y);
      }
      ZDif<R1, 0, +1>::run(v);
      const i64 col = (dst.row0 + jl) * (i64)a.icn + ic;
#pragma unroll
      for (int q = 0; q < R1; q++) {
        const int n = j + q * R2;
        double2 e = v[zf_brev(q, B1)];
        if (TWO) {  // lanes l (parity 0: E[n]) and l ^ 4 (parity 1: O[n]) hold the same column
          if (par) {
            const double2 w = stw[n];
            e = zmul(e, w.x, -w.y);
          }
          double2 o;
          o.x = __shfl_xor_sync(0xffffffffu, e.x, 4);
          o.y = __shfl_xor_sync(0xffffffffu, e.y, 4);
          e = par ? zsub(o, e) : zadd(e, o);
        }
        const int nn = n + N * par;
        if (!colok) continue;
        if (PEER) dst.base[nn >> dst.shift][(i64)(nn & dst.mask) * dst.plane_c + col] = e;
        else dst.base[0][(i64)nn * dst.plane_c + col] = e;
      }
    }
  }
}

// k_zfused_p with the landing zone and the exchange buffer SEPARATED, so that the tensor load of the next column
// group is in flight during the whole transform of the current one instead of during its last pass only.  Round 1's
// ncu capture of k_zfused_p showed the FP64 pipe and DRAM each ~45 % busy but never at the same time: one [N][8]
// double2 buffer (128 KB at N = 1024) was both, and a second one does not fit.  Here the exchange moves real and
// imaginary parts in two half-steps through a [N][8] buffer of DOUBLES (64 KB): 128 KB landing zone + 64 KB
// exchange + 24 KB tables = 216 KB.  The exchange rows are skewed by 8 doubles per 32 rows so that the four rows a
// warp touches in the scatter half-step fall on two bank groups (2 wavefronts per 256 bytes: the minimum).
template <int R1, int R2, bool PEER>
__global__ void __launch_bounds__(8 * (R1 > R2 ? R1 : R2), 1)
k_zfused_s(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ZDst dst, ColFftArgs a) {
  constexpr int C = 8, N = R1 * R2, RM = R1 > R2 ? R1 : R2, B1 = zf_log2(R1), B2 = zf_log2(R2);
  constexpr int RB = N < 256 ? N : 256, NOPS = N / RB;
  extern __shared__ __align__(128) unsigned char cf_smem[];
  double2* L = (double2*)cf_smem;                 // landing zone [N][C]
  double* E = (double*)(L + N * C);               // exchange [N][C] doubles, skewed
  double2* stw = (double2*)(E + N * C + C * (N / 32));
  double* ss3 = (double*)(stw + N);
  __shared__ __align__(8) unsigned long long full;
  const int tid = threadIdx.x, c = tid % C, j = tid / C;
  auto EZ = [&](int r) { return r * C + c + C * (r >> 5); };
  for (int t = tid; t < N; t += blockDim.x) {
    stw[t] = __ldg(a.tw + t);
    ss3[t] = __ldg(a.s3 + t);
  }
  if (tid == 0) mbar_init(&full, 1);
  __syncthreads();
  auto issue_load = [&](int g) {  // thread 0 only
    const int og = g / a.nicg, icg = g - og * a.nicg;
    mbar_expect_tx(&full, (unsigned)(N * C * sizeof(double2)));
#pragma unroll
    for (int o = 0; o < NOPS; o++) tma_load_3d(L + o * RB * C, &tmap, icg * (2 * C), og + a.outer0, o * RB, &full);
  };
  int g = blockIdx.x;
  if (tid == 0 && g < a.ngroups) issue_load(g);
  unsigned phase = 0u;
  double2 v[RM];
  for (; g < a.ngroups; g += gridDim.x) {
    const int og = g / a.nicg, icg = g - og * a.nicg, jl = og + a.outer0;
    const int ic = icg * C + c;
    const bool colok = ic < a.icn;
    double sxy = 1.0;
    if (colok) sxy = __ldg(a.s1 + ic) + __ldg(a.s2 + a.j0 + jl);
    const bool dc_col = (ic == 0 && a.j0 + jl == 0);
    mbar_wait(&full, phase);
    phase ^= 1u;
    // ---- forward pass A: radix R1 over rows j + r*R2, read from the landing zone
    if (j < R2) {
#pragma unroll
      for (int r = 0; r < R1; r++) v[r] = L[(j + r * R2) * C + c];
    }
    __syncthreads();  // the landing zone is in registers: the next group's tensor load may overwrite it
    if (tid == 0 && g + (int)gridDim.x < a.ngroups) issue_load(g + gridDim.x);
    if (j < R2) ZDif<R1, 0, -1>::run(v);
    // ---- exchange 1 (two half-steps): pass-A thread j' writes (j'*R1 + q), pass-B thread j reads (j + r*R1)
    if (j < R2) {
#pragma unroll
      for (int q = 0; q < R1; q++) E[EZ(j * R1 + q)] = v[zf_brev(q, B1)].x;
    }
    __syncthreads();
    if (j < R1) {
#pragma unroll
      for (int r = 0; r < R2; r++) v[r].x = E[EZ(j + r * R1)];
    }
    __syncthreads();
    if (j < R2) {
#pragma unroll
      for (int q = 0; q < R1; q++) E[EZ(j * R1 + q)] = v[zf_brev(q, B1)].y;
    }
    __syncthreads();
    if (j < R1) {
#pragma unroll
      for (int r = 0; r < R2; r++) v[r].y = E[EZ(j + r * R1)];
      // ---- forward pass B, Green's function, inverse pass A'
#pragma unroll
      for (int r = 1; r < R2; r++) {
        const double2 w = stw[r * j];
        v[r] = zmul(v[r], w.x, w.y);
      }
      ZDif<R2, 0, -1>::run(v);
#pragma unroll
      for (int q = 0; q < R2; q++) {  // X[kz] *= scale / Ginv   (ParticleMesh.jl:494-501)
        const int kz = j + q * R1;
        const double ginv = -4.0 * (sxy + ss3[kz]);
        double f = (ginv != 0.0) ? a.scale * zf_rcp(ginv) : a.scale;
        if (dc_col && kz == 0) f = 0.0;
        double2& e = v[zf_brev(q, B2)];
        e.x *= f;
        e.y *= f;
      }
      ZDit<R2, 0, +1>::run(v);
    }
    __syncthreads();  // every read of exchange 1 is done
    // ---- exchange 2: pass-A' thread j' writes (j'*R2 + r), pass-B' thread j reads (j + r*R2)
    if (j < R1) {
#pragma unroll
      for (int r = 0; r < R2; r++) E[EZ(j * R2 + r)] = v[r].x;
    }
    __syncthreads();
    if (j < R2) {
#pragma unroll
      for (int r = 0; r < R1; r++) v[r].x = E[EZ(j + r * R2)];
    }
    __syncthreads();
    if (j < R1) {
#pragma unroll
      for (int r = 0; r < R2; r++) E[EZ(j * R2 + r)] = v[r].y;
    }
    __syncthreads();
    // ---- inverse pass B': radix R1 -> rows j + q*R2, stored straight from registers
    if (j < R2) {
#pragma unroll
      for (int r = 0; r < R1; r++) v[r].y = E[EZ(j + r * R2)];
#pragma unroll
      for (int r = 1; r < R1; r++) {
        const double2 w = stw[r * j];
        v[r] = zmul(v[r], w.x, -w.y);
      }
      ZDif<R1, 0, +1>::run(v);
      const i64 col = (dst.row0 + jl) * (i64)a.icn + ic;
      double2* const base0 = dst.base[0] + col;   // in place: one base for every row (no per-row constant-bank lookup)
#pragma unroll
      for (int q = 0; q < R1; q++) {
        const int n = j + q * R2;
        if (!colok) continue;
        if (PEER) dst.base[n >> dst.shift][(i64)(n & dst.mask) * dst.plane_c + col] = v[zf_brev(q, B1)];
        else base0[(i64)n * dst.plane_c] = v[zf_brev(q, B1)];
      }
    }
    // the next iteration's first barrier (after its landing-zone reads) separates these exchange reads from the
    // next exchange writes.  (Tried: handing the result to the TMA engine through the exchange buffer, two half-group
    // tensor stores -- 5.25-5.55 ms against 5.0-5.3 for these register stores, profiles/r02_probe_logs/r2n_probe.log.)
  }
}

// ---- host side -------------------------------------------------------------------------------
static int colfft_tensor_map(nnpm_ctx* ctx, CUtensorMap* tm, double* base, i64 n_mid, i64 n_outer, i64 mid_stride_b,
                             i64 outer_stride_b, bool ypass, i64 N, int inner = 8) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(ctx, NNPM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint32_t rb = (cuuint32_t)(N < 256 ? N : 256);
  cuuint64_t dims[3] = {(cuuint64_t)(2 * ctx->icn), (cuuint64_t)n_mid, (cuuint64_t)n_outer};
  cuuint64_t strides[2] = {(cuuint64_t)mid_stride_b, (cuuint64_t)outer_stride_b};
  cuuint32_t box[3] = {(cuuint32_t)inner, ypass ? rb : 1u, ypass ? 1u : rb};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((encode_fn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, NNPM_ERR_CUDA, "cuTensorMapEncodeTiled (column FFT) failed (%d)", (int)r);
  return NNPM_OK;
}

static int colfft_twiddles(nnpm_ctx* ctx, double2** tw, i64 N) {
  if (*tw) return NNPM_OK;  // W_N^m = exp(-2 pi i m / N), m < N, in extended precision on the host
  std::vector<double2> h((size_t)N);
  for (i64 m = 0; m < N; m++) {
    const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)N;
    h[m] = make_double2((double)cosl(ang), (double)sinl(ang));
  }
  CKR(dalloc(ctx, tw, (size_t)N));
  CK(cudaMemcpyAsync(*tw, h.data(), (size_t)N * sizeof(double2), cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return NNPM_OK;
}

template <int R1, int R2, int MODE, bool PEER>
static int colfft_go(nnpm_ctx* ctx, const CUtensorMap& tm, const ColFftArgs& a, const PeerMaps* pm) {
  constexpr int N = R1 * R2, RM = R1 > R2 ? R1 : R2;
  const size_t smem = (size_t)2 * N * 4 * 16 + (size_t)R1 * (R2 + 1) * 4 * 16 + (size_t)N * 16;
  CK(cudaFuncSetAttribute(k_colfft<R1, R2, MODE, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = a.ngroups < ctx->nsm ? a.ngroups : ctx->nsm;
  k_colfft<R1, R2, MODE, PEER><<<grid, 4 * RM, smem, ctx->st>>>(tm, a, *pm, ctx->P, (int)ctx->njl, (int)ctx->g.kz0);
  LAUNCHED(ctx);
  KCHECK();
  return NNPM_OK;
}

template <int MODE, bool PEER>
static int colfft_dispatch(nnpm_ctx* ctx, i64 N, const CUtensorMap& tm, const ColFftArgs& a, const PeerMaps* pm) {
  switch (N) {
    case 64: return colfft_go<8, 8, MODE, PEER>(ctx, tm, a, pm);
    case 128: return colfft_go<16, 8, MODE, PEER>(ctx, tm, a, pm);
    case 256: return colfft_go<16, 16, MODE, PEER>(ctx, tm, a, pm);
    case 512: return colfft_go<32, 16, MODE, PEER>(ctx, tm, a, pm);
    case 1024: return colfft_go<32, 32, MODE, PEER>(ctx, tm, a, pm);
  }
  return fail(ctx, NNPM_ERR_STATE, "column FFT does not support N = %lld", (long long)N);
}

static bool colfft_supported_n(i64 N) { int r1, r2; return zfused_factor(N, &r1, &r2); }

// y pass over the owned planes, in place in the mesh (forward after the x r2c, inverse before the x c2r).
// to_peers (forward only, nranks > 1): the output goes straight into every rank's slab-transposed spectrum
// (ctx->peer_maps) instead -- the forward all-to-all fused into the pass.
static int colfft_y(nnpm_ctx* ctx, bool inverse, i64 plane0 = 0, i64 nplanes = -1, bool to_peers = false) {
  const Geom& g = ctx->g;
  if (!ctx->have_tm_y) {
    CKR(colfft_tensor_map(ctx, &ctx->tm_y, ctx->mesh + g.plane * g.glo, g.N2, g.nzl, ctx->icn * 16, g.plane * 8, true, g.N2));
    ctx->have_tm_y = true;
  }
  CKR(colfft_twiddles(ctx, &ctx->ytw, g.N2));
  ColFftArgs a;
  memset(&a, 0, sizeof a);
  if (nplanes < 0) nplanes = g.nzl;
  a.nicg = (int)((ctx->icn + 3) / 4);
  a.ngroups = (int)(nplanes * a.nicg);
  a.outer0 = (int)plane0;
  a.icn = (int)ctx->icn;
  a.tw = ctx->ytw;
  a.scale = 1.0;
  a.run = ctx->opt_cfrun_y;
  static const PeerMaps no_peers = {};
  if (to_peers) {
    if (inverse || !ctx->peer_maps) return fail(ctx, NNPM_ERR_STATE, "fused forward transpose: peer tensor maps missing");
    if (!NNPM_EXP_PEER) return fail(ctx, NNPM_ERR_STATE, "fused forward transpose: built without -DNNPM_EXPERIMENTS");
    return colfft_dispatch<CF_FWD, NNPM_EXP_PEER>(ctx, g.N2, ctx->tm_y, a, ctx->peer_maps);
  }
  return inverse ? colfft_dispatch<CF_INV, false>(ctx, g.N2, ctx->tm_y, a, &no_peers)
                 : colfft_dispatch<CF_FWD, false>(ctx, g.N2, ctx->tm_y, a, &no_peers);
}

// tensor maps over every rank's transposed spectrum T_d[k][jl][ic] (peer memory), box = (4 columns, <= 256 rows jl, 1 plane)
static int colfft_peer_maps(nnpm_ctx* ctx) {
  if (ctx->peer_maps) return NNPM_OK;
  const Geom& g = ctx->g;
  if (ctx->P > 8) return fail(ctx, NNPM_ERR_STATE, "fused transposes support nranks <= 8");
  ctx->peer_maps = new PeerMaps();
  memset(ctx->peer_maps, 0, sizeof(PeerMaps));
  const i64 njl = ctx->njl;
  for (int d = 0; d < ctx->P; d++)
    CKR(colfft_tensor_map(ctx, &ctx->peer_maps->m[d], ctx->peer_meshT[d], njl, g.N3, ctx->icn * 16, njl * ctx->icn * 16, true, njl));
  return NNPM_OK;
}

// persistent k_zfused with TMA prefetch over the local rows jl in [jlbeg, jlend)
static bool zfused_p_two(const nnpm_ctx* ctx) {  // N3 = 2 x (two-pass length): 2048 always, smaller lengths as a test hook
  int r1, r2;
  const i64 N3 = ctx->g.N3;
  if (N3 % 2 || !zfused_factor(N3 / 2, &r1, &r2)) return false;
  return !zfused_factor(N3, &r1, &r2) || ctx->opt_zsplit;
}
static void zdst_inplace(nnpm_ctx* ctx, double2* T, ZDst* d) {
  memset(d, 0, sizeof *d);
  d->base[0] = T;
  d->shift = 31;
  d->mask = 0x7fffffff;
  d->plane_c = ctx->njl * ctx->icn;
  d->row0 = 0;
}
// fused backward slab transpose: rows go straight into the owning rank's mesh slab (peer memory)
static bool zdst_peers(nnpm_ctx* ctx, ZDst* d) {
  const Geom& g = ctx->g;
  if (!ctx->peer_ok || ctx->P > 8 || (g.nzl & (g.nzl - 1))) return false;
  memset(d, 0, sizeof *d);
  for (int r = 0; r < ctx->P; r++) d->base[r] = (double2*)(ctx->peer_mesh[r] + g.plane * g.glo);
  int sh = 0;
  while ((1ll << sh) < g.nzl) sh++;
  d->shift = sh;
  d->mask = (int)g.nzl - 1;
  d->plane_c = g.plane / 2;
  d->row0 = ctx->njl * ctx->r;
  return true;
}

static int zfused_p_launch(nnpm_ctx* ctx, double2* T, double scale, i64 jlbeg, i64 jlend, const ZDst* dst = nullptr) {
  const Geom& g = ctx->g;
  const bool two = zfused_p_two(ctx);
  const int cr = two ? 4 : 8;  // real columns per group
  if (!two && !ctx->have_tm_z8) {
    CKR(colfft_tensor_map(ctx, &ctx->tm_z8, (double*)T, ctx->njl, g.N3, ctx->icn * 16, ctx->njl * ctx->icn * 16, false, g.N3, 16));
    ctx->have_tm_z8 = true;
  }
  if (two && !ctx->have_tm_z4) {
    CKR(colfft_tensor_map(ctx, &ctx->tm_z4, (double*)T, ctx->njl, g.N3, ctx->icn * 16, ctx->njl * ctx->icn * 16, false, g.N3, 8));
    ctx->have_tm_z4 = true;
  }
  CKR(colfft_twiddles(ctx, &ctx->ztw, g.N3));
  ColFftArgs a;
  memset(&a, 0, sizeof a);
  a.nicg = (int)((ctx->icn + cr - 1) / cr);
  a.ngroups = (int)((jlend - jlbeg) * a.nicg);
  a.outer0 = (int)jlbeg;
  a.icn = (int)ctx->icn;
  a.j0 = ctx->njl * ctx->r;
  a.s1 = ctx->s1; a.s2 = ctx->s2; a.s3 = ctx->s3;
  a.tw = ctx->ztw;
  a.scale = scale;
  ZDst d;
  if (dst) d = *dst; else zdst_inplace(ctx, T, &d);
#define ZP(R1_, R2_, TWO_)                                                                                             \
  do {                                                                                                                 \
    const size_t smem = (size_t)R1_ * R2_ * (8 * 16 + (TWO_ ? 2 : 1) * (16 + 8));                                        \
    const int grid = a.ngroups < ctx->nsm ? a.ngroups : ctx->nsm;                                                       \
    if (dst) {                                                                                                         \
      CK(cudaFuncSetAttribute(k_zfused_p<R1_, R2_, TWO_, NNPM_EXP_PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_zfused_p<R1_, R2_, TWO_, NNPM_EXP_PEER><<<grid, 8 * (R1_ > R2_ ? R1_ : R2_), smem, ctx->st>>>(TWO_ ? ctx->tm_z4 : ctx->tm_z8, d, a); \
    } else {                                                                                                           \
      CK(cudaFuncSetAttribute(k_zfused_p<R1_, R2_, TWO_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_zfused_p<R1_, R2_, TWO_, false><<<grid, 8 * (R1_ > R2_ ? R1_ : R2_), smem, ctx->st>>>(TWO_ ? ctx->tm_z4 : ctx->tm_z8, d, a); \
    }                                                                                                                  \
    LAUNCHED(ctx);                                                                                                     \
    KCHECK();                                                                                                          \
    return NNPM_OK;                                                                                                    \
  } while (0)
  if (two) {
    switch (g.N3 / 2) {
      case 64: ZP(8, 8, true);
      case 128: ZP(16, 8, true);
      case 256: ZP(16, 16, true);
      case 512: ZP(32, 16, true);
      case 1024: ZP(32, 32, true);
    }
  } else {
    switch (g.N3) {
      case 64: ZP(8, 8, false);
      case 128: ZP(16, 8, false);
      case 256: ZP(16, 16, false);
      case 512: ZP(32, 16, false);
      case 1024: ZP(32, 32, false);
    }
  }
#undef ZP
  return fail(ctx, NNPM_ERR_STATE, "persistent fused z pass does not support N3 = %lld", (long long)g.N3);
}

// split-exchange persistent fused z pass (k_zfused_s) over the local rows jl in [jlbeg, jlend); N3 = 64 .. 1024.
// dst == nullptr: in place in T.
static int zfused_s_launch(nnpm_ctx* ctx, double2* T, double scale, i64 jlbeg, i64 jlend, const ZDst* dst = nullptr) {
  const Geom& g = ctx->g;
  if (!ctx->have_tm_z8) {
    CKR(colfft_tensor_map(ctx, &ctx->tm_z8, (double*)T, ctx->njl, g.N3, ctx->icn * 16, ctx->njl * ctx->icn * 16, false, g.N3, 16));
    ctx->have_tm_z8 = true;
  }
  CKR(colfft_twiddles(ctx, &ctx->ztw, g.N3));
  ColFftArgs a;
  memset(&a, 0, sizeof a);
  a.nicg = (int)((ctx->icn + 7) / 8);
  a.ngroups = (int)((jlend - jlbeg) * a.nicg);
  a.outer0 = (int)jlbeg;
  a.icn = (int)ctx->icn;
  a.j0 = ctx->njl * ctx->r;
  a.s1 = ctx->s1; a.s2 = ctx->s2; a.s3 = ctx->s3;
  a.tw = ctx->ztw;
  a.scale = scale;
  ZDst d;
  if (dst) d = *dst; else zdst_inplace(ctx, T, &d);
#define ZS(R1_, R2_)                                                                                                   \
  do {                                                                                                                 \
    constexpr int N_ = R1_ * R2_;                                                                                      \
    const size_t smem = (size_t)N_ * 8 * 16 + ((size_t)N_ * 8 + 8 * (N_ / 32)) * 8 + (size_t)N_ * 24;                    \
    const int grid = a.ngroups < ctx->nsm ? a.ngroups : ctx->nsm;                                                       \
    if (dst) {                                                                                                         \
      CK(cudaFuncSetAttribute(k_zfused_s<R1_, R2_, NNPM_EXP_PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
      k_zfused_s<R1_, R2_, NNPM_EXP_PEER><<<grid, 8 * (R1_ > R2_ ? R1_ : R2_), smem, ctx->st>>>(ctx->tm_z8, d, a);                 \
    } else {                                                                                                           \
      CK(cudaFuncSetAttribute(k_zfused_s<R1_, R2_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
      k_zfused_s<R1_, R2_, false><<<grid, 8 * (R1_ > R2_ ? R1_ : R2_), smem, ctx->st>>>(ctx->tm_z8, d, a);                \
    }                                                                                                                  \
    LAUNCHED(ctx);                                                                                                     \
    KCHECK();                                                                                                          \
    return NNPM_OK;                                                                                                    \
  } while (0)
  switch (g.N3) {
    case 64: ZS(8, 8);
    case 128: ZS(16, 8);
    case 256: ZS(16, 16);
    case 512: ZS(32, 16);
    case 1024: ZS(32, 32);
  }
#undef ZS
  return fail(ctx, NNPM_ERR_STATE, "split-exchange fused z pass does not support N3 = %lld", (long long)g.N3);
}
