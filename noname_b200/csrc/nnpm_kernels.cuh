// nnpm_kernels.cuh -- device kernels of the particle-mesh step (sm_100a).
//
// Arithmetic follows src/ParticleMesh.jl of the reference operation by operation.  In the un-fused operator kernels
// (k_deposit weights, k_accel, k_interp_acc, k_axpy_rn, k_kick_comoving, k_push) the __dmul_rn/__dadd_rn/__dsub_rn/
// __ddiv_rn intrinsics pin the reference's evaluation order (they are never contracted into FMAs), so every
// per-particle quantity is bit-identical to the Julia/oracle value for the same inputs.  The tiled hot kernels relax
// this in two documented places, both inside the 1e-12 single-step budget: k_deposit_tile sums unit-mass weights in
// 2^-44 fixed point before the flush, and k_push_tile's gather (cic_grad_gather_fma) forms each corner weight once and
// accumulates with FMAs (a few ulp).  Summation order (atomics) and the FFT differ from the reference at rounding level.
#pragma once
#include <cuda.h>
#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

typedef long long i64;
typedef unsigned long long u64;

struct Geom {
  i64 N1, N2, N3;   // global mesh (TopgridDimensions); N3 == 1 for rank-2 runs
  i64 rowp;         // padded row length in doubles = 2*(N1/2+1)  (in-place r2c layout)
  i64 plane;        // N2*rowp doubles per k-plane
  i64 kz0, nzl;     // first owned plane, number of owned planes (slab along k)
  int glo;          // ghost planes kept below the slab (2 when nranks>1, else 0)
  int nplanes;      // planes allocated = nzl + (nranks>1 ? 5 : 0)
  double fN1, fN2, fN3;
};

// ------------------------------------------------------------------------------------------
// reference-order scalar helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double wrap01(double x) {  // make_periodic, ParticleMesh.jl:409-415
  if (x < 0.) x = __dadd_rn(x, 1.0);
  if (x > 1.) x = __dsub_rn(x, 1.0);
  return x;
}

// ii = floor(x*N)+1 ; d = x*N - ii + 1 ; ip1 = ii==N ? 1 : ii+1   (ParticleMesh.jl:386-388),
// returned 0-based; index N (x == 1.0) and strays wrap periodically.
__device__ __forceinline__ void cell_of(double x, double fN, i64 N, i64& i0, i64& ip, double& d) {
  double t = __dmul_rn(x, fN);
  double fl = floor(t);
  d = __dadd_rn(__dsub_rn(t, fl + 1.0), 1.0);
  i64 z = (i64)fl;
  if (z >= N) z -= N;
  if (z < 0) z += N;
  if (z >= N || z < 0) z = ((z % N) + N) % N;
  i0 = z;
  ip = (z == N - 1) ? 0 : z + 1;
}

__device__ __forceinline__ i64 wrap_idx(i64 z, i64 N) {
  if (z >= N) z -= N;
  if (z < 0) z += N;
  if (z >= N || z < 0) z = ((z % N) + N) % N;
  return z;
}


// 32-bit variants for the hot kernels (N < 2^30; in-plane offsets rowp*N2 < 2^31):
// fl = floor(x*N) kept as a double for the offset, index from F2I.
__device__ __forceinline__ void cell_of32(double x, double fN, int N, int& i0, int& ip, double& d) {
  const double t = __dmul_rn(x, fN);
  const double fl = floor(t);
  d = __dadd_rn(__dsub_rn(t, fl + 1.0), 1.0);
  int z = __double2int_rz(fl);
  if (z >= N) z -= N;   // x == 1.0 (make_periodic leaves it, ParticleMesh.jl:411) -> index N -> 0
  if (z < 0) z += N;
  if ((unsigned)z >= (unsigned)N) z = ((z % N) + N) % N;  // strays (never taken after make_periodic)
  i0 = z;
  ip = (z == N - 1) ? 0 : z + 1;
}
// Branch-free variants for the tiled hot loops.  wrap01_sel == wrap01 bit for bit except that -0.0
// becomes +0.0; cell_fast returns floor(x*N) in [0, N] un-wrapped (index N, i.e. x == 1.0, is folded
// by the caller's tile-relative wrap) and the reference's offset d = (x*N - ii) + 1, ii = floor+1.
__device__ __forceinline__ double wrap01_sel(double x) {
  x = __dadd_rn(x, __hiloint2double(x < 0. ? 0x3FF00000 : 0, 0));
  return __dsub_rn(x, __hiloint2double(x > 1. ? 0x3FF00000 : 0, 0));
}
__device__ __forceinline__ void cell_fast(double x, double fN, int& i0, double& d) {
  const double t = __dmul_rn(x, fN);
  i0 = __double2int_rd(t);
  d = __dadd_rn(__dsub_rn(t, (double)i0 + 1.0), 1.0);
}
__device__ __forceinline__ int tile_rel(int i, int org, int N) {  // i - org folded into [0, N)
  int r = i - org;
  r += (r < 0) ? N : 0;
  r -= (r >= N) ? N : 0;
  return r;
}
__device__ __forceinline__ int local_plane32(int k, const Geom& g) {
  int d = k - (int)g.kz0;
  if (g.glo) {
    if (d >= (int)g.nzl + 3) d -= (int)g.N3;
    if (d < -2) d += (int)g.N3;
    if (d < -2 || d >= (int)g.nzl + 3) return -1;
  }
  return d + g.glo;
}

// local plane of global plane k: p = k - kz0 + glo, unwrapped periodically into the slab's
// ghost window.  Returns -1 when the plane is not held by this rank.
__device__ __forceinline__ int local_plane(i64 k, const Geom& g) {
  i64 d = k - g.kz0;
  if (g.glo) {
    if (d >= g.nzl + 3) d -= g.N3;
    if (d < -2) d += g.N3;
    if (d < -2 || d >= g.nzl + 3) return -1;
  }
  return (int)(d + g.glo);
}

// running maximum without fmax()'s NaN plumbing: fmax compiles to DSETP.MAX + two selects + NaN-quieting LOP3s and
// register moves (~8 SASS instructions; nine of them per particle were 18 % of the tiled push); compare + select is
// three.  A NaN candidate is ignored either way.
__device__ __forceinline__ double dmax_nn(double cur, double cand) { return cand > cur ? cand : cur; }

__device__ __forceinline__ i64 ord_encode(double v) {  // monotone double -> int64
  i64 k = __double_as_longlong(v);
  return k >= 0 ? k : (k ^ 0x7FFFFFFFFFFFFFFFLL);
}

// max of three values over the block, then one global atomicMax each.  The doubles are compared
// through their monotone int64 encoding so the warp step is two 32-bit REDUX ops per value
// (hi word signed, then lo word unsigned among the lanes holding the max hi word).
__device__ __forceinline__ i64 warp_max_key(i64 key) {
  const unsigned full = 0xffffffffu;
  const int hi = (int)(key >> 32);
  const int mhi = __reduce_max_sync(full, hi);
  const unsigned lo = (hi == mhi) ? (unsigned)key : 0u;
  const unsigned mlo = __reduce_max_sync(full, lo);
  return ((i64)mhi << 32) | (i64)mlo;
}
__device__ __forceinline__ void block_max3(double a, double b, double c, i64* out) {
  __shared__ i64 s[3];
  if (threadIdx.x < 3) s[threadIdx.x] = (i64)0x8000000000000000ull;
  __syncthreads();
  const i64 ka = warp_max_key(ord_encode(a)), kb = warp_max_key(ord_encode(b)), kc = warp_max_key(ord_encode(c));
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&s[0], ka);
    atomicMax(&s[1], kb);
    atomicMax(&s[2], kc);
  }
  __syncthreads();
  if (threadIdx.x < 3) atomicMax(&out[threadIdx.x], s[threadIdx.x]);
}

struct PartPtrs {
  double* x[3];
  double* v[3];
};

// ------------------------------------------------------------------------------------------
// host AoS (rank,n)  <->  device SoA
// ------------------------------------------------------------------------------------------
template <int RANK>
__global__ void k_aos2soa(const double* __restrict__ aos, double* d0, double* d1, double* d2, i64 n) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= n) return;
  d0[t] = aos[RANK * t + 0];
  d1[t] = aos[RANK * t + 1];
  if (RANK == 3) d2[t] = aos[RANK * t + 2];
}
template <int RANK>
__global__ void k_soa2aos(double* __restrict__ aos, const double* d0, const double* d1, const double* d2,
                          i64 n) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= n) return;
  aos[RANK * t + 0] = d0[t];
  aos[RANK * t + 1] = d1[t];
  if (RANK == 3) aos[RANK * t + 2] = d2[t];
}
__global__ void k_iota_u32(uint32_t* ids, i64 n, u64 first) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t < n) ids[t] = (uint32_t)(first + (u64)t);
}
__global__ void k_ids_to_dest(uint32_t* dest, const uint32_t* ids, uint32_t first, i64 n) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t < n) dest[t] = ids[t] - first;
}
__global__ void k_u32_to_i64(i64* out, const uint32_t* in, i64 n) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t < n) out[t] = (i64)in[t];
}

// ------------------------------------------------------------------------------------------
// un-fused particle operators
// ------------------------------------------------------------------------------------------
__global__ void k_make_periodic(double* a, i64 n) {  // ParticleMesh.jl:409-415
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t < n) a[t] = wrap01(a[t]);
}
__global__ void k_axpy_rn(double* x, const double* v, double dt, i64 n) {  // drift :723-728, kick :731-736
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t < n) x[t] = __dadd_rn(x[t], __dmul_rn(dt, v[t]));
}
// comoving kick, ParticleMesh.jl:800-807:  vmid = v + (pa*dt)/2 ; v += (-vmid*(adot/a) + pa/a)*dt
__device__ __forceinline__ double kick_comoving_rn(double v, double pa, double dt, double a, double hub) {
  double vmid = __dadd_rn(v, __dmul_rn(__dmul_rn(pa, dt), 0.5));
  double t = __dadd_rn(__dmul_rn(-vmid, hub), __ddiv_rn(pa, a));
  return __dadd_rn(v, __dmul_rn(t, dt));
}
__global__ void k_kick_comoving(double* v, const double* pa, double dt, double a, double hub, i64 n) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t < n) v[t] = kick_comoving_rn(v[t], pa[t], dt, a, hub);
}

// ------------------------------------------------------------------------------------------
// K1 deposit (global-atomic form).  ParticleMesh.jl:289-303, 326-407.
//   PRE: apply the leading half-drift x + c0*v (comoving step, :854) without storing it.
//   Always applies make_periodic (:766 / :856) on the fly.
//   FIXED: accumulate llrint(w * 2^B) into the int64 view of the mesh (deterministic).
// ------------------------------------------------------------------------------------------
template <int RANK, int INTERP, bool FIXED, bool PRE>
__global__ void __launch_bounds__(256)
k_deposit(PartPtrs p, i64 np, const uint32_t* __restrict__ list, const u64* __restrict__ list_count, double c0, Geom g,
          double* __restrict__ mesh, double m, double fscale, u64* __restrict__ viol) {
  // work items: particles [0, np), or (list != nullptr) the indices list[0 .. *list_count) -- the
  // tiled deposit's stragglers -- walked with a grid stride
  const i64 total = list ? (i64)*list_count : np;
  for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
  const i64 n = list ? (i64)list[t] : t;
  double x0 = p.x[0][n], x1 = p.x[1][n], x2 = (RANK == 3) ? p.x[2][n] : 0.0;
  if (PRE) {
    x0 = __dadd_rn(x0, __dmul_rn(c0, p.v[0][n]));
    x1 = __dadd_rn(x1, __dmul_rn(c0, p.v[1][n]));
    if (RANK == 3) x2 = __dadd_rn(x2, __dmul_rn(c0, p.v[2][n]));
  }
  x0 = wrap01(x0);
  x1 = wrap01(x1);
  if (RANK == 3) x2 = wrap01(x2);

  auto add = [&](i64 idx, double w) {
    if (FIXED) {
      i64 q = __double2ll_rn(__dmul_rn(w, fscale));
      atomicAdd((u64*)mesh + idx, (u64)q);
    } else {
      atomicAdd(mesh + idx, w);
    }
  };

  if (INTERP == 0) {  // NGP, :326-353
    i64 i, j, k = 0;
    if (RANK == 2) {
      i = (i64)floor(__dmul_rn(x0, g.fN1));
      j = (i64)floor(__dmul_rn(x1, g.fN2));
    } else {
      i = (i64)floor(__dadd_rn(__dmul_rn(x0, g.fN1), 1.0)) - 1;
      j = (i64)floor(__dadd_rn(__dmul_rn(x1, g.fN2), 1.0)) - 1;
      k = wrap_idx((i64)floor(__dadd_rn(__dmul_rn(x2, g.fN3), 1.0)) - 1, g.N3);
    }
    i = wrap_idx(i, g.N1);
    j = wrap_idx(j, g.N2);
    int pl = (RANK == 3) ? local_plane(k, g) : g.glo;
    if (pl < 0) { atomicAdd(viol, 1ull); continue; }
    add(i + g.rowp * j + g.plane * pl, FIXED ? 1.0 : m);
    continue;
  }

  int i, ip, j, jp;
  double dx, dy;
  cell_of32(x0, g.fN1, (int)g.N1, i, ip, dx);
  cell_of32(x1, g.fN2, (int)g.N2, j, jp, dy);
  const double mm = FIXED ? 1.0 : m;
  const double wx0 = __dmul_rn(mm, __dsub_rn(1.0, dx)), wx1 = __dmul_rn(mm, dx);
  const double ody = __dsub_rn(1.0, dy);
  const unsigned r0 = (unsigned)((int)g.rowp * j), r1 = (unsigned)((int)g.rowp * jp);
  if (RANK == 2) {  // :376-379
    const i64 off = g.plane * g.glo;
    add(off + (r0 + i), __dmul_rn(wx0, ody));
    add(off + (r1 + i), __dmul_rn(wx0, dy));
    add(off + (r0 + ip), __dmul_rn(wx1, ody));
    add(off + (r1 + ip), __dmul_rn(wx1, dy));
    continue;
  }
  int k, kp;
  double dz;
  cell_of32(x2, g.fN3, (int)g.N3, k, kp, dz);
  const int p0 = local_plane32(k, g), p1 = local_plane32(kp, g);
  if ((p0 | p1) < 0) { atomicAdd(viol, 1ull); continue; }
  const double odz = __dsub_rn(1.0, dz);
  const i64 o0 = g.plane * (i64)p0, o1 = g.plane * (i64)p1;
  // :395-402
  add(o0 + (r0 + i), __dmul_rn(__dmul_rn(wx0, ody), odz));
  add(o0 + (r1 + i), __dmul_rn(__dmul_rn(wx0, dy), odz));
  add(o0 + (r0 + ip), __dmul_rn(__dmul_rn(wx1, ody), odz));
  add(o0 + (r1 + ip), __dmul_rn(__dmul_rn(wx1, dy), odz));
  add(o1 + (r0 + i), __dmul_rn(__dmul_rn(wx0, ody), dz));
  add(o1 + (r1 + i), __dmul_rn(__dmul_rn(wx0, dy), dz));
  add(o1 + (r0 + ip), __dmul_rn(__dmul_rn(wx1, ody), dz));
  add(o1 + (r1 + ip), __dmul_rn(__dmul_rn(wx1, dy), dz));
}
}

// int64 fixed-point accumulations -> f64 density, in place: rho = m * (Q * 2^-B)
__global__ void k_fixed_to_f64(double* mesh, i64 n, double inv_scale, double m) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= n) return;
  i64 q = ((const i64*)mesh)[t];
  mesh[t] = __dmul_rn(m, __dmul_rn((double)q, inv_scale));
}

// dst += src over n elements; FIXED adds the int64 views (ghost-plane fold)
template <bool FIXED>
__global__ void k_add_planes(double* dst, const double* src, i64 n) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (FIXED) ((i64*)dst)[t] += ((const i64*)src)[t];
  else dst[t] = __dadd_rn(dst[t], src[t]);
}

// ------------------------------------------------------------------------------------------
// K3 (cuFFT-composed form): Green's function multiply in k space, ParticleMesh.jl:482-502 /
// :451-466.  Spectrum layout [k][jl][ic] (k slowest), jl = local j range starting at j0
// (slab-transposed when nranks > 1).  s1/s2/s3 = sin^2(k_d/2) tables built on the host with
// the reference's expressions.  scale = 1/(N1*N2*N3) (unnormalised cuFFT inverse) * 1/N1^2.
// ------------------------------------------------------------------------------------------
__global__ void k_green(double2* __restrict__ rt, i64 icn, i64 njl, i64 j0, i64 nk, const double* __restrict__ s1,
                        const double* __restrict__ s2, const double* __restrict__ s3, int rank3, double scale) {
  i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  i64 jl = blockIdx.y;
  if (i >= icn) return;
  const double sxy = __dadd_rn(s1[i], s2[j0 + jl]);
  for (i64 k = blockIdx.z; k < nk; k += gridDim.z) {
    double ginv = rank3 ? __dmul_rn(-4.0, __dadd_rn(sxy, s3[k])) : __dmul_rn(-4.0, sxy);
    i64 idx = i + icn * (jl + njl * k);
    double2 val = rt[idx];
    if (ginv != 0.0) {
      val.x = __dmul_rn(__ddiv_rn(val.x, ginv), scale);
      val.y = __dmul_rn(__ddiv_rn(val.y, ginv), scale);
    } else {
      val.x = __dmul_rn(val.x, scale);
      val.y = __dmul_rn(val.y, scale);
    }
    if (i == 0 && j0 + jl == 0 && k == 0) { val.x = 0.0; val.y = 0.0; }  // rt[1,1,1] = 0, :501
    rt[idx] = val;
  }
}

// slab transpose helpers (nranks > 1).  A: [kl][j][ic] (local planes, all j);
// S: [d][kl][jl][ic] (block d goes to rank d).
__global__ void k_pack_fwd(const double2* __restrict__ A, double2* __restrict__ S, i64 icn, i64 N2, i64 njl,
                           i64 nzl, i64 aplane_c) {
  i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  i64 j = blockIdx.y, kl = blockIdx.z;
  if (i >= icn) return;
  i64 d = j / njl, jl = j - d * njl;
  S[i + icn * (jl + njl * (kl + nzl * d))] = A[i + icn * j + aplane_c * kl];
}
__global__ void k_unpack_bwd(double2* __restrict__ A, const double2* __restrict__ S, i64 icn, i64 N2, i64 njl,
                             i64 nzl, i64 aplane_c) {
  i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  i64 j = blockIdx.y, kl = blockIdx.z;
  if (i >= icn) return;
  i64 d = j / njl, jl = j - d * njl;
  A[i + icn * j + aplane_c * kl] = S[i + icn * (jl + njl * (kl + nzl * d))];
}

// ------------------------------------------------------------------------------------------
// compute_acceleration materialised (parity path only), ParticleMesh.jl:631-637, 549-573:
// acc[c,i,j,q] = -(N1 * ((phi[+1 along c] - phi[-1 along c]) / 2)), component-fastest.
// q runs over nq planes starting at global plane kq0 (owned planes, +-1 when nranks > 1).
// ------------------------------------------------------------------------------------------
template <int RANK>
__global__ void k_accel(double* __restrict__ acc, const double* __restrict__ phi, Geom g, i64 kq0, i64 nq) {
  i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  i64 j = blockIdx.y, q = blockIdx.z;
  if (i >= g.N1) return;
  i64 k = wrap_idx(kq0 + q, g.N3);
  i64 ip1 = (i == g.N1 - 1) ? 0 : i + 1, im1 = (i == 0) ? g.N1 - 1 : i - 1;
  i64 jp1 = (j == g.N2 - 1) ? 0 : j + 1, jm1 = (j == 0) ? g.N2 - 1 : j - 1;
  int pc = (RANK == 3) ? local_plane(k, g) : g.glo;
  const double* P = phi + g.plane * pc;
  double* o = acc + 3 * (i + g.N1 * (j + g.N2 * q));
  o[0] = -__dmul_rn(g.fN1, __dmul_rn(__dsub_rn(P[ip1 + g.rowp * j], P[im1 + g.rowp * j]), 0.5));
  o[1] = -__dmul_rn(g.fN1, __dmul_rn(__dsub_rn(P[i + g.rowp * jp1], P[i + g.rowp * jm1]), 0.5));
  if (RANK == 3) {
    i64 kp1 = (k == g.N3 - 1) ? 0 : k + 1, km1 = (k == 0) ? g.N3 - 1 : k - 1;
    const double* Pp = phi + g.plane * local_plane(kp1, g);
    const double* Pm = phi + g.plane * local_plane(km1, g);
    o[2] = -__dmul_rn(g.fN1, __dmul_rn(__dsub_rn(Pp[i + g.rowp * j], Pm[i + g.rowp * j]), 0.5));
  } else {
    o[2] = 0.0;  // third component stays as allocated (zeros), ParticleMesh.jl:754
  }
}

// interpVecToPoints from a materialised acc (parity path only), ParticleMesh.jl:640-721.
template <int RANK, int INTERP>
__global__ void k_interp_acc(double* pa0, double* pa1, double* pa2, const double* __restrict__ acc, PartPtrs p,
                             i64 np, Geom g, i64 kq0, i64 nq, int bugz) {
  i64 n = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (n >= np) return;
  double x0 = p.x[0][n], x1 = p.x[1][n], x2 = (RANK == 3) ? p.x[2][n] : 0.0;
  auto A = [&](int m, i64 i, i64 j, i64 k) -> double {
    i64 q = 0;
    if (RANK == 3) {
      q = k - kq0;
      if (q < 0) q += g.N3;
      if (q >= g.N3) q -= g.N3;
      if (q >= nq) q = nq - 1;  // not held: clamp (halo violation is reported by the fused path)
    }
    return acc[m + 3 * (i + g.N1 * (j + g.N2 * q))];
  };
  if (INTERP == 0) {  // :642-652
    i64 i = wrap_idx((i64)floor(__dmul_rn(x0, g.fN1)), g.N1);
    i64 j = wrap_idx((i64)floor(__dmul_rn(x1, g.fN2)), g.N2);
    i64 k = 0;
    if (RANK == 3) k = bugz ? wrap_idx((i64)floor(x2) * g.N3, g.N3) : wrap_idx((i64)floor(__dmul_rn(x2, g.fN3)), g.N3);
    pa0[n] = A(0, i, j, k);
    pa1[n] = A(1, i, j, k);
    if (RANK == 3) pa2[n] = A(2, i, j, k);
    return;
  }
  i64 i, ip, j, jp;
  double dx, dy;
  cell_of(x0, g.fN1, g.N1, i, ip, dx);
  cell_of(x1, g.fN2, g.N2, j, jp, dy);
  const double odx = __dsub_rn(1.0, dx), ody = __dsub_rn(1.0, dy);
  if (RANK == 2) {  // :686-689
    double* out[2] = {pa0, pa1};
#pragma unroll
    for (int m = 0; m < 2; m++) {
      double s = __dmul_rn(__dmul_rn(A(m, i, j, 0), odx), ody);
      s = __dadd_rn(s, __dmul_rn(__dmul_rn(A(m, ip, j, 0), dx), ody));
      s = __dadd_rn(s, __dmul_rn(__dmul_rn(A(m, i, jp, 0), odx), dy));
      s = __dadd_rn(s, __dmul_rn(__dmul_rn(A(m, ip, jp, 0), dx), dy));
      out[m][n] = s;
    }
    return;
  }
  i64 k, kp;
  double dz;
  if (bugz) {  // literal :706-708:  k = floor(x3)*Ng3 + 1 ; kp1 = k==Ng3 ? 1 : k+1 ; dz = x3*Ng3 - k + 1
    i64 k1 = (i64)floor(x2) * g.N3 + 1;
    dz = __dadd_rn(__dsub_rn(__dmul_rn(x2, g.fN3), (double)k1), 1.0);
    k = wrap_idx(k1 - 1, g.N3);
    kp = (k1 == g.N3) ? 0 : wrap_idx(k1, g.N3);
  } else {
    cell_of(x2, g.fN3, g.N3, k, kp, dz);
  }
  const double odz = __dsub_rn(1.0, dz);
  double* out[3] = {pa0, pa1, pa2};
#pragma unroll
  for (int m = 0; m < 3; m++) {  // :710-717
    double s = __dmul_rn(__dmul_rn(__dmul_rn(A(m, i, j, k), odx), ody), odz);
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(A(m, ip, j, k), dx), ody), odz));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(A(m, i, jp, k), odx), dy), odz));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(A(m, ip, jp, k), dx), dy), odz));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(A(m, i, j, kp), odx), ody), dz));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(A(m, ip, j, kp), dx), ody), dz));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(A(m, i, jp, kp), odx), dy), dz));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dmul_rn(A(m, ip, jp, kp), dx), dy), dz));
    out[m][n] = s;
  }
}

// ------------------------------------------------------------------------------------------
// compute_acceleration (:631-637) + cic_interp (:676-721) fused: the finite-difference gradient is
// evaluated at the 4 / 8 CIC corners from phi values delivered by ld(a, b, c), a/b/c in 0..3
// indexing cells (i-1, i, ip, ip+1) x (j-1, j, jp, jp+1) x (k-1, k, kp, kp+1).
//   -(N1*((hi-lo)/2)) == (hi-lo) * (-N1/2) bit for bit (the halving and negation are exact), so
//   gfac = -N1/2; weights multiply left to right, corners add in the reference's order.
// ------------------------------------------------------------------------------------------
template <int RANK, class LD>
__device__ __forceinline__ void cic_grad_gather(const LD& ld, double dx, double dy, double dz, double gfac, double* pa) {
  const double odx = __dsub_rn(1.0, dx), ody = __dsub_rn(1.0, dy);
  const double wxs[2] = {odx, dx}, wys[2] = {ody, dy};
  if (RANK == 2) {
    double c[2][4], ylo[2], yhi[2];
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int a = 0; a < 4; a++) c[b][a] = ld(a, 1 + b, 0);
#pragma unroll
    for (int a = 0; a < 2; a++) { ylo[a] = ld(1 + a, 0, 0); yhi[a] = ld(1 + a, 3, 0); }
    // order :687-688  (i,j) (ip,j) (i,jp) (ip,jp)
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int a = t & 1, b = t >> 1;
      const double gx = __dmul_rn(__dsub_rn(c[b][a + 2], c[b][a]), gfac);
      const double hi = b ? yhi[a] : c[1][a + 1], lo = b ? c[0][a + 1] : ylo[a];
      const double gy = __dmul_rn(__dsub_rn(hi, lo), gfac);
      const double tx = __dmul_rn(__dmul_rn(gx, wxs[a]), wys[b]);
      const double ty = __dmul_rn(__dmul_rn(gy, wxs[a]), wys[b]);
      s0 = t ? __dadd_rn(s0, tx) : tx;
      s1 = t ? __dadd_rn(s1, ty) : ty;
    }
    pa[0] = s0; pa[1] = s1;
  } else {
    const double odz = __dsub_rn(1.0, dz);
    const double wzs[2] = {odz, dz};
    double c[2][2][4], ylo[2][2], yhi[2][2], zlo[2][2], zhi[2][2];
#pragma unroll
    for (int cz = 0; cz < 2; cz++)
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int a = 0; a < 4; a++) c[cz][b][a] = ld(a, 1 + b, 1 + cz);
#pragma unroll
    for (int cz = 0; cz < 2; cz++)
#pragma unroll
      for (int a = 0; a < 2; a++) { ylo[cz][a] = ld(1 + a, 0, 1 + cz); yhi[cz][a] = ld(1 + a, 3, 1 + cz); }
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int a = 0; a < 2; a++) { zlo[b][a] = ld(1 + a, 1 + b, 0); zhi[b][a] = ld(1 + a, 1 + b, 3); }
    // order :710-717: (i,j,k)(ip,j,k)(i,jp,k)(ip,jp,k) then the kp four
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const int a = t & 1, b = (t >> 1) & 1, cz = t >> 2;
      const double gx = __dmul_rn(__dsub_rn(c[cz][b][a + 2], c[cz][b][a]), gfac);
      const double yh = b ? yhi[cz][a] : c[cz][1][a + 1], yl = b ? c[cz][0][a + 1] : ylo[cz][a];
      const double gy = __dmul_rn(__dsub_rn(yh, yl), gfac);
      const double zh = cz ? zhi[b][a] : c[1][b][a + 1], zl = cz ? c[0][b][a + 1] : zlo[b][a];
      const double gz = __dmul_rn(__dsub_rn(zh, zl), gfac);
      const double tx = __dmul_rn(__dmul_rn(__dmul_rn(gx, wxs[a]), wys[b]), wzs[cz]);
      const double ty = __dmul_rn(__dmul_rn(__dmul_rn(gy, wxs[a]), wys[b]), wzs[cz]);
      const double tz = __dmul_rn(__dmul_rn(__dmul_rn(gz, wxs[a]), wys[b]), wzs[cz]);
      s0 = t ? __dadd_rn(s0, tx) : tx;
      s1 = t ? __dadd_rn(s1, ty) : ty;
      s2 = t ? __dadd_rn(s2, tz) : tz;
    }
    pa[0] = s0; pa[1] = s1; pa[2] = s2;
  }
}

// ------------------------------------------------------------------------------------------
// K5 fused push: [leading half-drift recomputed] -> make_periodic -> gradient of phi taken on
// the fly (compute_acceleration, :631-637) -> back-interpolation (:640-721) -> kick (:731-736 or
// :800-807) -> drift(s) (:723-728) -> make_periodic -> max reductions (:782, :875-878).
// phi is read straight from the padded in-place FFT buffer; acc and pa are never materialised.
//   COMOVING: x' = wrap(x + c0*v); v' = kick_comoving; x'' = wrap(x' + c2*v')
//   static  : x' = wrap(x);        x1 = x' + c1*v; v' = v + dt*pa; x'' = x1 + c1*v'
// ------------------------------------------------------------------------------------------
// Particle migration bookkeeping shared with nnpm_comm.cuh.  Counter record of one rank: [0, 64) particles packed per
// destination, then the slots below; MC_REC entries of it are all-gathered (one row per rank) behind the local record.
enum { MC_HOLES = 64, MC_OVER = 65, MC_NLO = 66, MC_NFILL = 67, MC_NP = 68, MC_CAP = 69, MC_REC = 72, MC_ALL = 80 };
#define MIG_HOLE 0xFFFFFFFFu
#define MIG_REC 7  // doubles per migrating particle record: x[3], v[3], id bits

// nranks > 1: the push itself classifies every particle it has just moved -- it knows the new z -- and packs the
// leavers into the per-destination send segments (the separate classification pass over all particles that round 1
// ran after the push cost 2.3 ms per step at 2 GPUs, 0.6 ms at 8).
struct MigArgs {
  int on, P, r, kz0, nzl, N3;
  double fN3;
  i64 seg_cap;
  double* send;
  uint32_t* holes;
  unsigned long long* cnt;
  uint32_t* ids;
};

struct PushParams {
  double c0, c1, c2, dt, a, hub;  // hub = adot/a evaluated on the host as the reference does
  int bugz;
  MigArgs mg;
};

// leaver test + packing for particle n with final position xn / velocity vn (all RANK == 3)
__device__ __forceinline__ void mig_emit(const MigArgs& mg, i64 n, const double* xn, const double* vn) {
  int k = __double2int_rd(__dmul_rn(wrap01(xn[2]), mg.fN3));
  if (k >= mg.N3) k -= mg.N3;                 // x == 1.0 (make_periodic leaves it) -> plane 0
  if (k < 0) k += mg.N3;
  if ((unsigned)(k - mg.kz0) < (unsigned)mg.nzl) return;   // stays
  const int owner = k / mg.nzl;
  const unsigned long long slot = atomicAdd(&mg.cnt[owner], 1ull);
  if ((i64)slot >= mg.seg_cap) { atomicAdd(&mg.cnt[MC_OVER], 1ull); return; }  // stays for now; a classification round follows
  double* rec = mg.send + ((i64)owner * mg.seg_cap + (i64)slot) * MIG_REC;
  rec[0] = xn[0]; rec[1] = xn[1]; rec[2] = xn[2];
  rec[3] = vn[0]; rec[4] = vn[1]; rec[5] = vn[2];
  rec[6] = __longlong_as_double((i64)mg.ids[n]);
  mg.ids[n] = MIG_HOLE;
  mg.holes[atomicAdd(&mg.cnt[MC_HOLES], 1ull)] = (uint32_t)n;
}

#define PHI(ii, jj, pp) __ldg(phi + (ii) + g.rowp * (jj) + g.plane * (i64)(pp))

// kick + drift(s) + wrap of one particle, shared by the global and the tiled push
template <int RANK, bool COMOVING>
__device__ __forceinline__ void push_finish(const PartPtrs& p, i64 n, const double* x, const double* v, const double* pa,
                                            const PushParams& pp, double& mxa, double& mxv, double& mxp) {
  double xo[3] = {0.0, 0.0, 0.0}, vo[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < RANK; d++) {
    double vn, xn;
    if (COMOVING) {
      vn = kick_comoving_rn(v[d], pa[d], pp.dt, pp.a, pp.hub);
      xn = wrap01(__dadd_rn(x[d], __dmul_rn(pp.c2, vn)));
    } else {
      double x1 = __dadd_rn(x[d], __dmul_rn(pp.c1, v[d]));
      vn = __dadd_rn(v[d], __dmul_rn(pp.dt, pa[d]));
      xn = __dadd_rn(x1, __dmul_rn(pp.c1, vn));
    }
    p.x[d][n] = xn;
    p.v[d][n] = vn;
    xo[d] = xn; vo[d] = vn;
    mxa = dmax_nn(mxa, fabs(vn));
    mxv = dmax_nn(mxv, vn);
    mxp = dmax_nn(mxp, pa[d]);
  }
  if (RANK == 3 && pp.mg.on) mig_emit(pp.mg, n, xo, vo);
}

// Global-gather push.  Work items: particles [base, base+count) or, when list != nullptr, the
// particle indices list[0 .. *list_count) (the tiled push's stragglers).
template <int RANK, int INTERP, bool COMOVING>
__global__ void __launch_bounds__(256, 2)
k_push(PartPtrs p, i64 base, i64 count, const uint32_t* __restrict__ list, const u64* __restrict__ list_count, Geom g,
       const double* __restrict__ phi, PushParams pp, i64* __restrict__ red, u64* __restrict__ viol) {
  double mxa = -INFINITY, mxv = -INFINITY, mxp = -INFINITY;
  const i64 total = list ? (i64)*list_count : count;
  for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const i64 n = list ? (i64)list[t] : base + t;
    double x[3], v[3], pa[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < RANK; d++) {
      x[d] = p.x[d][n];
      v[d] = p.v[d][n];
      if (COMOVING) x[d] = __dadd_rn(x[d], __dmul_rn(pp.c0, v[d]));
      x[d] = wrap01(x[d]);
    }
    bool ok = true;
    if (INTERP == 0) {
      i64 i = wrap_idx((i64)floor(__dmul_rn(x[0], g.fN1)), g.N1);
      i64 j = wrap_idx((i64)floor(__dmul_rn(x[1], g.fN2)), g.N2);
      i64 ip1 = (i == g.N1 - 1) ? 0 : i + 1, im1 = (i == 0) ? g.N1 - 1 : i - 1;
      i64 jp1 = (j == g.N2 - 1) ? 0 : j + 1, jm1 = (j == 0) ? g.N2 - 1 : j - 1;
      int pc = g.glo, pu = 0, pd = 0;
      if (RANK == 3) {
        i64 k = pp.bugz ? wrap_idx((i64)floor(x[2]) * g.N3, g.N3) : wrap_idx((i64)floor(__dmul_rn(x[2], g.fN3)), g.N3);
        i64 kp1 = (k == g.N3 - 1) ? 0 : k + 1, km1 = (k == 0) ? g.N3 - 1 : k - 1;
        pc = local_plane(k, g); pu = local_plane(kp1, g); pd = local_plane(km1, g);
        ok = pc >= 0 && pu >= 0 && pd >= 0;
      }
      if (ok) {
        pa[0] = -__dmul_rn(g.fN1, __dmul_rn(__dsub_rn(PHI(ip1, j, pc), PHI(im1, j, pc)), 0.5));
        pa[1] = -__dmul_rn(g.fN1, __dmul_rn(__dsub_rn(PHI(i, jp1, pc), PHI(i, jm1, pc)), 0.5));
        if (RANK == 3) pa[2] = -__dmul_rn(g.fN1, __dmul_rn(__dsub_rn(PHI(i, j, pu), PHI(i, j, pd)), 0.5));
      }
    } else {
      const int N1 = (int)g.N1, N2 = (int)g.N2, rowp = (int)g.rowp;
      int ia[4], ro[4];  // i-1, i, ip, ip+1 ; row offsets rowp*(j-1, j, jp, jp+1)
      double dx, dy, dz = 0.0;
      {
        int j1, j2;
        cell_of32(x[0], g.fN1, N1, ia[1], ia[2], dx);
        cell_of32(x[1], g.fN2, N2, j1, j2, dy);
        ia[0] = (ia[1] == 0) ? N1 - 1 : ia[1] - 1;
        ia[3] = (ia[2] == N1 - 1) ? 0 : ia[2] + 1;
        ro[0] = rowp * ((j1 == 0) ? N2 - 1 : j1 - 1);
        ro[1] = rowp * j1;
        ro[2] = rowp * j2;
        ro[3] = rowp * ((j2 == N2 - 1) ? 0 : j2 + 1);
      }
      const double gfac = -0.5 * g.fN1;
      if (RANK == 2) {
        const double* P = phi + g.plane * (i64)g.glo;
        cic_grad_gather<2>([&](int a, int b, int) { return __ldg(P + (unsigned)(ro[b] + ia[a])); }, dx, dy, 0.0, gfac, pa);
      } else {
        const int N3 = (int)g.N3;
        int k, kp;
        if (pp.bugz) {  // literal :706-708
          i64 k1 = (i64)floor(x[2]) * g.N3 + 1;
          dz = __dadd_rn(__dsub_rn(__dmul_rn(x[2], g.fN3), (double)k1), 1.0);
          k = (int)wrap_idx(k1 - 1, g.N3);
          kp = (k1 == g.N3) ? 0 : (int)wrap_idx(k1, g.N3);
        } else {
          cell_of32(x[2], g.fN3, N3, k, kp, dz);
        }
        const int km1 = (k == 0) ? N3 - 1 : k - 1, kp2 = (kp == N3 - 1) ? 0 : kp + 1;
        const int pl[4] = {local_plane32(km1, g), local_plane32(k, g), local_plane32(kp, g), local_plane32(kp2, g)};
        ok = (pl[0] | pl[1] | pl[2] | pl[3]) >= 0;
        if (ok) {
          const double* P[4] = {phi + g.plane * (i64)pl[0], phi + g.plane * (i64)pl[1], phi + g.plane * (i64)pl[2],
                                phi + g.plane * (i64)pl[3]};
          cic_grad_gather<3>([&](int a, int b, int c) { return __ldg(P[c] + (unsigned)(ro[b] + ia[a])); }, dx, dy, dz, gfac, pa);
        }
      }
    }
    if (!ok) atomicAdd(viol, 1ull);
    push_finish<RANK, COMOVING>(p, n, x, v, pa, pp, mxa, mxv, mxp);
  }
  block_max3(mxa, mxv, mxp, red);
}

// ------------------------------------------------------------------------------------------
// Tiled push (CIC): one CTA per tile of TX x TY x TZ cells of tile-sorted particles.  The phi box
// the tile's particles can touch -- cells [-H-1, T+H+2) around the tile, H = cells a particle may
// have strayed since the sort -- is staged in shared memory once; every particle then takes its 32
// (16 in 2-D) stencil values from shared memory.  Particles outside the box go to `slow_list` and
// are pushed afterwards by k_push's global-gather path.  Arithmetic is cic_grad_gather in both.
// ------------------------------------------------------------------------------------------
#define NNPM_TX 32
#define NNPM_TY 8
#define NNPM_TZ 8
#define NNPM_TH 1
struct TileGeom {
  int ntx, nty, ntz;   // tiles per axis (z: over the slab's owned planes)
  int tx, ty, tz;      // tile extent in cells (clamped to the mesh)
};

// Work list of the tiled kernels, rebuilt by every sort: one entry per chunk of at most NNPM_CHUNK
// particles of one tile, so that a tile holding a collapsed halo (10^3..10^5 x the mean occupancy in
// the clustered late-time state) is spread over many CTAs instead of serialising one, and empty tiles
// cost nothing.  Entries are ordered along a banded raster: tiles x fastest, then y within bands of
// NNPM_YB tile rows, then z, then the next band -- the halo planes a tile shares with its z neighbour
// are touched again after NNPM_YB*ntx tiles (a few MB, still in L2) instead of after a whole plane
// of tiles (~100 MB at 1024^2 cells per plane).
#define NNPM_YB 8
#define NNPM_CHUNK 4096
struct WorkList {
  const uint32_t* tile;  // tile index of chunk w
  const uint32_t* beg;   // first particle of chunk w
  const uint32_t* count; // number of chunks (device scalar)
};
__device__ __forceinline__ int tile_of_raster(const TileGeom& tg, int rpos) {
  const int L = rpos / tg.ntx;
  const int ttx = rpos - L * tg.ntx;
  const int band = L / (NNPM_YB * tg.ntz);
  const int rows = min(NNPM_YB, tg.nty - band * NNPM_YB);
  const int r = L - band * NNPM_YB * tg.ntz;
  const int ttz = r / rows;
  const int tty = band * NNPM_YB + (r - ttz * rows);
  return ttx + tg.ntx * (tty + tg.nty * ttz);
}
// chunks per raster position (exclusive-scanned afterwards); nw[ntiles] = 0 receives the total
__global__ void k_work_count(TileGeom tg, i64 ntiles, const uint32_t* __restrict__ tile_off, uint32_t* __restrict__ nw) {
  const i64 r = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (r > ntiles) return;
  if (r == ntiles) { nw[r] = 0u; return; }
  const int t = tile_of_raster(tg, (int)r);
  const uint32_t cnt = tile_off[t + 1] - tile_off[t];
  nw[r] = (cnt + NNPM_CHUNK - 1) / NNPM_CHUNK;
}
__global__ void k_work_fill(TileGeom tg, i64 ntiles, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ woff,
                            uint32_t* __restrict__ wtile, uint32_t* __restrict__ wbeg) {
  const i64 r = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (r >= ntiles) return;
  const int t = tile_of_raster(tg, (int)r);
  const uint32_t n0 = tile_off[t], n1 = tile_off[t + 1];
  uint32_t w = woff[r];
  for (uint32_t b = n0; b < n1; b += NNPM_CHUNK, w++) { wtile[w] = (uint32_t)t; wbeg[w] = b; }
}

// Fast-path gather: same operator as cic_grad_gather with the products re-associated so that the
// FP64 pipe (16 lanes per SM sub-partition -- a co-limit of this kernel next to HBM) sees 72
// instead of 141 operations: corner weights w = wx*wy*wz*(-N1/2) are formed once and shared by
// the three components, and the corner sums use fused multiply-adds.  Differences from the
// reference's evaluation order are a few ulp of the largest term (parity budget: 1e-12).
template <int RANK, class LD>
__device__ __forceinline__ void cic_grad_gather_fma(const LD& ld, double dx, double dy, double dz, double gfac, double* pa) {
  const double odx = 1.0 - dx, ody = 1.0 - dy;
  const double wx[2] = {odx * gfac, dx * gfac}, wy[2] = {ody, dy};
  if (RANK == 2) {
    double c[2][4], ylo[2], yhi[2];
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int a = 0; a < 4; a++) c[b][a] = ld(a, 1 + b, 0);
#pragma unroll
    for (int a = 0; a < 2; a++) { ylo[a] = ld(1 + a, 0, 0); yhi[a] = ld(1 + a, 3, 0); }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int a = t & 1, b = t >> 1;
      const double w = wx[a] * wy[b];
      const double hi = b ? yhi[a] : c[1][a + 1], lo = b ? c[0][a + 1] : ylo[a];
      s0 = fma(c[b][a + 2] - c[b][a], w, s0);
      s1 = fma(hi - lo, w, s1);
    }
    pa[0] = s0; pa[1] = s1;
  } else {
    const double wz[2] = {1.0 - dz, dz};
    double c[2][2][4], ylo[2][2], yhi[2][2], zlo[2][2], zhi[2][2];
#pragma unroll
    for (int cz = 0; cz < 2; cz++)
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int a = 0; a < 4; a++) c[cz][b][a] = ld(a, 1 + b, 1 + cz);
#pragma unroll
    for (int cz = 0; cz < 2; cz++)
#pragma unroll
      for (int a = 0; a < 2; a++) { ylo[cz][a] = ld(1 + a, 0, 1 + cz); yhi[cz][a] = ld(1 + a, 3, 1 + cz); }
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int a = 0; a < 2; a++) { zlo[b][a] = ld(1 + a, 1 + b, 0); zhi[b][a] = ld(1 + a, 1 + b, 3); }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const int a = t & 1, b = (t >> 1) & 1, cz = t >> 2;
      const double w = wx[a] * wy[b] * wz[cz];
      const double yh = b ? yhi[cz][a] : c[cz][1][a + 1], yl = b ? c[cz][0][a + 1] : ylo[cz][a];
      const double zh = cz ? zhi[b][a] : c[1][b][a + 1], zl = cz ? c[0][b][a + 1] : zlo[b][a];
      s0 = fma(c[cz][b][a + 2] - c[cz][b][a], w, s0);
      s1 = fma(yh - yl, w, s1);
      s2 = fma(zh - zl, w, s2);
    }
    pa[0] = s0; pa[1] = s1; pa[2] = s2;
  }
}

// correctly rounded p/a from the correctly rounded reciprocal ra = RN(1/a) (Markstein): three
// operations instead of the ~12 of a full division; equals __ddiv_rn(p, a) for every a whose
// significand is not all ones.
__device__ __forceinline__ double div_by_scalar(double p, double a, double ra) {
  const double q = p * ra;
  const double r = fma(-q, a, p);
  return fma(r, ra, q);
}

template <int RANK, bool COMOVING>
__device__ __forceinline__ void push_finish_fast(const PartPtrs& p, i64 n, const double* x, const double* v, const double* pa,
                                                 const PushParams& pp, double ra, double& mxa, double& mxv, double& mxp) {
  double xo[3] = {0.0, 0.0, 0.0}, vo[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < RANK; d++) {
    double vn, xn;
    if (COMOVING) {  // ParticleMesh.jl:800-807 in the reference's operation order; only pa/a is restructured
      const double vmid = __dadd_rn(v[d], __dmul_rn(__dmul_rn(pa[d], pp.dt), 0.5));
      const double t = __dadd_rn(__dmul_rn(-vmid, pp.hub), div_by_scalar(pa[d], pp.a, ra));
      vn = __dadd_rn(v[d], __dmul_rn(t, pp.dt));
      xn = wrap01(__dadd_rn(x[d], __dmul_rn(pp.c2, vn)));
    } else {
      double x1 = __dadd_rn(x[d], __dmul_rn(pp.c1, v[d]));
      vn = __dadd_rn(v[d], __dmul_rn(pp.dt, pa[d]));
      xn = __dadd_rn(x1, __dmul_rn(pp.c1, vn));
    }
    p.x[d][n] = xn;
    p.v[d][n] = vn;
    xo[d] = xn; vo[d] = vn;
    mxa = dmax_nn(mxa, fabs(vn));
    mxv = dmax_nn(mxv, vn);
    mxp = dmax_nn(mxp, pa[d]);
  }
  if (RANK == 3 && pp.mg.on) mig_emit(pp.mg, n, xo, vo);
}

#define NNPM_BXU (NNPM_TX + 2 * NNPM_TH + 3)  // 37 box columns a tile's particles can touch
// Row pitch of the box in shared memory, a multiple of 16 doubles (one 128-byte bank period): cell-sorted particles walk
// a mesh row, then the next one; with pitch = 0 (mod 16) the bank of a lane's cell keeps advancing cyclically across
// the row (and plane: BX * BY = 0 mod 16 too) break, so 32 lanes on ~32 consecutive cells stay at two lanes per 8-byte
// bank -- the minimum for LDS.64 -- instead of piling the two row segments onto the same banks (pitch 38: 3.2 wavefronts
// per gather, ncu r2f).  The 11 padding columns are filled by the same tensor copy and never read.
#ifndef NNPM_BX
#define NNPM_BX 48
#endif
#define NNPM_BY (NNPM_TY + 2 * NNPM_TH + 3)
#define NNPM_BZ (NNPM_TZ + 2 * NNPM_TH + 3)

__device__ __forceinline__ int wrap_small(int z, int N) {  // z within a few periods of [0, N)
  if (z < 0) { z += N; if (z < 0) z = ((z % N) + N) % N; }
  else if (z >= N) { z -= N; if (z >= N) z %= N; }
  return z;
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier primitives, sm_90+/sm_100a PTX ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NNPM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NNPM_DONE;\n"
      "bra NNPM_WAIT;\n"
      "NNPM_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}


// Pull [n0, n1) of every particle array the kernel reads into L2 with one bulk prefetch per array
// (cp.async.bulk.prefetch.L2: the copy engine fetches the whole 16-32 KB run as one sequential DRAM burst,
// no registers, no LSU traffic); the kernel's own 8-byte loads then hit L2.  Issued by threads 0..NARR-1.
template <int RANK, bool WITH_V>
__device__ __forceinline__ void l2_prefetch_particles(const PartPtrs& p, i64 n0, i64 n1) {
  constexpr int NARR = WITH_V ? 2 * RANK : RANK;
  const int a = threadIdx.x;
  if (a >= NARR || n1 <= n0) return;
  const double* base = p.x[0];  // select chain: a dynamic index would put PartPtrs on the stack
#pragma unroll
  for (int d = 1; d < RANK; d++) base = (a == d) ? p.x[d] : base;
  if (WITH_V) {
#pragma unroll
    for (int d = 0; d < RANK; d++) base = (a == RANK + d) ? p.v[d] : base;
  }
  const u64 lo = (u64)(base + n0) & ~15ull, hi = (u64)(base + n1) & ~15ull;  // 16-byte units inside the run
  if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned)(hi - lo)) : "memory");
}

template <int RANK, bool COMOVING>
__global__ void __launch_bounds__(256, 2)
k_push_tile(const __grid_constant__ CUtensorMap tmap, PartPtrs p, i64 np, Geom g, TileGeom tg, WorkList wl,
            const uint32_t* __restrict__ tile_off, const double* __restrict__ phi, PushParams pp, i64* __restrict__ red,
            uint32_t* __restrict__ slow_list, u64* __restrict__ slow_count, int pf) {
  extern __shared__ __align__(128) double box[];  // [BZ][BY][BX]
  __shared__ __align__(8) unsigned long long mbar;
  constexpr int BX = NNPM_BX, BY = NNPM_BY, BZ = (RANK == 3) ? NNPM_BZ : 1, BXY = BX * BY;
  if (blockIdx.x >= *wl.count) return;  // the grid is sized for the worst case
  const int t = (int)wl.tile[blockIdx.x];
  const i64 n0 = wl.beg[blockIdx.x];
  i64 n1 = n0 + NNPM_CHUNK;
  if (n1 > (i64)tile_off[t + 1]) n1 = tile_off[t + 1];
  if (n1 > np) n1 = np;
  if (n0 >= n1) return;  // uniform across the block
  if (pf) l2_prefetch_particles<RANK, true>(p, n0, n1);  // this item's x, v
  const int ttx = t % tg.ntx, tty = (t / tg.ntx) % tg.nty, ttz = t / (tg.ntx * tg.nty);
  const int N1 = (int)g.N1, N2 = (int)g.N2, N3 = (int)g.N3;
  const int ti0 = ttx * tg.tx, tj0 = tty * tg.ty, tk0 = ttz * tg.tz;  // tk0: plane within the slab
  // Stage the phi box.  Interior tiles (box inside the mesh, no periodic wrap): ONE TMA tensor copy
  // issued by one thread, completion on an mbarrier.  Tiles whose box wraps: cp.async row by row.
  const int bx0 = ti0 - NNPM_TH - 1, by0 = tj0 - NNPM_TH - 1;
  const int bz0 = (RANK == 3) ? tk0 - NNPM_TH - 1 + g.glo : g.glo;  // local plane of the box origin
  const bool interior = bx0 >= 0 && bx0 + NNPM_BXU <= N1 && by0 >= 0 && by0 + BY <= N2 &&   // columns past N1 (padding only) are zero-filled by TMA
                        (RANK == 2 || g.glo != 0 || (bz0 >= 0 && bz0 + BZ <= N3));
  if (interior) {
    if (threadIdx.x == 0) {
      mbar_init(&mbar, 1);
      mbar_expect_tx(&mbar, (unsigned)(BX * BY * BZ * sizeof(double)));
      tma_load_3d(box, &tmap, bx0, by0, bz0, &mbar);
    }
  } else {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int gi0 = wrap_small(bx0 + lane, N1);
    int gi1 = wrap_small(bx0 + lane + 32, N1);
    for (int row = warp; row < BY * BZ; row += 8) {
      const int byi = row % BY, bzi = row / BY;  // compile-time divisors
      const int gj = wrap_small(by0 + byi, N2);
      int lp = g.glo;
      bool valid = true;
      if (RANK == 3) {
        lp = bz0 + bzi;
        if (g.glo) valid = lp >= 0 && lp < g.nplanes;
        else lp = wrap_small(lp, N3);
      }
      const double* src = phi + g.plane * (i64)lp + (unsigned)((int)g.rowp * gj);
      double* dst = box + row * BX;
      if (valid) {
        __pipeline_memcpy_async(dst + lane, src + gi0, 8);
        if (lane + 32 < NNPM_BXU) __pipeline_memcpy_async(dst + lane + 32, src + gi1, 8);
      } else {
        dst[lane] = 0.0;
        if (lane + 32 < NNPM_BXU) dst[lane + 32] = 0.0;
      }
    }
    __pipeline_commit();
  }
  // the first particle's loads are in flight while the box arrives
  const double* __restrict__ px0 = p.x[0] + n0;
  const double* __restrict__ px1 = p.x[1] + n0;
  const double* __restrict__ px2 = (RANK == 3) ? p.x[2] + n0 : nullptr;
  const double* __restrict__ pv0 = p.v[0] + n0;
  const double* __restrict__ pv1 = p.v[1] + n0;
  const double* __restrict__ pv2 = (RANK == 3) ? p.v[2] + n0 : nullptr;
  const int cnt = (int)(n1 - n0);
  // two register sets (A, B) alternate: while one particle is processed the next one's loads are in
  // flight, without register-to-register copies
  auto fetch = [&](double* x, double* v, int q) {
    if (q < cnt) {
      x[0] = px0[q]; x[1] = px1[q]; v[0] = pv0[q]; v[1] = pv1[q];
      if (RANK == 3) { x[2] = px2[q]; v[2] = pv2[q]; }
    }
  };
  int q = threadIdx.x;
  double xa[3] = {0.0, 0.0, 0.0}, va[3] = {0.0, 0.0, 0.0}, xb[3] = {0.0, 0.0, 0.0}, vb[3] = {0.0, 0.0, 0.0};
  fetch(xa, va, q);
  if (interior) {
    __syncthreads();        // mbarrier initialised by thread 0 before anyone polls it
    mbar_wait(&mbar, 0);
  } else {
    __pipeline_wait_prior(0);
    __syncthreads();
  }
  const double gfac = -0.5 * g.fN1;
  const double ra = 1.0 / pp.a;
  const int oi = ti0 - NNPM_TH, oj = tj0 - NNPM_TH, ok = (int)g.kz0 + tk0 - NNPM_TH;  // mesh cell of halo cell 0
  const unsigned fx = tg.tx + 2 * NNPM_TH, fy = tg.ty + 2 * NNPM_TH, fz = tg.tz + 2 * NNPM_TH;
  double mxa = -INFINITY, mxv = -INFINITY, mxp = -INFINITY;
  auto process = [&](double* x, double* v, int q) {
    double pa[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < RANK; d++) {
      if (COMOVING) x[d] = __dadd_rn(x[d], __dmul_rn(pp.c0, v[d]));
      x[d] = wrap01_sel(x[d]);
    }
    int i, j, k = 0;
    double dx, dy, dz = 0.0;
    cell_fast(x[0], g.fN1, i, dx);
    cell_fast(x[1], g.fN2, j, dy);
    const unsigned ri = (unsigned)tile_rel(i, oi, N1), rj = (unsigned)tile_rel(j, oj, N2);
    unsigned rk = 0;
    bool fast = ri < fx && rj < fy;
    if (RANK == 3) {
      cell_fast(x[2], g.fN3, k, dz);
      rk = (unsigned)tile_rel(k, ok, N3);
      fast = fast && rk < fz;
    }
    if (!fast) {
      slow_list[atomicAdd(slow_count, 1ull)] = (uint32_t)(n0 + q);
    } else {
      const double* b0 = box + (rk * BXY + rj * BX + ri);
      if (RANK == 2) cic_grad_gather_fma<2>([&](int a, int b, int) { return b0[b * BX + a]; }, dx, dy, 0.0, gfac, pa);
      else cic_grad_gather_fma<3>([&](int a, int b, int c) { return b0[c * BXY + b * BX + a]; }, dx, dy, dz, gfac, pa);
      push_finish_fast<RANK, COMOVING>(p, n0 + q, x, v, pa, pp, ra, mxa, mxv, mxp);
    }
  };
  while (q < cnt) {
    fetch(xb, vb, q + 256);
    process(xa, va, q);
    q += 256;
    if (q >= cnt) break;
    fetch(xa, va, q + 256);
    process(xb, vb, q);
    q += 256;
  }
  block_max3(mxa, mxv, mxp, red);
}

#undef PHI

// ------------------------------------------------------------------------------------------
// K1 tiled deposit (CIC): one CTA per tile of tile-sorted particles.  The tile's particles scatter
// into a shared-memory box covering cells [-H, T+H+1) around the tile (H = NNPM_TH cells a particle
// may have strayed since the sort, +1 for the CIC upper neighbour); the box is then flushed to the
// mesh with one global reduction per box cell -- (T+3)^3/T^3 ~ 2 global atomics per cell instead of
// 8 per particle.
//   64-bit shared-memory atomics are compare-and-swap loops on sm_100 (ATOMS.CAST.SPIN, for f64 and
// u64 alike; measured slower than the global REDG path), so the box accumulates unit-mass weights in
// 64-bit fixed point (2^-B resolution, B = fixed_bits) built from NATIVE 32-bit shared atomics: add
// the low word (ATOMS.ADD returning the old value), derive the carry, add high word + carry.  The
// sum is exact integer arithmetic, hence independent of the order of the adds.
//   FIXED : the flush adds the integers into the int64 mesh (bit-exact run to run, and identical to
//           the global fixed-point path);  otherwise the flush converts to m * q * 2^-B and uses
//           f64 global reductions.  Particles outside the box take the global-atomic path directly.
// ------------------------------------------------------------------------------------------
#define NNPM_DX (NNPM_TX + 2 * NNPM_TH + 2)  // box x origin = tile origin - H - 1: even, so rows are 16-byte aligned
#define NNPM_DY (NNPM_TY + 2 * NNPM_TH + 1)
#define NNPM_DZ (NNPM_TZ + 2 * NNPM_TH + 1)

// bulk asynchronous reduction shared -> global (TMA engine): dst[0..n) += src[0..n), n*8 bytes, both
// 16-byte aligned.  PTX ISA 8.0+, sm_90+: cp.reduce.async.bulk ... .add.f64 / .add.u64
template <bool FIXED>
__device__ __forceinline__ void bulk_reduce_add(void* gdst, const void* ssrc, unsigned bytes) {
  if (FIXED)
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
  else
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

// BULK: flush rows with cp.reduce.async.bulk (needs N1 even and >= NNPM_DX); otherwise per-cell atomics.
// UNR: particles whose loads are issued together before any of them is scattered (1 or 2): the loop has no
// software prefetch, so UNR = 2 doubles the bytes each thread keeps in flight (at ~12 more registers).
template <int RANK, bool FIXED, bool PRE, int UNR>
__global__ void __launch_bounds__(256, 5)
k_deposit_tile(PartPtrs p, i64 np, double c0, Geom g, TileGeom tg, WorkList wl, const uint32_t* __restrict__ tile_off,
               double* __restrict__ mesh, double m, double fscale, uint32_t* __restrict__ slow_list,
               u64* __restrict__ slow_count, int bulk, int pf) {
  constexpr int DX = NNPM_DX, DY = NNPM_DY, DZ = (RANK == 3) ? NNPM_DZ : 1, DXY = DX * DY, NBOX = DXY * DZ;
  // accumulators: the low and the high 32-bit words of the cells' 64-bit sums in TWO arrays (lo[NBOX], then hi[NBOX]), so that
  // the lanes of a warp -- on consecutive cells once the particles are cell-sorted -- hit consecutive 4-byte banks: as
  // interleaved (lo, hi) pairs every atomic touched only the even (or only the odd) banks, a built-in 2-way conflict on
  // top of the same-cell pairs (ncu r2e: 4.1 wavefronts per shared atomic).  Converted to one f64 / u64 per cell before the flush.
  constexpr int NBOXP = (NBOX + 3) & ~3;
  __shared__ __align__(16) uint32_t acc[2 * NBOXP];
  uint32_t* const acc_lo = acc;
  uint32_t* const acc_hi = acc + NBOXP;
  if (blockIdx.x >= *wl.count) return;  // the grid is sized for the worst case
  const int t = (int)wl.tile[blockIdx.x];
  const i64 n0 = wl.beg[blockIdx.x];
  i64 n1 = n0 + NNPM_CHUNK;
  if (n1 > (i64)tile_off[t + 1]) n1 = tile_off[t + 1];
  if (n1 > np) n1 = np;
  if (n0 >= n1) return;  // uniform across the block
  if (pf) l2_prefetch_particles<RANK, PRE>(p, n0, n1);  // lands while the box is being zeroed
  const int ttx = t % tg.ntx, tty = (t / tg.ntx) % tg.nty, ttz = t / (tg.ntx * tg.nty);
  for (int c = threadIdx.x; c < NBOXP / 2; c += 256) ((uint4*)acc)[c] = make_uint4(0u, 0u, 0u, 0u);
  const int N1 = (int)g.N1, N2 = (int)g.N2, N3 = (int)g.N3;
  const int bi0 = ttx * tg.tx - NNPM_TH - 1, bj0 = tty * tg.ty - NNPM_TH;  // mesh cell of box cell (0,0,0)
  const int bk0 = (int)g.kz0 + ttz * tg.tz - NNPM_TH;                       // global plane
  const unsigned fx = tg.tx + 2 * NNPM_TH, fy = tg.ty + 2 * NNPM_TH, fz = tg.tz + 2 * NNPM_TH;
  const double* __restrict__ px0 = p.x[0] + n0;
  const double* __restrict__ px1 = p.x[1] + n0;
  const double* __restrict__ px2 = (RANK == 3) ? p.x[2] + n0 : nullptr;
  const double* __restrict__ pv0 = p.v[0] + n0;
  const double* __restrict__ pv1 = p.v[1] + n0;
  const double* __restrict__ pv2 = (RANK == 3) ? p.v[2] + n0 : nullptr;
  const int cnt = (int)(n1 - n0);
  __syncthreads();
  // 64-bit fixed-point add from two native 32-bit shared atomics (the low word returns the old value)
  auto sadd = [&](int idx, u64 q) {
    const uint32_t lo = (uint32_t)q;
    const uint32_t old = atomicAdd(&acc_lo[idx], lo);
    atomicAdd(&acc_hi[idx], (uint32_t)(q >> 32) + (old > ~lo ? 1u : 0u));
  };
  const int lane = threadIdx.x & 31;
  // One particle per lane (valid = false: the lane has none).  Weights -> 2^B fixed point -> two atomics per CIC corner.
  // AGG (work items of tiles holding >= 4096 particles: collapsed halos of the clustered late-time state, BASELINE config 4):
  // warp-aggregated atomics.  Cell-sorted particles of one cell sit in adjacent lanes; when the warp holds a run of >= 8
  // of them, the runs of equal cells are summed by a segmented shuffle reduction (integer sums: exact, order-independent,
  // so the fixed-point mode stays bit-exact) and only the head lane of every run touches shared memory.  Without it a warp of
  // 32 same-cell particles serialises every one of its 16 atomics 32-fold: the cell-order sort alone made the clustered
  // deposit 29.9 ms against 15.9.  The run detection costs ~30 % more instructions, so sparse tiles take the plain path.
  auto scatter = [&](auto AGG, bool valid, int q, double x0, double x1, double x2) {
    constexpr bool agg_path = decltype(AGG)::value;
    int i, j, k = 0;
    double dx, dy, dz = 0.0;
    cell_fast(wrap01_sel(x0), g.fN1, i, dx);
    cell_fast(wrap01_sel(x1), g.fN2, j, dy);
    const unsigned ri = (unsigned)tile_rel(i, bi0, N1), rj = (unsigned)tile_rel(j, bj0, N2);
    unsigned rk = 0;
    bool fast = (ri - 1u) < fx && rj < fy;   // box column 0 is alignment padding
    if (RANK == 3) {
      cell_fast(wrap01_sel(x2), g.fN3, k, dz);
      rk = (unsigned)tile_rel(k, bk0, N3);
      fast = fast && rk < fz;
    }
    if (valid && !fast) slow_list[atomicAdd(slow_count, 1ull)] = (uint32_t)(n0 + q);  // strayed beyond the box: the global-atomic kernel takes it
    const bool on = valid && fast;
    if (!agg_path && !on) return;
    // unit-mass weights in the reference's order (1-dx)*(1-dy)*(1-dz) ... (ParticleMesh.jl:376-379,
    // :395-402), pre-scaled by 2^B (exact: a power of two commutes with every rounding)
    const double wx0 = __dmul_rn(__dsub_rn(1.0, dx), fscale), wx1 = __dmul_rn(dx, fscale);
    const double ody = __dsub_rn(1.0, dy);
    const double w00 = __dmul_rn(wx0, ody), w01 = __dmul_rn(wx0, dy), w10 = __dmul_rn(wx1, ody), w11 = __dmul_rn(wx1, dy);
    const int b = (int)(rk * DXY + rj * DX + ri);
    bool agg = false, writer = on;
    int runend = 31;
    if (agg_path) {
      // runs of equal cells in lane order
      const int key = on ? b : (int)(0x40000000 | lane);
      const int prev = __shfl_up_sync(0xffffffffu, key, 1);
      const unsigned follow = __ballot_sync(0xffffffffu, lane > 0 && key == prev);   // bit l: lane l continues lane l-1's run
      const unsigned t1 = follow & (follow >> 1), t2 = t1 & (t1 >> 2);
      agg = (t2 & (t1 >> 4) & (follow >> 6)) != 0u;   // some run has >= 8 lanes (7 consecutive followers): warp-uniform
      if (agg) {
        const unsigned heads_above = ~follow & ~((2u << lane) - 1u);             // run heads in lanes > lane
        runend = heads_above ? __ffs(heads_above) - 2 : 31;                      // last lane of my run
        writer = on && !((follow >> lane) & 1u);                                 // the head of every run adds the run's sums
      }
    }
    // one CIC corner for the whole warp: weight -> fixed point, (segmented sum over the run,) two atomics
    auto corner = [&](int idx, double w) {
      u64 qq = (u64)__double2ll_rn(w);
      if (agg_path && agg) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const u64 o = __shfl_down_sync(0xffffffffu, qq, d);
          if (lane + d <= runend) qq += o;
        }
      }
      if (writer) sadd(idx, qq);
    };
    if (RANK == 2) {
      corner(b, w00); corner(b + DX, w01); corner(b + 1, w10); corner(b + DX + 1, w11);
    } else {
      const double odz = __dsub_rn(1.0, dz);
      corner(b, __dmul_rn(w00, odz)); corner(b + DX, __dmul_rn(w01, odz));
      corner(b + 1, __dmul_rn(w10, odz)); corner(b + DX + 1, __dmul_rn(w11, odz));
      corner(b + DXY, __dmul_rn(w00, dz)); corner(b + DXY + DX, __dmul_rn(w01, dz));
      corner(b + DXY + 1, __dmul_rn(w10, dz)); corner(b + DXY + DX + 1, __dmul_rn(w11, dz));
    }
  };
  // warp-uniform trip count: every lane of a warp runs the same iterations (the aggregated path's shuffles need all 32 lanes)
  auto run = [&](auto AGG) {
    for (int q0 = (int)(threadIdx.x & ~31u); q0 < cnt; q0 += 256 * UNR) {
      const int q = q0 + lane;
      double x0[UNR], x1[UNR], x2[UNR], v0[UNR], v1[UNR], v2[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const int qu = q + 256 * u;
        x0[u] = x1[u] = x2[u] = v0[u] = v1[u] = v2[u] = 0.0;
        if (qu < cnt) {
          x0[u] = px0[qu]; x1[u] = px1[qu];
          if (RANK == 3) x2[u] = px2[qu];
          if (PRE) {
            v0[u] = pv0[qu]; v1[u] = pv1[qu];
            if (RANK == 3) v2[u] = pv2[qu];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const int qu = q + 256 * u;
        if (q0 + 256 * u >= cnt) break;   // warp-uniform
        if (PRE) {
          x0[u] = __dadd_rn(x0[u], __dmul_rn(c0, v0[u]));
          x1[u] = __dadd_rn(x1[u], __dmul_rn(c0, v1[u]));
          if (RANK == 3) x2[u] = __dadd_rn(x2[u], __dmul_rn(c0, v2[u]));
        }
        scatter(AGG, qu < cnt, qu, x0[u], x1[u], x2[u]);
      }
    }
  };
  if (tile_off[t + 1] - tile_off[t] >= 4096u) run(std::true_type{});   // block-uniform
  else run(std::false_type{});
  __syncthreads();
  // flush: box cell (bx,by,bz) -> mesh cell (bi0 + bx, bj0 + by, bk0 + bz), periodic.  Box extents are
  // <= N + 3, so one fold brings every coordinate into [0, N).
  const int ey = tg.ty + 2 * NNPM_TH + 1, ez = (RANK == 3) ? tg.tz + 2 * NNPM_TH + 1 : 1;
  const double inv_scale = 1.0 / fscale;  // exact: fscale = 2^B
  u64* acc64 = (u64*)acc;
  if (bulk) {
    // (lo[c], hi[c]) -> one 8-byte value per cell, in place through registers: u64 sums (FIXED) or m * q * 2^-B
    constexpr int PER = (NBOX + 255) / 256;
    u64 qv[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int c = threadIdx.x + 256 * k;
      qv[k] = c < NBOX ? ((u64)acc_hi[c] << 32) | (u64)acc_lo[c] : 0ull;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int c = threadIdx.x + 256 * k;
      if (c < NBOX) {
        if (FIXED) acc64[c] = qv[k];
        else ((double*)acc)[c] = __dmul_rn(m, __dmul_rn((double)(i64)qv[k], inv_scale));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy reads
    __syncthreads();
    const int row = threadIdx.x;  // one row (DX cells, 16-byte aligned) per thread: one or two bulk reductions
    if (row < ey * ez) {
      const int by = row % ey, bz = row / ey;
      int lp = g.glo;
      if (RANK == 3) lp = local_plane32(tile_rel(bk0 + bz, 0, N3), g);
      if (lp >= 0) {
        double* drow = mesh + g.plane * (i64)lp + (unsigned)((int)g.rowp * tile_rel(bj0 + by, 0, N2));
        const double* srow = (const double*)acc + (bz * DXY + by * DX);
        const int start = tile_rel(bi0, 0, N1);                 // even: bi0 and N1 are even
        const int len1 = (N1 - start) < DX ? (N1 - start) : DX;  // cells before the periodic wrap
        bulk_reduce_add<FIXED>(drow + start, srow, (unsigned)len1 * 8u);
        if (len1 < DX) bulk_reduce_add<FIXED>(drow, srow + len1, (unsigned)(DX - len1) * 8u);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the reads
    }
    return;
  }
  const int ex = tg.tx + 2 * NNPM_TH + 2;
  for (int c = threadIdx.x; c < NBOX; c += 256) {
    const u64 q = ((u64)acc_hi[c] << 32) | (u64)acc_lo[c];
    if (!q) continue;
    const int bx = c % DX, by = (c / DX) % DY, bz = c / DXY;  // compile-time divisors
    if (bx >= ex || by >= ey || bz >= ez) continue;
    int lp = g.glo;
    if (RANK == 3) {
      lp = local_plane32(tile_rel(bk0 + bz, 0, N3), g);
      if (lp < 0) continue;  // cannot hold mass: such a particle went to the slow list
    }
    double* dst = mesh + g.plane * (i64)lp + (unsigned)((int)g.rowp * tile_rel(bj0 + by, 0, N2)) + tile_rel(bi0 + bx, 0, N1);
    if (FIXED) atomicAdd((u64*)dst, q);
    else atomicAdd(dst, __dmul_rn(m, __dmul_rn((double)(i64)q, inv_scale)));
  }
}

// ------------------------------------------------------------------------------------------
// K0 counting sort by tile of NNPM_TX x NNPM_TY x NNPM_TZ cells (histogram -> scan -> scatter).
// After the scan `hist` holds the tile offsets the tiled deposit / push kernels index by.
// ------------------------------------------------------------------------------------------
// cpt > 0 (cell-order sort): the bin is (tile, cell within the tile), x fastest -- after the sort the particles of a
// tile are ordered by cell, so the 32 lanes of a warp work on ~32 consecutive cells of one mesh row: shared-memory
// gathers and atomics of the tiled kernels become conflict-free runs instead of scattered accesses.  cpt == 0: tile bins.
template <int RANK>
__global__ void __launch_bounds__(256)
k_sort_count(PartPtrs p, i64 np, Geom g, TileGeom tg, int cpt, uint32_t* __restrict__ hist,
             uint32_t* __restrict__ key, uint32_t* __restrict__ rnk) {
  // four particles per thread, their returning atomics all in flight before the first result is used: the kernel is
  // bound by the round trip of that atomic (ncu: long_scoreboard), not by its 34 GB of traffic
  constexpr int U = 4;
  const i64 base = blockIdx.x * (i64)(256 * U) + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t t[U], r[U];
  unsigned peers[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const i64 n = base + u * 256;
    t[u] = 0xFFFFFFFFu;  // lanes past the end form their own group and do nothing
    if (n < np) {
      int i, j, k, dummy;
      double d;
      cell_of32(wrap01(p.x[0][n]), g.fN1, (int)g.N1, i, dummy, d);
      cell_of32(wrap01(p.x[1][n]), g.fN2, (int)g.N2, j, dummy, d);
      int kl = 0;
      if (RANK == 3) {
        cell_of32(wrap01(p.x[2][n]), g.fN3, (int)g.N3, k, dummy, d);
        kl = k - (int)g.kz0;
        if (kl < 0) kl = 0;              // strays (not yet migrated) sort into the boundary tiles
        if (kl >= (int)g.nzl) kl = (int)g.nzl - 1;
      }
      const int ti = i / tg.tx, tj = j / tg.ty, tk = kl / tg.tz;
      t[u] = (uint32_t)(ti + tg.ntx * (tj + tg.nty * tk));
      if (cpt) t[u] = t[u] * (uint32_t)cpt + (uint32_t)((i - ti * tg.tx) + tg.tx * ((j - tj * tg.ty) + tg.ty * (kl - tk * tg.tz)));
    }
  }
  // warp-aggregated histogram: particles arrive nearly sorted (lattice ICs, previous sorts, collapsed halos), so a
  // warp often holds few distinct bins -- one atomic per distinct bin instead of one per lane.
#pragma unroll
  for (int u = 0; u < U; u++) {
    peers[u] = __match_any_sync(0xffffffffu, t[u]);
    const int leader = __ffs(peers[u]) - 1;
    r[u] = 0;
    if (lane == leader && t[u] != 0xFFFFFFFFu) r[u] = atomicAdd(&hist[t[u]], (uint32_t)__popc(peers[u]));
  }
#pragma unroll
  for (int u = 0; u < U; u++) {
    const i64 n = base + u * 256;
    const int leader = __ffs(peers[u]) - 1;
    const uint32_t b = __shfl_sync(peers[u], r[u], leader);
    if (n < np) {
      key[n] = t[u];
      rnk[n] = b + (uint32_t)__popc(peers[u] & ((1u << lane) - 1u));
    }
  }
}
// tile offsets out of the (scanned) bin offsets of a cell-order sort: toff[t] = boff[t * cpt], toff[ntiles] = np
__global__ void k_tile_offsets(uint32_t* __restrict__ toff, const uint32_t* __restrict__ boff, i64 ntiles, int cpt, uint32_t np) {
  const i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t < ntiles) toff[t] = boff[t * (i64)cpt];
  else if (t == ntiles) toff[t] = np;
}

// exclusive scan of uint32, reduce-then-scan: blocks of SCAN_ITEMS = 256 threads x 16 items.
//   k_scan_reduce: bsum[b] = sum of block b;   (the block sums are scanned recursively)
//   k_scan_apply : data[i] = bprefix[b] + exclusive prefix within block b  (thread-sequential 16 items, warp shuffles,
//                  one shared-memory hop across the 8 warps)
#define SCAN_ITEMS 4096
__device__ __forceinline__ uint32_t scan_block_excl(uint32_t mine, uint32_t* total) {   // exclusive prefix of `mine` over 256 threads
  __shared__ uint32_t wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    if (w < warp) base += wsum[w];
    tot += wsum[w];
  }
  __syncthreads();
  if (total) *total = tot;
  return base + inc - mine;
}
__global__ void __launch_bounds__(256) k_scan_reduce(const uint32_t* __restrict__ data, uint32_t* __restrict__ bsum, i64 n) {
  const i64 b0 = blockIdx.x * (i64)SCAN_ITEMS + threadIdx.x * 16;
  uint32_t acc = 0;
  if (b0 + 16 <= n && ((size_t)(data + b0) & 15) == 0) {
    const uint4* q = (const uint4*)(data + b0);
#pragma unroll
    for (int u = 0; u < 4; u++) { const uint4 v = q[u]; acc += v.x + v.y + v.z + v.w; }
  } else {
    for (int u = 0; u < 16; u++) if (b0 + u < n) acc += data[b0 + u];
  }
  uint32_t tot;
  scan_block_excl(acc, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(256) k_scan_apply(uint32_t* __restrict__ data, const uint32_t* __restrict__ bprefix, i64 n) {
  const i64 b0 = blockIdx.x * (i64)SCAN_ITEMS + threadIdx.x * 16;
  uint32_t v[16];
  const bool fast = b0 + 16 <= n && ((size_t)(data + b0) & 15) == 0;
  if (fast) {
    const uint4* q = (const uint4*)(data + b0);
#pragma unroll
    for (int u = 0; u < 4; u++) { const uint4 t = q[u]; v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w; }
  } else {
#pragma unroll
    for (int u = 0; u < 16; u++) v[u] = (b0 + u < n) ? data[b0 + u] : 0u;
  }
  uint32_t acc = 0;
#pragma unroll
  for (int u = 0; u < 16; u++) acc += v[u];
  uint32_t run = scan_block_excl(acc, nullptr) + (bprefix ? bprefix[blockIdx.x] : 0u);
#pragma unroll
  for (int u = 0; u < 16; u++) { const uint32_t t = v[u]; v[u] = run; run += t; }
  if (fast) {
    uint4* q = (uint4*)(data + b0);
#pragma unroll
    for (int u = 0; u < 4; u++) q[u] = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
  } else {
#pragma unroll
    for (int u = 0; u < 16; u++) if (b0 + u < n) data[b0 + u] = v[u];
  }
}
__global__ void k_sort_dest(uint32_t* __restrict__ key, const uint32_t* __restrict__ rnk,
                            const uint32_t* __restrict__ off, i64 np) {
  i64 n = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (n < np) key[n] = off[key[n]] + rnk[n];
}
template <typename T>
__global__ void k_permute(T* __restrict__ out, const T* __restrict__ in, const uint32_t* __restrict__ dest, i64 np) {
  i64 n = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (n < np) out[dest[n]] = in[n];
}

// ------------------------------------------------------------------------------------------
// device-side synthetic ICs: lattice (ParticleMesh.jl:257-287) + plane-wave Zel'dovich field
// ------------------------------------------------------------------------------------------
struct ICModes {
  int n;
  double amp;
  int kx[64], ky[64], kz[64];
  double ph[64];
};
template <int RANK>
__global__ void k_init_synth(PartPtrs p, uint32_t* ids, i64 nloc, i64 first, i64 P1, i64 P2, i64 P3, ICModes M,
                             double vfac) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= nloc) return;
  i64 gidx = first + t;
  i64 i = gidx % P1, j = (gidx / P1) % P2, k = gidx / (P1 * P2);
  double q[3];
  q[0] = __ddiv_rn(__dsub_rn((double)(i + 1), 0.5), (double)P1);
  q[1] = __ddiv_rn(__dsub_rn((double)(j + 1), 0.5), (double)P2);
  q[2] = (RANK == 3) ? __ddiv_rn(__dsub_rn((double)(k + 1), 0.5), (double)P3) : 0.0;
  double psi[3] = {0.0, 0.0, 0.0};
  for (int mIdx = 0; mIdx < M.n; mIdx++) {
    double k2 = (double)(M.kx[mIdx] * M.kx[mIdx] + M.ky[mIdx] * M.ky[mIdx] + M.kz[mIdx] * M.kz[mIdx]);
    double arg = M.kx[mIdx] * q[0] + M.ky[mIdx] * q[1] + M.kz[mIdx] * q[2] + M.ph[mIdx];
    double s = sinpi(2.0 * arg);
    psi[0] += (M.amp * M.kx[mIdx] / k2) * s;
    psi[1] += (M.amp * M.ky[mIdx] / k2) * s;
    psi[2] += (M.amp * M.kz[mIdx] / k2) * s;
  }
#pragma unroll
  for (int d = 0; d < RANK; d++) {
    p.x[d][t] = wrap01(q[d] + psi[d]);
    p.v[d][t] = vfac * psi[d];
  }
  ids[t] = (uint32_t)gidx;
}

// Clustered late-time stand-in (BASELINE config 4): even-hash particles stay on
// the lattice, the others fall into `nblobs` Gaussian blobs of width sigma with a b^-1/2 mass function
// (blob rank b in [1, nblobs]); centres, membership and offsets are pure functions of (seed, global
// particle index), so any decomposition generates the same particles.  Mirrored by
// oracle/pm_ref.py:synthetic_clustered.
__device__ __forceinline__ u64 mix64(u64 z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(u64 h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
template <int RANK>
__global__ void k_init_clustered(PartPtrs p, uint32_t* ids, i64 nloc, i64 first, i64 P1, i64 P2, i64 P3, u64 seed, i64 nblobs,
                                 double sigma, double vdisp) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= nloc) return;
  const i64 gidx = first + t;
  const i64 i = gidx % P1, j = (gidx / P1) % P2, k = gidx / (P1 * P2);
  double x[3];
  x[0] = __ddiv_rn(__dsub_rn((double)(i + 1), 0.5), (double)P1);
  x[1] = __ddiv_rn(__dsub_rn((double)(j + 1), 0.5), (double)P2);
  x[2] = (RANK == 3) ? __ddiv_rn(__dsub_rn((double)(k + 1), 0.5), (double)P3) : 0.0;
  const u64 h0 = mix64(seed ^ mix64((u64)gidx));
  double v[3] = {0.0, 0.0, 0.0};
  if (h0 & 1ull) {
    const double u = u01(mix64(h0 + 1));
    const double sq = 1.0 + u * (sqrt((double)nblobs) - 1.0);
    i64 b = (i64)(sq * sq);
    if (b > nblobs) b = nblobs;
    const u64 hb = mix64(seed * 0x100000001B3ull + (u64)b);
#pragma unroll
    for (int d = 0; d < RANK; d++) {
      const double c = u01(mix64(hb + 11 * (d + 1)));
      const double u1 = u01(mix64(h0 + 101 * (d + 1))), u2 = u01(mix64(h0 + 211 * (d + 1)));
      const double r = sqrt(-2.0 * log(u1));
      x[d] = c + sigma * r * cospi(2.0 * u2);
      x[d] -= floor(x[d]);
      v[d] = vdisp * r * sinpi(2.0 * u2);
    }
  }
#pragma unroll
  for (int d = 0; d < RANK; d++) {
    p.x[d][t] = x[d];
    p.v[d][t] = v[d];
  }
  ids[t] = (uint32_t)gidx;
}
