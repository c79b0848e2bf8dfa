// nnpm.cu -- C ABI of libnnpm.so (see include/nnpm.h): context, device-resident state,
// cuFFT plans, step sequencing.  Kernels live in nnpm_kernels.cuh, the slab-decomposed
// multi-GPU plumbing (NCCL) in nnpm_comm.cuh.
#include "../../include/nnpm.h"

#include <cuda_runtime.h>
#include <cufft.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "nnpm_kernels.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct NcclApi;  // nnpm_comm.cuh

struct nnpm_ctx {
  nnpm_config cfg;
  int R;  // particle rank (2|3)
  Geom g;
  i64 icn;       // N1/2+1
  int P, r;      // slab decomposition
  i64 njl;       // N2/P: local j range after the slab transpose
  cudaStream_t st;
  bool own_stream;
  std::string err;
  size_t dev_bytes;
  std::vector<void*> allocs;

  // mesh: nplanes padded planes; rho and phi share it (rho is consumed by the solve)
  double* mesh;
  double* meshT;  // nranks>1: transposed spectrum [k][jl][ic]
  double* meshS;  // nranks>1: send/recv staging   [d][kl][jl][ic]
  double *s1, *s2, *s3;  // sin^2(k/2) tables
  int field_state;       // what mesh currently holds: 0 nothing, 1 rho, 2 phi
  double* rho_keep;      // "keep_rho": copy of the owned planes' density taken by the fused step before the solve
  bool rho_keep_valid;
  int opt_keep_rho;

  // particles (SoA)
  PartPtrs pp;
  double* spare;       // rotating spare array for out-of-place permutes
  uint32_t* ids;
  uint32_t* ids_spare;
  i64 cap, np;
  double m;
  bool ids_dirty;      // particle order != original order
  i64 first_id;        // global index of the first uploaded particle (ids = first_id + upload position)

  // lazily allocated parity buffers
  double* acc;
  i64 acc_k0, acc_nq;
  double* pa[3];
  bool acc_valid, pa_valid;

  // staging for host<->device AoS copies
  double* stage;
  i64 stage_elems;

  // sort scratch
  uint32_t *hist, *hist_fine, *bsum, *bsum2, *skey, *srnk;   // hist: tile offsets; hist_fine: (tile, cell) bin offsets of the cell-order sort
  i64 bsum_cap;
  int opt_sort_cells;
  uint32_t *wcnt, *wtile, *wbeg;  // work list of the tiled kernels (k_work_count / k_work_fill)
  i64 wcap;
  TileGeom tg;
  i64 ntiles;
  i64 sorted_np;     // particles [0, sorted_np) are covered by the tile offsets in `hist` (0 = not sorted)
  i64 last_untiled;  // stragglers of the last tiled push

  CUtensorMap tmap_phi;  // 3-D tiled tensor map over the mesh (rowp x N2 x nplanes doubles), box = push tile box
  bool have_tmap;

  // reductions
  i64* d_red;  // [0..2] encoded maxima, [3] violation counter
  i64* h_red;  // pinned

  // cuFFT
  cufftHandle plan2f, plan2b, planz, plan3f, plan3b;
  bool have2, havez, have3;
  int opt_fft3d, opt_zfused, opt_tile_deposit, opt_tile_push, opt_bulk_flush;
  void* fftwork;
  size_t fftwork_bytes;

  // timing
  cudaEvent_t ev[16];
  cudaEvent_t ev_fft[8];  // x fwd [0,1], y fwd [1,2], y inv [4,5], x inv [5,6]  (timing = 1)
  int step_count;
  int launches;

  // comm
  NcclApi* nccl;
  void* comm;
  bool comm_borrowed;   // temporary lattice context of nnpm_init_grf: the communicator belongs to its parent
  // migration scratch (nranks > 1)
  double* mig_send;
  double* mig_recv;
  i64 mig_cap;
  uint32_t* mig_holes;
  uint32_t* mig_lo;
  double* ghost_rx;
  double* red_buf;
  i64* mig_cnt;    // device counters
  i64* h_cnt;      // pinned
  i64 last_migrated;
  double* pk_buf;
  double* ic_partial;   // per-block partial sums of nnpm_init_grf's displacement norm
  double2* ztw;  // z-FFT twiddles W_N3^m
  // NVLink peer-memory transposes (nranks > 1): every rank's mesh / meshT mapped through CUDA IPC
  double* peer_mesh[64];
  double* peer_meshT[64];
  bool peer_ok;
  int opt_peer;   // un-fused transposes: 0 = NCCL send/recv, 1 = copy-engine peer copies, 2 = SM bulk-copy kernel, 3 = copy engines + SM kernel for the last chunk (default)
  int* bar_buf;
  // TMA-staged column FFT engine (nnpm_colfft.cuh)
  CUtensorMap tm_y, tm_z8, tm_z4;
  bool have_tm_y, have_tm_z8, have_tm_z4, have1;
  int opt_zsplit;  // test hook: run N3 <= 1024 through the radix-2 wrapper of the persistent fused z pass
  double2* ytw;
  cufftHandle plan1f, plan1b;  // batched 1-D x transforms (r2c / c2r) over the owned rows
  int opt_pipeline;
  cudaEvent_t ev_chunk[4];
  int opt_l2pf_push, opt_l2pf_deposit, opt_colfft, opt_cfrun_y, opt_rowfft, opt_fuse_transpose, opt_mig_in_push, nsm;
  double2 *xtwM, *xtwN;         // x-pass twiddles W_M^m, W_N^k (nnpm_rowfft.cuh)
  unsigned* f2_sync;            // work queue, fault flag and per-plane counters of the fused 2-D transform (k_fft2d)
  int opt_fft2d, opt_fft2d_block;
  struct PeerMaps* peer_maps;   // tensor maps over every rank's transposed spectrum (fused forward transpose)
  cudaStream_t st_hi;     // highest-priority side stream of the SM transport ("peer_transpose" = 2)
  cudaEvent_t ev_hi;
  int opt_xfer_ctas;      // its grid (0 = one CTA per SM)
  cudaStream_t st2[8];    // side streams: the blocks of a transpose go out on several copy engines at once
  cudaEvent_t ev_fork, ev_join[8];
  int opt_ce_streams, opt_pipe_chunks;
  int pipe_nch, ce_ns;    // chunks / side streams in effect for the current solve
};

static thread_local std::string g_create_err;

static int fail(nnpm_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_err = buf;
  return code;
}

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail(ctx, NNPM_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define CKF(call)                                                                                    \
  do {                                                                                               \
    cufftResult r_ = (call);                                                                         \
    if (r_ != CUFFT_SUCCESS)                                                                         \
      return fail(ctx, NNPM_ERR_CUFFT, "%s:%d %s -> cufft error %d", __FILE__, __LINE__, #call, (int)r_); \
  } while (0)
#define CKR(call)                    \
  do {                               \
    int rc_ = (call);                \
    if (rc_ != NNPM_OK) return rc_;  \
  } while (0)
#define LAUNCHED(ctx) ((ctx)->launches++)
#define KCHECK() CK(cudaGetLastError())

static inline unsigned nblk(i64 n, int b) { return (unsigned)((n + b - 1) / b); }

template <typename T>
static int dalloc(nnpm_ctx* ctx, T** p, size_t count) {
  void* q = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess)
    return fail(ctx, NNPM_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  ctx->allocs.push_back(q);
  ctx->dev_bytes += bytes;
  *p = (T*)q;
  return NNPM_OK;
}

#include "nnpm_comm.cuh"
#include "nnpm_zfft.cuh"
#include "nnpm_colfft.cuh"
#include "nnpm_rowfft.cuh"
#include "nnpm_xfer.cuh"

// ------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------
extern "C" int nnpm_version(void) { return NNPM_VERSION; }

extern "C" const char* nnpm_last_error(const nnpm_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

static int grow_fftwork(nnpm_ctx* ctx, size_t need) {
  if (need <= ctx->fftwork_bytes) return NNPM_OK;
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, need);
  if (e != cudaSuccess) return fail(ctx, NNPM_ERR_CUDA, "cudaMalloc(cuFFT work area, %zu bytes): %s", need, cudaGetErrorString(e));
  if (ctx->fftwork) {
    CK(cudaStreamSynchronize(ctx->st));
    cudaFree(ctx->fftwork);
    ctx->dev_bytes -= ctx->fftwork_bytes;
  }
  ctx->fftwork = q;
  ctx->fftwork_bytes = need;
  ctx->dev_bytes += need;
  cufftHandle hs[7] = {ctx->plan2f, ctx->plan2b, ctx->planz, ctx->plan3f, ctx->plan3b, ctx->plan1f, ctx->plan1b};
  bool on[7] = {ctx->have2, ctx->have2, ctx->havez, ctx->have3, ctx->have3, ctx->have1, ctx->have1};
  for (int i = 0; i < 7; i++)
    if (on[i]) CKF(cufftSetWorkArea(hs[i], ctx->fftwork));
  return NNPM_OK;
}

static int new_plan(nnpm_ctx* ctx, cufftHandle* h, int rank, long long* n, long long* inembed, long long istride,
                    long long idist, long long* onembed, long long ostride, long long odist, cufftType type,
                    long long batch, size_t* ws) {
  CKF(cufftCreate(h));
  CKF(cufftSetAutoAllocation(*h, 0));
  CKF(cufftMakePlanMany64(*h, rank, n, inembed, istride, idist, onembed, ostride, odist, type, batch, ws));
  CKF(cufftSetStream(*h, ctx->st));
  return NNPM_OK;
}

// local transforms: batched 2-D r2c/c2r over the owned planes (+ a strided 1-D z transform, used
// by the power spectrum and by the un-fused z path).  All plans share one work area.
static int make_plans(nnpm_ctx* ctx) {
  const Geom& g = ctx->g;
  size_t ws = 0, wmax = 0;
  long long n[2] = {g.N2, g.N1};
  long long inembed[2] = {g.N2, g.rowp};
  long long onembed[2] = {g.N2, ctx->icn};
  CKR(new_plan(ctx, &ctx->plan2f, 2, n, inembed, 1, g.plane, onembed, 1, g.plane / 2, CUFFT_D2Z, g.nzl, &ws));
  wmax = ws > wmax ? ws : wmax;
  CKR(new_plan(ctx, &ctx->plan2b, 2, n, onembed, 1, g.plane / 2, inembed, 1, g.plane, CUFFT_Z2D, g.nzl, &ws));
  wmax = ws > wmax ? ws : wmax;
  ctx->have2 = true;
  {  // x-only transforms for the [cuFFT x pass + TMA column-FFT y pass] form of the 2-D transform
    long long n1[1] = {g.N1};
    long long e1[1] = {g.rowp}, e2[1] = {ctx->icn};
    CKR(new_plan(ctx, &ctx->plan1f, 1, n1, e1, 1, g.rowp, e2, 1, ctx->icn, CUFFT_D2Z, g.N2 * g.nzl, &ws));
    wmax = ws > wmax ? ws : wmax;
    CKR(new_plan(ctx, &ctx->plan1b, 1, n1, e2, 1, ctx->icn, e1, 1, g.rowp, CUFFT_Z2D, g.N2 * g.nzl, &ws));
    wmax = ws > wmax ? ws : wmax;
    ctx->have1 = true;
  }
  if (g.N3 > 1) {
    long long nz[1] = {g.N3};
    long long stride = ctx->njl * ctx->icn;  // [k][jl][ic]: k has stride njl*icn
    long long emb[1] = {g.N3};
    CKR(new_plan(ctx, &ctx->planz, 1, nz, emb, stride, 1, emb, stride, 1, CUFFT_Z2Z, stride, &ws));
    wmax = ws > wmax ? ws : wmax;
    ctx->havez = true;
  }
  CKR(grow_fftwork(ctx, wmax ? wmax : 8));
  if (getenv("NNPM_VERBOSE"))
    fprintf(stderr, "[nnpm] rank %d/%d mesh %lldx%lldx%lld slab %lld planes, cuFFT work %.1f MB\n", ctx->r, ctx->P,
            g.N1, g.N2, g.N3, g.nzl, wmax / 1048576.0);
  return NNPM_OK;
}

static int ensure_plan3(nnpm_ctx* ctx) {
  if (ctx->have3) return NNPM_OK;
  const Geom& g = ctx->g;
  size_t ws = 0, wmax = 0;
  long long n[3] = {g.N3, g.N2, g.N1};
  long long inembed[3] = {g.N3, g.N2, g.rowp};
  long long onembed[3] = {g.N3, g.N2, ctx->icn};
  CKR(new_plan(ctx, &ctx->plan3f, 3, n, inembed, 1, g.plane * g.N3, onembed, 1, g.plane / 2 * g.N3, CUFFT_D2Z, 1, &ws));
  wmax = ws > wmax ? ws : wmax;
  CKR(new_plan(ctx, &ctx->plan3b, 3, n, onembed, 1, g.plane / 2 * g.N3, inembed, 1, g.plane * g.N3, CUFFT_Z2D, 1, &ws));
  wmax = ws > wmax ? ws : wmax;
  ctx->have3 = true;
  if (wmax > ctx->fftwork_bytes) CKR(grow_fftwork(ctx, wmax));
  else { CKF(cufftSetWorkArea(ctx->plan3f, ctx->fftwork)); CKF(cufftSetWorkArea(ctx->plan3b, ctx->fftwork)); }
  return NNPM_OK;
}

static void sin2_table(std::vector<double>& t, i64 N, bool wrapneg, bool round_half) {
  // kx = 2pi*(i-1)/N ; wrapped to 2pi*(i-N-1)/N above N/2+1 for the 2nd/3rd axes
  // (ParticleMesh.jl:451-458, 483-494); value = sin(k/2)^2
  t.resize(N);
  i64 half = round_half ? (i64)nearbyint((double)N / 2.0) : N / 2;
  for (i64 i = 0; i < N; i++) {
    double k = 2.0 * M_PI * (double)i / (double)N;
    if (wrapneg && (i + 1) > half + 1) k = 2.0 * M_PI * (double)(i - N) / (double)N;
    double s = sin(k / 2);
    t[i] = s * s;
  }
}

// TMA descriptor of the mesh for the tiled push: doubles, dims (N1, N2, nplanes) with the padded
// strides of the in-place FFT layout, box = (NNPM_BX, NNPM_BY, NNPM_BZ|1), no swizzle, OOB -> 0.
static int make_tensor_map(nnpm_ctx* ctx) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(ctx, NNPM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  const Geom& g = ctx->g;
  cuuint64_t dims[3] = {(cuuint64_t)g.N1, (cuuint64_t)g.N2, (cuuint64_t)g.nplanes};
  cuuint64_t strides[2] = {(cuuint64_t)g.rowp * 8, (cuuint64_t)g.plane * 8};
  cuuint32_t box[3] = {NNPM_BX, NNPM_BY, (cuuint32_t)(ctx->R == 3 ? NNPM_BZ : 1)};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((encode_fn)fn)(&ctx->tmap_phi, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, ctx->mesh, dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, NNPM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  ctx->have_tmap = true;
  return NNPM_OK;
}

extern "C" int nnpm_create(nnpm_ctx** out, const nnpm_config* cfg) {
  nnpm_ctx* ctx = nullptr;
  if (!out || !cfg) return fail(ctx, NNPM_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->struct_size != (int32_t)sizeof(nnpm_config))
    return fail(ctx, NNPM_ERR_ARG, "nnpm_config.struct_size %d != %zu (ABI mismatch)", cfg->struct_size,
                sizeof(nnpm_config));
  if (cfg->rank != 2 && cfg->rank != 3) return fail(ctx, NNPM_ERR_ARG, "rank must be 2 or 3");
  const i64 N1 = cfg->ngrid[0], N2 = cfg->ngrid[1], N3 = cfg->ngrid[2];
  if (N1 < 2 || N2 < 2 || N3 < 1) return fail(ctx, NNPM_ERR_ARG, "ngrid must be >= (2,2,1)");
  if (cfg->rank == 2 && N3 != 1) return fail(ctx, NNPM_ERR_ARG, "rank 2 needs ngrid[2] == 1");
  if (cfg->rank == 3 && N3 < 2) return fail(ctx, NNPM_ERR_ARG, "rank 3 needs ngrid[2] >= 2");
  const int P = cfg->nranks < 1 ? 1 : cfg->nranks;
  if (cfg->rank_id < 0 || cfg->rank_id >= P) return fail(ctx, NNPM_ERR_ARG, "rank_id out of range");
  if (P > 64) return fail(ctx, NNPM_ERR_ARG, "nranks must be <= 64");
  if (P > 1 && (cfg->rank != 3 || N3 % P || N2 % P || N3 / P < 4))
    return fail(ctx, NNPM_ERR_ARG, "nranks>1 needs rank 3, N3 %% P == 0, N2 %% P == 0 and N3/P >= 4");
  if (cfg->npart_total < 0 || cfg->npart_total >= (1ll << 32))
    return fail(ctx, NNPM_ERR_ARG, "npart_total must be < 2^32 (particle ids are uint32)");
  if ((cfg->deposit != 0 && cfg->deposit != 1) || (cfg->backinterp != 0 && cfg->backinterp != 1))
    return fail(ctx, NNPM_ERR_ARG, "deposit/backinterp must be NNPM_INTERP_NONE or NNPM_INTERP_CIC");
  if (cfg->bugcompat_interp_z && P > 1) return fail(ctx, NNPM_ERR_ARG, "bugcompat_interp_z needs nranks == 1");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(ctx, NNPM_ERR_CUDA, "no CUDA device (%s): libnnpm has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(ctx, NNPM_ERR_ARG, "device %d out of range", cfg->device);

  ctx = new nnpm_ctx();
  nnpm_ctx* c = ctx;
  memset(&c->g, 0, sizeof c->g);
  c->cfg = *cfg;
  c->cfg.nranks = P;
  if (c->cfg.fixed_bits <= 0) c->cfg.fixed_bits = 44;
  c->R = cfg->rank;
  c->P = P;
  c->r = cfg->rank_id;
  c->icn = N1 / 2 + 1;
  c->njl = N2 / P;
  Geom& g = c->g;
  g.N1 = N1; g.N2 = N2; g.N3 = N3;
  g.rowp = 2 * c->icn;
  g.plane = N2 * g.rowp;
  g.nzl = N3 / P;
  g.kz0 = g.nzl * c->r;
  g.glo = P > 1 ? 2 : 0;
  g.nplanes = (int)(g.nzl + (P > 1 ? 5 : 0));
  g.fN1 = (double)N1; g.fN2 = (double)N2; g.fN3 = (double)N3;
  c->dev_bytes = 0;
  c->mesh = c->meshT = c->meshS = nullptr;
  c->rho_keep = nullptr; c->rho_keep_valid = false; c->opt_keep_rho = 0;
  c->spare = nullptr; c->ids = c->ids_spare = nullptr;
  c->acc = nullptr; c->pa[0] = c->pa[1] = c->pa[2] = nullptr;
  c->acc_valid = c->pa_valid = false;
  c->stage = nullptr; c->stage_elems = 0;
  c->hist = c->hist_fine = c->bsum = c->bsum2 = c->skey = c->srnk = nullptr; c->bsum_cap = 0; c->opt_sort_cells = 1;
  c->wcnt = c->wtile = c->wbeg = nullptr; c->wcap = 0;
  c->have2 = c->havez = c->have3 = false;
  c->have_tmap = false;
  c->opt_fft3d = 0; c->opt_zfused = 3; c->opt_tile_deposit = 1; c->opt_tile_push = 1; c->opt_bulk_flush = 1;
  c->fftwork = nullptr; c->fftwork_bytes = 0;
  c->step_count = 0; c->launches = 0;
  c->nccl = nullptr; c->comm = nullptr; c->comm_borrowed = false;
  c->mig_send = c->mig_recv = nullptr; c->mig_cap = 0; c->mig_holes = nullptr; c->mig_cnt = nullptr; c->h_cnt = nullptr;
  c->last_migrated = 0;
  c->pk_buf = nullptr; c->ic_partial = nullptr;
  c->ztw = nullptr;
  for (int i = 0; i < 64; i++) { c->peer_mesh[i] = nullptr; c->peer_meshT[i] = nullptr; }
  c->peer_ok = false; c->opt_peer = 3; c->bar_buf = nullptr;
  c->ev_fork = nullptr;
  for (int i = 0; i < 8; i++) { c->st2[i] = nullptr; c->ev_join[i] = nullptr; }
  c->st_hi = nullptr; c->ev_hi = nullptr; c->opt_xfer_ctas = 0;
  c->opt_ce_streams = 0; c->opt_pipe_chunks = 0;   // 0 = automatic (see do_solve)
  c->pipe_nch = 4; c->ce_ns = 4;
  c->opt_pipeline = 1;
  for (int i = 0; i < 4; i++) c->ev_chunk[i] = nullptr;
  c->have_tm_y = c->have_tm_z8 = c->have_tm_z4 = c->have1 = false; c->opt_zsplit = 0; c->ytw = nullptr;
  c->opt_l2pf_push = 1; c->opt_l2pf_deposit = 1; c->opt_colfft = 1; c->opt_cfrun_y = 1; c->opt_rowfft = 1; c->opt_fuse_transpose = 0; c->opt_mig_in_push = 1;
  c->xtwM = c->xtwN = nullptr; c->peer_maps = nullptr; c->nsm = 148;
  c->f2_sync = nullptr; c->opt_fft2d = 0; c->opt_fft2d_block = 4;
  c->mig_lo = nullptr; c->ghost_rx = nullptr; c->red_buf = nullptr;
  c->np = 0; c->m = 0.0; c->ids_dirty = false; c->field_state = 0; c->first_id = 0;
  c->own_stream = false;
  c->h_red = nullptr;
  for (int i = 0; i < 16; i++) c->ev[i] = nullptr;
  for (int i = 0; i < 8; i++) c->ev_fft[i] = nullptr;

  auto bail = [&](int rc) {
    g_create_err = c->err;
    nnpm_destroy(c);
    return rc;
  };
#define CKB(call) do { int rc2_ = [&]() -> int { call; return NNPM_OK; }(); if (rc2_ != NNPM_OK) return bail(rc2_); } while (0)

  CKB(CK(cudaSetDevice(cfg->device)));
  CKB(CK(cudaDeviceGetAttribute(&c->nsm, cudaDevAttrMultiProcessorCount, cfg->device)));
  if (cfg->stream) c->st = (cudaStream_t)cfg->stream;
  else { CKB(CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking))); c->own_stream = true; }

  c->cap = cfg->part_capacity > 0 ? cfg->part_capacity : (i64)((double)cfg->npart_total / P * (P > 1 ? 1.25 : 1.0)) + (P > 1 ? 4096 : 0);
  if (c->cap < 1) c->cap = 1;
  CKB(CKR(dalloc(c, &c->mesh, (size_t)g.plane * g.nplanes)));
  if (P > 1) {
    CKB(CKR(dalloc(c, &c->meshT, (size_t)g.plane * g.nzl)));
    // meshS (staging of the grouped ncclSend / ncclRecv transposes) is allocated on first use: the peer-memory paths
    // never need it, and at 2048^3 on 2 GPUs it would be another 34 GB
  }
  for (int d = 0; d < 3; d++) { c->pp.x[d] = nullptr; c->pp.v[d] = nullptr; }
  for (int d = 0; d < c->R; d++) {
    // +2: the bulk-copy push rounds its last 16-byte unit up and may read one element past np
    CKB(CKR(dalloc(c, &c->pp.x[d], (size_t)c->cap + 2)));
    CKB(CKR(dalloc(c, &c->pp.v[d], (size_t)c->cap + 2)));
  }
  CKB(CKR(dalloc(c, &c->ids, (size_t)c->cap)));
  CKB(CKR(dalloc(c, &c->d_red, 8)));
  CKB(CK(cudaHostAlloc((void**)&c->h_red, 8 * sizeof(i64), cudaHostAllocDefault)));
  CKB(CK(cudaHostAlloc((void**)&c->h_cnt, 72 * 64 * sizeof(i64), cudaHostAllocDefault)));  // MC_REC * 64 (nnpm_comm.cuh)
  {
    std::vector<double> t1, t2, t3;
    sin2_table(t1, N1, false, false);
    sin2_table(t2, N2, true, c->R == 3);
    sin2_table(t3, N3, true, true);
    CKB(CKR(dalloc(c, &c->s1, (size_t)N1)));
    CKB(CKR(dalloc(c, &c->s2, (size_t)N2)));
    CKB(CKR(dalloc(c, &c->s3, (size_t)N3)));
    CKB(CK(cudaMemcpy(c->s1, t1.data(), N1 * 8, cudaMemcpyHostToDevice)));
    CKB(CK(cudaMemcpy(c->s2, t2.data(), N2 * 8, cudaMemcpyHostToDevice)));
    CKB(CK(cudaMemcpy(c->s3, t3.data(), N3 * 8, cudaMemcpyHostToDevice)));
  }
  CKB(CKR(make_tensor_map(c)));
  CKB(CKR(make_plans(c)));
  for (int i = 0; i < 16; i++) CKB(CK(cudaEventCreate(&c->ev[i])));
  for (int i = 0; i < 8; i++) CKB(CK(cudaEventCreate(&c->ev_fft[i])));
  c->tg.tx = (int)(N1 < NNPM_TX ? N1 : NNPM_TX);
  c->tg.ty = (int)(N2 < NNPM_TY ? N2 : NNPM_TY);
  c->tg.tz = c->R == 3 ? (int)(g.nzl < NNPM_TZ ? g.nzl : NNPM_TZ) : 1;
  c->tg.ntx = (int)((N1 + c->tg.tx - 1) / c->tg.tx);
  c->tg.nty = (int)((N2 + c->tg.ty - 1) / c->tg.ty);
  c->tg.ntz = (int)((g.nzl + c->tg.tz - 1) / c->tg.tz);
  c->ntiles = (i64)c->tg.ntx * c->tg.nty * c->tg.ntz;
  c->sorted_np = 0;
  c->last_untiled = 0;
  CKB(CK(cudaMemsetAsync(c->mesh, 0, (size_t)g.plane * g.nplanes * 8, c->st)));
  CKB(CK(cudaStreamSynchronize(c->st)));
#undef CKB
  *out = c;
  return NNPM_OK;
}

extern "C" int nnpm_destroy(nnpm_ctx* ctx) {
  if (!ctx) return NNPM_OK;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->st) cudaStreamSynchronize(ctx->st);
  comm_destroy(ctx);
  if (ctx->have2) { cufftDestroy(ctx->plan2f); cufftDestroy(ctx->plan2b); }
  if (ctx->have1) { cufftDestroy(ctx->plan1f); cufftDestroy(ctx->plan1b); }
  for (int i = 0; i < 4; i++) if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
  if (ctx->havez) cufftDestroy(ctx->planz);
  if (ctx->have3) { cufftDestroy(ctx->plan3f); cufftDestroy(ctx->plan3b); }
  for (void* p : ctx->allocs) cudaFree(p);
  if (ctx->fftwork) cudaFree(ctx->fftwork);  // allocated outside ctx->allocs (grow_fftwork)
  delete ctx->peer_maps;
  if (ctx->h_red) cudaFreeHost(ctx->h_red);
  if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
  for (int i = 0; i < 16; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 8; i++) if (ctx->ev_fft[i]) cudaEventDestroy(ctx->ev_fft[i]);
  if (ctx->own_stream && ctx->st) cudaStreamDestroy(ctx->st);
  delete ctx;
  return NNPM_OK;
}

extern "C" int64_t nnpm_device_bytes(const nnpm_ctx* ctx) { return ctx ? (int64_t)ctx->dev_bytes : 0; }
extern "C" int64_t nnpm_local_count(const nnpm_ctx* ctx) { return ctx ? ctx->np : 0; }

extern "C" int nnpm_host_alloc(void** ptr, uint64_t bytes) {
  if (!ptr) return NNPM_ERR_ARG;
  cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 8, cudaHostAllocDefault);
  if (e != cudaSuccess) { g_create_err = cudaGetErrorString(e); return NNPM_ERR_CUDA; }
  return NNPM_OK;
}
extern "C" int nnpm_host_free(void* ptr) {
  if (ptr && cudaFreeHost(ptr) != cudaSuccess) return NNPM_ERR_CUDA;
  return NNPM_OK;
}

// ------------------------------------------------------------------------------------------
// particle upload / download
// ------------------------------------------------------------------------------------------
static int ensure_stage(nnpm_ctx* ctx) {
  if (ctx->stage) return NNPM_OK;
  i64 want = ctx->cap * ctx->R;
  const i64 maxe = 48ll << 20;  // 384 MB of doubles
  ctx->stage_elems = want < maxe ? want : maxe;
  if (ctx->stage_elems < 3) ctx->stage_elems = 3;
  return dalloc(ctx, &ctx->stage, (size_t)ctx->stage_elems);
}

static int upload_aos(nnpm_ctx* ctx, const double* aos, double* d0, double* d1, double* d2, i64 n) {
  const int R = ctx->R;
  const i64 chunk = ctx->stage_elems / R;
  for (i64 b = 0; b < n; b += chunk) {
    i64 cn = (n - b) < chunk ? (n - b) : chunk;
    CK(cudaMemcpyAsync(ctx->stage, aos + b * R, (size_t)cn * R * 8, cudaMemcpyHostToDevice, ctx->st));
    if (R == 3) k_aos2soa<3><<<nblk(cn, 256), 256, 0, ctx->st>>>(ctx->stage, d0 + b, d1 + b, d2 + b, cn);
    else k_aos2soa<2><<<nblk(cn, 256), 256, 0, ctx->st>>>(ctx->stage, d0 + b, d1 + b, nullptr, cn);
    LAUNCHED(ctx);
  }
  KCHECK();
  return NNPM_OK;
}

static int download_aos(nnpm_ctx* ctx, double* aos, const double* d0, const double* d1, const double* d2, i64 n) {
  const int R = ctx->R;
  const i64 chunk = ctx->stage_elems / R;
  for (i64 b = 0; b < n; b += chunk) {
    i64 cn = (n - b) < chunk ? (n - b) : chunk;
    if (R == 3) k_soa2aos<3><<<nblk(cn, 256), 256, 0, ctx->st>>>(ctx->stage, d0 + b, d1 + b, d2 + b, cn);
    else k_soa2aos<2><<<nblk(cn, 256), 256, 0, ctx->st>>>(ctx->stage, d0 + b, d1 + b, nullptr, cn);
    LAUNCHED(ctx);
    CK(cudaMemcpyAsync(aos + b * R, ctx->stage, (size_t)cn * R * 8, cudaMemcpyDeviceToHost, ctx->st));
    // the single staging buffer is reused by the next chunk: the stream orders kernel after copy
  }
  KCHECK();
  return NNPM_OK;
}

extern "C" int nnpm_upload_particles(nnpm_ctx* ctx, const double* x_aos, const double* v_aos, int64_t n,
                                     int64_t first_id, double m) {
  if (!ctx) return NNPM_ERR_ARG;
  if (n < 0 || (n > 0 && !x_aos)) return fail(ctx, NNPM_ERR_ARG, "bad particle array");
  if (n > ctx->cap) return fail(ctx, NNPM_ERR_CAPACITY, "upload of %lld particles exceeds capacity %lld", (long long)n, (long long)ctx->cap);
  CK(cudaSetDevice(ctx->cfg.device));
  CKR(ensure_stage(ctx));
  CKR(upload_aos(ctx, x_aos, ctx->pp.x[0], ctx->pp.x[1], ctx->pp.x[2], n));
  if (v_aos) CKR(upload_aos(ctx, v_aos, ctx->pp.v[0], ctx->pp.v[1], ctx->pp.v[2], n));
  else for (int d = 0; d < ctx->R; d++) CK(cudaMemsetAsync(ctx->pp.v[d], 0, (size_t)n * 8, ctx->st));
  if (n) { k_iota_u32<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->ids, n, (u64)first_id); LAUNCHED(ctx); KCHECK(); }
  ctx->np = n;
  ctx->m = m;
  ctx->first_id = first_id;
  ctx->sorted_np = 0;
  ctx->ids_dirty = false;
  ctx->acc_valid = ctx->pa_valid = false;
  CK(cudaStreamSynchronize(ctx->st));
  return NNPM_OK;
}

static int ensure_spare(nnpm_ctx* ctx) {
  if (!ctx->spare) CKR(dalloc(ctx, &ctx->spare, (size_t)ctx->cap + 2));
  if (!ctx->ids_spare) CKR(dalloc(ctx, &ctx->ids_spare, (size_t)ctx->cap));
  return NNPM_OK;
}

// apply out[dest[n]] = in[n] to the six particle arrays and the ids, rotating the spare buffer
static int permute_all(nnpm_ctx* ctx, const uint32_t* dest) {
  const i64 np = ctx->np;
  CKR(ensure_spare(ctx));
  for (int d = 0; d < ctx->R; d++) {
    k_permute<double><<<nblk(np, 256), 256, 0, ctx->st>>>(ctx->spare, ctx->pp.x[d], dest, np);
    LAUNCHED(ctx);
    std::swap(ctx->spare, ctx->pp.x[d]);
    k_permute<double><<<nblk(np, 256), 256, 0, ctx->st>>>(ctx->spare, ctx->pp.v[d], dest, np);
    LAUNCHED(ctx);
    std::swap(ctx->spare, ctx->pp.v[d]);
  }
  k_permute<uint32_t><<<nblk(np, 256), 256, 0, ctx->st>>>(ctx->ids_spare, ctx->ids, dest, np);
  LAUNCHED(ctx);
  std::swap(ctx->ids_spare, ctx->ids);
  KCHECK();
  return NNPM_OK;
}

extern "C" int nnpm_download_particles(nnpm_ctx* ctx, double* x_aos, double* v_aos, int64_t* ids_out) {
  if (!ctx) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  CKR(ensure_stage(ctx));
  const i64 np = ctx->np;
  if (!ids_out && ctx->P == 1 && ctx->ids_dirty && np) {
    // restore the reference's particle order: out[id[n] - first_id] = in[n]  (ids are a permutation of
    // first_id .. first_id + np - 1); permute_all rotates ids_spare, so the destination list lives in the sort scratch
    if (!ctx->skey) CKR(dalloc(ctx, &ctx->skey, (size_t)ctx->cap));
    k_ids_to_dest<<<nblk(np, 256), 256, 0, ctx->st>>>(ctx->skey, ctx->ids, (uint32_t)ctx->first_id, np);
    LAUNCHED(ctx);
    CKR(permute_all(ctx, ctx->skey));
    ctx->ids_dirty = false;
    ctx->sorted_np = 0;
  }
  if (x_aos) CKR(download_aos(ctx, x_aos, ctx->pp.x[0], ctx->pp.x[1], ctx->pp.x[2], np));
  if (v_aos) CKR(download_aos(ctx, v_aos, ctx->pp.v[0], ctx->pp.v[1], ctx->pp.v[2], np));
  if (ids_out && np) {
    i64* tmp = (i64*)ctx->stage;
    const i64 chunk = ctx->stage_elems;
    for (i64 b = 0; b < np; b += chunk) {
      i64 cn = (np - b) < chunk ? (np - b) : chunk;
      k_u32_to_i64<<<nblk(cn, 256), 256, 0, ctx->st>>>(tmp, ctx->ids + b, cn);
      LAUNCHED(ctx);
      CK(cudaMemcpyAsync(ids_out + b, tmp, (size_t)cn * 8, cudaMemcpyDeviceToHost, ctx->st));
    }
    KCHECK();
  }
  CK(cudaStreamSynchronize(ctx->st));
  return NNPM_OK;
}

extern "C" int nnpm_init_synthetic(nnpm_ctx* ctx, int kind, const int64_t pdims[3], uint64_t seed, int32_t nmodes,
                                   int32_t kmax, double disp_rms, double vfac, double m) {
  if (!ctx || !pdims) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  const i64 P1 = pdims[0], P2 = pdims[1], P3 = ctx->R == 3 ? pdims[2] : 1;
  if (P1 * P2 * P3 != ctx->cfg.npart_total) return fail(ctx, NNPM_ERR_ARG, "prod(pdims) != npart_total");
  if (kind != NNPM_IC_CLUSTERED && (nmodes > 64 || nmodes < 0)) return fail(ctx, NNPM_ERR_ARG, "nmodes must be in [0,64]");
  if (ctx->P > 1 && P3 % ctx->P) return fail(ctx, NNPM_ERR_ARG, "pdims[2] must be divisible by nranks");
  // this rank generates lattice planes [r*P3/P, (r+1)*P3/P)
  const i64 per = P3 / ctx->P * P1 * P2;
  const i64 first = per * ctx->r;
  if (per > ctx->cap) return fail(ctx, NNPM_ERR_CAPACITY, "synthetic ICs need capacity %lld", (long long)per);
  ICModes M;
  memset(&M, 0, sizeof M);
  if (kind == NNPM_IC_ZELDOVICH_MODES && nmodes > 0) {
    // splitmix64 stream; mirrored by oracle/pm_ref.py:synthetic_modes
    u64 s = seed;
    auto nxt = [&]() {
      s += 0x9E3779B97F4A7C15ull;
      u64 z = s;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      return z ^ (z >> 31);
    };
    if (kmax < 1) return fail(ctx, NNPM_ERR_ARG, "kmax must be >= 1");
    std::vector<int> kx, ky, kz;
    std::vector<double> ph;
    while ((int)kx.size() < nmodes) {
      int a = (int)(nxt() % (u64)(2 * kmax + 1)) - kmax;
      int b = (int)(nxt() % (u64)(2 * kmax + 1)) - kmax;
      int c = (int)(nxt() % (u64)(2 * kmax + 1)) - kmax;
      double p = (double)(nxt() >> 11) * (1.0 / 9007199254740992.0);
      if (a == 0 && b == 0 && c == 0) continue;
      kx.push_back(a); ky.push_back(b); kz.push_back(c); ph.push_back(p);
    }
    double norm2 = 0.0;
    for (int i = 0; i < nmodes; i++) {
      if (ctx->R == 2) { kz[i] = 0; if (kx[i] == 0 && ky[i] == 0) continue; }
      M.kx[M.n] = kx[i]; M.ky[M.n] = ky[i]; M.kz[M.n] = kz[i]; M.ph[M.n] = ph[i];
      norm2 += 0.5 / (double)(kx[i] * kx[i] + ky[i] * ky[i] + kz[i] * kz[i]);
      M.n++;
    }
    M.amp = M.n ? disp_rms * sqrt((double)ctx->R) / sqrt(norm2) : 0.0;
  }
  if (per && kind == NNPM_IC_CLUSTERED) {
    // number of blobs = nmodes * kmax (two int32 factors so that 10^5 and more fit), disp_rms = blob
    // sigma (box units), vfac = velocity dispersion
    const i64 nblobs = kmax > 0 ? (i64)nmodes * (i64)kmax : (i64)nmodes;
    if (nblobs < 1) return fail(ctx, NNPM_ERR_ARG, "clustered ICs need nmodes*kmax >= 1 blobs");
    if (ctx->R == 3) k_init_clustered<3><<<nblk(per, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, per, first, P1, P2, P3, seed, nblobs, disp_rms, vfac);
    else k_init_clustered<2><<<nblk(per, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, per, first, P1, P2, P3, seed, nblobs, disp_rms, vfac);
    LAUNCHED(ctx);
    KCHECK();
  } else if (per) {
    if (ctx->R == 3) k_init_synth<3><<<nblk(per, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, per, first, P1, P2, P3, M, vfac);
    else k_init_synth<2><<<nblk(per, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, per, first, P1, P2, P3, M, vfac);
    LAUNCHED(ctx);
    KCHECK();
  }
  ctx->np = per;
  ctx->m = m;
  ctx->first_id = 0;
  ctx->sorted_np = 0;
  ctx->ids_dirty = ctx->P > 1;
  ctx->acc_valid = ctx->pa_valid = false;
  CK(cudaStreamSynchronize(ctx->st));
  if (ctx->P > 1) return nnpm_redistribute(ctx);
  return NNPM_OK;
}

// ------------------------------------------------------------------------------------------
// K0 sort
// ------------------------------------------------------------------------------------------
// exclusive scan of n uint32 in place (up to SCAN_ITEMS^3 entries); bsum / bsum2 hold the block sums of the two levels
static int scan_u32(nnpm_ctx* ctx, uint32_t* data, i64 n) {
  const i64 nb1 = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
  const i64 nb2 = (nb1 + SCAN_ITEMS - 1) / SCAN_ITEMS;
  if (nb2 > SCAN_ITEMS) return fail(ctx, NNPM_ERR_ARG, "scan too large");
  if (nb1 > ctx->bsum_cap) {
    CKR(dalloc(ctx, &ctx->bsum, (size_t)nb1 + 16));   // (the previous, smaller buffer stays in ctx->allocs until destroy)
    ctx->bsum_cap = nb1;
  }
  if (!ctx->bsum2) CKR(dalloc(ctx, &ctx->bsum2, (size_t)SCAN_ITEMS + 16));
  if (nb1 == 1) {
    k_scan_apply<<<1, 256, 0, ctx->st>>>(data, nullptr, n);
    LAUNCHED(ctx);
  } else {
    k_scan_reduce<<<(unsigned)nb1, 256, 0, ctx->st>>>(data, ctx->bsum, n);
    if (nb2 == 1) {
      k_scan_apply<<<1, 256, 0, ctx->st>>>(ctx->bsum, nullptr, nb1);
      ctx->launches += 2;
    } else {
      k_scan_reduce<<<(unsigned)nb2, 256, 0, ctx->st>>>(ctx->bsum, ctx->bsum2, nb1);
      k_scan_apply<<<1, 256, 0, ctx->st>>>(ctx->bsum2, nullptr, nb2);
      k_scan_apply<<<(unsigned)nb2, 256, 0, ctx->st>>>(ctx->bsum, ctx->bsum2, nb1);
      ctx->launches += 4;
    }
    k_scan_apply<<<(unsigned)nb1, 256, 0, ctx->st>>>(data, ctx->bsum, n);
    LAUNCHED(ctx);
  }
  KCHECK();
  return NNPM_OK;
}

extern "C" int nnpm_sort(nnpm_ctx* ctx) {
  if (!ctx) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  const i64 np = ctx->np;
  if (np == 0) return NNPM_OK;
  if (!ctx->hist) CKR(dalloc(ctx, &ctx->hist, (size_t)ctx->ntiles + 1));
  if (!ctx->skey) CKR(dalloc(ctx, &ctx->skey, (size_t)ctx->cap));
  if (!ctx->srnk) CKR(dalloc(ctx, &ctx->srnk, (size_t)ctx->cap));
  // cell-order sort ("sort_cells", default): bins = (tile, cell in tile); needs a counter per mesh cell of the slab
  const int cpt = ctx->tg.tx * ctx->tg.ty * ctx->tg.tz;
  const i64 nbins = ctx->ntiles * (i64)cpt;
  const bool fine = ctx->opt_sort_cells && nbins <= (1ll << 30) && np < (1ll << 32);
  uint32_t* bins = ctx->hist;
  i64 nb = ctx->ntiles;
  if (fine) {
    if (!ctx->hist_fine) CKR(dalloc(ctx, &ctx->hist_fine, (size_t)nbins + 16));
    bins = ctx->hist_fine;
    nb = nbins;
  }
  CK(cudaMemsetAsync(bins, 0, (size_t)(nb + 1) * 4, ctx->st));
  if (ctx->R == 3) k_sort_count<3><<<nblk(np, 1024), 256, 0, ctx->st>>>(ctx->pp, np, ctx->g, ctx->tg, fine ? cpt : 0, bins, ctx->skey, ctx->srnk);
  else k_sort_count<2><<<nblk(np, 1024), 256, 0, ctx->st>>>(ctx->pp, np, ctx->g, ctx->tg, fine ? cpt : 0, bins, ctx->skey, ctx->srnk);
  LAUNCHED(ctx);
  CKR(scan_u32(ctx, bins, nb));   // bins[nb] is not needed by k_sort_dest
  k_sort_dest<<<nblk(np, 256), 256, 0, ctx->st>>>(ctx->skey, ctx->srnk, bins, np);
  LAUNCHED(ctx);
  if (fine) k_tile_offsets<<<nblk(ctx->ntiles + 1, 256), 256, 0, ctx->st>>>(ctx->hist, bins, ctx->ntiles, cpt, (uint32_t)np);
  else k_tile_offsets<<<nblk(ctx->ntiles + 1, 256), 256, 0, ctx->st>>>(ctx->hist, bins, ctx->ntiles, 1, (uint32_t)np);   // sets hist[ntiles] = np
  LAUNCHED(ctx);
  KCHECK();
  CKR(permute_all(ctx, ctx->skey));
  // work list: chunks of <= NNPM_CHUNK particles per tile along the banded raster
  if (!ctx->wcnt) {
    ctx->wcap = ctx->ntiles + ctx->cap / NNPM_CHUNK + 2;
    CKR(dalloc(ctx, &ctx->wcnt, (size_t)ctx->ntiles + 1));
    CKR(dalloc(ctx, &ctx->wtile, (size_t)ctx->wcap));
    CKR(dalloc(ctx, &ctx->wbeg, (size_t)ctx->wcap));
  }
  k_work_count<<<nblk(ctx->ntiles + 1, 256), 256, 0, ctx->st>>>(ctx->tg, ctx->ntiles, ctx->hist, ctx->wcnt);
  LAUNCHED(ctx);
  CKR(scan_u32(ctx, ctx->wcnt, ctx->ntiles + 1));
  k_work_fill<<<nblk(ctx->ntiles, 256), 256, 0, ctx->st>>>(ctx->tg, ctx->ntiles, ctx->hist, ctx->wcnt, ctx->wtile, ctx->wbeg);
  LAUNCHED(ctx);
  KCHECK();
  ctx->ids_dirty = true;
  ctx->pa_valid = false;
  ctx->sorted_np = np;
  return NNPM_OK;
}

// ------------------------------------------------------------------------------------------
// operators
// ------------------------------------------------------------------------------------------
extern "C" int nnpm_make_periodic(nnpm_ctx* ctx) {
  if (!ctx) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  for (int d = 0; d < ctx->R && ctx->np; d++) { k_make_periodic<<<nblk(ctx->np, 256), 256, 0, ctx->st>>>(ctx->pp.x[d], ctx->np); LAUNCHED(ctx); }
  KCHECK();
  return NNPM_OK;
}
extern "C" int nnpm_drift(nnpm_ctx* ctx, double dt) {
  if (!ctx) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  for (int d = 0; d < ctx->R && ctx->np; d++) { k_axpy_rn<<<nblk(ctx->np, 256), 256, 0, ctx->st>>>(ctx->pp.x[d], ctx->pp.v[d], dt, ctx->np); LAUNCHED(ctx); }
  KCHECK();
  return NNPM_OK;
}
extern "C" int nnpm_kick(nnpm_ctx* ctx, double dt) {
  if (!ctx) return NNPM_ERR_ARG;
  if (!ctx->pa_valid) return fail(ctx, NNPM_ERR_STATE, "nnpm_kick needs pa: call nnpm_interp or nnpm_set_field(PA) first");
  CK(cudaSetDevice(ctx->cfg.device));
  for (int d = 0; d < ctx->R && ctx->np; d++) { k_axpy_rn<<<nblk(ctx->np, 256), 256, 0, ctx->st>>>(ctx->pp.v[d], ctx->pa[d], dt, ctx->np); LAUNCHED(ctx); }
  KCHECK();
  return NNPM_OK;
}
extern "C" int nnpm_kick_comoving(nnpm_ctx* ctx, double dt, double a, double dadt) {
  if (!ctx) return NNPM_ERR_ARG;
  if (!ctx->pa_valid) return fail(ctx, NNPM_ERR_STATE, "nnpm_kick_comoving needs pa: call nnpm_interp first");
  CK(cudaSetDevice(ctx->cfg.device));
  const double hub = dadt / a;  // (dadt/a), ParticleMesh.jl:806
  for (int d = 0; d < ctx->R && ctx->np; d++) { k_kick_comoving<<<nblk(ctx->np, 256), 256, 0, ctx->st>>>(ctx->pp.v[d], ctx->pa[d], dt, a, hub, ctx->np); LAUNCHED(ctx); }
  KCHECK();
  return NNPM_OK;
}

template <int RANK, int INTERP, bool FIXED>
static void launch_deposit(nnpm_ctx* ctx, bool pre, double c0, u64* viol, i64 base) {
  const i64 cnt = base < 0 ? 0 : ctx->np - base;
  const double fscale = ldexp(1.0, ctx->cfg.fixed_bits);
  PartPtrs q = ctx->pp;
  for (int d = 0; d < ctx->R && base > 0; d++) { q.x[d] += base; q.v[d] += base; }
  const bool stragglers = base < 0;  // the tiled kernel's device-side list instead of a range
  const uint32_t* list = stragglers ? ctx->srnk : nullptr;
  const u64* lcount = (const u64*)(ctx->d_red + 5);
  const unsigned grid = stragglers ? (unsigned)ctx->nsm * 4 : nblk(cnt, 256);
  if (stragglers) q = ctx->pp;
  if (pre) k_deposit<RANK, INTERP, FIXED, true><<<grid, 256, 0, ctx->st>>>(q, cnt, list, lcount, c0, ctx->g, ctx->mesh, ctx->m, fscale, viol);
  else k_deposit<RANK, INTERP, FIXED, false><<<grid, 256, 0, ctx->st>>>(q, cnt, list, lcount, c0, ctx->g, ctx->mesh, ctx->m, fscale, viol);
}

// zero + scatter + ghost fold (+ fixed -> f64).  pre: apply x + c0*v on the fly.
static int do_deposit(nnpm_ctx* ctx, bool pre, double c0) {
  const Geom& g = ctx->g;
  const bool fixed = ctx->cfg.deposit_mode == NNPM_DEPOSIT_FIXED;
  const int interp = ctx->cfg.deposit;
  CK(cudaMemsetAsync(ctx->mesh, 0, (size_t)g.plane * g.nplanes * 8, ctx->st));
  LAUNCHED(ctx);
  u64* viol = (u64*)(ctx->d_red + 3);
  const double fscale = ldexp(1.0, ctx->cfg.fixed_bits);
  // tile-sorted particles [0, covered): shared-memory tiled scatter; the rest: global atomics
  i64 covered = 0;
  if (interp == NNPM_INTERP_CIC && ctx->opt_tile_deposit && ctx->sorted_np > 0 && ctx->np > 0) {
    covered = ctx->sorted_np < ctx->np ? ctx->sorted_np : ctx->np;
    CK(cudaMemsetAsync(ctx->d_red + 5, 0, 8, ctx->st));
    const int bulk = ctx->opt_bulk_flush && (g.N1 % 2 == 0) && g.N1 >= NNPM_DX && ctx->tg.tx == NNPM_TX;
    const WorkList wl = {ctx->wtile, ctx->wbeg, ctx->wcnt + ctx->ntiles};
    const unsigned wgrid = (unsigned)(ctx->ntiles + ctx->sorted_np / NNPM_CHUNK + 1);  // >= number of chunks
#define DT(RK, FX, PR)                                                                                                   \
  k_deposit_tile<RK, FX, PR, 2><<<wgrid, 256, 0, ctx->st>>>(ctx->pp, covered, c0, ctx->g, ctx->tg, wl, ctx->hist, ctx->mesh, ctx->m, fscale, ctx->srnk, (u64*)(ctx->d_red + 5), bulk, ctx->opt_l2pf_deposit)
    if (ctx->R == 3) { if (fixed) { if (pre) DT(3, true, true); else DT(3, true, false); } else { if (pre) DT(3, false, true); else DT(3, false, false); } }
    else { if (fixed) { if (pre) DT(2, true, true); else DT(2, true, false); } else { if (pre) DT(2, false, true); else DT(2, false, false); } }
#undef DT
    LAUNCHED(ctx);
    // stragglers (device-side count) through the global-atomic kernel
    if (ctx->R == 3) { if (fixed) launch_deposit<3, 1, true>(ctx, pre, c0, viol, -1); else launch_deposit<3, 1, false>(ctx, pre, c0, viol, -1); }
    else { if (fixed) launch_deposit<2, 1, true>(ctx, pre, c0, viol, -1); else launch_deposit<2, 1, false>(ctx, pre, c0, viol, -1); }
    LAUNCHED(ctx);
    KCHECK();
  }
  if (ctx->np > covered) {
#define DEP(RK, IT)                                                                     \
  do {                                                                                  \
    if (fixed) launch_deposit<RK, IT, true>(ctx, pre, c0, viol, covered);                \
    else launch_deposit<RK, IT, false>(ctx, pre, c0, viol, covered);                     \
  } while (0)
    if (ctx->R == 3) { if (interp) DEP(3, 1); else DEP(3, 0); }
    else { if (interp) DEP(2, 1); else DEP(2, 0); }
#undef DEP
    LAUNCHED(ctx);
    KCHECK();
  }
  if (ctx->P > 1) CKR(comm_fold_rho_ghosts(ctx, fixed));
  if (fixed) {
    const i64 n = g.plane * g.nzl;
    k_fixed_to_f64<<<nblk(n, 256), 256, 0, ctx->st>>>(ctx->mesh + g.plane * g.glo, n, ldexp(1.0, -ctx->cfg.fixed_bits), ctx->m);
    LAUNCHED(ctx);
    KCHECK();
  }
  ctx->field_state = 1;
  ctx->acc_valid = false;
  return NNPM_OK;
}

extern "C" int nnpm_deposit(nnpm_ctx* ctx) {
  if (!ctx) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemsetAsync(ctx->d_red + 3, 0, 8, ctx->st));
  return do_deposit(ctx, false, 0.0);
}

// x pass over (a sub-range of) the owned planes: own row-FFT kernels, or cuFFT's batched 1-D plans
static int fft_x(nnpm_ctx* ctx, bool inverse) {
  const Geom& g = ctx->g;
  double* own = ctx->mesh + g.plane * g.glo;
  if (ctx->opt_rowfft && rowfft_supported_n(g.N1)) return rowfft_x(ctx, inverse);
  if (inverse) CKF(cufftExecZ2D(ctx->plan1b, (cufftDoubleComplex*)own, own));
  else CKF(cufftExecD2Z(ctx->plan1f, own, (cufftDoubleComplex*)own));
  LAUNCHED(ctx);
  return NNPM_OK;
}

// the fused z pass [z-FFT, Green's function, z-iFFT] over the local rows [jlbeg, jlend) of T
static int fft_z_green(nnpm_ctx* ctx, double2* T, double scale, i64 jlbeg, i64 jlend, const ZDst* dst) {
  if (ctx->opt_zfused == 3 && !zfused_p_two(ctx)) return zfused_s_launch(ctx, T, scale, jlbeg, jlend, dst);
  return zfused_p_launch(ctx, T, scale, jlbeg, jlend, dst);
}

// forward transform of the owned planes, in place: rho -> spectrum.  Returns the spectrum
// pointer T with layout [k][jl][ic] (jl = local j range of this rank; all of j when P == 1).
// fuse (nranks > 1): the y pass stores straight into the peers' transposed spectra (no separate all-to-all).
static int fft_forward(nnpm_ctx* ctx, double2** Tout, bool z_too, cudaEvent_t* comm_t, bool fuse) {
  const Geom& g = ctx->g;
  double* own = ctx->mesh + g.plane * g.glo;
  if (ctx->opt_fft3d && ctx->P == 1 && g.N3 > 1) {
    CKR(ensure_plan3(ctx));
    CKF(cufftExecD2Z(ctx->plan3f, own, (cufftDoubleComplex*)own));
    LAUNCHED(ctx);
    *Tout = (double2*)own;
    return NNPM_OK;
  }
  const bool tim = comm_t != nullptr;  // the fused step passes its event array when timing = 1
  if (ctx->opt_fft2d && !fuse && ctx->opt_colfft && ctx->opt_rowfft && fft2d_supported(ctx)) {
    CKR(fft2d_xy(ctx, false));   // x and y passes in one kernel, intermediate in L2
  } else if (ctx->opt_colfft && colfft_supported_n(g.N2)) {
    if (tim) CK(cudaEventRecord(ctx->ev_fft[0], ctx->st));
    CKR(fft_x(ctx, false));
    if (tim) CK(cudaEventRecord(ctx->ev_fft[1], ctx->st));
    CKR(colfft_y(ctx, false, 0, -1, fuse));
    if (tim) CK(cudaEventRecord(ctx->ev_fft[2], ctx->st));
  } else {
    CKF(cufftExecD2Z(ctx->plan2f, own, (cufftDoubleComplex*)own));
    LAUNCHED(ctx);
  }
  double2* T = (double2*)own;
  if (ctx->P > 1) {
    if (comm_t) CK(cudaEventRecord(comm_t[0], ctx->st));
    if (fuse) CKR(comm_barrier(ctx));   // every rank's remote stores have landed
    else CKR(comm_transpose_fwd(ctx));
    if (comm_t) CK(cudaEventRecord(comm_t[1], ctx->st));
    T = (double2*)ctx->meshT;
  }
  if (z_too && ctx->havez) { CKF(cufftExecZ2Z(ctx->planz, (cufftDoubleComplex*)T, (cufftDoubleComplex*)T, CUFFT_FORWARD)); LAUNCHED(ctx); }
  *Tout = T;
  return NNPM_OK;
}

// fused_bwd: the z pass already stored its output into the owners' slabs; only the barrier remains
static int fft_inverse(nnpm_ctx* ctx, bool z_too, cudaEvent_t* comm_t, bool fused_bwd) {
  const Geom& g = ctx->g;
  double* own = ctx->mesh + g.plane * g.glo;
  if (ctx->opt_fft3d && ctx->P == 1 && g.N3 > 1) {
    CKF(cufftExecZ2D(ctx->plan3b, (cufftDoubleComplex*)own, own));
    LAUNCHED(ctx);
    return NNPM_OK;
  }
  double2* T = ctx->P > 1 ? (double2*)ctx->meshT : (double2*)own;
  if (z_too && ctx->havez) { CKF(cufftExecZ2Z(ctx->planz, (cufftDoubleComplex*)T, (cufftDoubleComplex*)T, CUFFT_INVERSE)); LAUNCHED(ctx); }
  if (ctx->P > 1) {
    if (comm_t) CK(cudaEventRecord(comm_t[2], ctx->st));
    if (fused_bwd) CKR(comm_barrier(ctx));
    else CKR(comm_transpose_bwd(ctx));
    if (comm_t) CK(cudaEventRecord(comm_t[3], ctx->st));
  }
  if (ctx->opt_fft2d && ctx->opt_colfft && ctx->opt_rowfft && fft2d_supported(ctx)) {
    CKR(fft2d_xy(ctx, true));
  } else if (ctx->opt_colfft && colfft_supported_n(g.N2)) {
    const bool tim = comm_t != nullptr;
    if (tim) CK(cudaEventRecord(ctx->ev_fft[4], ctx->st));
    CKR(colfft_y(ctx, true));
    if (tim) CK(cudaEventRecord(ctx->ev_fft[5], ctx->st));
    CKR(fft_x(ctx, true));
    if (tim) CK(cudaEventRecord(ctx->ev_fft[6], ctx->st));
  } else {
    CKF(cufftExecZ2D(ctx->plan2b, (cufftDoubleComplex*)own, own));
    LAUNCHED(ctx);
  }
  return NNPM_OK;
}

// Un-fused multi-GPU solve with the copy-engine transposes overlapped with the transforms ("fuse_transpose" = 0,
// "pipeline" >= 1): the owned planes are transformed in NCH chunks, and as soon as a chunk's 2-D transform is done its
// block of the forward transpose goes out on the copy engines (side streams) while the next chunk is computed; the
// fused z pass is chunked over jl the same way for the backward transpose.
static int do_solve_pipelined(nnpm_ctx* ctx, cudaEvent_t* ev, double scale) {
  const Geom& g = ctx->g;
  const int NCH = ctx->pipe_nch;
  const i64 pc = g.nzl / NCH, jc = ctx->njl / NCH;
  if (!ctx->ev_chunk[0])
    for (int i = 0; i < 4; i++) CK(cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
  for (int c = 0; c < NCH; c++) {
    CKR(rowfft_x(ctx, false, pc * c, pc));
    CKR(colfft_y(ctx, false, pc * c, pc));
    CK(cudaEventRecord(ctx->ev_chunk[c], ctx->st));
    CKR(comm_chunk_ce(ctx, true, c, NCH, ctx->ev_chunk[c]));
  }
  if (ev) CK(cudaEventRecord(ev[0], ctx->st));
  CKR(comm_join_barrier(ctx));
  if (ev) { CK(cudaEventRecord(ev[1], ctx->st)); CK(cudaEventRecord(ev[4], ctx->st)); }
  for (int c = 0; c < NCH; c++) {
    CKR(fft_z_green(ctx, (double2*)ctx->meshT, scale, jc * c, jc * (c + 1), nullptr));
    CK(cudaEventRecord(ctx->ev_chunk[c], ctx->st));
    CKR(comm_chunk_ce(ctx, false, c, NCH, ctx->ev_chunk[c]));
  }
  if (ev) { CK(cudaEventRecord(ev[5], ctx->st)); CK(cudaEventRecord(ev[2], ctx->st)); }
  CKR(comm_join_barrier(ctx));
  if (ev) CK(cudaEventRecord(ev[3], ctx->st));
  CKR(colfft_y(ctx, true));
  CKR(fft_x(ctx, true));
  return NNPM_OK;
}

// compute_potential, ParticleMesh.jl:417-505: rho (in the mesh) -> phi (in the mesh).
// ev (optional, 8 events): [0..3] around the transposes (fused: around the barriers that close them), [4],[5]
// around the z pass.
static int do_solve(nnpm_ctx* ctx, cudaEvent_t* ev) {
  const Geom& g = ctx->g;
  const double scale = 1.0 / ((double)g.N1 * (double)g.N2 * (double)g.N3) * 1 / ((double)g.N1 * (double)g.N1);
  const bool fft3d = ctx->opt_fft3d && ctx->P == 1 && g.N3 > 1;
  const bool fusedz = !fft3d && ctx->opt_zfused && g.N3 > 1 && zfused_supported(ctx);
  const bool ycol = ctx->opt_colfft && colfft_supported_n(g.N2);
  if (ctx->P > 1) NEED_COMM();
  // NVLink transposes fused into the transform kernels: forward into the y pass' tensor stores (needs the TMA
  // column engine), backward into the z pass' register stores (needs the fused z pass and a power-of-two slab)
  ZDst zd;
  const bool fuse = NNPM_EXP_PEER && ctx->P > 1 && ctx->opt_fuse_transpose && ctx->peer_ok && ctx->P <= 8;
  const bool fuse_fwd = fuse && ycol;
  const bool fuse_bwd = fuse && fusedz && zdst_peers(ctx, &zd);
  if (fuse_fwd) CKR(colfft_peer_maps(ctx));
  // chunked copy-engine pipeline (un-fused path only).  Measured at 1024^3: P = 2, four chunks: solve 17.5 -> 15.5 ms
  // (round 1); P = 8: two chunks on eight side streams 11.63 -> 11.25 ms per step, four chunks 11.51, eight streams
  // without chunking 12.40 (profiles/bench_r02_variants_n8.txt).  Automatic: four chunks at P = 2, two beyond.
  ctx->ce_ns = ctx->opt_ce_streams > 0 ? ctx->opt_ce_streams : 4;
  const int nch = ctx->opt_pipe_chunks > 0 ? ctx->opt_pipe_chunks : (ctx->P == 2 ? 4 : 2);
  const bool pipe = ctx->opt_pipeline >= 1;
  if (ctx->P > 1 && !fuse && pipe && ctx->peer_ok && ctx->opt_peer >= 1 && fusedz && ycol && rowfft_supported_n(g.N1) && ctx->opt_rowfft &&
      g.nzl % nch == 0 && ctx->njl % nch == 0 && g.nzl / nch >= 1) {
    ctx->pipe_nch = nch;
    ctx->ce_ns = ctx->opt_ce_streams > 0 ? ctx->opt_ce_streams : (ctx->P > 4 ? 8 : 4);
    CKR(do_solve_pipelined(ctx, ev, scale));
    CKR(comm_fill_phi_ghosts(ctx));
    ctx->field_state = 2;
    ctx->acc_valid = false;
    return NNPM_OK;
  }
  double2* T = nullptr;
  CKR(fft_forward(ctx, &T, !fusedz, ev, fuse_fwd));
  if (ev) CK(cudaEventRecord(ev[4], ctx->st));
  if (fusedz) {
    CKR(fft_z_green(ctx, T, scale, 0, ctx->njl, fuse_bwd ? &zd : nullptr));
  } else {
    dim3 grid(nblk(ctx->icn, 128), (unsigned)ctx->njl, (unsigned)(g.N3 < 64 ? g.N3 : 64));
    k_green<<<grid, 128, 0, ctx->st>>>(T, ctx->icn, ctx->njl, ctx->njl * ctx->r, g.N3, ctx->s1, ctx->s2, ctx->s3, ctx->R == 3, scale);
    LAUNCHED(ctx);
    KCHECK();
  }
  if (ev && fusedz) CK(cudaEventRecord(ev[5], ctx->st));
  CKR(fft_inverse(ctx, !fusedz, ev, fuse_bwd));
  if (ev && !fusedz) CK(cudaEventRecord(ev[5], ctx->st));
  if (ctx->P > 1) CKR(comm_fill_phi_ghosts(ctx));
  ctx->field_state = 2;
  ctx->acc_valid = false;
  return NNPM_OK;
}

extern "C" int nnpm_solve(nnpm_ctx* ctx) {
  if (!ctx) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  if (ctx->field_state != 1) return fail(ctx, NNPM_ERR_STATE, "nnpm_solve needs rho in the mesh (deposit or set_field(RHO) first)");
  return do_solve(ctx, nullptr);
}

static int ensure_acc(nnpm_ctx* ctx) {
  if (ctx->acc) return NNPM_OK;
  const Geom& g = ctx->g;
  ctx->acc_nq = g.nzl + (ctx->P > 1 ? 2 : 0);
  ctx->acc_k0 = g.kz0 - (ctx->P > 1 ? 1 : 0);
  return dalloc(ctx, &ctx->acc, (size_t)3 * g.N1 * g.N2 * ctx->acc_nq);
}
static int ensure_pa(nnpm_ctx* ctx) {
  for (int d = 0; d < ctx->R; d++)
    if (!ctx->pa[d]) CKR(dalloc(ctx, &ctx->pa[d], (size_t)ctx->cap));
  return NNPM_OK;
}

extern "C" int nnpm_accel(nnpm_ctx* ctx) {
  if (!ctx) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  if (ctx->field_state != 2) return fail(ctx, NNPM_ERR_STATE, "nnpm_accel needs phi in the mesh (solve or set_field(PHI) first)");
  CKR(ensure_acc(ctx));
  const Geom& g = ctx->g;
  dim3 grid(nblk(g.N1, 128), (unsigned)g.N2, (unsigned)ctx->acc_nq);
  if (ctx->R == 3) k_accel<3><<<grid, 128, 0, ctx->st>>>(ctx->acc, ctx->mesh, g, ctx->acc_k0, ctx->acc_nq);
  else k_accel<2><<<grid, 128, 0, ctx->st>>>(ctx->acc, ctx->mesh, g, ctx->acc_k0, ctx->acc_nq);
  LAUNCHED(ctx);
  KCHECK();
  ctx->acc_valid = true;
  return NNPM_OK;
}

extern "C" int nnpm_interp(nnpm_ctx* ctx) {
  if (!ctx) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  if (!ctx->acc_valid) return fail(ctx, NNPM_ERR_STATE, "nnpm_interp needs acc (nnpm_accel or set_field(ACC) first)");
  CKR(ensure_pa(ctx));
  const i64 np = ctx->np;
  const int bugz = ctx->cfg.bugcompat_interp_z;
  if (np) {
#define ITP(RK, IT) k_interp_acc<RK, IT><<<nblk(np, 256), 256, 0, ctx->st>>>(ctx->pa[0], ctx->pa[1], ctx->pa[2], ctx->acc, ctx->pp, np, ctx->g, ctx->acc_k0, ctx->acc_nq, bugz)
    if (ctx->R == 3) { if (ctx->cfg.backinterp) ITP(3, 1); else ITP(3, 0); }
    else { if (ctx->cfg.backinterp) ITP(2, 1); else ITP(2, 0); }
#undef ITP
    LAUNCHED(ctx);
    KCHECK();
  }
  ctx->pa_valid = true;
  return NNPM_OK;
}

// ------------------------------------------------------------------------------------------
// field get / set (Julia layout, local slab)
// ------------------------------------------------------------------------------------------
extern "C" int nnpm_get_field(nnpm_ctx* ctx, int which, double* host_out) {
  if (!ctx || !host_out) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  const Geom& g = ctx->g;
  if (which == NNPM_FIELD_RHO || which == NNPM_FIELD_PHI) {
    const double* src = ctx->mesh + g.plane * g.glo;
    if (which == NNPM_FIELD_RHO && ctx->field_state != 1 && ctx->rho_keep_valid) src = ctx->rho_keep;  // kept by the last fused step
    else if (ctx->field_state != (which == NNPM_FIELD_RHO ? 1 : 2))
      return fail(ctx, NNPM_ERR_STATE, "mesh does not currently hold %s%s", which == NNPM_FIELD_RHO ? "rho" : "phi",
                  which == NNPM_FIELD_RHO ? " (the solve consumed it; set option keep_rho before the step)" : "");
    CK(cudaMemcpy2DAsync(host_out, (size_t)g.N1 * 8, src, (size_t)g.rowp * 8, (size_t)g.N1 * 8,
                         (size_t)g.N2 * g.nzl, cudaMemcpyDeviceToHost, ctx->st));
  } else if (which == NNPM_FIELD_ACC) {
    if (!ctx->acc_valid) return fail(ctx, NNPM_ERR_STATE, "acc not computed (call nnpm_accel)");
    const i64 off = 3 * g.N1 * g.N2 * (g.kz0 - ctx->acc_k0);
    CK(cudaMemcpyAsync(host_out, ctx->acc + off, (size_t)3 * g.N1 * g.N2 * g.nzl * 8, cudaMemcpyDeviceToHost, ctx->st));
  } else if (which == NNPM_FIELD_PA) {
    if (!ctx->pa_valid) return fail(ctx, NNPM_ERR_STATE, "pa not computed (call nnpm_interp)");
    CKR(ensure_stage(ctx));
    CKR(download_aos(ctx, host_out, ctx->pa[0], ctx->pa[1], ctx->pa[2], ctx->np));
  } else {
    return fail(ctx, NNPM_ERR_ARG, "unknown field %d", which);
  }
  CK(cudaStreamSynchronize(ctx->st));
  return NNPM_OK;
}

extern "C" int nnpm_set_field(nnpm_ctx* ctx, int which, const double* host_in) {
  if (!ctx || !host_in) return NNPM_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  const Geom& g = ctx->g;
  if (which == NNPM_FIELD_RHO || which == NNPM_FIELD_PHI) {
    CK(cudaMemsetAsync(ctx->mesh, 0, (size_t)g.plane * g.nplanes * 8, ctx->st));
    CK(cudaMemcpy2DAsync(ctx->mesh + g.plane * g.glo, (size_t)g.rowp * 8, host_in, (size_t)g.N1 * 8, (size_t)g.N1 * 8,
                         (size_t)g.N2 * g.nzl, cudaMemcpyHostToDevice, ctx->st));
    ctx->field_state = which == NNPM_FIELD_RHO ? 1 : 2;
    ctx->acc_valid = false;
    if (which == NNPM_FIELD_PHI && ctx->P > 1) CKR(comm_fill_phi_ghosts(ctx));
  } else if (which == NNPM_FIELD_ACC) {
    if (ctx->P > 1) return fail(ctx, NNPM_ERR_ARG, "set_field(ACC) needs nranks == 1");
    CKR(ensure_acc(ctx));
    CK(cudaMemcpyAsync(ctx->acc, host_in, (size_t)3 * g.N1 * g.N2 * g.nzl * 8, cudaMemcpyHostToDevice, ctx->st));
    ctx->acc_valid = true;
  } else if (which == NNPM_FIELD_PA) {
    CKR(ensure_pa(ctx));
    CKR(ensure_stage(ctx));
    CKR(upload_aos(ctx, host_in, ctx->pa[0], ctx->pa[1], ctx->pa[2], ctx->np));
    ctx->pa_valid = true;
  } else {
    return fail(ctx, NNPM_ERR_ARG, "unknown field %d", which);
  }
  CK(cudaStreamSynchronize(ctx->st));
  return NNPM_OK;
}

// ------------------------------------------------------------------------------------------
// fused steps
// ------------------------------------------------------------------------------------------
__global__ void k_reset_red(i64* red) {
  if (threadIdx.x < 3) red[threadIdx.x] = (i64)0x8000000000000000ull;
  if (threadIdx.x == 3 || threadIdx.x == 4) red[threadIdx.x] = 0;
}

static inline double ord_decode(i64 k) {
  i64 b = k >= 0 ? k : (k ^ 0x7FFFFFFFFFFFFFFFLL);
  double d;
  memcpy(&d, &b, 8);
  return d;
}

template <int RANK, int INTERP>
static int launch_push(nnpm_ctx* ctx, bool comoving, const PushParams& pp) {
  const i64 np = ctx->np;
  u64* viol = (u64*)(ctx->d_red + 3);
  u64* slow_count = (u64*)(ctx->d_red + 4);
  auto global_push = [&](i64 base, i64 count, const uint32_t* list, unsigned grid) {
    if (comoving) k_push<RANK, INTERP, true><<<grid, 256, 0, ctx->st>>>(ctx->pp, base, count, list, slow_count, ctx->g, ctx->mesh, pp, ctx->d_red, viol);
    else k_push<RANK, INTERP, false><<<grid, 256, 0, ctx->st>>>(ctx->pp, base, count, list, slow_count, ctx->g, ctx->mesh, pp, ctx->d_red, viol);
    LAUNCHED(ctx);
  };
  const bool tiled = INTERP == 1 && ctx->opt_tile_push && ctx->sorted_np > 0 && !ctx->cfg.bugcompat_interp_z;
  if (!tiled) {
    global_push(0, np, nullptr, nblk(np, 256));
    KCHECK();
    return NNPM_OK;
  }
  // one CTA per work item, phi box staged by TMA, x / v by register loads behind a bulk L2 prefetch.  The variants
  // measured slower (round 1: DMA-warp particle streams, persistent double-buffered boxes, deeper register prefetch,
  // profiles/r01_probe_logs; round 2: a streamed separable gather at 3 / 4 CTAs per SM, 26.3 / 32.6 ms against 25.9 on
  // tile-sorted particles and 24.1 against 23.5 - 24.0 on cell-sorted ones: occupancy is not the lever,
  // profiles/r02_probe_logs) are no longer compiled.
  const size_t smem = (size_t)NNPM_BX * NNPM_BY * (RANK == 3 ? NNPM_BZ : 1) * sizeof(double);
  const i64 covered = ctx->sorted_np < np ? ctx->sorted_np : np;
  const WorkList wl = {ctx->wtile, ctx->wbeg, ctx->wcnt + ctx->ntiles};
  const unsigned wgrid = (unsigned)(ctx->ntiles + ctx->sorted_np / NNPM_CHUNK + 1);  // >= number of chunks
#define LT(CM)                                                                                                 \
  do {                                                                                                         \
    CK(cudaFuncSetAttribute(k_push_tile<RANK, CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    k_push_tile<RANK, CM><<<wgrid, 256, smem, ctx->st>>>(ctx->tmap_phi, ctx->pp, covered, ctx->g, ctx->tg, wl, ctx->hist, \
                                                         ctx->mesh, pp, ctx->d_red, ctx->srnk, slow_count, ctx->opt_l2pf_push); \
  } while (0)
  if (comoving) LT(true); else LT(false);
#undef LT
  LAUNCHED(ctx);
  global_push(0, 0, ctx->srnk, (unsigned)ctx->nsm * 4);         // stragglers (device-side count)
  if (np > covered) global_push(covered, np - covered, nullptr, nblk(np - covered, 256));  // arrivals since the sort
  KCHECK();
  return NNPM_OK;
}

static int do_step(nnpm_ctx* ctx, bool comoving, double dt, double a_d1, double a_k, double adot_k, double a_d2,
                   nnpm_stats* stats) {
  CK(cudaSetDevice(ctx->cfg.device));
  const bool tim = ctx->cfg.timing != 0;
  const int launches0 = ctx->launches;
  cudaEvent_t* ev = ctx->ev;
  // counting sort on the configured cadence, or early when > 1/16 of the particles left their tile box
  const bool do_sort = ctx->cfg.sort_every > 0 && ctx->np > 0 &&
                       ((ctx->step_count % ctx->cfg.sort_every) == 0 || ctx->last_untiled * 16 > ctx->np);
  if (tim) CK(cudaEventRecord(ev[0], ctx->st));
  if (do_sort) CKR(nnpm_sort(ctx));
  if (tim) CK(cudaEventRecord(ev[1], ctx->st));
  k_reset_red<<<1, 32, 0, ctx->st>>>(ctx->d_red);
  LAUNCHED(ctx);
  // reference scalars, evaluated exactly as the Julia expressions
  PushParams pp;
  pp.dt = dt;
  pp.bugz = ctx->cfg.bugcompat_interp_z;
  if (comoving) {
    pp.c0 = dt / 2 / a_d1;   // drift(x,v,dt/2/a)  ParticleMesh.jl:854
    pp.c1 = 0.0;
    pp.c2 = dt / 2 / a_d2;   // :867
    pp.a = a_k;
    pp.hub = adot_k / a_k;   // (dadt/a)  :806
  } else {
    pp.c0 = 0.0;
    pp.c1 = dt / 2;          // drift(x,v,dt/2)  :776,:778
    pp.c2 = 0.0;
    pp.a = 1.0;
    pp.hub = 0.0;
  }
  // nranks > 1: the push packs this step's leavers itself (mig_emit); the counters must be clean before it runs
  const bool mig_in_push = ctx->P > 1 && ctx->comm && ctx->R == 3 && ctx->opt_mig_in_push;
  memset(&pp.mg, 0, sizeof pp.mg);
  if (mig_in_push) { CKR(migrate_reset(ctx)); pp.mg = migrate_args(ctx); }
  CKR(do_deposit(ctx, comoving, pp.c0));
  ctx->rho_keep_valid = false;
  if (ctx->opt_keep_rho) {  // gd[1].d["ρD"] of the reference holds this deposit until the next one (ParticleMesh.jl:767, :857)
    const size_t n = (size_t)ctx->g.plane * ctx->g.nzl;
    if (!ctx->rho_keep) CKR(dalloc(ctx, &ctx->rho_keep, n));
    CK(cudaMemcpyAsync(ctx->rho_keep, ctx->mesh + ctx->g.plane * ctx->g.glo, n * 8, cudaMemcpyDeviceToDevice, ctx->st));
    ctx->rho_keep_valid = true;
  }
  if (tim) CK(cudaEventRecord(ev[2], ctx->st));
  CKR(do_solve(ctx, tim ? ev + 8 : nullptr));
  if (tim) CK(cudaEventRecord(ev[3], ctx->st));
  if (ctx->np) {
#define PSH(RK, IT) CKR((launch_push<RK, IT>(ctx, comoving, pp)))
    if (ctx->R == 3) { if (ctx->cfg.backinterp) PSH(3, 1); else PSH(3, 0); }
    else { if (ctx->cfg.backinterp) PSH(2, 1); else PSH(2, 0); }
#undef PSH
  }
  if (tim) CK(cudaEventRecord(ev[4], ctx->st));
  ctx->last_migrated = 0;
  if (ctx->P > 1) CKR(comm_migrate(ctx, mig_in_push));
  if (tim) CK(cudaEventRecord(ev[5], ctx->st));
  CK(cudaMemcpyAsync(ctx->h_red, ctx->d_red, 5 * sizeof(i64), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  ctx->step_count++;
  ctx->last_untiled = ctx->h_red[4] + (ctx->sorted_np > 0 && ctx->np > ctx->sorted_np ? ctx->np - ctx->sorted_np : 0);
  ctx->ids_dirty = ctx->ids_dirty || ctx->P > 1;
  ctx->pa_valid = false;
  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->max_abs_v = ctx->np ? ord_decode(ctx->h_red[0]) : 0.0;
    stats->max_v = ctx->np ? ord_decode(ctx->h_red[1]) : -INFINITY;
    stats->max_pa = ctx->np ? ord_decode(ctx->h_red[2]) : -INFINITY;
    stats->n_halo_violations = ctx->h_red[3];
    stats->n_untiled = ctx->h_red[4];
    stats->npart_local = ctx->np;
    stats->n_migrated = ctx->last_migrated;
    stats->kernel_launches = ctx->launches - launches0;
    if (tim) {
      float ms;
      cudaEventElapsedTime(&ms, ev[0], ev[1]); stats->ms[NNPM_MS_SORT] = ms;
      cudaEventElapsedTime(&ms, ev[1], ev[2]); stats->ms[NNPM_MS_DEPOSIT] = ms;
      cudaEventElapsedTime(&ms, ev[2], ev[3]); stats->ms[NNPM_MS_SOLVE] = ms;
      cudaEventElapsedTime(&ms, ev[3], ev[4]); stats->ms[NNPM_MS_PUSH] = ms;
      cudaEventElapsedTime(&ms, ev[4], ev[5]); stats->ms[NNPM_MS_EXCHANGE] = ms;
      if (ctx->P > 1) {
        float m1, m2;
        cudaEventElapsedTime(&m1, ev[8], ev[9]);
        cudaEventElapsedTime(&m2, ev[10], ev[11]);
        stats->ms[NNPM_MS_COMM] = m1 + m2;
      }
      cudaEventElapsedTime(&ms, ev[12], ev[13]);
      stats->ms[NNPM_MS_FFT_Z] = ms;  // un-fused path: includes the inverse 2-D transforms' z part only
      stats->ms[NNPM_MS_FFT_XY] = stats->ms[NNPM_MS_SOLVE] - ms - stats->ms[NNPM_MS_COMM];
      // x / y split of the 2-D transforms (only recorded on the un-pipelined path with separate x and y kernels)
      float a1, a2;
      if (!(ctx->opt_fft2d && fft2d_supported(ctx))) {
        if (cudaEventElapsedTime(&a1, ctx->ev_fft[0], ctx->ev_fft[1]) == cudaSuccess && cudaEventElapsedTime(&a2, ctx->ev_fft[5], ctx->ev_fft[6]) == cudaSuccess)
          stats->ms[NNPM_MS_FFT_X] = a1 + a2;
        if (cudaEventElapsedTime(&a1, ctx->ev_fft[1], ctx->ev_fft[2]) == cudaSuccess && cudaEventElapsedTime(&a2, ctx->ev_fft[4], ctx->ev_fft[5]) == cudaSuccess)
          stats->ms[NNPM_MS_FFT_Y] = a1 + a2;
      }
      cudaGetLastError();  // events never recorded (cuFFT 2-D path): not an error
    }
  }
  if (ctx->h_red[3] != 0)
    return fail(ctx, NNPM_ERR_STATE, "%lld particles left the slab ghost region within one half-drift (dt too large for the slab decomposition)", (long long)ctx->h_red[3]);
  return NNPM_OK;
}

extern "C" int nnpm_step_static(nnpm_ctx* ctx, double dt, nnpm_stats* stats) {
  if (!ctx) return NNPM_ERR_ARG;
  return do_step(ctx, false, dt, 1.0, 1.0, 0.0, 1.0, stats);
}
extern "C" int nnpm_step_comoving(nnpm_ctx* ctx, double dt, double a_d1, double a_k, double adot_k, double a_d2,
                                  nnpm_stats* stats) {
  if (!ctx) return NNPM_ERR_ARG;
  return do_step(ctx, true, dt, a_d1, a_k, adot_k, a_d2, stats);
}

#include "nnpm_pk.cuh"
#include "nnpm_ic.cuh"

extern "C" int nnpm_set_option(nnpm_ctx* ctx, const char* name, double value) {
  if (!ctx || !name) return NNPM_ERR_ARG;
  const int v = (int)value;
  if (!strcmp(name, "fft3d")) {
    if (v && ctx->P > 1) return fail(ctx, NNPM_ERR_ARG, "fft3d needs nranks == 1");
    ctx->opt_fft3d = v;
  } else if (!strcmp(name, "zfused")) ctx->opt_zfused = v;
  else if (!strcmp(name, "tile_deposit")) ctx->opt_tile_deposit = v;
  else if (!strcmp(name, "tile_push")) ctx->opt_tile_push = v;
  else if (!strcmp(name, "bulk_flush")) ctx->opt_bulk_flush = v;
  else if (!strcmp(name, "sort_cells")) ctx->opt_sort_cells = v;
  else if (!strcmp(name, "colfft")) ctx->opt_colfft = v;
  else if (!strcmp(name, "rowfft")) ctx->opt_rowfft = v;
  else if (!strcmp(name, "fft2d")) {
    if (v && !NNPM_EXP_PEER) return fail(ctx, NNPM_ERR_ARG, "option fft2d needs a library built with -DNNPM_EXPERIMENTS (make EXPERIMENTS=1)");
    ctx->opt_fft2d = v;
  }
  else if (!strcmp(name, "fft2d_block")) ctx->opt_fft2d_block = v;
  else if (!strcmp(name, "pipeline")) ctx->opt_pipeline = v;
  else if (!strcmp(name, "pipeline_chunks")) ctx->opt_pipe_chunks = v < 0 ? 0 : (v > 4 ? 4 : v);
  else if (!strcmp(name, "ce_streams")) ctx->opt_ce_streams = v < 0 ? 0 : (v > 8 ? 8 : v);
  else if (!strcmp(name, "fuse_transpose")) {
    if (v && !NNPM_EXP_PEER) return fail(ctx, NNPM_ERR_ARG, "option fuse_transpose needs a library built with -DNNPM_EXPERIMENTS (make EXPERIMENTS=1)");
    ctx->opt_fuse_transpose = v;
  }
  else if (!strcmp(name, "mig_in_push")) ctx->opt_mig_in_push = v;
  else if (!strcmp(name, "peer_transpose")) ctx->opt_peer = v < 0 ? 3 : v;
  else if (!strcmp(name, "xfer_ctas")) ctx->opt_xfer_ctas = v;
  else if (!strcmp(name, "zsplit")) ctx->opt_zsplit = v;
  else if (!strcmp(name, "l2pf_push")) ctx->opt_l2pf_push = v;
  else if (!strcmp(name, "l2pf_deposit")) ctx->opt_l2pf_deposit = v;
  else if (!strcmp(name, "cfrun_y")) ctx->opt_cfrun_y = v < 1 ? 1 : v;
  else if (!strcmp(name, "keep_rho")) ctx->opt_keep_rho = v;
  else return fail(ctx, NNPM_ERR_ARG, "unknown option '%s'", name);
  return NNPM_OK;
}
