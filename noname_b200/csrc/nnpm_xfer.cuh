// nnpm_xfer.cuh -- SM transport of the slab transposes ("peer_transpose" = 2): the same strided blocks the copy
// engines move in comm_transpose_ce / comm_chunk_ce (nnpm_comm.cuh), moved by one kernel instead.  A CTA is a single
// thread driving the TMA unit: 8 KB pieces travel  local HBM -> shared memory (cp.async.bulk, completion on an
// mbarrier) -> the destination rank's mesh (cp.async.bulk shared -> global on the peer mapping, i.e. NVLink writes in
// full-size packets), through a 4-slot ring, two loads and one store in flight beyond the current piece.  Consecutive
// pieces of a CTA go to consecutive destination ranks, so every NVLink port is loaded all the time; the rank's own
// block is one more destination.  32 KB of shared memory and one warp per CTA: the kernel shares the SMs with the
// transform kernels of the chunked pipeline (highest-priority stream, so its CTAs are placed first).
// This is the exchange step of compute_potential's distributed FFT (the reference has no multi-GPU path:
// SURVEY.md section 8e); results are bit-identical to the copy-engine and NCCL forms (pure data movement).
#pragma once

#define XF_PIECE 8192u
#define XF_STAGES 4

struct XferArgs {
  const char* src[16];   // per destination: the first row of the block that goes there
  char* dst[16];         // per destination: where that row lands (peer mapping, or local for the own block)
  size_t spitch, dpitch; // row strides in bytes
  unsigned width, rows;  // bytes per row (multiple of 16), rows per block
  unsigned ppr, ndst;    // pieces per row, destinations
};

__global__ void __launch_bounds__(32) k_peer_xfer(const __grid_constant__ XferArgs a) {
  extern __shared__ __align__(128) unsigned char xf_ring[];
  __shared__ unsigned long long bar[XF_STAGES];
  if (threadIdx.x != 0) return;
  for (int s = 0; s < XF_STAGES; s++) mbar_init(&bar[s], 1);
  const u64 total = (u64)a.ndst * a.rows * a.ppr, G = gridDim.x;
  const u64 n = total > blockIdx.x ? (total - blockIdx.x + G - 1) / G : 0;
  const u64 pol = l2_policy_evict_first();   // read once, never again
  auto locate = [&](u64 i, const char*& s, char*& d, unsigned& bytes) {
    const u64 q = blockIdx.x + i * G;
    const unsigned o = (unsigned)(q % a.ndst);
    const u64 rest = q / a.ndst;
    const unsigned seg = (unsigned)(rest % a.ppr);
    const u64 row = rest / a.ppr;
    const unsigned off = seg * XF_PIECE;
    bytes = a.width - off < XF_PIECE ? a.width - off : XF_PIECE;
    s = a.src[o] + row * a.spitch + off;
    d = a.dst[o] + row * a.dpitch + off;
  };
  auto load = [&](u64 i) {
    const char* s; char* d; unsigned bytes;
    locate(i, s, d, bytes);
    const int st = (int)(i % XF_STAGES);
    mbar_expect_tx(&bar[st], bytes);
    bulk_load_1d_hint(xf_ring + (size_t)st * XF_PIECE, s, bytes, &bar[st], pol);
  };
  for (u64 j = 0; j < 2 && j < n; j++) load(j);
  for (u64 i = 0; i < n; i++) {
    if (i + 2 < n) {
      // slot (i+2) % 4 was read by the store of piece i-2: all but the latest store must have left shared memory
      if (i >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      load(i + 2);
    }
    mbar_wait(&bar[i % XF_STAGES], (unsigned)((i / XF_STAGES) & 1));
    const char* s; char* d; unsigned bytes;
    locate(i, s, d, bytes);
    bulk_store_1d(d, xf_ring + (size_t)(i % XF_STAGES) * XF_PIECE, bytes);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the remote writes are complete before the kernel ends
}

// chunk c of nch of a transpose (nch = 1: all of it), on stream s.  Geometry as in comm_chunk_ce.
static int comm_xfer_sm(nnpm_ctx* ctx, bool fwd, int c, int nch, cudaStream_t s) {
  const Geom& g = ctx->g;
  if (ctx->P > 16) return fail(ctx, NNPM_ERR_ARG, "peer_transpose = 2 supports up to 16 ranks");
  const size_t chunk = (size_t)ctx->njl * ctx->icn * 16;
  const size_t plane_b = (size_t)g.plane * 8;
  const size_t glo_b = (size_t)g.plane * g.glo * 8;
  XferArgs a;
  a.ndst = (unsigned)ctx->P;
  if (fwd) {
    const size_t pc = (size_t)g.nzl / nch;
    a.spitch = plane_b; a.dpitch = chunk; a.width = (unsigned)chunk; a.rows = (unsigned)pc;
    for (int o = 0; o < ctx->P; o++) {
      const int d = (ctx->r + o) % ctx->P;
      a.src[o] = (const char*)ctx->mesh + glo_b + plane_b * pc * c + chunk * d;
      a.dst[o] = (char*)ctx->peer_meshT[d] + chunk * ((size_t)g.kz0 + pc * c);
    }
  } else {
    const size_t wc = chunk / nch;
    a.spitch = chunk; a.dpitch = plane_b; a.width = (unsigned)wc; a.rows = (unsigned)g.nzl;
    for (int o = 0; o < ctx->P; o++) {
      const int d = (ctx->r + o) % ctx->P;
      a.src[o] = (const char*)ctx->meshT + chunk * (size_t)g.nzl * d + wc * c;
      a.dst[o] = (char*)ctx->peer_mesh[d] + glo_b + chunk * ctx->r + wc * c;
    }
  }
  if (a.width % 16) return fail(ctx, NNPM_ERR_ARG, "peer_transpose = 2: row width must be a multiple of 16 bytes");
  a.ppr = (a.width + XF_PIECE - 1) / XF_PIECE;
  const unsigned grid = (unsigned)(ctx->opt_xfer_ctas > 0 ? ctx->opt_xfer_ctas : ctx->nsm);
  k_peer_xfer<<<grid, 32, XF_STAGES * XF_PIECE, s>>>(a);
  LAUNCHED(ctx);
  KCHECK();
  return NNPM_OK;
}
