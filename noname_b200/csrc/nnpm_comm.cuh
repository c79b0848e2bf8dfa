// nnpm_comm.cuh -- slab-decomposed multi-GPU plumbing (NCCL over NVLink), included by nnpm.cu.
//
// The reference has no parallelism (Readme.md:53); everything here is new.  One process drives one
// GPU / one slab of k-planes.  NCCL is bound at run time with dlopen("libnccl.so.2") so that, inside
// a torch.distributed process, the library shares torch's already-loaded NCCL instead of pulling in
// a second copy; nothing here is needed (or loaded) when nranks == 1.
//
//   comm_fold_rho_ghosts   deposit ghost planes -> owning neighbour (p2p send/recv + add)
//   comm_fill_phi_ghosts   potential ghost planes <- neighbours (p2p)
//   comm_transpose_fwd/bwd slab transpose of the half spectrum = all-to-all (grouped send/recv)
//   comm_migrate           particles that left the slab -> new owner (count exchange + p2p)
#pragma once
#include <nccl.h>

struct NcclApi {
  void* h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
};

static NcclApi* nccl_api(std::string* why) {
  static NcclApi api;
  static bool tried = false, ok = false;
  static std::string err;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char* n : names) {
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) {
      err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
    } else {
      ok = true;
#define NSYM(field, name)                                      \
  do {                                                         \
    *(void**)(&api.field) = dlsym(api.h, name);                \
    if (!api.field) { ok = false; err = std::string("missing NCCL symbol ") + name; } \
  } while (0)
      NSYM(GetUniqueId, "ncclGetUniqueId");
      NSYM(CommInitRank, "ncclCommInitRank");
      NSYM(CommDestroy, "ncclCommDestroy");
      NSYM(Send, "ncclSend");
      NSYM(Recv, "ncclRecv");
      NSYM(AllReduce, "ncclAllReduce");
      NSYM(AllGather, "ncclAllGather");
      NSYM(GroupStart, "ncclGroupStart");
      NSYM(GroupEnd, "ncclGroupEnd");
      NSYM(GetErrorString, "ncclGetErrorString");
#undef NSYM
    }
  }
  if (!ok && why) *why = err;
  return ok ? &api : nullptr;
}

#define CKN(call)                                                                                      \
  do {                                                                                                 \
    ncclResult_t n_ = (call);                                                                          \
    if (n_ != ncclSuccess)                                                                             \
      return fail(ctx, NNPM_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, ctx->nccl->GetErrorString(n_)); \
  } while (0)
#define COMM ((ncclComm_t)ctx->comm)
#define NEED_COMM()                                                                                   \
  do {                                                                                                \
    if (!ctx->comm) return fail(ctx, NNPM_ERR_STATE, "nranks > 1 but nnpm_comm_init was not called"); \
  } while (0)

extern "C" int nnpm_comm_unique_id(uint8_t id_out[128]) {
  std::string why;
  NcclApi* a = nccl_api(&why);
  if (!a) { g_create_err = why; return NNPM_ERR_NCCL; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (a->GetUniqueId(&id) != ncclSuccess) { g_create_err = "ncclGetUniqueId failed"; return NNPM_ERR_NCCL; }
  memcpy(id_out, &id, 128);
  return NNPM_OK;
}


static int comm_allreduce(nnpm_ctx* ctx, double* host, int n, ncclRedOp_t op);

// Map every rank's mesh and meshT into this process (CUDA IPC) so that the slab transposes can store
// straight into the peers' buffers over NVLink.  All ranks agree (allreduce-min) on whether it worked;
// if not, the grouped ncclSend/ncclRecv all-to-all is used.
static int peer_setup(nnpm_ctx* ctx) {
  const int P = ctx->P, r = ctx->r;
  ctx->peer_ok = false;
  if (getenv("NNPM_NO_PEER")) return NNPM_OK;
  struct Rec { cudaIpcMemHandle_t mesh, meshT; };
  static_assert(sizeof(Rec) == 128, "two 64-byte IPC handles");
  Rec mine;
  int ok = 1;
  if (cudaIpcGetMemHandle(&mine.mesh, ctx->mesh) != cudaSuccess) ok = 0;
  if (cudaIpcGetMemHandle(&mine.meshT, ctx->meshT) != cudaSuccess) ok = 0;
  cudaGetLastError();
  uint8_t* d_all = nullptr;
  CKR(dalloc(ctx, &d_all, (size_t)128 * (P + 1)));
  CK(cudaMemcpyAsync(d_all + 128 * P, &mine, 128, cudaMemcpyHostToDevice, ctx->st));
  CKN(ctx->nccl->AllGather(d_all + 128 * P, d_all, 128, ncclUint8, COMM, ctx->st));
  std::vector<Rec> all((size_t)P);
  CK(cudaMemcpyAsync(all.data(), d_all, (size_t)128 * P, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  for (int d = 0; d < P && ok; d++) {
    if (d == r) { ctx->peer_mesh[d] = ctx->mesh; ctx->peer_meshT[d] = ctx->meshT; continue; }
    void *pm = nullptr, *pt = nullptr;
    if (cudaIpcOpenMemHandle(&pm, all[d].mesh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&pt, all[d].meshT, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      ok = 0;
      cudaGetLastError();
      break;
    }
    ctx->peer_mesh[d] = (double*)pm;
    ctx->peer_meshT[d] = (double*)pt;
  }
  double flag = ok ? 1.0 : 0.0;
  CKR(comm_allreduce(ctx, &flag, 1, ncclMin));
  ctx->peer_ok = flag > 0.5;
  if (getenv("NNPM_VERBOSE")) fprintf(stderr, "[nnpm] rank %d: NVLink peer-memory transposes %s\n", r, ctx->peer_ok ? "enabled" : "unavailable (NCCL send/recv)");
  return NNPM_OK;
}

extern "C" int nnpm_comm_init(nnpm_ctx* ctx, const uint8_t id[128]) {
  if (!ctx || !id) return NNPM_ERR_ARG;
  if (ctx->P == 1) return NNPM_OK;
  if (ctx->comm) return fail(ctx, NNPM_ERR_STATE, "communicator already initialised");
  std::string why;
  ctx->nccl = nccl_api(&why);
  if (!ctx->nccl) return fail(ctx, NNPM_ERR_NCCL, "%s", why.c_str());
  CK(cudaSetDevice(ctx->cfg.device));
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclComm_t comm;
  CKN(ctx->nccl->CommInitRank(&comm, ctx->P, uid, ctx->r));
  ctx->comm = (void*)comm;
  // ghost-plane receive buffer (5 planes), migration buffers, counters
  CKR(dalloc(ctx, &ctx->ghost_rx, (size_t)ctx->g.plane * 5));
  i64 want = ctx->cap / 8;
  if (want < 65536) want = 65536;
  ctx->mig_cap = (want / ctx->P) * ctx->P;
  CKR(dalloc(ctx, &ctx->mig_send, (size_t)ctx->mig_cap * MIG_REC));
  CKR(dalloc(ctx, &ctx->mig_recv, (size_t)ctx->mig_cap * MIG_REC));
  CKR(dalloc(ctx, &ctx->mig_holes, (size_t)ctx->mig_cap));
  CKR(dalloc(ctx, &ctx->mig_lo, (size_t)ctx->mig_cap));
  CKR(dalloc(ctx, &ctx->mig_cnt, (size_t)(MC_ALL + MC_REC * 64)));  // local record (MC_* slots), then the all-gathered records of every rank
  CKR(dalloc(ctx, &ctx->bar_buf, 4));
  CK(cudaMemsetAsync(ctx->bar_buf, 0, 16, ctx->st));
  CKR(peer_setup(ctx));
  CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < 8; i++) {
    CK(cudaStreamCreateWithFlags(&ctx->st2[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
  }
  {  // the SM transport of the transposes ("peer_transpose" = 2) runs beside the transforms: its CTAs go first
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&ctx->st_hi, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&ctx->ev_hi, cudaEventDisableTiming));
  }
  return NNPM_OK;
}

static void comm_destroy(nnpm_ctx* ctx) {
  if (ctx->comm_borrowed) { ctx->comm = nullptr; return; }
  for (int d = 0; d < ctx->P && d < 64; d++) {
    if (d == ctx->r) continue;
    if (ctx->peer_mesh[d]) cudaIpcCloseMemHandle(ctx->peer_mesh[d]);
    if (ctx->peer_meshT[d]) cudaIpcCloseMemHandle(ctx->peer_meshT[d]);
    ctx->peer_mesh[d] = ctx->peer_meshT[d] = nullptr;
  }
  for (int i = 0; i < 8; i++) {
    if (ctx->st2[i]) { cudaStreamSynchronize(ctx->st2[i]); cudaStreamDestroy(ctx->st2[i]); ctx->st2[i] = nullptr; }
    if (ctx->ev_join[i]) { cudaEventDestroy(ctx->ev_join[i]); ctx->ev_join[i] = nullptr; }
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  ctx->ev_fork = nullptr;
  if (ctx->st_hi) { cudaStreamSynchronize(ctx->st_hi); cudaStreamDestroy(ctx->st_hi); ctx->st_hi = nullptr; }
  if (ctx->ev_hi) { cudaEventDestroy(ctx->ev_hi); ctx->ev_hi = nullptr; }
  if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy(COMM);
  ctx->comm = nullptr;
}

// ------------------------------------------------------------------------------------------
// ghost planes
// ------------------------------------------------------------------------------------------
// Deposit touches local planes 1 .. nzl+3 (particles may have half-drifted one plane out of the
// slab, and CIC reaches one plane up): fold plane 1 into the lower neighbour's top owned plane and
// planes nzl+2, nzl+3 into the upper neighbour's first two owned planes.
static int comm_fold_rho_ghosts(nnpm_ctx* ctx, bool fixed) {
  NEED_COMM();
  const Geom& g = ctx->g;
  const int up = (ctx->r + 1) % ctx->P, dn = (ctx->r + ctx->P - 1) % ctx->P;
  const size_t pl = (size_t)g.plane;
  double* below = ctx->mesh + pl * 1;                 // 1 plane  -> dn
  double* above = ctx->mesh + pl * (g.glo + g.nzl);   // 2 planes -> up
  double* rx_from_up = ctx->ghost_rx;                 // 1 plane (up's "below")
  double* rx_from_dn = ctx->ghost_rx + pl;            // 2 planes (dn's "above")
  CKN(ctx->nccl->GroupStart());
  CKN(ctx->nccl->Send(below, pl, ncclDouble, dn, COMM, ctx->st));
  CKN(ctx->nccl->Send(above, 2 * pl, ncclDouble, up, COMM, ctx->st));
  CKN(ctx->nccl->Recv(rx_from_up, pl, ncclDouble, up, COMM, ctx->st));
  CKN(ctx->nccl->Recv(rx_from_dn, 2 * pl, ncclDouble, dn, COMM, ctx->st));
  CKN(ctx->nccl->GroupEnd());
  ctx->launches += 1;
  double* top = ctx->mesh + pl * (g.glo + g.nzl - 1);
  double* first = ctx->mesh + pl * g.glo;
  if (fixed) {
    k_add_planes<true><<<nblk(pl, 256), 256, 0, ctx->st>>>(top, rx_from_up, pl);
    k_add_planes<true><<<nblk(2 * pl, 256), 256, 0, ctx->st>>>(first, rx_from_dn, 2 * pl);
  } else {
    k_add_planes<false><<<nblk(pl, 256), 256, 0, ctx->st>>>(top, rx_from_up, pl);
    k_add_planes<false><<<nblk(2 * pl, 256), 256, 0, ctx->st>>>(first, rx_from_dn, 2 * pl);
  }
  ctx->launches += 2;
  KCHECK();
  return NNPM_OK;
}

// The push gathers phi from planes k-1 .. k+2 of particles that may sit one plane outside the slab:
// 2 ghost planes below (the lower neighbour's top two) and 3 above (the upper neighbour's first three).
static int comm_fill_phi_ghosts(nnpm_ctx* ctx) {
  NEED_COMM();
  const Geom& g = ctx->g;
  const int up = (ctx->r + 1) % ctx->P, dn = (ctx->r + ctx->P - 1) % ctx->P;
  const size_t pl = (size_t)g.plane;
  double* own = ctx->mesh + pl * g.glo;
  CKN(ctx->nccl->GroupStart());
  CKN(ctx->nccl->Send(own + pl * (g.nzl - 2), 2 * pl, ncclDouble, up, COMM, ctx->st));  // my top two -> up's below ghosts
  CKN(ctx->nccl->Send(own, 3 * pl, ncclDouble, dn, COMM, ctx->st));                      // my first three -> dn's above ghosts
  CKN(ctx->nccl->Recv(ctx->mesh, 2 * pl, ncclDouble, dn, COMM, ctx->st));
  CKN(ctx->nccl->Recv(own + pl * g.nzl, 3 * pl, ncclDouble, up, COMM, ctx->st));
  CKN(ctx->nccl->GroupEnd());
  ctx->launches += 1;
  return NNPM_OK;
}

// ------------------------------------------------------------------------------------------
// slab transpose of the half spectrum: A[kl][j][ic] (all j, local k) <-> T[k][jl][ic] (all k, local j)
// ------------------------------------------------------------------------------------------
static int comm_all_to_all(nnpm_ctx* ctx, const double* src, double* dst) {
  const Geom& g = ctx->g;
  const size_t blk = (size_t)g.nzl * ctx->njl * ctx->icn * 2;  // doubles per (src rank, dst rank) block
  CKN(ctx->nccl->GroupStart());
  for (int o = 1; o < ctx->P; o++) {
    const int to = (ctx->r + o) % ctx->P, from = (ctx->r + ctx->P - o) % ctx->P;
    CKN(ctx->nccl->Send(src + blk * to, blk, ncclDouble, to, COMM, ctx->st));
    CKN(ctx->nccl->Recv(dst + blk * from, blk, ncclDouble, from, COMM, ctx->st));
  }
  CKN(ctx->nccl->GroupEnd());
  CK(cudaMemcpyAsync(dst + blk * ctx->r, src + blk * ctx->r, blk * 8, cudaMemcpyDeviceToDevice, ctx->st));
  ctx->launches += 2;
  return NNPM_OK;
}

// stream-ordered barrier over all ranks: remote stores of the kernels before it are complete (kernel
// boundary) before any rank's later work starts
static int comm_barrier(nnpm_ctx* ctx) {
  CKN(ctx->nccl->AllReduce(ctx->bar_buf, ctx->bar_buf, 1, ncclInt32, ncclSum, COMM, ctx->st));
  ctx->launches += 1;
  return NNPM_OK;
}

// Copy-engine form of the same transposes.  For a fixed destination rank d the block is one strided
// 2-D copy: nzl chunks of njl*icn complex numbers, contiguous on both sides --
//   forward   src A[kl][d*njl ..][:]  (chunk stride = padded plane)  ->  T_d[kz0+kl][:][:]  (contiguous)
//   backward  src T[d*nzl+kl][:][:]   (contiguous)                   ->  A_d[kl][j0 ..][:]  (plane stride)
// so the whole all-to-all is P-1 peer cudaMemcpy2DAsync calls on the NVLink copy engines (no SM
// time), spread over up to four side streams so several copy engines run at once.
static int comm_xfer_sm(nnpm_ctx* ctx, bool fwd, int c, int nch, cudaStream_t s);   // nnpm_xfer.cuh
static int comm_transpose_ce(nnpm_ctx* ctx, bool fwd) {
  const Geom& g = ctx->g;
  if (ctx->opt_peer >= 2) {   // SM transport: one kernel on the compute stream moves every block
    CKR(comm_xfer_sm(ctx, fwd, 0, 1, ctx->st));
    return comm_barrier(ctx);
  }
  const size_t chunk = (size_t)ctx->njl * ctx->icn * 16;      // bytes of one (plane, j-block) chunk
  const size_t plane_b = (size_t)g.plane * 8;
  const size_t glo_b = (size_t)g.plane * g.glo * 8;
  const int ns = ctx->P < ctx->ce_ns ? ctx->P : ctx->ce_ns;
  CK(cudaEventRecord(ctx->ev_fork, ctx->st));
  for (int i = 0; i < ns; i++) CK(cudaStreamWaitEvent(ctx->st2[i], ctx->ev_fork, 0));
  for (int o = 0; o < ctx->P; o++) {
    const int d = (ctx->r + o) % ctx->P;
    cudaStream_t s = ctx->st2[o % ns];
    if (fwd) {
      const char* src = (const char*)ctx->mesh + glo_b + chunk * d;
      char* dst = (char*)ctx->peer_meshT[d] + chunk * (size_t)g.kz0;
      CK(cudaMemcpy2DAsync(dst, chunk, src, plane_b, chunk, (size_t)g.nzl, cudaMemcpyDeviceToDevice, s));
    } else {
      const char* src = (const char*)ctx->meshT + chunk * (size_t)g.nzl * d;
      char* dst = (char*)ctx->peer_mesh[d] + glo_b + chunk * ctx->r;
      CK(cudaMemcpy2DAsync(dst, plane_b, src, chunk, chunk, (size_t)g.nzl, cudaMemcpyDeviceToDevice, s));
    }
    ctx->launches += 1;
  }
  for (int i = 0; i < ns; i++) {
    CK(cudaEventRecord(ctx->ev_join[i], ctx->st2[i]));
    CK(cudaStreamWaitEvent(ctx->st, ctx->ev_join[i], 0));
  }
  return comm_barrier(ctx);
}

// Pipelined form: chunk c of nch.  forward: the planes [c*nzl/nch, (c+1)*nzl/nch) of this rank (all j);
// backward: the rows jl in [c*njl/nch, (c+1)*njl/nch) of T (all k).  The copies are enqueued on the side
// streams behind `after` (recorded by the caller on the compute stream when the chunk's data is ready),
// so they run on the copy engines while the next chunk is being transformed.
static int comm_chunk_ce(nnpm_ctx* ctx, bool fwd, int c, int nch, cudaEvent_t after) {
  const Geom& g = ctx->g;
  // 2: every chunk by the SM transport; 3: only the last one, which no transform is left to run beside
  if (ctx->opt_peer == 2 || (ctx->opt_peer == 3 && c == nch - 1)) {
    CK(cudaStreamWaitEvent(ctx->st_hi, after, 0));
    return comm_xfer_sm(ctx, fwd, c, nch, ctx->st_hi);
  }
  const size_t chunk = (size_t)ctx->njl * ctx->icn * 16;
  const size_t plane_b = (size_t)g.plane * 8;
  const size_t glo_b = (size_t)g.plane * g.glo * 8;
  const int ns = ctx->P < ctx->ce_ns ? ctx->P : ctx->ce_ns;
  for (int i = 0; i < ns; i++) CK(cudaStreamWaitEvent(ctx->st2[i], after, 0));
  for (int o = 0; o < ctx->P; o++) {
    const int d = (ctx->r + o) % ctx->P;
    cudaStream_t s = ctx->st2[o % ns];
    if (fwd) {
      const size_t pc = (size_t)g.nzl / nch;  // planes per chunk
      const char* src = (const char*)ctx->mesh + glo_b + plane_b * pc * c + chunk * d;
      char* dst = (char*)ctx->peer_meshT[d] + chunk * ((size_t)g.kz0 + pc * c);
      CK(cudaMemcpy2DAsync(dst, chunk, src, plane_b, chunk, pc, cudaMemcpyDeviceToDevice, s));
    } else {
      const size_t wc = chunk / nch;          // bytes of the jl sub-block of one k row
      const char* src = (const char*)ctx->meshT + chunk * (size_t)g.nzl * d + wc * c;
      char* dst = (char*)ctx->peer_mesh[d] + glo_b + chunk * ctx->r + wc * c;
      CK(cudaMemcpy2DAsync(dst, plane_b, src, chunk, wc, (size_t)g.nzl, cudaMemcpyDeviceToDevice, s));
    }
    ctx->launches += 1;
  }
  return NNPM_OK;
}
static int comm_join_barrier(nnpm_ctx* ctx) {
  if (ctx->opt_peer >= 2) {
    CK(cudaEventRecord(ctx->ev_hi, ctx->st_hi));
    CK(cudaStreamWaitEvent(ctx->st, ctx->ev_hi, 0));
    if (ctx->opt_peer == 2) return comm_barrier(ctx);
  }
  const int ns = ctx->P < ctx->ce_ns ? ctx->P : ctx->ce_ns;
  for (int i = 0; i < ns; i++) {
    CK(cudaEventRecord(ctx->ev_join[i], ctx->st2[i]));
    CK(cudaStreamWaitEvent(ctx->st, ctx->ev_join[i], 0));
  }
  return comm_barrier(ctx);
}

static int comm_transpose_fwd(nnpm_ctx* ctx) {
  NEED_COMM();
  const Geom& g = ctx->g;
  if (ctx->peer_ok && ctx->opt_peer) return comm_transpose_ce(ctx, true);
  if (!ctx->meshS) CKR(dalloc(ctx, &ctx->meshS, (size_t)g.plane * g.nzl));
  const double2* A = (const double2*)(ctx->mesh + g.plane * g.glo);
  dim3 grid(nblk(ctx->icn, 128), (unsigned)g.N2, (unsigned)g.nzl);
  k_pack_fwd<<<grid, 128, 0, ctx->st>>>(A, (double2*)ctx->meshS, ctx->icn, g.N2, ctx->njl, g.nzl, g.plane / 2);
  LAUNCHED(ctx);
  KCHECK();
  return comm_all_to_all(ctx, ctx->meshS, ctx->meshT);
}

static int comm_transpose_bwd(nnpm_ctx* ctx) {
  NEED_COMM();
  const Geom& g = ctx->g;
  if (ctx->peer_ok && ctx->opt_peer) return comm_transpose_ce(ctx, false);
  if (!ctx->meshS) CKR(dalloc(ctx, &ctx->meshS, (size_t)g.plane * g.nzl));
  CKR(comm_all_to_all(ctx, ctx->meshT, ctx->meshS));
  double2* A = (double2*)(ctx->mesh + g.plane * g.glo);
  dim3 grid(nblk(ctx->icn, 128), (unsigned)g.N2, (unsigned)g.nzl);
  k_unpack_bwd<<<grid, 128, 0, ctx->st>>>(A, (const double2*)ctx->meshS, ctx->icn, g.N2, ctx->njl, g.nzl, g.plane / 2);
  LAUNCHED(ctx);
  KCHECK();
  return NNPM_OK;
}

// ------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------
static int comm_allreduce(nnpm_ctx* ctx, double* host, int n, ncclRedOp_t op) {
  NEED_COMM();
  if (n > 2048) return fail(ctx, NNPM_ERR_ARG, "allreduce too large");
  if (!ctx->red_buf) CKR(dalloc(ctx, &ctx->red_buf, 2048));
  CK(cudaMemcpyAsync(ctx->red_buf, host, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->st));
  CKN(ctx->nccl->AllReduce(ctx->red_buf, ctx->red_buf, (size_t)n, ncclDouble, op, COMM, ctx->st));
  CK(cudaMemcpyAsync(host, ctx->red_buf, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  ctx->launches += 1;
  return NNPM_OK;
}
static int comm_allreduce_sum(nnpm_ctx* ctx, double* host, int n) { return comm_allreduce(ctx, host, n, ncclSum); }

extern "C" int nnpm_reduce_stats(nnpm_ctx* ctx, nnpm_stats* s) {
  if (!ctx || !s) return NNPM_ERR_ARG;
  if (ctx->P == 1) return NNPM_OK;
  CK(cudaSetDevice(ctx->cfg.device));
  double mx[3 + NNPM_MS_COUNT] = {s->max_abs_v, s->max_v, s->max_pa};
  for (int i = 0; i < NNPM_MS_COUNT; i++) mx[3 + i] = s->ms[i];
  CKR(comm_allreduce(ctx, mx, 3 + NNPM_MS_COUNT, ncclMax));
  double sm[2] = {(double)s->n_halo_violations, (double)s->n_migrated};
  CKR(comm_allreduce(ctx, sm, 2, ncclSum));
  s->max_abs_v = mx[0]; s->max_v = mx[1]; s->max_pa = mx[2];
  for (int i = 0; i < NNPM_MS_COUNT; i++) s->ms[i] = (float)mx[3 + i];
  s->n_halo_violations = (int64_t)sm[0];
  s->n_migrated = (int64_t)sm[1];
  return NNPM_OK;
}

// ------------------------------------------------------------------------------------------
// particle migration
// ------------------------------------------------------------------------------------------
__global__ void k_mig_set(unsigned long long* cnt, unsigned long long np, unsigned long long cap) {
  cnt[MC_NP] = np;
  cnt[MC_CAP] = cap;
}

// owner slab of every particle; leavers are packed into the per-destination segment of `send`,
// their slot is recorded as a hole and their id set to MIG_HOLE.
__global__ void k_mig_classify(PartPtrs p, uint32_t* __restrict__ ids, i64 np, Geom g, int P, int r, i64 seg_cap,
                               double* __restrict__ send, uint32_t* __restrict__ holes, unsigned long long* cnt) {
  i64 n = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (n >= np) return;
  const double x2 = wrap01(p.x[2][n]);
  i64 k = wrap_idx((i64)floor(__dmul_rn(x2, g.fN3)), g.N3);
  const int owner = (int)(k / g.nzl);
  if (owner == r) return;
  const unsigned long long slot = atomicAdd(&cnt[owner], 1ull);
  if ((i64)slot >= seg_cap) { atomicAdd(&cnt[MC_OVER], 1ull); return; }  // stays; caller runs another round
  double* rec = send + ((i64)owner * seg_cap + (i64)slot) * MIG_REC;
  rec[0] = p.x[0][n]; rec[1] = p.x[1][n]; rec[2] = p.x[2][n];
  rec[3] = p.v[0][n]; rec[4] = p.v[1][n]; rec[5] = p.v[2][n];
  rec[6] = __longlong_as_double((i64)ids[n]);
  ids[n] = MIG_HOLE;
  holes[atomicAdd(&cnt[MC_HOLES], 1ull)] = (uint32_t)n;
}
__global__ void k_mig_holes_lo(const uint32_t* __restrict__ holes, i64 nholes, i64 np_keep, uint32_t* __restrict__ lo,
                               unsigned long long* cnt) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t < nholes && (i64)holes[t] < np_keep) lo[atomicAdd(&cnt[MC_NLO], 1ull)] = holes[t];
}
__global__ void k_mig_fill(PartPtrs p, uint32_t* __restrict__ ids, i64 np_keep, i64 np, const uint32_t* __restrict__ lo,
                           unsigned long long* cnt) {
  i64 t = np_keep + blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= np || ids[t] == MIG_HOLE) return;
  const i64 d = lo[atomicAdd(&cnt[MC_NFILL], 1ull)];
#pragma unroll
  for (int c = 0; c < 3; c++) { p.x[c][d] = p.x[c][t]; p.v[c][d] = p.v[c][t]; }
  ids[d] = ids[t];
}
__global__ void k_mig_unpack(PartPtrs p, uint32_t* __restrict__ ids, i64 base, const double* __restrict__ recv, i64 n) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double* rec = recv + t * MIG_REC;
  const i64 d = base + t;
  p.x[0][d] = rec[0]; p.x[1][d] = rec[1]; p.x[2][d] = rec[2];
  p.v[0][d] = rec[3]; p.v[1][d] = rec[4]; p.v[2][d] = rec[5];
  ids[d] = (uint32_t)__double_as_longlong(rec[6]);
}

// zero the counter record and note np / capacity in it (before a classification, by the push or by k_mig_classify)
static int migrate_reset(nnpm_ctx* ctx) {
  unsigned long long* cnt = (unsigned long long*)ctx->mig_cnt;
  CK(cudaMemsetAsync(cnt, 0, MC_ALL * 8, ctx->st));
  k_mig_set<<<1, 1, 0, ctx->st>>>(cnt, (unsigned long long)ctx->np, (unsigned long long)ctx->cap);
  LAUNCHED(ctx);
  return NNPM_OK;
}
static MigArgs migrate_args(nnpm_ctx* ctx) {
  MigArgs mg;
  memset(&mg, 0, sizeof mg);
  if (ctx->P > 1 && ctx->comm && ctx->R == 3) {
    mg.on = 1; mg.P = ctx->P; mg.r = ctx->r; mg.kz0 = (int)ctx->g.kz0; mg.nzl = (int)ctx->g.nzl; mg.N3 = (int)ctx->g.N3;
    mg.fN3 = ctx->g.fN3;
    mg.seg_cap = ctx->mig_cap / ctx->P;
    mg.send = ctx->mig_send; mg.holes = ctx->mig_holes; mg.cnt = (unsigned long long*)ctx->mig_cnt; mg.ids = ctx->ids;
  }
  return mg;
}

// one migration round; *leftover = particles (all ranks) that did not fit the send segments.
// classified: the push already packed this step's leavers (mig_emit) -- skip the classification pass.
static int migrate_round(nnpm_ctx* ctx, i64* leftover, bool classified) {
  NEED_COMM();
  const Geom& g = ctx->g;
  const int P = ctx->P, r = ctx->r;
  const i64 seg_cap = ctx->mig_cap / P;
  unsigned long long* cnt = (unsigned long long*)ctx->mig_cnt;
  const i64 np = ctx->np;
  if (!classified) CKR(migrate_reset(ctx));
  if (np && !classified) {
    k_mig_classify<<<nblk(np, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, np, g, P, r, seg_cap, ctx->mig_send, ctx->mig_holes, cnt);
    LAUNCHED(ctx);
    KCHECK();
  }
  // everyone learns everyone's record: row s of the matrix = counts sent by rank s, its holes, np and capacity
  CKN(ctx->nccl->AllGather(cnt, cnt + MC_ALL, MC_REC, ncclUint64, COMM, ctx->st));
  LAUNCHED(ctx);
  i64* h = ctx->h_cnt;  // pinned, MC_REC * 64 entries
  CK(cudaMemcpyAsync(h, cnt + MC_ALL, (size_t)MC_REC * P * 8, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  auto sent = [&](int s, int d) { i64 c = h[MC_REC * s + d]; return c < seg_cap ? c : seg_cap; };
  i64 over = 0;
  for (int s = 0; s < P; s++) over += h[MC_REC * s + MC_OVER];
  *leftover = over;
  // capacity is decided COLLECTIVELY from the gathered records, before any rank posts a send or a receive: a rank
  // bailing out on its own would leave its peers hanging in NCCL
  for (int t = 0; t < P; t++) {
    i64 rx = 0;
    for (int s = 0; s < P; s++) if (s != t) rx += sent(s, t);
    const i64 after = h[MC_REC * t + MC_NP] - h[MC_REC * t + MC_HOLES] + rx;
    if (after > h[MC_REC * t + MC_CAP])
      return fail(ctx, NNPM_ERR_CAPACITY, "rank %d would hold %lld particles after migration, capacity %lld (every rank stops here; raise part_capacity)",
                  t, (long long)after, (long long)h[MC_REC * t + MC_CAP]);
  }
  const i64 nholes = h[MC_REC * r + MC_HOLES];
  i64 nrecv = 0;
  for (int s = 0; s < P; s++) if (s != r) nrecv += sent(s, r);
  const i64 np_keep = np - nholes;
  // data exchange
  CKN(ctx->nccl->GroupStart());
  i64 roff = 0;
  std::vector<i64> roffs(P, 0);
  for (int s = 0; s < P; s++) if (s != r) { roffs[s] = roff; roff += sent(s, r); }
  for (int o = 1; o < P; o++) {
    const int to = (r + o) % P, from = (r + P - o) % P;
    if (sent(r, to)) CKN(ctx->nccl->Send(ctx->mig_send + (i64)to * seg_cap * MIG_REC, (size_t)sent(r, to) * MIG_REC, ncclDouble, to, COMM, ctx->st));
    if (sent(from, r)) CKN(ctx->nccl->Recv(ctx->mig_recv + roffs[from] * MIG_REC, (size_t)sent(from, r) * MIG_REC, ncclDouble, from, COMM, ctx->st));
  }
  CKN(ctx->nccl->GroupEnd());
  LAUNCHED(ctx);
  // close the holes below np_keep with the survivors above it, then append the arrivals
  if (nholes) {
    k_mig_holes_lo<<<nblk(nholes, 256), 256, 0, ctx->st>>>(ctx->mig_holes, nholes, np_keep, ctx->mig_lo, cnt);
    k_mig_fill<<<nblk(np - np_keep, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, np_keep, np, ctx->mig_lo, cnt);
    ctx->launches += 2;
  }
  if (nrecv) {
    k_mig_unpack<<<nblk(nrecv, 256), 256, 0, ctx->st>>>(ctx->pp, ctx->ids, np_keep, ctx->mig_recv, nrecv);
    LAUNCHED(ctx);
  }
  KCHECK();
  ctx->np = np_keep + nrecv;
  ctx->last_migrated += nholes;
  ctx->ids_dirty = true;
  return NNPM_OK;
}

static int comm_migrate(nnpm_ctx* ctx, bool classified = false) {
  i64 left = 0;
  CKR(migrate_round(ctx, &left, classified));
  for (int it = 0; left > 0 && it < 64; it++) CKR(migrate_round(ctx, &left, false));
  if (left > 0) return fail(ctx, NNPM_ERR_CAPACITY, "particle migration did not converge (%lld left)", (long long)left);
  return NNPM_OK;
}

extern "C" int nnpm_redistribute(nnpm_ctx* ctx) {
  if (!ctx) return NNPM_ERR_ARG;
  if (ctx->P == 1) return NNPM_OK;
  CK(cudaSetDevice(ctx->cfg.device));
  ctx->last_migrated = 0;
  CKR(comm_migrate(ctx));
  CK(cudaStreamSynchronize(ctx->st));
  return NNPM_OK;
}
