// nnpm_comm.cuh -- slab-decomposed multi-GPU plumbing (NCCL over NVLink), included by nnpm.cu.
// PLACEHOLDER for the first single-GPU bring-up; replaced by the NCCL implementation.
#pragma once
struct NcclApi {};
static void comm_destroy(nnpm_ctx*) {}
static int comm_fold_rho_ghosts(nnpm_ctx* ctx, bool) { return fail(ctx, NNPM_ERR_NCCL, "multi-GPU path not built"); }
static int comm_fill_phi_ghosts(nnpm_ctx* ctx) { return fail(ctx, NNPM_ERR_NCCL, "multi-GPU path not built"); }
static int comm_transpose_fwd(nnpm_ctx* ctx) { return fail(ctx, NNPM_ERR_NCCL, "multi-GPU path not built"); }
static int comm_transpose_bwd(nnpm_ctx* ctx) { return fail(ctx, NNPM_ERR_NCCL, "multi-GPU path not built"); }
static int comm_allreduce_sum(nnpm_ctx* ctx, double*, int) { return fail(ctx, NNPM_ERR_NCCL, "multi-GPU path not built"); }
static int comm_migrate(nnpm_ctx* ctx) { return fail(ctx, NNPM_ERR_NCCL, "multi-GPU path not built"); }
extern "C" int nnpm_comm_unique_id(uint8_t id_out[128]) { (void)id_out; return NNPM_ERR_NCCL; }
extern "C" int nnpm_comm_init(nnpm_ctx* ctx, const uint8_t id[128]) { (void)id; return fail(ctx, NNPM_ERR_NCCL, "multi-GPU path not built"); }
extern "C" int nnpm_redistribute(nnpm_ctx* ctx) { if (!ctx) return NNPM_ERR_ARG; if (ctx->P == 1) return NNPM_OK; return fail(ctx, NNPM_ERR_NCCL, "multi-GPU path not built"); }
extern "C" int nnpm_reduce_stats(nnpm_ctx* ctx, nnpm_stats* s) { if (!ctx || !s) return NNPM_ERR_ARG; if (ctx->P == 1) return NNPM_OK; return fail(ctx, NNPM_ERR_NCCL, "multi-GPU path not built"); }
