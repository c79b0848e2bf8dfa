// nnpm_zfft.cuh -- K3: fused [z-FFT -> Green's function multiply -> z-iFFT] in one HBM pass.
// PLACEHOLDER for the first bring-up: reports "unsupported" so the cuFFT z path is used.
#pragma once
static bool zfused_supported(nnpm_ctx*) { return false; }
static int zfused_launch(nnpm_ctx* ctx, double2*, double) { return fail(ctx, NNPM_ERR_STATE, "fused z pass not built"); }
