/* nnpm.h -- C ABI of libnnpm.so, the B200-native particle-mesh gravity step for NoName.
 *
 * The reference (yipihey/NoName, pure Julia) has no FFI of its own; its extension mechanism is
 * function-name dispatch from the .conf file (src/NoName.jl:82-83) onto free functions that
 * mutate plain Float64 arrays (src/ParticleMesh.jl).  Each entry point below names the reference
 * function (file:line under /root/reference) it replaces.  A Julia host binds them with `ccall`
 * (julia/NoNameB200.jl); the Python host in noname_b200/ binds them with ctypes.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; every call returns 0 on success or a
 *     negative nnpm_status; nnpm_last_error(ctx) gives the message (owned by ctx).
 *   - host arrays use the reference's Julia column-major layout AS IS:
 *       mesh   rho/phi[i,j,k]   -> double[i + N1*(j + N2*k)]
 *       acc[c,i,j,k]            -> double[c + 3*(i + N1*(j + N2*k))]   (first dim always 3)
 *       particles x/v/pa[d,n]   -> double[d + rank*n]                  (rank = 2 or 3)
 *   - one host thread drives one context; one context drives ONE GPU (one process per GPU).
 *     With nranks > 1 the mesh is slab-decomposed along k (the slowest Julia index): rank r owns
 *     planes [r*N3/P, (r+1)*N3/P) and the particles whose floor(x3*N3) falls in them; mesh
 *     get/set calls then move only the local slab.
 *   - the device state (SoA x,v, mesh) stays resident between calls; host copies happen only in
 *     upload/download/get/set calls.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     NNPM_ERR_CUDA.
 *
 * Not provided, deliberately (SURVEY.md section 8b proposed them): a `gradient = spectral` switch -- the fused push
 * takes the reference's centred difference (deriv1_3d, :549-573) on the fly from the staged phi box, which is both
 * bit-faithful to the reference and free of the 25.8 GB acc array; multiplying the spectrum by -i sin(k_d) N gives the
 * same numbers to 6e-16 (tests/test_oracle_kats.py, kat5) at the price of three more inverse transforms.  And a
 * `rho_mean` statistic -- the mean is never needed (the DC bin is zeroed, :501) and is m * npart / ncell exactly.
 */
#ifndef NNPM_H
#define NNPM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNPM_VERSION 2

typedef struct nnpm_ctx nnpm_ctx;

typedef enum {
  NNPM_OK = 0,
  NNPM_ERR_ARG = -1,      /* bad argument / unsupported configuration */
  NNPM_ERR_CUDA = -2,     /* CUDA runtime error (incl. "no device") */
  NNPM_ERR_CUFFT = -3,
  NNPM_ERR_NCCL = -4,
  NNPM_ERR_CAPACITY = -5, /* particle capacity exceeded on this rank */
  NNPM_ERR_STATE = -6     /* call sequence error (e.g. interp before accel) */
} nnpm_status;

/* ParticleDepositInterpolation / ParticleBackInterpolation = none | cic  (src/default.conf:32-33) */
enum { NNPM_INTERP_NONE = 0, NNPM_INTERP_CIC = 1 };
/* deposit accumulation.  ATOMIC (default): f64 reductions into the mesh; on tile-sorted particles the tile's
 * contributions are first summed in shared memory as 64-bit fixed point (unit-mass weights, 2^-fixed_bits resolution)
 * and converted to f64 at the flush, so results differ from a pure f64 sum by <= 2^-fixed_bits per contribution
 * (inside the 1e-12 single-step budget).  FIXED: the mesh itself accumulates int64 fixed point -- bit-exact run to
 * run, under any particle order and any slab decomposition. */
enum { NNPM_DEPOSIT_ATOMIC = 0, NNPM_DEPOSIT_FIXED = 1 };
/* fields for nnpm_get_field / nnpm_set_field */
enum { NNPM_FIELD_RHO = 0, NNPM_FIELD_PHI = 1, NNPM_FIELD_ACC = 2, NNPM_FIELD_PA = 3 };
/* synthetic initial conditions generated on the device (nnpm_init_synthetic; nnpm_init_grf and nnpm_init_pancake
 * have entry points of their own) */
enum { NNPM_IC_LATTICE = 0, NNPM_IC_ZELDOVICH_MODES = 1, NNPM_IC_CLUSTERED = 2 };

typedef struct {
  int32_t struct_size;      /* = sizeof(nnpm_config), for ABI checks */
  int32_t rank;             /* particle dimensionality: 2 or 3 (ParticleMesh.jl:105) */
  int64_t ngrid[3];         /* TopgridDimensions N1,N2,N3 (N3 = 1 when rank == 2) */
  int64_t npart_total;      /* prod(ParticleDimensions): particles over all ranks; < 2^32 (original-order ids are uint32) */
  int64_t part_capacity;    /* per-rank particle capacity; 0 = npart_total/nranks * 1.25 + 4096 */
  int32_t deposit;          /* NNPM_INTERP_* */
  int32_t backinterp;       /* NNPM_INTERP_* */
  int32_t deposit_mode;     /* NNPM_DEPOSIT_* */
  int32_t fixed_bits;       /* fractional bits of the fixed-point deposit (0 = 44) */
  int32_t sort_every;       /* counting-sort particles by cell every n steps (0 = never) */
  int32_t bugcompat_interp_z; /* 1 = reproduce ParticleMesh.jl:706/:647 literally (rank 3) */
  int32_t device;           /* CUDA device ordinal */
  int32_t nranks;           /* slab decomposition width P (1 = single GPU) */
  int32_t rank_id;          /* this context's slab index in [0,P) */
  int32_t timing;           /* 1 = record per-phase CUDA-event times into nnpm_stats.ms */
  void *stream;             /* cudaStream_t to run on; NULL = the library creates its own */
} nnpm_config;

enum {
  NNPM_MS_DEPOSIT = 0, /* zero + (drift) + scatter + ghost-plane add (+ fixed->f64) */
  NNPM_MS_SOLVE = 1,   /* forward FFTs, Green multiply, inverse FFTs, transposes, ghost planes */
  NNPM_MS_PUSH = 2,    /* gradient + gather + kick + drift + wrap + reductions */
  NNPM_MS_SORT = 3,    /* counting sort (when it ran this step) */
  NNPM_MS_EXCHANGE = 4,/* particle migration between slabs */
  NNPM_MS_COMM = 5,    /* of SOLVE: time inside the all-to-all transposes */
  NNPM_MS_FFT_Z = 6,   /* of SOLVE: the z pass (z-FFT, Green multiply, z-iFFT) */
  NNPM_MS_FFT_XY = 7,  /* of SOLVE: the two batched 2-D transforms */
  NNPM_MS_FFT_X = 8,   /* of FFT_XY: the x passes (r2c + c2r); 0 when not separately timed */
  NNPM_MS_FFT_Y = 9,   /* of FFT_XY: the y passes (forward + inverse); 0 when not separately timed (chunked multi-rank pipeline) */
  NNPM_MS_COUNT = 10
};

/* per-step reductions the reference prints / uses for dt control
 * (src/ParticleMesh.jl:782, :875-878) */
typedef struct {
  double max_abs_v;   /* maximum(abs(v)) over this rank (the host adds 1e-30) */
  double max_v;       /* maximum(v), signed */
  double max_pa;      /* maximum(pa), signed */
  int64_t n_halo_violations; /* particles whose half-drift left the slab's ghost region (must be 0) */
  int64_t npart_local;/* particles on this rank after the step */
  int64_t n_migrated; /* particles this rank sent away this step */
  int64_t n_untiled;  /* particles the tiled kernels handed to the global-memory path (strayed > 1 cell from their tile since the sort) */
  float ms[NNPM_MS_COUNT]; /* CUDA-event milliseconds per phase (timing=1), else 0 */
  int32_t kernel_launches; /* kernels + cuFFT execs + NCCL ops launched by this call */
  int32_t pad;
} nnpm_stats;

/* ---- lifetime ------------------------------------------------------------------------- */
int nnpm_version(void);
/* Replaces the buffer allocation at src/ParticleMesh.jl:754-757 / :833-838 and the
 * GridPatch/GridData/particle-dict containers (src/GridTypes.jl:4-16, ParticleMesh.jl:112-124). */
int nnpm_create(nnpm_ctx **out, const nnpm_config *cfg);
int nnpm_destroy(nnpm_ctx *ctx);
const char *nnpm_last_error(const nnpm_ctx *ctx); /* ctx may be NULL: error of a failed create */

/* Multi-GPU (new; the reference has no parallelism, Readme.md:53).  Rank 0 calls
 * nnpm_comm_unique_id, the host broadcasts the 128 bytes by any means (torch.distributed, MPI,
 * a file), then every rank calls nnpm_comm_init.  Not needed when nranks == 1. */
int nnpm_comm_unique_id(uint8_t id_out[128]);
int nnpm_comm_init(nnpm_ctx *ctx, const uint8_t id[128]);

/* pinned host memory for fast upload/download (cudaHostAlloc) */
int nnpm_host_alloc(void **ptr, uint64_t bytes);
int nnpm_host_free(void *ptr);

/* ---- particle state --------------------------------------------------------------------- */
/* p["x"], p["v"], p["m"] (ParticleMesh.jl:122-124): n particles in Julia (rank,n) layout from
 * this rank's host; any particles may be given to any rank -- nnpm_redistribute routes them to
 * the owning slab.  first_id is the global index of x_aos[:,0] (original order is kept as an id
 * so that download returns the reference's ordering). */
int nnpm_upload_particles(nnpm_ctx *ctx, const double *x_aos, const double *v_aos, int64_t n,
                          int64_t first_id, double m);
/* Route every particle to the rank owning floor(x3*N3) (NCCL p2p); no-op when nranks == 1. */
int nnpm_redistribute(nnpm_ctx *ctx);
int64_t nnpm_local_count(const nnpm_ctx *ctx);
/* Copy this rank's particles back: x_aos/v_aos (rank,n_local) and their global ids (may be
 * NULL).  With ids_out == NULL and nranks == 1 the particles are written in ORIGINAL order. */
int nnpm_download_particles(nnpm_ctx *ctx, double *x_aos, double *v_aos, int64_t *ids_out);
/* Device-side synthetic ICs (lattice of ParticleMesh.jl:257-287, optionally displaced by a
 * seeded plane-wave Zel'dovich field; oracle/pm_ref.py:synthetic_zeldovich is the checker).
 * pdims = ParticleDimensions; v = vfac * displacement.
 * kind == NNPM_IC_CLUSTERED (BASELINE config 4 stand-in): half the particles stay on the lattice, half
 * fall into nmodes*kmax Gaussian blobs of width disp_rms (box units) with a b^-1/2 mass function and
 * velocity dispersion vfac; oracle/pm_ref.py:synthetic_clustered is the checker. */
int nnpm_init_synthetic(nnpm_ctx *ctx, int kind, const int64_t pdims[3], uint64_t seed,
                        int32_t nmodes, int32_t kmax, double disp_rms, double vfac, double m);
/* PowerSpectrumParticles (src/ParticleMesh.jl:209-254) with scaleFreePS (src/InputPowerSpectra.jl:26), on the
 * device and SEEDED: lattice + Zel'dovich displacement psi = irfft(c_d) on the particle lattice pdims with
 *   c_d(k) = sqrt(P(|k|)) (g1 - i g2) / 2 / k^2 * k_d,  P(k) = k^ps_index,  k in rad/cell via fftfreq (:202-206),
 * k = 0 skipped, (g1, g2) standard normal from a counter-based generator keyed by (seed, mode index) -- any slab
 * decomposition draws the same field.  Deviations from the (non-parsing, unseeded) reference function, as
 * SURVEY.md section 8(d) prescribes: one complex amplitude per mode shared by the components instead of a fresh
 * randn(2) per component; the planes i = 0 and i = N1/2 are made Hermitian so that the field is real whatever the
 * c2r transform does with redundant bins; psi is rescaled to a per-component RMS of disp_rms (box units); and
 * velocities v = vfac * psi are set (the reference leaves them "still to implement", :251; vfac = adot/a gives the
 * growing mode as in ZeldovichPancakeParticles :193-194).  The inverse transform runs through the library's own
 * FFT passes (and slab transposes when nranks > 1).  oracle/pm_ref.py:synthetic_grf is the checker. */
int nnpm_init_grf(nnpm_ctx *ctx, const int64_t pdims[3], uint64_t seed, double ps_index, double disp_rms,
                  double vfac, double m);
/* The same generator with multiplePeaksPS(n, norm, k, lnwidth) (src/InputPowerSpectra.jl:6-24, P at :28-36;
 * doc/InputPowerSpectra.md), the spectrum run/PM/particle_mesh.conf names:
 *   P(k) = k^ps_index  where |k - kpeak[p]| / k < lnwidth[p] / 2 for some p < npeaks,  0 elsewhere
 * (k = |k| in rad/cell, as the reference evaluates it, :229-231).  1 <= npeaks <= NNPM_MAX_PEAKS.  A spectrum whose
 * windows contain no mode of the lattice gives zero displacements.  oracle/pm_ref.py:synthetic_grf(peaks=...) checks it. */
#define NNPM_MAX_PEAKS 8
int nnpm_init_grf_peaks(nnpm_ctx *ctx, const int64_t pdims[3], uint64_t seed, double ps_index, int32_t npeaks,
                        const double *kpeak, const double *lnwidth, double disp_rms, double vfac, double m);
/* ZeldovichPancakeParticles (src/ParticleMesh.jl:177-198) on the device: lattice, then
 *   v1 = vamp * sin(2 pi x1),  x1 += xamp * sin(2 pi x1),  make_periodic;
 * the host passes xamp = 2/5 A a and vamp = 2/5 A adot with A, a, adot evaluated as the reference does. */
int nnpm_init_pancake(nnpm_ctx *ctx, const int64_t pdims[3], double xamp, double vamp, double m);
/* Counting sort of the local particles by cell (new; K0 of the design). */
int nnpm_sort(nnpm_ctx *ctx);

/* ---- operator-granular entry points (un-fused; for field-level parity) -------------------- */
int nnpm_make_periodic(nnpm_ctx *ctx);                 /* make_periodic  ParticleMesh.jl:409-415 */
int nnpm_deposit(nnpm_ctx *ctx);                       /* deposit        ParticleMesh.jl:289-303,326-407 */
int nnpm_solve(nnpm_ctx *ctx);                         /* compute_potential ParticleMesh.jl:417-505 (mean removal :768-769 folded into the zeroed DC bin) */
int nnpm_accel(nnpm_ctx *ctx);                         /* compute_acceleration ParticleMesh.jl:631-637,507-573 (materialises acc) */
int nnpm_interp(nnpm_ctx *ctx);                        /* interpVecToPoints ParticleMesh.jl:640-721 (materialises pa from acc) */
int nnpm_drift(nnpm_ctx *ctx, double dt);              /* drift          ParticleMesh.jl:723-728 */
int nnpm_kick(nnpm_ctx *ctx, double dt);               /* kick (static)  ParticleMesh.jl:731-736 */
int nnpm_kick_comoving(nnpm_ctx *ctx, double dt, double a, double dadt); /* kick ParticleMesh.jl:800-807 */

/* gd[1].d["ρD"], gd[1].d["Φ"], acc, pa -- Julia layout, local slab / local particles. */
int nnpm_get_field(nnpm_ctx *ctx, int which, double *host_out);
int nnpm_set_field(nnpm_ctx *ctx, int which, const double *host_in);

/* ---- fused steps (the hot path) ----------------------------------------------------------- */
/* Body of evolvePM's loop, src/ParticleMesh.jl:766-778, plus the reductions of :782. */
int nnpm_step_static(nnpm_ctx *ctx, double dt, nnpm_stats *stats);
/* Body of evolvePMCosmology's loop, src/ParticleMesh.jl:853-868, plus the reductions of
 * :875-878.  a_d1 = a(t+dt/4), a_k = a(t+dt/2), adot_k = da/dt(t+dt/2), a_d2 = a(t+3dt/4)
 * come from the host cosmology (src/cosmology.jl:19-42). */
int nnpm_step_comoving(nnpm_ctx *ctx, double dt, double a_d1, double a_k, double adot_k,
                       double a_d2, nnpm_stats *stats);
/* Global (all-rank) maxima of the last step's stats: NCCL allreduce(max) when nranks > 1. */
int nnpm_reduce_stats(nnpm_ctx *ctx, nnpm_stats *stats);

/* Device bytes currently allocated by this context (for sizing reports). */
int64_t nnpm_device_bytes(const nnpm_ctx *ctx);

/* powerSpectrum, src/ParticleMesh.jl:65-100 (+ fftfreq :202-206), on the device-resident
 * particles: deposit -> delta = rho/mean - 1 -> |fft(delta)|^2 * prod(box)/prod(dims)^2 ->
 * shell bins n = round(|k|/dk), dk = 2pi/max(dims).  karray_out/power_out have lenk =
 * round(max(dims)/2) entries; with nranks > 1 every rank receives the global result.
 * Overwrites the mesh (rho/phi).  box = (1,1,1), the only domain the reference tests. */
int nnpm_power_spectrum(nnpm_ctx *ctx, double *karray_out, double *power_out, int32_t lenk);

/* Runtime tuning knobs (new; no reference counterpart).  Unknown names -> NNPM_ERR_ARG.  The defaults are the
 * fastest measured paths; the other values are the library / un-fused compositions the tests cross-check against.
 *   "rowfft"       1 = own row-FFT kernels for the x passes (r2c / c2r, bulk-copy staged; default when N1/2 is
 *                  64..1024), 0 = cuFFT batched 1-D plans
 *   "fft2d"        1 = x and y passes of a plane in ONE persistent kernel with the intermediate held in L2 (work queue
 *                  of row / column items, per-plane completion counters; N1 = N2 in 128..1024); 0 = separate kernels
 *                  (default).  Measured slower than the separate passes: only built with -DNNPM_EXPERIMENTS
 *                  (make EXPERIMENTS=1); the default library returns NNPM_ERR_ARG for 1
 *   "fft2d_block"  planes per block of that kernel (default 4: ~34 MB of intermediate in flight at 1024^2)
 *   "colfft"       1 = TMA-staged column FFT for the y passes (default when N2 is 64..1024), 0 = cuFFT 2-D plans
 *   "zfused"       3 = fused z-FFT * Green * z-iFFT kernel, persistent, next column group loaded by TMA during the
 *                  whole transform (separate landing zone + half-size split exchange; default; N3 = 64..1024),
 *                  2 = the same with one buffer (load overlaps the last pass only; the only form for N3 = 2048),
 *                  0 = cuFFT z + Green kernel
 *   "fft3d"        1 = single cuFFT 3-D plan instead of the passes above (nranks == 1 only)
 *   "tile_deposit" 1 = shared-memory tiled deposit on tile-sorted particles (default), 0 = global atomics
 *   "tile_push"    1 = push gathers phi from a shared-memory tile on sorted particles, 0 = L2 gather
 *   "sort_cells"   1 = the counting sort orders the particles of a tile by cell (bins = mesh cells of the slab; default
 *                  while the slab has <= 2^30 cells), 0 = by tile only (arrival order within a tile)
 *   "bulk_flush"   1 = tiled deposit flushes its box with cp.reduce.async.bulk (default), 0 = per-cell atomics
 *   "l2pf_push"    1 = each CTA of the tiled push pulls its work item's x, v runs into L2 with one
 *                  cp.async.bulk.prefetch.L2 per array before it stages its phi box (default), 0 = off
 *   "l2pf_deposit" 1 = the same in the tiled deposit (default), 0 = off
 *   "fuse_transpose" nranks > 1: 1 = the slab transposes are fused into the transform kernels: the forward y pass
 *                  stores each column group straight into the owning ranks' transposed spectra through per-peer
 *                  TMA tensor maps, the z pass stores its rows straight into the owning ranks' slabs (NVLink peer
 *                  memory, no separate all-to-all pass).  Bit-identical to the other paths, but measured SLOWER:
 *                  the 64 / 128-byte pieces such stores put on NVLink reach ~380 GB/s against ~600 GB/s for the
 *                  copy engines' 1 MB runs; 0 = separate transposes (default).  Only built with
 *                  -DNNPM_EXPERIMENTS (make EXPERIMENTS=1); the default library returns NNPM_ERR_ARG for 1
 *   "peer_transpose" un-fused transposes over the CUDA-IPC peer mappings: 1 = copy-engine peer copies; 2 = one SM
 *                  kernel whose single-thread CTAs relay 8 KB pieces HBM -> shared memory -> peer mesh with bulk
 *                  asynchronous copies (nnpm_xfer.cuh); 3 = copy engines for the chunks of the pipeline that travel
 *                  beside a transform and the SM kernel for the last chunk of each transpose, which nothing is left
 *                  to run beside (default: on the all-to-all of 8 x B200 the copy engines reach ~510 GB/s per GPU,
 *                  the SM kernel ~630; profiles/bench_r02_transport_variants.txt); 0 = grouped ncclSend/ncclRecv
 *                  (also the fallback when CUDA IPC mapping of the peers' meshes fails)
 *   "xfer_ctas"    grid of that SM kernel; 0 = one CTA per SM (default)
 *   "pipeline"     un-fused copy-engine transposes: 1 / 2 = overlap them with chunked transforms whenever the slab
 *                  divides into the chunks (default), 0 = never
 *   "pipeline_chunks" chunks of that pipeline, 1..4; 0 = automatic (default: 4 at nranks = 2, else 2)
 *   "ce_streams"   side streams the peer copies of a transpose are spread over, 1..8; 0 = automatic (default: 8 when
 *                  pipelined over more than 4 ranks, else 4)
 *   "mig_in_push"  nranks > 1: 1 = the push kernels classify the particles they move and pack the leavers for
 *                  migration themselves (default), 0 = separate classification pass over all particles
 *   "zsplit"       1 = run N3 in 128..1024 through the radix-2 wrapper of the persistent fused z pass (the path
 *                  N3 = 2048 always takes); test hook
 *   "keep_rho"     1 = the fused steps copy the deposited density before the solve consumes it, so that
 *                  nnpm_get_field(NNPM_FIELD_RHO) after the step returns what gd[1].d["ρD"] holds in the reference
 *                  (the host sets it only for steps after which an output is due); 0 = off (default)
 *   "cfrun_y"      launch-shape experiment of the y pass */
int nnpm_set_option(nnpm_ctx *ctx, const char *name, double value);

#ifdef __cplusplus
}
#endif
#endif /* NNPM_H */
