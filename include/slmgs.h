/* slmgs.h -- C ABI of libslmgs.so: the B200-native GS / WGS hologram loop.
 *
 * This is the drop-in boundary for ONE hot path of slmsuite (v0.4.1 @ 39243f08):
 * `Hologram.optimize` / `SpotHologram.optimize` -> `optimize_gs`
 * (slmsuite/holography/algorithms/_hologram.py:1351-1368, :1427-1493).  slmsuite has no FFI
 * of its own: its only backend seam is the module alias `cp` (algorithms/_header.py:16-32)
 * and subclass overrides of `_nearfield2farfield` / `_gs_farfield_routines` /
 * `_farfield2nearfield` / `_update_weights`.  Each entry point below cites the reference
 * method it replaces; INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all host buffers are C-contiguous and borrowed for the
 *     duration of the call (nothing is retained);
 *   - images are exchanged in the REFERENCE's centred (fftshift-ed) convention, shape
 *     [batch][H][W]; SLM-sized arrays are [batch][h][w]; the library stores them rolled;
 *   - every function returns 0 on success or a negative slmgs_status; the message is
 *     available from slmgs_last_error(ctx);
 *   - a context is bound to one device and one stream and is NOT thread-safe; different
 *     contexts may be driven from different threads;
 *   - compute entry points are asynchronous on the context's stream; getters synchronise.
 */
#ifndef SLMGS_H
#define SLMGS_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SLMGS_API __attribute__((visibility("default")))
#else
#define SLMGS_API
#endif

typedef struct slmgs_ctx slmgs_ctx;

typedef enum {
    SLMGS_OK = 0,
    SLMGS_ERR_INVALID = -1, /* bad argument / unsupported shape */
    SLMGS_ERR_CUDA = -2,    /* CUDA runtime error */
    SLMGS_ERR_OOM = -3,     /* allocation failed */
    SLMGS_ERR_STATE = -4,   /* call sequence error (e.g. constrain without forward) */
    SLMGS_ERR_NCCL = -5     /* NCCL could not be loaded / a collective failed (slmgs_comm_last_error) */
} slmgs_status;

/* ALGORITHM_DEFAULTS keys, algorithms/_header.py:53-71 */
typedef enum {
    SLMGS_GS = 0,
    SLMGS_WGS_LEONARDO = 1,
    SLMGS_WGS_KIM = 2,
    SLMGS_WGS_NOGRETTE = 3,
    SLMGS_WGS_WU = 4,
    SLMGS_WGS_TANH = 5
} slmgs_method;

/* how the far-field phase is obtained in the constraint, _hologram.py:1556-1605 */
typedef enum {
    SLMGS_PHASE_COMPUTE = 0,       /* phase_ff = angle(farfield), not kept */
    SLMGS_PHASE_COMPUTE_STORE = 1, /* phase_ff = angle(farfield), kept (WGS-Kim: iteration that fixes) */
    SLMGS_PHASE_STORED = 2         /* use the stored phase_ff (fixed_phase) */
} slmgs_phase_mode;

/* flags of one iteration (subset of Hologram.flags that the device needs) */
typedef struct {
    int method;             /* slmgs_method */
    int update_weights;     /* "WGS" in method and iter > 0, _hologram.py:1552 */
    int phase_mode;         /* slmgs_phase_mode */
    float feedback_exponent; /* flags["feedback_exponent"] */
    float feedback_factor;   /* flags["feedback_factor"] */
    int mraf;               /* target has NaN noise region, _hologram.py:1495-1548 */
    int mraf_has_factor;    /* flags["mraf_factor"] is not None */
    float mraf_factor;
    int feedback;           /* 0: pixel feedback ("computational"); 1: per-spot window feedback ("computational_spot") */
    int spot_width;         /* spot_integration_width_knm, _spots.py:1292-1297 (feedback == 1) */
    int zero_weights;       /* MRAF zero-region accumulator in use (hasattr(self, "zero_weights")), _hologram.py:1511-1515 */
    float zero_factor;      /* flags.get("zero_factor", 1), :1615 */
} slmgs_params;

SLMGS_API int slmgs_version(void);
/* message of the last failure on ctx (ctx == NULL: last failure of slmgs_create) */
SLMGS_API const char* slmgs_last_error(const slmgs_ctx* ctx);

/* Hologram.__init__ state, _hologram.py:196-478.  shape = (H, W) powers of two in [16, 8192];
 * slm_shape = (h, w) <= shape, any parity; batch >= 1 independent holograms. */
SLMGS_API int slmgs_create(slmgs_ctx** out, int device, int batch, int H, int W, int h, int w);
SLMGS_API int slmgs_destroy(slmgs_ctx* ctx);
SLMGS_API int slmgs_sync(slmgs_ctx* ctx);
/* device pointer of the [batch][h][w] float32 phase buffer (for an NCCL all-gather of final phases) */
SLMGS_API void* slmgs_phase_device_ptr(slmgs_ctx* ctx);
SLMGS_API void* slmgs_stream(slmgs_ctx* ctx);
/* PCI bus id ("0000:1b:00.0", NUL terminated) of CUDA device `device`: lets a multi-GPU host place its threads and pinned
 * buffers on the NUMA node next to the GPU (slmsuite_b200/_lib.py bind_host_to_device).  Host utility of the one-process-
 * per-GPU deployment; the reference is single GPU and has no counterpart. */
SLMGS_API int slmgs_device_pci_bus_id(int device, char* out, int len);

/* ---- state upload / download ------------------------------------------------------------ */
SLMGS_API int slmgs_set_phase(slmgs_ctx*, const float* phase);            /* reset_phase, :570-601; [B][h][w] */
SLMGS_API int slmgs_get_phase(slmgs_ctx*, float* phase);                  /* raw phase (get_phase adds pi on the host, :786-811) */
SLMGS_API int slmgs_set_amp_scalar(slmgs_ctx*, float amp);                /* amp=None -> 1/sqrt(h w), :401-402 */
SLMGS_API int slmgs_set_amp_array(slmgs_ctx*, const float* amp, int per_hologram); /* [h][w] or [B][h][w], already L2-normalised, :404-405 */
SLMGS_API int slmgs_set_propagation(slmgs_ctx*, const float* kernel);     /* [h][w] or NULL, :408-415 */
SLMGS_API int slmgs_set_target(slmgs_ctx*, const float* target, int shared); /* normalised target, NaN = MRAF noise, :760-766; [B][H][W] or [H][W] */
SLMGS_API int slmgs_get_target(slmgs_ctx*, float* target);
SLMGS_API int slmgs_reset_weights(slmgs_ctx*);                            /* weights = nan_to_num(target, nan=0), :603-614 */
SLMGS_API int slmgs_set_weights(slmgs_ctx*, const float* weights);        /* set_weights, :820-840 */
SLMGS_API int slmgs_get_weights(slmgs_ctx*, float* weights);
SLMGS_API int slmgs_set_phase_ff(slmgs_ctx*, const float* phase_ff);
SLMGS_API int slmgs_get_phase_ff(slmgs_ctx*, float* phase_ff);
SLMGS_API int slmgs_get_amp_ff(slmgs_ctx*, float* amp_ff);
SLMGS_API int slmgs_get_farfield(slmgs_ctx*, float* farfield_c64);        /* interleaved re/im, ortho-scaled */

/* "next" row of the path: the gray levels SLM.set_phase() would display for get_phase(),
 * hardware/slms/slm.py:636-690 + _phase2gray :695-743 (phase_scaling == 1): out[b][h][w] uint8 (bitdepth <= 8) or
 * uint16; correction = optional float64 [h][w] wavefront correction (source["phase"]) or NULL. */
SLMGS_API int slmgs_get_phase_gray(slmgs_ctx*, int bitdepth, const double* correction, void* out);

/* "next" row: what a simulated camera sees, SimulatedCamera._get_image_hw,
 * hardware/cameras/simulated.py:344-402: out[b][i] = clip(|farfield|^2 sampled at (ky[i], kx[i]) with
 * scipy.ndimage.map_coordinates(order=0) semantics (nearest pixel, 0 outside [0, len-1]) * scale, clip_max), cast to
 * out_kind (0 float32, 1 uint8, 2 uint16; clip_max < 0 = no clipping).  Coordinates are float64 pixel coordinates of
 * the centred far field (the reference's knm_cam, simulated.py:180-185); the grid stays on the device between calls.
 * The far field is recomputed from the current phase (this refreshes amp_ff like get_farfield, _hologram.py:922-929). */
SLMGS_API int slmgs_set_sample_grid(slmgs_ctx*, long long n, const double* ky, const double* kx);
SLMGS_API int slmgs_sample_intensity(slmgs_ctx*, float scale, float clip_max, int out_kind, void* out);

/* ---- fused loop -------------------------------------------------------------------------- */
/* optimize_gs with callback=None and no per-iteration statistics (:1465-1493):
 * n_iter iterations, then _populate_results (:934-949).  params[i] are the flags of iteration i
 * (the host runs the WGS-Kim state machine of :1556-1585, which only depends on the iteration
 * count in this mode).  GS, WGS-Leonardo/Kim/Wu/tanh with pixel feedback and MRAF with GS run two kernels
 * per iteration; weight updates with a global dependency inside the iteration (WGS-Nogrette's mean, per-spot
 * window feedback, MRAF + WGS) run a forward column pass for |farfield| first, then the update kernels, then
 * the fused kernels.  Callbacks and per-iteration statistics go through the stepped entry points. */
SLMGS_API int slmgs_run(slmgs_ctx*, const slmgs_params* params, int n_iter, int populate);

/* Sparse far field in slmgs_run (mode 1 = automatic, the default; 0 = always dense).  The constrained far field
 * weights * exp(i phase_ff) (_hologram.py:1601-1605) is zero wherever weights == 0, so column tiles with all-zero
 * weights (no MRAF noise pixel, no spot-feedback window) are skipped by the column kernels and their columns are
 * neither stored nor loaded by the row kernels: identical results, several times faster on spot targets (the
 * motivation of the reference's CompressedSpotHologram, _spots.py:222-241).  The occupancy is recomputed on the
 * device after every target / weights upload, per hologram of a batch.  slmgs_sparse_info: out[0] = last slmgs_run
 * was sparse, out[1] = active column tiles (of the hologram with the most), out[2] = column tiles. */
SLMGS_API int slmgs_set_sparse(slmgs_ctx*, int mode);
SLMGS_API int slmgs_sparse_info(const slmgs_ctx*, int* out3);

/* ---- stepped loop (callbacks, statistics, Nogrette, spot feedback, MRAF + WGS) ------------- */
SLMGS_API int slmgs_forward(slmgs_ctx*);                                   /* _nearfield2farfield + _midloop_cleaning, :1038-1056, :951-959 */
SLMGS_API int slmgs_update_weights(slmgs_ctx*, const slmgs_params*);       /* Hologram._update_weights, pixel feedback, :1914-1922 -> :1822-1879 */
SLMGS_API int slmgs_set_spots(slmgs_ctx*, int n, const int* x, const int* y, const float* spot_amp); /* spot_knm_rounded, spot_amp, _spots.py:1490-1546 */
SLMGS_API int slmgs_update_weights_spot(slmgs_ctx*, const slmgs_params*, int width); /* SpotHologram._update_weights, _spots.py:1573-1624 */
SLMGS_API int slmgs_constrain_inverse(slmgs_ctx*, const slmgs_params*);    /* _gs_farfield_routines (constraint part) + _farfield2nearfield, :1587-1653, :1058-1073 */
SLMGS_API int slmgs_populate(slmgs_ctx*);                                  /* _populate_results, :934-949 */

/* ---- MultiplaneHologram ("next" row; _multiplane.py:255-286) ---------------------
 * N child contexts (same slm_shape, same device, any padded shape) share one near-field phase.  Per iteration each
 * child runs slmgs_forward, its own weight update, then slmgs_constrain_accumulate, which applies the far-field
 * constraint, transforms back and ADDS weight * nearfield[crop] * exp(-i kernel) into `sum` ([batch][h][w] complex64
 * on the device, obtained from slmgs_nearfield_sum_ptr of any child); slmgs_extract_phase_from_sum then sets
 * phase = arctan2(sum) in every child.  Children must share a stream (slmgs_share_stream). */
SLMGS_API int slmgs_share_stream(slmgs_ctx* ctx, slmgs_ctx* leader);
SLMGS_API void* slmgs_nearfield_sum_ptr(slmgs_ctx* ctx);
SLMGS_API int slmgs_constrain_accumulate(slmgs_ctx*, const slmgs_params*, float weight, void* sum, int first);
SLMGS_API int slmgs_extract_phase_from_sum(slmgs_ctx*, const void* sum);
/* one fused iteration of a child (no callback / statistics): row first + fused column kernel + accumulating row inverse */
SLMGS_API int slmgs_run_accumulate(slmgs_ctx*, const slmgs_params*, float weight, void* sum, int first);

/* ---- statistics ---------------------------------------------------------------------------- */
/* _calculate_stats(amp_ff, target) pieces, _stats.py:7-116.  out[b][8] =
 * {sum f^2, nansum t^2, nansum t f, ratio min, ratio max, err min, err max, err sum} and
 * out2[b][2] = {err sum of squares, masked count}.  The host finishes the formulas. */
SLMGS_API int slmgs_stats_pixel(slmgs_ctx*, double* out8, double* out2);
/* analysis.take(amp_ff^2, centres, width, centered, integrate), analysis/__init__.py:61-204:
 * out[b][n] float64 window powers; also total[b] = sum(amp_ff^2) */
SLMGS_API int slmgs_window_power(slmgs_ctx*, int n, const int* x, const int* y, int width, double* out, double* total);

/* device-side snapshot / restore of the near-field phase (re-run from the same start without H2D) */
SLMGS_API int slmgs_save_phase(slmgs_ctx*);
SLMGS_API int slmgs_restore_phase(slmgs_ctx*);

/* ---- timing (CUDA events on the context's stream; torch.cuda.Event cannot see this stream) --- */
SLMGS_API int slmgs_timer_start(slmgs_ctx*);
SLMGS_API int slmgs_timer_stop(slmgs_ctx*, float* ms); /* synchronises */
/* per-kernel profile: when enabled, every FFT kernel launch of slmgs_run / stepped calls is bracketed by
 * events.  slmgs_profile_read synchronises and returns, for class k (0 row first, 1 row fused, 2 row last,
 * 3 column forward, 4 column fused, 5 column inverse): ms[k] = summed duration, count[k] = launches; resets. */
SLMGS_API int slmgs_profile_enable(slmgs_ctx*, int on);
SLMGS_API int slmgs_profile_read(slmgs_ctx*, float* ms6, int* count6);

/* ---- introspection (tests / bench) ---------------------------------------------------------- */
/* number of kernels this context has launched since creation */
SLMGS_API long long slmgs_launch_count(const slmgs_ctx*);
/* launch geometry chosen for the two fused kernels: out[0..3] = row threads, row grid x, col threads, col grid x */
SLMGS_API int slmgs_launch_geometry(const slmgs_ctx*, int* out4);
/* time n back-to-back launches of one kernel with CUDA events on the context's stream.
 * which: 0 = row fused, 1 = column fused (GS), 2 = row first, 3 = column forward (populate). ms_out = average ms per launch */
SLMGS_API int slmgs_time_kernel(slmgs_ctx*, int which, int n, float* ms_out);

/* ---- compressed spot hologram ("next" row) --------------------------------------------
 * CompressedSpotHologram, slmsuite/holography/algorithms/_spots.py:178-1019: instead of a DFT grid every spot n owns a
 * phase kernel phi_n(pix) = sum_d spot_zernike[d, n] Z_d(x_pix, y_pix) (_spots.py:595-636) and the maps of the GS loop
 * are direct sums over pixels / spots (the reference's NumPy matmul pair :767-824 / :887-915 and its CUDA pair
 * toolbox/cuda.cu:95-288).  The host evaluates the M <= 10 basis functions once: mono[m][pix] = Z_m(x_pix, y_pix)
 * (float64, aperture-scaled grid; any functions whose weighted sum is the phase) and cw[m][n] = spot_zernike[m, n]
 * (float64, radians).
 * Targets / weights / far field are N-vectors; the target may hold NaN (MRAF noise point) and 0 (null point).
 * slmgs_params as for slmgs_run (feedback is always the computed spot amplitude, _spots.py:950-989). */
typedef struct slmgs_comp slmgs_comp;
SLMGS_API const char* slmgs_comp_last_error(const slmgs_comp*);
SLMGS_API int slmgs_comp_create(slmgs_comp** out, int device, int h, int w, int n_spots, int n_basis);
SLMGS_API int slmgs_comp_destroy(slmgs_comp*);
SLMGS_API int slmgs_comp_sync(slmgs_comp*);
SLMGS_API long long slmgs_comp_launch_count(const slmgs_comp*);
SLMGS_API int slmgs_comp_set_basis(slmgs_comp*, const double* mono, const double* cw);   /* _build_cupy_kernel_batched, :595-636 */
SLMGS_API int slmgs_comp_set_phase(slmgs_comp*, const float* phase);                     /* [h][w] */
SLMGS_API int slmgs_comp_get_phase(slmgs_comp*, float* phase);
SLMGS_API int slmgs_comp_set_amp_scalar(slmgs_comp*, float amp);
SLMGS_API int slmgs_comp_set_amp_array(slmgs_comp*, const float* amp);                   /* [h][w] */
SLMGS_API int slmgs_comp_set_target(slmgs_comp*, const float* target);                   /* [N], set_target :917-948 */
SLMGS_API int slmgs_comp_set_weights(slmgs_comp*, const float* weights);
SLMGS_API int slmgs_comp_get_weights(slmgs_comp*, float* weights);
SLMGS_API int slmgs_comp_set_phase_ff(slmgs_comp*, const float* phase_ff);
SLMGS_API int slmgs_comp_get_phase_ff(slmgs_comp*, float* phase_ff);
SLMGS_API int slmgs_comp_get_amp_ff(slmgs_comp*, float* amp_ff);
SLMGS_API int slmgs_comp_get_farfield(slmgs_comp*, float* farfield_c64);                 /* [N] interleaved re/im, normalised (:822) */
SLMGS_API int slmgs_comp_forward(slmgs_comp*, int populate);                              /* _nearfield2farfield :677-708 + amp_ff; populate: also phase_ff (_populate_results) */
SLMGS_API int slmgs_comp_run(slmgs_comp*, const slmgs_params* params, int n_iter, int populate); /* optimize_gs with the compressed maps */
/* The loop in pieces, for ONE hologram whose pixels are sharded over several GPUs (each rank: a context over its slab
 * of SLM rows, all N spots).  near -> far sums over pixels, so the accumulators are partial sums: all-reduce the
 * [N][2] float64 buffer at slmgs_comp_facc_ptr (16 N bytes) between slmgs_comp_near2far and
 * slmgs_comp_constrain_far2near (or slmgs_comp_finalize for a forward / _populate_results); every rank then runs the
 * identical N-vector stage and projects its own slab. */
SLMGS_API int slmgs_comp_near2far(slmgs_comp*);
SLMGS_API void* slmgs_comp_facc_ptr(slmgs_comp*);
SLMGS_API void* slmgs_comp_stream(slmgs_comp*);
SLMGS_API int slmgs_comp_constrain_far2near(slmgs_comp*, const slmgs_params*);
SLMGS_API int slmgs_comp_finalize(slmgs_comp*, int populate);
SLMGS_API int slmgs_comp_timer(slmgs_comp*, int start, float* ms);                        /* CUDA events on the context's stream */

/* ---- multi-GPU: the one collective of the sharded batch path --------------------------------------
 * A batch of independent holograms shards across GPUs by hologram with no communication inside the loop
 * (the reference has no batch: a list of Hologram objects); the job ends with ONE all-gather of the final
 * near-field phases.  NCCL is loaded with dlopen inside the library (no torch.distributed).  Rank 0 creates the unique
 * id, the caller distributes its 128 bytes (slmsuite_b200/comm.py: TCP rendezvous on MASTER_ADDR / MASTER_PORT), every
 * rank creates its communicator.  One process per GPU. */
typedef struct slmgs_comm slmgs_comm;
SLMGS_API const char* slmgs_comm_last_error(void);
SLMGS_API int slmgs_comm_nccl_version(void);                                 /* ncclGetVersion, -1 if NCCL is unavailable */
SLMGS_API int slmgs_comm_unique_id(unsigned char* out128);                   /* ncclGetUniqueId */
SLMGS_API int slmgs_comm_create(slmgs_comm** out, const unsigned char* id128, int rank, int world, int device);
SLMGS_API int slmgs_comm_destroy(slmgs_comm*);
/* every rank contributes per_rank holograms of `elems` floats (ctx holds n_local <= per_rank: a short or empty last shard
 * is zero padded; ctx may be NULL when n_local == 0) and receives world * per_rank of them: out_host (or NULL) gets the
 * gathered array, *out_dev (or NULL) its device address (owned by the communicator), *ms (or NULL) the device time of
 * the collective.  Runs on the context's stream behind the loop's kernels. */
SLMGS_API int slmgs_allgather_phase(slmgs_ctx*, slmgs_comm*, int n_local, int per_rank, long long elems,
                                    float* out_host, void** out_dev, float* ms);
/* in-place sum over ranks of `count` float64 values at a device address, ordered on `stream` (the 16 N-byte exchange of a
 * pixel-sharded compressed spot hologram, slmgs_comp_facc_ptr) */
SLMGS_API int slmgs_comm_allreduce_f64(slmgs_comm*, void* dev_ptr, long long count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLMGS_H */
