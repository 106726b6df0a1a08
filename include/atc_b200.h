/*
 * atc_b200.h — C ABI of the B200-native batched ATC approach-control environment step.
 *
 * This is the drop-in boundary for the reference's hot path AtcGym.step()/reset()
 * (/root/reference/envs/atc/atc_gym.py:128-192, :337-365 and what they call in envs/atc/model.py).
 * The reference is pure Python and has no FFI; the binding a maintainer adds is the ctypes stub shown in
 * INTEGRATION.md (it is what atc_reinforcement_learning_b200/_native.py does).
 *
 * Rules of the boundary
 *   - extern "C", plain pointers and sizes, no torch / C++ types in any signature.
 *   - The CALLER owns every device buffer (PyTorch allocates them); the library allocates only its private copy
 *     of the static sector at atc_create() and never allocates or frees afterwards, so atc_step()/atc_rollout()
 *     are CUDA-graph capturable.  Kernels run on the caller-supplied stream (cudaStream_t passed as void*).
 *   - No exceptions, no stdout: every entry point returns 0 on success or a negative AtcStatus;
 *     atc_last_error() returns the message of the last failure on that handle (or the global one for atc_create).
 *   - Invalid *actions* are not errors: like the reference (atc_gym.py:312-315) they cost the env -1.0 each.
 *   - One handle per device; calls on one handle must be serialised by the caller; distinct handles are independent.
 *
 * Layouts (all row-major, contiguous)
 *   state        double [5][n_env*n_ac]        planes x, y, h, phi, v        (reference: Airplane attributes, model.py:32-40)
 *   actions      float  [n_env][n_ac][3]       v, h, phi in [-1, 1] (or MultiDiscrete indices as floats)
 *   obs, raw_obs float  [n_env][n_ac][10]      atc_gym.py:262-277 ; raw_obs = info["original_state"]
 *   reward       float  [n_env]                sum over the env's aircraft (atc_gym.py:137-185)
 *   done         uint8  [n_env]
 *   term         int32  [n_env]                bits 0-7 env code, bits 8+3a..10+3a per-aircraft code (AtcTermCode)
 *   rollout buffers carry a leading [T] dimension.
 */
#ifndef ATC_B200_H
#define ATC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATC_ABI_VERSION 5
#define ATC_MAX_AIRCRAFT 8
#define ATC_MAX_MVA 31
#define ATC_OBS_DIM 10

typedef enum AtcStatus {
    ATC_OK = 0,
    ATC_ERR_INVALID_ARGUMENT = -1,
    ATC_ERR_CUDA = -2,
    ATC_ERR_UNSUPPORTED = -3
} AtcStatus;

/* per-aircraft / per-env termination codes (reference branches: atc_gym.py:149-153, 156-161, 163-169, 171-173) */
typedef enum AtcTermCode {
    ATC_TERM_RUNNING = 0,
    ATC_TERM_BELOW_MVA = 1,
    ATC_TERM_LEFT_AIRSPACE = 2,
    ATC_TERM_CAPTURED = 3,
    ATC_TERM_TIMEOUT = 4,
    ATC_TERM_SEPARATION = 5
} AtcTermCode;

/* Static sector, compiled on the host (atc_reinforcement_learning_b200/sector.py).  All pointers are HOST pointers;
 * atc_create() copies what it needs to the device.  Replaces scenarios.py:7-207 + the constants derived in
 * Corridor.__init__ (model.py:155-186) and AtcGym.__init__ (atc_gym.py:49-58, 88-110). */
typedef struct AtcSectorDesc {
    int32_t n_mva;                 /* <= ATC_MAX_MVA */
    int32_t n_vertices;            /* total closed-ring vertices */
    const double *ring_xy;         /* [n_vertices][2], rings closed, reference vertex order */
    const int32_t *ring_off;       /* [n_mva + 1] */
    const double *mva_height;      /* [n_mva] ft */
    const double *mva_bounds;      /* [n_mva][4] minx, miny, maxx, maxy (model.py:268) */
    double runway_x, runway_y, runway_h, phi_to_runway;
    double faf[2];                 /* model.py:172 */
    double normal[2];              /* model.py:165 */
    double tri_h[8], tri_1[8], tri_2[8];   /* closed 4-vertex rings: corridor_horizontal / corridor1 / corridor2 */
    double sin_to_runway, cos_to_runway;   /* rot_matrix(phi_to_runway).[0,1]^T  (model.py:219) */
    double glide_tan;              /* tan(3 deg)  (model.py:205) */
    double bbox[4];                /* union bounds (model.py:294-306) */
    double world_max_distance;     /* atc_gym.py:58 */
    double faf_mva;                /* atc_gym.py:49 */
    float norm_min[ATC_OBS_DIM];   /* atc_gym.py:88-98 */
    float norm_max[ATC_OBS_DIM];   /* atc_gym.py:99-110 */
    int32_t n_entry;               /* scenarios.py:192-207 */
    const double *entry_xyphi;     /* [n_entry][3] */
    const int32_t *level_off;      /* [n_entry + 1] */
    const int32_t *levels;         /* flight levels (x100 ft) */
    /* exact MVA lookup accelerator: uniform grid whose cell (0, 0) starts at (grid_x0, grid_y0), two
       cells outside the bbox, and which ends two cells beyond it; the outermost ring of cells must be "outside" (0).
       The kernels pick the cell from float32 coordinates, so every cell's entry must hold on the cell grown by the
       float32 index error (sector.py `margin`). */
    int32_t grid_nx, grid_ny;
    double grid_inv_cell;
    double grid_x0, grid_y0;
    const uint16_t *grid_cell;     /* [grid_ny][grid_nx]; bit15 clear: 0 = outside, k = polygon k-1 for the whole cell;
                                      bit15 set: bits 0-14 = index into grid_prog_off (a cell an edge passes near) */
    int32_t n_mixed;               /* entries of grid_prog_off */
    int32_t n_prog;                /* entries of grid_prog */
    const uint32_t *grid_prog_off; /* [n_mixed]: bits 26-30 = candidate polygons, bits 0-25 = offset into grid_prog */
    const uint16_t *grid_prog;     /* per candidate polygon, in list order: header (bits 0-4 polygon, bit 5 parity of
                                      the edges that always cross, bit 6 bbox test needed, bits 8-15 edge count)
                                      followed by the ring-vertex indices of the edges to test exactly */
    const double *grid_line;       /* [n_mixed][4]: a, b, c of the single boundary line crossing the cell (a = b = 0:
                                      none) with a*a + b*b = 1, and two int32 in the 4th double: answer (polygon + 1,
                                      0 = outside) on the positive / negative side.  Points farther than 1e-9 nm from
                                      the line take the answer, the others run the exact program. */
    /* wind extension (not in the reference; README.md:64) — NULL / 0 = calm */
    int32_t wind_gx, wind_gy;
    const float *wind;             /* [wind_gy][wind_gx][2] knots (east, north), nodes on the bbox corners */
    /* optional compact grid, NULL = none: a coarse, two-level copy of the accelerator small enough
       to live in the shared memory of one SM (at most atc_compact_grid_budget() bytes for cells + lines); the rollout
       kernel with one CTA per SM reads it instead of the fine grid.  Origin / padding conventions as the fine grid;
       the kernel takes the float32 index at 1/8 of the cell size (the sub-block resolution) and shifts it down by 3.
       coarse cell: bit 15 clear = polygon + 1 for the whole cell (0 = outside); bit 15 set, line id (bits 0-6) < 126 =
       one boundary line crosses the cell, bits 7-10 / 11-14 = answer (polygon + 1) on its positive / negative side;
       line id 126 / 127 = refined by the 8 x 8 sub-block number (id & 1) * 256 + bits 7-14 (a block number >=
       cgrid_n_blocks: undecidable).  Sub-block cells: same encoding, line id 127 = undecidable.  Points within 1e-9 nm
       of their line and undecidable cells are resolved through the fine grid. */
    int32_t cgrid_nx, cgrid_ny;
    double cgrid_inv_cell;         /* 1 / size of the SUB-block cells (= 8 / coarse cell size) */
    double cgrid_x0, cgrid_y0;
    const uint16_t *cgrid_cell;    /* [cgrid_ny * cgrid_nx] coarse cells, then [cgrid_n_blocks][8][8] sub-blocks */
    int32_t cgrid_n_blocks;        /* <= 511 */
    int32_t n_cline;               /* <= 126 */
    const double *cline;           /* [n_cline][4]: a, b, c (a*a + b*b = 1), 0 */
} AtcSectorDesc;

/* model.py:132-145 SimParameters + batch geometry */
typedef struct AtcSimParams {
    double timestep;               /* seconds */
    int32_t reward_shaping;
    int32_t normalize_state;
    int32_t discrete_action_space;
    int32_t normalize_reset_obs;   /* 0 = reference behaviour: reset() returns the RAW observation (atc_gym.py:351,365) */
    int32_t n_env;
    int32_t n_aircraft;            /* 1..ATC_MAX_AIRCRAFT */
    int32_t track_actions;         /* maintain last_action / actions_taken (atc_gym.py:305-311) */
    int32_t exact_math;            /* 1: observation and reward shaping in float64 like the reference (slow);
                                      0: float32 observation/shaping, float64 state and decisions (default) */
    uint64_t seed;                 /* spawn RNG key */
    int64_t env_index_base;        /* global index of local env 0 (multi-GPU sharding) */
} AtcSimParams;

/* Persistent per-env state, DEVICE pointers owned by the caller. */
typedef struct AtcBuffers {
    double *state;                 /* [5][n_env*n_ac] */
    double *last_action;           /* [3][n_env*n_ac] (track_actions) or NULL */
    int32_t *timesteps;            /* [n_env]  atc_gym.py:39,135 */
    int32_t *episodes;             /* [n_env]  episodes started (spawn RNG counter) */
    double *ep_return;             /* [n_env]  total_reward (atc_gym.py:196) */
    int32_t *actions_taken;        /* [n_env] (track_actions) or NULL */
    double *last_ep_return;        /* [n_env]  return of the last finished episode (what the NCCL gather ships) */
    int32_t *last_ep_len;          /* [n_env] */
    int32_t *win_ring;             /* [n_env]  bit k = outcome of the k-th last finished episode (atc_gym.py:359-363) */
} AtcBuffers;

/* Per-call inputs/outputs, DEVICE pointers (atc_step / atc_rollout) or HOST pointers (the *_host variants). */
typedef struct AtcStepIO {
    const float *actions;          /* [T][n_env][n_ac][3] */
    float *obs;                    /* [T][n_env][n_ac][10] */
    float *raw_obs;                /* same shape or NULL */
    float *reward;                 /* [T][n_env] */
    uint8_t *done;                 /* [T][n_env] */
    int32_t *term;                 /* [T][n_env] or NULL */
} AtcStepIO;

typedef struct AtcHandle AtcHandle;

int atc_abi_version(void);

/* Bytes of shared memory the one-CTA-per-SM rollout kernel can give to a compact grid (cells + lines). */
int64_t atc_compact_grid_budget(void);

/* Replaces AtcGym.__init__ (atc_gym.py:28-115) for a batch: validates, copies the sector to `device`. */
int atc_create(const AtcSectorDesc *sector, const AtcSimParams *params, int device, AtcHandle **out);
int atc_destroy(AtcHandle *h);

/* Replaces AtcGym.reset (atc_gym.py:337-365) for the envs whose mask byte is non-zero (mask NULL = all).
 * spawn NULL: entry point / level drawn by the counter-based device RNG; else explicit [n_env][n_ac][5]
 * (x, y, h, phi, v) device array — the hook parity tests use to inject the reference's own spawn choices.
 * obs (may be NULL) receives the reset observation of the reset envs only. */
int atc_reset(AtcHandle *h, const AtcBuffers *b, const uint8_t *mask, const double *spawn, float *obs, void *stream);

/* Replaces AtcGym.step (atc_gym.py:128-192): one fused kernel launch advancing every env by one step.
 * autoreset != 0: finished envs are re-spawned inside the same launch and obs holds their reset observation
 * (VecEnv convention); raw_obs always holds the pre-reset state. */
int atc_step(AtcHandle *h, const AtcBuffers *b, const AtcStepIO *io, int autoreset, void *stream);

/* T fused steps in ONE launch, state held in registers, autoreset on; io buffers carry the leading [T]. */
int atc_rollout(AtcHandle *h, const AtcBuffers *b, const AtcStepIO *io, int n_steps, void *stream);

/* End-to-end variants: io holds HOST pointers (pinned for full speed).  Actions are copied host->device, the
 * step/rollout runs, results are copied device->host, all on `stream`, using the device staging buffers in
 * `dev_io` (same shapes, caller-owned).  Returns after the stream is synchronised. */
int atc_step_host(AtcHandle *h, const AtcBuffers *b, const AtcStepIO *host_io, const AtcStepIO *dev_io, int autoreset,
                  void *stream);
int atc_rollout_host(AtcHandle *h, const AtcBuffers *b, const AtcStepIO *host_io, const AtcStepIO *dev_io, int n_steps,
                     void *stream);

/* Geometry probes used by the parity tests (device arrays): MVA height in ft or -1 outside (model.py:282-292),
 * and Runway.inside_corridor (model.py:188-231).  xy [n][2]; xyhphi [n][4]. */
int atc_query_mva(AtcHandle *h, int n, const double *xy, int32_t *out, void *stream);
int atc_query_corridor(AtcHandle *h, int n, const double *xyhphi, uint8_t *out, void *stream);

/* Next-row component: running observation statistics + normalisation + finiteness check on the
 * device, the role stable-baselines' VecNormalize / VecCheckNan play around the env in the reference's tuner
 * (/root/reference/learning/tune_hyperparameters.py:94-97).  Stateless: all buffers are caller-owned device memory.
 *   x        float [n_rows][dim]
 *   rms      double [2*dim + 1]   running mean[dim], var[dim], count   (initialise: mean 0, var 1, count 1e-4)
 *   scratch  double [2*dim + 2]   zero-initialised by the caller once; left zeroed by every call
 *   nonfinite int32 [1]           set to 1 if x holds a NaN / Inf (never cleared by the library)
 * atc_obs_stats_update merges the batch moments into rms (Chan et al. parallel update, float64);
 * atc_obs_normalize writes clip((x - mean) / sqrt(var + epsilon), -clip, clip). */
int atc_obs_stats_update(const float *x, int64_t n_rows, int32_t dim, double *rms, double *scratch, int32_t *nonfinite,
                         void *stream);
int atc_obs_normalize(const float *x, int64_t n_rows, int32_t dim, const double *rms, double epsilon, double clip,
                      float *out, void *stream);

/* Fused VecNormalize + VecCheckNan (same next-row component; csrc/atc_vecnorm.cu): ONE cooperative launch applies
 * stable-baselines' VecNormalize.step_wait to n_steps consecutive env steps — per step: ret = ret * gamma + reward;
 * running moments of the observation rows and of ret updated with the step's batch (training); obs_out =
 * clip((obs - mean) / sqrt(var + epsilon)), reward_out = clip(reward / sqrt(ret_var + epsilon)); ret[done] = 0 —
 * as two streaming passes over the rows with one grid-wide barrier in between (the batch totals of the steps are
 * independent; only their 22-double merge is sequential).
 * All buffers are caller-owned device memory on `device`; obs_out may alias obs_in, reward_out may alias reward_in.
 *   obs_in / obs_out     float [n_steps][n_env][n_aircraft][10], 16-byte aligned
 *   reward_in / _out     float [n_steps][n_env], done uint8 [n_steps][n_env]; reward_in = NULL skips the reward path
 *                        (the reset() observation: stable-baselines updates and normalises the observation only)
 * The state persists between calls; initialise obs_rms = {0 x10, 1 x10, 1e-4}, ret_rms = {0, 1, 1e-4}, ret and sync 0.
 * n_steps <= atc_vecnorm_max_steps(device) per call (longer rollouts: several calls, the state carries over). */
typedef struct AtcVecNormState {
    double *obs_rms;       /* [21] running mean[10], var[10], count */
    double *ret_rms;       /* [3]  running mean, var, count of the discounted return */
    double *ret;           /* [n_env] discounted return accumulators (may be NULL when reward_in is never given) */
    double *scratch;       /* [scratch_doubles] work space, contents irrelevant between calls */
    int64_t scratch_doubles; /* >= atc_vecnorm_scratch_doubles(device, n_steps, n_env, n_aircraft) */
    uint32_t *sync;        /* [4] grid barrier (count, generation), barrier time-out flag, reserved */
    int32_t *nonfinite;    /* [1] set to 1 when an observation or reward is NaN / Inf (never cleared by the library) */
} AtcVecNormState;
typedef struct AtcVecNormParams {
    int32_t training, norm_obs, norm_reward, reserved;
    double clip_obs, clip_reward, gamma, epsilon;
} AtcVecNormParams;
int atc_vecnorm_run(const AtcVecNormState *st, const AtcVecNormParams *p, int32_t n_steps, int64_t n_env, int32_t n_aircraft,
                    const float *obs_in, float *obs_out, const float *reward_in, float *reward_out, const uint8_t *done,
                    int device, void *stream);
int64_t atc_vecnorm_scratch_doubles(int device, int32_t n_steps, int64_t n_env, int32_t n_aircraft);   /* -1: bad arguments */
int32_t atc_vecnorm_max_steps(int device);
/* The launch geometry atc_vecnorm_run would choose on a machine with n_sm SMs and ctas_per_sm co-resident CTAs per SM (no
 * GPU needed; tests): out = {CTAs, slabs per step, columns per slab, CTAs running the return recurrence, floats per
 * load (2 / 4), scratch doubles}. */
int atc_vecnorm_plan(int n_sm, int ctas_per_sm, int32_t n_steps, int64_t n_env, int32_t n_aircraft, int64_t out[6]);
const char *atc_vecnorm_last_error(void);

/* Next-row component: headless replacement of AtcGym.render(mode='rgb_array')
 * (/root/reference/envs/atc/atc_gym.py:367-552, themes.py) — same layout (10 px padding around the sector bbox, scale
 * from the width; pass height = (int)((bbox_y1 - bbox_y0) * scale) + 20 for the reference's aspect), same elements and
 * colours; the text labels are stamped separately by atc_render_text.  rgb: device uint8 [height][width][3], row 0 = north.  trail_xy: device double [n_trail][2]
 * past positions drawn as dots; heads_xy: device double [n_heads][2] current aircraft positions drawn as symbols. */
int atc_render(AtcHandle *h, uint8_t *rgb, int width, int height, const double *trail_xy, int n_trail,
               const double *heads_xy, int n_heads, void *stream);

/* Text labels over a rendered image (csrc/atc_text.cu) — the reference's pyglet labels (rendering.py:7-23): reward lines
 * (atc_gym.py:404-412) and the name / "FL  speed" lines next to every aircraft (atc_gym.py:436-443).  Screen coordinates
 * with the origin bottom-left like the reference's viewer, (x, y) = the label's top-left corner; ColorScheme.label;
 * 5 x 7 bitmap font (there is no font rasteriser here), lower case drawn as upper case.  labels: DEVICE array. */
#define ATC_TEXT_MAX 40
typedef struct AtcTextLabel {
    float x, y;
    int32_t bold, n;               /* n characters of text are drawn */
    char text[ATC_TEXT_MAX];
} AtcTextLabel;
int atc_render_text(uint8_t *rgb, int width, int height, const AtcTextLabel *labels, int n_labels, void *stream);

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t atc_launch_count(const AtcHandle *h);

/* Which kernel the last atc_step / atc_rollout (or the last chunk of a *_host call) launched, and in which shape —
 * so that a benchmark labels what actually ran instead of what it expected to run. */
typedef enum AtcKernelId {
    ATC_KERNEL_NONE = 0,
    ATC_KERNEL_STEP_FUSED = 1,        /* atc_step_kernel: one lane does everything (gym step(), short launches) */
    ATC_KERNEL_ROLLOUT_PIPE = 2,      /* atc_rollout_pipe_kernel<.., PAIRS = 1>: one mover + observer pair per CTA */
    ATC_KERNEL_ROLLOUT_PIPE_SM = 3    /* atc_rollout_pipe_kernel<.., PAIRS = 14>: one CTA per SM, MVA grid in shared memory */
} AtcKernelId;
typedef struct AtcLaunchInfo {
    int32_t kernel;                /* AtcKernelId */
    int32_t n_steps;               /* env steps fused into the launch */
    int32_t grid, block;           /* CTAs, threads per CTA */
    int32_t pairs_per_cta;         /* mover + observer warp pairs per CTA (0 for the fused kernel) */
    int32_t lanes_per_env;         /* G = next_pow2(n_aircraft) */
    int32_t wind, track_actions, exact_math, raw_obs;   /* template / output switches in effect */
    int32_t cfg;                   /* rollout kernels: 0 = run-time switches, 1 / 2 = common configuration compiled in
                                      (normalise + shaping + term + auto-reset; 2 also writes raw_obs) */
    int32_t reserved;
    int64_t dyn_smem_bytes;
} AtcLaunchInfo;
int atc_last_launch_info(const AtcHandle *h, AtcLaunchInfo *out);
const char *atc_last_error(const AtcHandle *h);   /* h may be NULL: last atc_create error */

#ifdef __cplusplus
}
#endif
#endif /* ATC_B200_H */
