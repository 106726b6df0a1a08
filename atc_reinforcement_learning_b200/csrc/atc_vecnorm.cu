// atc_vecnorm.cu — fused VecNormalize + VecCheckNan around the env step.
//
// What it replaces: the stable-baselines 2.8.0 wrappers the reference's tuner puts around the env
// (/root/reference/learning/tune_hyperparameters.py:94-97: VecCheckNan(VecNormalize(env, norm_obs, norm_reward, clip_obs,
// clip_reward, gamma))).  stable-baselines is not vendored under /root/reference (requirements.txt:117): the algorithm
// below follows its published VecNormalize.step_wait / RunningMeanStd.update_from_moments — parity UNPINNED, checked
// against the numpy restatement in oracle/vecnorm_ref.py.
//
// Per env step, in stable-baselines' order:
//     ret = ret * gamma + reward
//     obs_rms.update(obs)  (training)          obs_out = clip((obs - mean) / sqrt(var + eps), +-clip_obs)
//     ret_rms.update(ret)  (training)          reward_out = clip(reward / sqrt(ret_var + eps), +-clip_reward)
//     ret[done] = 0
// Step t is normalised with the moments AFTER the batches of steps 0 .. t have been merged, which looks sequential, but
// the batch totals of different steps do not depend on each other.  ONE cooperative launch therefore does T steps
// (T = 1 for step(), the rollout length for rollout()) in two streaming passes with a single grid-wide barrier:
//   pass 1   every (step, slab) tile is reduced to 20 partial sums (float64; 16-byte loads when the rows allow it),
//            and one thread per env runs the discounted-return recurrence over the T steps (per-step sums per CTA);
//   barrier
//   scan     every CTA folds the partial sums step by step into its own copy of the running moments (22 doubles per
//            step; the identical float64 update in every CTA) and keeps the moments of the steps whose tiles it owns;
//   pass 2   the same tiles again: normalise, clip, store (short launches find their rows in L2), rewards likewise.
// Traffic: obs read twice and written once, nothing else of size.  Round 1 used three launches per step plus eager torch
// arithmetic for the reward path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "atc_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kPairs = ATC_OBS_DIM / 2;                           // float2 columns of a row
constexpr int kActive = (kThreads / kPairs) * kPairs;             // 255: thread t always sees columns 2 (t % 5), + 1
constexpr int kObsAcc = 2 * ATC_OBS_DIM;                          // per tile: sum[10], sum of squares[10]
constexpr int kTot = kObsAcc + 2;                                 // per step: + sum, sum of squares of the return
constexpr int kMaxOwned = 32;                                     // tiles one CTA may own (T * S <= kMaxOwned * grid)
constexpr int kRetChunk = 128;                                    // steps the return recurrence stages at a time
constexpr int kScanChunk = 64;                                    // steps the scan stages at a time
constexpr unsigned kSpinLimit = 1u << 24;                         // a barrier that never completes sets sync[2], no hang

struct Args {
    AtcVecNormState st;
    AtcVecNormParams p;
    const float *obs_in;
    float *obs_out;
    const float *reward_in;
    float *reward_out;
    const uint8_t *done;
    double *part_obs;          // [n_steps][slabs][20]    per-tile sums (pass 1 -> scan)
    double *part_ret;          // [n_steps][ret_ctas][2]  per-CTA sums of the discounted return (pass 1 -> scan)
    double *rscale;            // [ret_ctas][n_steps]     reward scale per step, private to each env-owning CTA (scan -> pass 2)
    int64_t n_env, rows;       // rows = n_env * n_aircraft observation rows per step
    int64_t slab_cols;         // columns of V floats per slab (a multiple of 5)
    int32_t n_steps, slabs, ret_ctas;
};

// Generation barrier over the co-resident CTAs of a cooperative launch: sync[0] counts arrivals, sync[1] is the
// generation the last arriver bumps.  Bounded spin: on a time-out sync[2] is set (sticky) and the kernel carries on —
// the host wrapper raises — so a mis-launch cannot hang the device.
__device__ __forceinline__ void grid_barrier(unsigned *sync)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned gen = *reinterpret_cast<volatile unsigned *>(sync + 1);
        __threadfence();
        if (atomicAdd(sync, 1u) == gridDim.x - 1) {
            *reinterpret_cast<volatile unsigned *>(sync) = 0u;
            __threadfence();
            atomicAdd(sync + 1, 1u);
        } else {
            unsigned spins = *reinterpret_cast<volatile unsigned *>(sync + 2) ? kSpinLimit : 0u;
            while (*reinterpret_cast<volatile unsigned *>(sync + 1) == gen) {
                if (++spins > kSpinLimit) {
                    atomicExch(sync + 2, 1u);
                    break;
                }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// stable-baselines RunningMeanStd.update_from_moments on one feature (float64)
__device__ __forceinline__ void merge_moments(double &mean, double &var, double count, double b_sum, double b_sq, double b_count)
{
    const double b_mean = b_sum / b_count;
    double b_var = b_sq / b_count - b_mean * b_mean;
    b_var = b_var < 0.0 ? 0.0 : b_var;
    const double tot = count + b_count, delta = b_mean - mean;
    const double m2 = var * count + b_var * b_count + delta * delta * count * b_count / tot;
    mean = mean + delta * b_count / tot;
    var = m2 / tot;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, s);
    return v;
}

// V floats per load (2 or 4).  A row is 10 floats, so the columns of V floats repeat every 5 columns (10 or 20 floats):
// thread j always sees column j % 5, i.e. features (V (j % 5) + i) % 10, i < V.
template <int V> struct Vec;
template <> struct Vec<2> { typedef float2 type; };
template <> struct Vec<4> { typedef float4 type; };
__device__ __forceinline__ void unpack(const float2 &v, float f[2]) { f[0] = v.x; f[1] = v.y; }
__device__ __forceinline__ void unpack(const float4 &v, float f[4]) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
__device__ __forceinline__ void pack(float2 &v, const float f[2]) { v = make_float2(f[0], f[1]); }
__device__ __forceinline__ void pack(float4 &v, const float f[4]) { v = make_float4(f[0], f[1], f[2], f[3]); }

template <int V>
__global__ void __launch_bounds__(kThreads) atc_vecnorm_kernel(const __grid_constant__ Args a)
{
    typedef typename Vec<V>::type vec_t;
    // running obs mean[10], var[10], count [20]; return mean [21], var [22], count [23]
    __shared__ double s_run[2 * ATC_OBS_DIM + 4];
    __shared__ double s_own[kMaxOwned][2 * ATC_OBS_DIM];          // (mean, 1 / std) of the steps whose tiles this CTA owns
    __shared__ double s_stage[2 * 4 * kThreads];                  // pass 1: block reduce / per-step return sums; scan: totals
    __shared__ double s_col[kPairs][2 * 4];
    __shared__ int s_bad;
    static_assert(2 * 4 * kThreads >= kScanChunk * kTot && 2 * 4 * kThreads >= (kThreads / 32) * 2 * kRetChunk, "s_stage");
    constexpr int kCnt = 2 * ATC_OBS_DIM, kRm = kCnt + 1, kRv = kCnt + 2, kRc = kCnt + 3;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool training = a.p.training != 0, norm_obs = a.p.norm_obs != 0;
    const bool rewards = a.reward_in != nullptr, norm_rew = rewards && a.p.norm_reward != 0;
    const int T = a.n_steps, S = a.slabs, G = (int)gridDim.x, b = (int)blockIdx.x;
    const int64_t n_elem = a.rows * ATC_OBS_DIM, n_cols = n_elem / V;     // columns of V floats per step
    if (tid <= kCnt) s_run[tid] = a.st.obs_rms[tid];
    if (tid < 3) s_run[kRm + tid] = a.st.ret_rms[tid];
    if (tid == 0) s_bad = 0;
    bool bad = false;
    __syncthreads();

    // ------------------------------------------------------------------------------------------------ pass 1
    if (training && norm_obs) {
        for (int k = b; k < T * S; k += G) {
            const int t = k / S, s = k - t * S;
            const vec_t *xv = reinterpret_cast<const vec_t *>(a.obs_in + (size_t)t * n_elem);
            const int64_t q_lo = (int64_t)s * a.slab_cols, q_hi = min(q_lo + a.slab_cols, n_cols);
            double sm[V], sq[V];
#pragma unroll
            for (int i = 0; i < V; ++i) sm[i] = sq[i] = 0.0;
            if (tid < kActive) {
#pragma unroll 4
                for (int64_t q = q_lo + tid; q < q_hi; q += kActive) {
                    float f[V];
                    unpack(__ldcg(xv + q), f);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        bad |= !isfinite(f[i]);
                        sm[i] += (double)f[i];
                        sq[i] = fma((double)f[i], (double)f[i], sq[i]);
                    }
                }
            }
            // block reduce through shared memory: the 51 threads of each column, then the columns that hold a feature
#pragma unroll
            for (int i = 0; i < V; ++i) {
                s_stage[2 * V * tid + i] = sm[i];
                s_stage[2 * V * tid + V + i] = sq[i];
            }
            __syncthreads();
            if (tid < kPairs * 2 * V) {
                const int c = tid / (2 * V), comp = tid - c * 2 * V;
                double acc = 0.0;
                for (int jt = c; jt < kActive; jt += kPairs) acc += s_stage[2 * V * jt + comp];
                s_col[c][comp] = acc;
            }
            __syncthreads();
            if (tid < kObsAcc) {
                const int f = tid % ATC_OBS_DIM, hi = tid < ATC_OBS_DIM ? 0 : V;   // sum or sum of squares of feature f
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < kPairs; ++c)
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        if ((V * c + i) % ATC_OBS_DIM == f) acc += s_col[c][hi + i];
                a.part_obs[((size_t)t * S + s) * kObsAcc + tid] = acc;
            }
            __syncthreads();
        }
    }
    if (rewards && b < a.ret_ctas) {
        // one thread per env runs ret = ret * gamma + reward over the T steps; per-step sums of this CTA's envs
        const int64_t n_stride = (int64_t)a.ret_ctas * kThreads;
        const bool sums = training && norm_rew;
        double *const s_rt = s_stage + (tid >> 5) * 2 * kRetChunk;    // [warp][kRetChunk][2]: no atomics, summed at the flush
        for (int t0 = 0; t0 < T; t0 += kRetChunk) {
            const int tc = min(kRetChunk, T - t0);
            for (int i = tid; i < (kThreads / 32) * 2 * kRetChunk; i += kThreads) s_stage[i] = 0.0;
            __syncthreads();
            // (n - lane is the warp's first env: the lanes of a warp leave the loop together, the shuffles stay converged)
            for (int64_t n = (int64_t)b * kThreads + tid; n - lane < a.n_env; n += n_stride) {
                const bool live = n < a.n_env;
                double ret = live ? a.st.ret[n] : 0.0;
                for (int tb = 0; tb < tc; tb += 8) {
                    // rewards and done flags of 8 steps up front (independent loads), then the recurrence
                    float r8[8];
                    uint8_t d8[8];
                    double v8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const bool ok = live && tb + j < tc;
                        const size_t i = (size_t)(t0 + tb + j) * a.n_env + n;
                        r8[j] = ok ? __ldcg(a.reward_in + i) : 0.0f;
                        d8[j] = ok ? a.done[i] : (uint8_t)0;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v8[j] = 0.0;
                        if (tb + j < tc) {                        // uniform
                            bad |= !isfinite(r8[j]);
                            ret = ret * a.p.gamma + (double)r8[j];
                            v8[j] = live ? ret : 0.0;
                            if (d8[j]) ret = 0.0;
                        }
                    }
                    if (sums) {
                        // the 16 warp sums are independent of each other and of the recurrence
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const double ws = warp_sum(v8[j]), wq = warp_sum(v8[j] * v8[j]);
                            if (lane == 0 && tb + j < tc) {
                                s_rt[2 * (tb + j)] += ws;
                                s_rt[2 * (tb + j) + 1] += wq;
                            }
                        }
                    }
                }
                if (live) a.st.ret[n] = ret;
            }
            __syncthreads();
            if (sums)
                for (int i = tid; i < 2 * tc; i += kThreads) {
                    double acc = 0.0;
#pragma unroll
                    for (int w = 0; w < kThreads / 32; ++w) acc += s_stage[w * 2 * kRetChunk + i];
                    a.part_ret[((size_t)(t0 + (i >> 1)) * a.ret_ctas + b) * 2 + (i & 1)] = acc;
                }
            __syncthreads();
        }
    }

    // ------------------------------------------------------------------------------------------------ barrier + scan
    if (training) {
        grid_barrier(a.st.sync);
        const double rows_d = (double)a.rows, env_d = (double)a.n_env;
        int own = 0;                                              // this CTA's next tile is k = b + own * G
        for (int t0 = 0; t0 < T; t0 += kScanChunk) {
            const int tc = min(kScanChunk, T - t0);
            // totals of the chunk's steps into shared memory: item (step, slot) summed over the S slabs / the ret_ctas
            // CTAs; short launches have many slabs per step, so an item is split over `split` threads
            for (int i = tid; i < tc * kTot; i += kThreads) s_stage[i] = 0.0;
            __syncthreads();
            const int items = tc * kTot;
            int split = kThreads / items;
            split = split < 1 ? 1 : split;
            for (int w = tid; w < items * split; w += kThreads) {
                const int item = w / split, part = w - item * split;
                const int tl = item / kTot, f = item - tl * kTot;
                double acc = 0.0;
                if (f < kObsAcc) {
                    if (norm_obs) {
                        const double *src = a.part_obs + ((size_t)(t0 + tl) * S) * kObsAcc + f;
#pragma unroll 4
                        for (int s = part; s < S; s += split) acc += __ldcg(src + (size_t)s * kObsAcc);
                    }
                } else if (norm_rew) {
                    const double *src = a.part_ret + ((size_t)(t0 + tl) * a.ret_ctas) * 2 + (f - kObsAcc);
#pragma unroll 4
                    for (int c = part; c < a.ret_ctas; c += split) acc += __ldcg(src + (size_t)c * 2);
                }
                if (split == 1)
                    s_stage[item] = acc;
                else
                    atomicAdd(&s_stage[item], acc);
            }
            __syncthreads();
            // the sequential part, from shared memory
            for (int tl = 0; tl < tc; ++tl) {
                const int t = t0 + tl;
                const double *tot = s_stage + tl * kTot;
                if (norm_obs && tid < ATC_OBS_DIM)
                    merge_moments(s_run[tid], s_run[ATC_OBS_DIM + tid], s_run[kCnt], tot[tid], tot[ATC_OBS_DIM + tid], rows_d);
                if (norm_rew && tid == 32) merge_moments(s_run[kRm], s_run[kRv], s_run[kRc], tot[kObsAcc], tot[kObsAcc + 1], env_d);
                __syncthreads();
                if (tid == 0 && norm_obs) s_run[kCnt] += rows_d;
                if (tid == 32 && norm_rew) {
                    s_run[kRc] += env_d;
                    if (b < a.ret_ctas) a.rscale[(size_t)b * T + t] = 1.0 / sqrt(s_run[kRv] + a.p.epsilon);
                }
                // keep the moments of this step for every tile of it this CTA owns (tiles are numbered step-major)
                while (own < kMaxOwned && (b + own * G) / S == t) {
                    if (tid < ATC_OBS_DIM) {
                        s_own[own][tid] = s_run[tid];
                        s_own[own][ATC_OBS_DIM + tid] = 1.0 / sqrt(s_run[ATC_OBS_DIM + tid] + a.p.epsilon);
                    }
                    ++own;
                }
                __syncthreads();
            }
        }
    } else {
        // frozen moments: every owned tile uses the same ones
        if (tid < ATC_OBS_DIM)
            for (int o = 0; o < kMaxOwned; ++o) {
                s_own[o][tid] = s_run[tid];
                s_own[o][ATC_OBS_DIM + tid] = 1.0 / sqrt(s_run[ATC_OBS_DIM + tid] + a.p.epsilon);
            }
        __syncthreads();
    }

    // ------------------------------------------------------------------------------------------------ pass 2
    if (norm_obs || a.obs_in != a.obs_out) {
        int own = 0;
        for (int k = b; k < T * S; k += G, ++own) {
            const int t = k / S, s = k - t * S;
            const vec_t *xv = reinterpret_cast<const vec_t *>(a.obs_in + (size_t)t * n_elem);
            vec_t *yv = reinterpret_cast<vec_t *>(a.obs_out + (size_t)t * n_elem);
            const int64_t q_lo = (int64_t)s * a.slab_cols, q_hi = min(q_lo + a.slab_cols, n_cols);
            if (tid < kActive) {
                // (x - mean) / std in float64 like numpy on float32 obs and float64 moments, then cast
                double mu[V], rstd[V];
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const int f = (V * (tid % kPairs) + i) % ATC_OBS_DIM;
                    mu[i] = s_own[own][f];
                    rstd[i] = s_own[own][ATC_OBS_DIM + f];
                }
                const double c = a.p.clip_obs;
#pragma unroll 4
                for (int64_t q = q_lo + tid; q < q_hi; q += kActive) {
                    float f[V];
                    unpack(__ldcg(xv + q), f);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        bad |= !isfinite(f[i]);
                        if (norm_obs) {
                            double o = ((double)f[i] - mu[i]) * rstd[i];
                            o = o < -c ? -c : (o > c ? c : o);
                            f[i] = (float)o;
                        }
                    }
                    vec_t o;
                    pack(o, f);
                    __stcs(yv + q, o);
                }
            }
        }
    }
    if (rewards && b < a.ret_ctas) {
        const int64_t n_stride = (int64_t)a.ret_ctas * kThreads;
        const double frozen = 1.0 / sqrt(s_run[kRv] + a.p.epsilon), c = a.p.clip_reward;
        for (int64_t n = (int64_t)b * kThreads + tid; n < a.n_env; n += n_stride) {
            for (int tb = 0; tb < T; tb += 8) {
                float r8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r8[j] = tb + j < T ? __ldcg(a.reward_in + (size_t)(tb + j) * a.n_env + n) : 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (tb + j < T) {
                        // (rscale: written by this CTA's own thread 32 before the last __syncthreads of the scan)
                        const double rs = (training && norm_rew) ? a.rscale[(size_t)b * T + tb + j] : frozen;
                        bad |= !isfinite(r8[j]);
                        double o = r8[j];
                        if (norm_rew) {
                            o = (double)r8[j] * rs;
                            o = o < -c ? -c : (o > c ? c : o);
                        }
                        a.reward_out[(size_t)(tb + j) * a.n_env + n] = (float)o;
                    }
                }
            }
        }
    }
    if (bad) s_bad = 1;
    __syncthreads();
    if (tid == 0 && s_bad) *a.st.nonfinite = 1;
    if (b == 0 && training) {
        if (tid <= kCnt) a.st.obs_rms[tid] = s_run[tid];
        if (tid < 3) a.st.ret_rms[tid] = s_run[kRm + tid];
    }
}

thread_local char g_vn_error[256] = "";

int vn_fail(int code, const char *m)
{
    snprintf(g_vn_error, sizeof g_vn_error, "atc_vecnorm_run: %s", m);
    return code;
}

// launch geometry: co-resident CTAs (cooperative launch) and slabs per step
struct Plan {
    int grid, slabs, ret_ctas, vec;
    int64_t slab_cols, scratch_doubles;
};

// the geometry as a pure function of the machine (SMs, co-resident CTAs per SM) and the job — testable without a GPU
int plan_for(int n_sm, int per_sm_in, int32_t n_steps, int64_t n_env, int32_t n_aircraft, Plan *pl)
{
    const int per_sm = per_sm_in > 4 ? 4 : per_sm_in;
    const int64_t cap = (int64_t)n_sm * per_sm;
    const int64_t rows = n_env * n_aircraft;
    const int vec = (rows % 2 == 0) ? 4 : 2;                              // float4 needs every step to start 16-byte aligned
    const int64_t n_cols = rows * ATC_OBS_DIM / vec;
    // CTAs: ~8 loads per thread when the whole job is small (a single step: one CTA per SM or so — the barrier and
    // the scan grow with the CTA and slab counts), all that fit when it is big; never so few that a CTA owns too many tiles
    int64_t grid = (n_cols * n_steps + 8 * kActive - 1) / (8 * kActive);
    const int64_t need = ((int64_t)n_steps + kMaxOwned - 1) / kMaxOwned;
    grid = grid < need ? need : grid;
    grid = grid > cap ? cap : grid;
    // slabs per step: about one tile per CTA for short launches; for long ones the tile count that balances best
    const int64_t max_slabs = (n_cols / kPairs + 50) / 51;                // at least one pass of the 255 threads per slab
    int64_t best = 1;
    double best_cost = 1e30;
    for (int64_t sl = 1; sl <= 8 * grid / n_steps + 1 && sl <= max_slabs + 1; ++sl) {
        const int64_t s1 = sl > max_slabs ? (max_slabs < 1 ? 1 : max_slabs) : sl;
        const int64_t tiles = s1 * n_steps;
        if (tiles > (int64_t)kMaxOwned * grid) break;
        const double rounds = (double)((tiles + grid - 1) / grid);
        const double cost = rounds / ((double)tiles / (double)grid) + 1e-6 * (double)tiles;   // imbalance, then fewer tiles
        if (cost < best_cost) { best_cost = cost; best = s1; }
    }
    const int64_t slabs = best;
    if ((int64_t)n_steps * slabs > (int64_t)kMaxOwned * grid)
        return vn_fail(ATC_ERR_INVALID_ARGUMENT, "n_steps too large for one launch (atc_vecnorm_max_steps)");
    pl->grid = (int)grid;
    pl->slabs = (int)slabs;
    pl->vec = vec;
    pl->slab_cols = ((n_cols + slabs - 1) / slabs + kPairs - 1) / kPairs * kPairs;
    const int64_t rc = (n_env + kThreads - 1) / kThreads;
    pl->ret_ctas = (int)(rc < grid ? rc : grid);
    pl->scratch_doubles = (int64_t)n_steps * (slabs * kObsAcc + (int64_t)pl->ret_ctas * 3);
    return ATC_OK;
}

int make_plan(int device, int32_t n_steps, int64_t n_env, int32_t n_aircraft, Plan *pl)
{
    // the device queries cost tens of microseconds: remembered per thread for the device last asked about
    thread_local int c_device = -1, c_sm = 0, c_coop = 0, c_per_sm = 0;
    if (c_device != device) {
        int n_sm = 0, coop = 0, p2 = 0, p4 = 0;
        cudaError_t e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p2, atc_vecnorm_kernel<2>, kThreads, 0);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p4, atc_vecnorm_kernel<4>, kThreads, 0);
        if (e != cudaSuccess) return vn_fail(ATC_ERR_CUDA, cudaGetErrorString(e));
        c_device = device; c_sm = n_sm; c_coop = coop; c_per_sm = p2 < p4 ? p2 : p4;
    }
    if (!c_coop || c_per_sm < 1) return vn_fail(ATC_ERR_UNSUPPORTED, "the device does not support cooperative launches");
    return plan_for(c_sm, c_per_sm, n_steps, n_env, n_aircraft, pl);
}

}  // namespace

extern "C" {

const char *atc_vecnorm_last_error(void) { return g_vn_error; }

int atc_vecnorm_plan(int n_sm, int ctas_per_sm, int32_t n_steps, int64_t n_env, int32_t n_aircraft, int64_t out[6])
{
    if (!out || n_sm < 1 || ctas_per_sm < 1 || n_steps < 1 || n_env < 1 || n_aircraft < 1 || n_aircraft > ATC_MAX_AIRCRAFT)
        return vn_fail(ATC_ERR_INVALID_ARGUMENT, "atc_vecnorm_plan: bad arguments");
    Plan pl;
    const int rc = plan_for(n_sm, ctas_per_sm, n_steps, n_env, n_aircraft, &pl);
    if (rc != ATC_OK) return rc;
    out[0] = pl.grid; out[1] = pl.slabs; out[2] = pl.slab_cols; out[3] = pl.ret_ctas; out[4] = pl.vec; out[5] = pl.scratch_doubles;
    return ATC_OK;
}

int64_t atc_vecnorm_scratch_doubles(int device, int32_t n_steps, int64_t n_env, int32_t n_aircraft)
{
    if (n_steps < 1 || n_env < 1 || n_aircraft < 1 || n_aircraft > ATC_MAX_AIRCRAFT) return -1;
    Plan pl;
    if (cudaSetDevice(device) != cudaSuccess || make_plan(device, n_steps, n_env, n_aircraft, &pl) != ATC_OK) return -1;
    return pl.scratch_doubles;
}

int32_t atc_vecnorm_max_steps(int device)
{
    int n_sm = 0;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n_sm < 1) return -1;
    return kMaxOwned * n_sm;              // one resident CTA per SM is always possible (make_plan has the exact rule)
}

int atc_vecnorm_run(const AtcVecNormState *st, const AtcVecNormParams *p, int32_t n_steps, int64_t n_env, int32_t n_aircraft,
                    const float *obs_in, float *obs_out, const float *reward_in, float *reward_out, const uint8_t *done,
                    int device, void *stream)
{
    if (!st || !p || !obs_in || !obs_out) return vn_fail(ATC_ERR_INVALID_ARGUMENT, "state, params, obs_in and obs_out are required");
    if (!st->obs_rms || !st->ret_rms || !st->scratch || !st->sync || !st->nonfinite)
        return vn_fail(ATC_ERR_INVALID_ARGUMENT, "AtcVecNormState: NULL member");
    if (n_steps < 1 || n_env < 1 || n_aircraft < 1 || n_aircraft > ATC_MAX_AIRCRAFT) return vn_fail(ATC_ERR_INVALID_ARGUMENT, "bad sizes");
    if (reward_in && (!reward_out || !done || !st->ret))
        return vn_fail(ATC_ERR_INVALID_ARGUMENT, "reward_in needs reward_out, done and state.ret");
    if (!(p->clip_obs > 0.0) || !(p->clip_reward > 0.0) || !(p->epsilon >= 0.0))
        return vn_fail(ATC_ERR_INVALID_ARGUMENT, "clip_* must be > 0, epsilon >= 0");
    if ((reinterpret_cast<uintptr_t>(obs_in) | reinterpret_cast<uintptr_t>(obs_out)) & 15)
        return vn_fail(ATC_ERR_INVALID_ARGUMENT, "obs_in / obs_out must be 16-byte aligned");
    int cur = -1;
    cudaError_t e = cudaGetDevice(&cur);
    if (e == cudaSuccess && cur != device) e = cudaSetDevice(device);
    if (e != cudaSuccess) return vn_fail(ATC_ERR_CUDA, cudaGetErrorString(e));
    Plan pl;
    int rc = make_plan(device, n_steps, n_env, n_aircraft, &pl);
    if (rc != ATC_OK) return rc;
    if (st->scratch_doubles < pl.scratch_doubles)
        return vn_fail(ATC_ERR_INVALID_ARGUMENT, "scratch too small (atc_vecnorm_scratch_doubles)");
    Args a;
    a.st = *st; a.p = *p;
    a.obs_in = obs_in; a.obs_out = obs_out; a.reward_in = reward_in; a.reward_out = reward_out; a.done = done;
    a.n_env = n_env; a.rows = n_env * n_aircraft; a.n_steps = n_steps;
    a.slabs = pl.slabs; a.slab_cols = pl.slab_cols; a.ret_ctas = pl.ret_ctas;
    a.part_obs = st->scratch;
    a.part_ret = a.part_obs + (size_t)n_steps * pl.slabs * kObsAcc;
    a.rscale = a.part_ret + (size_t)n_steps * pl.ret_ctas * 2;
    void *params[] = {&a};
    const void *kern = pl.vec == 4 ? reinterpret_cast<const void *>(atc_vecnorm_kernel<4>)
                                   : reinterpret_cast<const void *>(atc_vecnorm_kernel<2>);
    e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)pl.grid), dim3(kThreads), params, 0, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return vn_fail(ATC_ERR_CUDA, cudaGetErrorString(e));
    return ATC_OK;
}

}  // extern "C"
