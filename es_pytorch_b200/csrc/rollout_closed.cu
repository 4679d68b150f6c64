// rollout_closed.cu -- antithetic pairs on the CLOSED-LOOP synthetic env (SURVEY.md section 8d's optional variant, reported
// separately from the open-loop headline): obs_{t+1} = tanh(A obs_t + B a_t), so the observation a policy sees depends on what
// it did and an episode cannot be batched over time like rollout_tc2.cu / rollout_f32x.cu do.  What replaces the reference here
// is the same loop as in the open-loop kernels -- Policy.pheno (src/core/policy.py:61-64), FeedForward.forward
// (src/nn/nn.py:42-50: clip((ob - mean) / std), Linear + Tanh x3), run_model's reward / position / saved observations
// (src/gym/gym_runner.py:33-67) -- with the env step inside it.
//
// Every step is a batch-1 matrix-vector product per policy, with weights that are unique per pair: the shape that is bound by
// where the WEIGHTS live.  Re-reading a pair's 117 KB slice from HBM at each of 1 000 steps is 1.2 PB per generation; so one
// CTA keeps one pair's perturbed weights on chip for the whole episode and the grid is persistent over the pairs:
//   * layer 1 (82 % of the weights): thread (o = tid / 8, s = tid % 8) owns the elements k = 8 j + s of row o; its + weights
//     live in REGISTERS (48), its - weights in shared memory in [j][tid] order (conflict-free); the normalised observations of
//     both signs are float2 in shared memory (8 distinct addresses per warp: broadcasts); the 8 partial sums of a row meet by
//     three shuffles;
//   * layers 2 / 3: both signs' weights in shared memory, same thread <-> element map;
//   * the env step: thread i owns observation i of both signs (A's diagonals and B transposed in shared memory), normalises
//     the new observation with its own mean / std (float64 like the reference: a float32 tensor minus a float64 ndarray), and
//     keeps the float32 column sums of the post-step observations for the ObStat of a rollout whose save_obs coin fell;
//   * reward (float32 dot in index order, summed in float64 like python's sum) and position by two threads outside the
//     observation range, concurrently with the env step.
// Four barriers per step; ~1 us per step for Humanoid shapes, i.e. ~1 ms per pair and SM.
#include <math.h>
#include "common.cuh"

namespace {

constexpr int CL_THREADS = 512;
constexpr int CL_H = 64;                   // hidden units per layer at most (thread rows)
constexpr int CL_A = 64;                   // action units at most (layer-3 rows)

struct ClParams {
    const float* table; long long table_len; const int64_t* idx; int n_pairs;
    const float* theta; float sigma;
    int obs, h1, h2, act, T, n_params;
    const double* ob_mean; const double* ob_std; double ob_clip;
    const float* obs0; const float* env_a; const float* env_b; int band;
    const float* crew; float pos_scale;
    const uint32_t* coins; double chance;
    double* fit_pos; double* fit_neg; int fit_stride;
    float* behv_pos; float* behv_neg;
    double* ob_sum; double* ob_sumsq; double* ob_count;
    int* err;
};

struct ClLayout {                          // offsets in floats into dynamic shared memory
    int wm1, w2p, w2m, w3p, w3m, bias, env_a, env_b, x2, o2, h1, h2, a2, prod, total;
};
__host__ __device__ inline ClLayout cl_layout(int J, int obs, int act, int band) {
    ClLayout L;
    int at = 0;
    const int obs_pad = 8 * J;
    L.wm1 = at; at += J * CL_THREADS;
    L.w2p = at; at += 8 * CL_THREADS;
    L.w2m = at; at += 8 * CL_THREADS;
    L.w3p = at; at += 8 * CL_THREADS;
    L.w3m = at; at += 8 * CL_THREADS;
    L.bias = at; at += 6 * CL_H;            // b1+ b1- b2+ b2- b3+ b3-
    L.env_a = at; at += band * obs;
    L.env_b = at; at += act * obs;
    at = (at + 1) & ~1;
    L.x2 = at; at += 2 * obs_pad;           // float2 (x+, x-) normalised observations, zero padded to 8 J
    L.o2 = at; at += 2 * 2 * obs;           // [2 buffers] float2 raw observations
    L.h1 = at; at += 2 * CL_H;
    L.h2 = at; at += 2 * CL_H;
    L.a2 = at; at += 2 * CL_A;
    L.prod = at; at += 2 * CL_A;
    L.total = at;
    return L;
}

__device__ __forceinline__ float cl_group_sum(float v) {       // sum over the 8 lanes that share a row
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
// clip((ob - mean) / std) in float64 like the reference (a float32 tensor minus a float64 ndarray), the quotient as a product
// with the rounded reciprocal: the float64 result can differ in its last bit, which survives the rounding to float32 once in
// ~2^28 values (and then by one float32 ulp) -- far inside the tolerance this variant is held to
__device__ __forceinline__ float cl_normalise(float o, double mean, double rstd, double clip) {
    double x = ((double)o - mean) * rstd;
    x = fmin(fmax(x, -clip), clip);
    return (float)x;
}
// tanh(x) = 1 - 2 / (1 + e^2x) with the fast exponential and division: absolute error ~1e-7 (the float32 kernels' tanhf is
// ~40 dependent instructions per call, and every phase of a step ends in one)
__device__ __forceinline__ float cl_tanh(float x) {
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, 1.f + e);
}

template <int J>
__global__ void __launch_bounds__(CL_THREADS, 1) rollout_closed_kernel(const ClParams p) {
    extern __shared__ __align__(16) float cl_smem[];
    const ClLayout L = cl_layout(J, p.obs, p.act, p.band);
    float* __restrict__ Wm1 = cl_smem + L.wm1;
    float* __restrict__ W2p = cl_smem + L.w2p;
    float* __restrict__ W2m = cl_smem + L.w2m;
    float* __restrict__ W3p = cl_smem + L.w3p;
    float* __restrict__ W3m = cl_smem + L.w3m;
    float* __restrict__ bias = cl_smem + L.bias;
    float* __restrict__ envA = cl_smem + L.env_a;
    float* __restrict__ envB = cl_smem + L.env_b;
    float2* __restrict__ x2 = reinterpret_cast<float2*>(cl_smem + L.x2);
    float2* __restrict__ o2 = reinterpret_cast<float2*>(cl_smem + L.o2);
    float2* __restrict__ h1v = reinterpret_cast<float2*>(cl_smem + L.h1);
    float2* __restrict__ h2v = reinterpret_cast<float2*>(cl_smem + L.h2);
    float2* __restrict__ a2 = reinterpret_cast<float2*>(cl_smem + L.a2);
    float2* __restrict__ prod = reinterpret_cast<float2*>(cl_smem + L.prod);

    const int tid = threadIdx.x, o = tid >> 3, s = tid & 7;
    const int obs = p.obs, h1 = p.h1, h2 = p.h2, act = p.act, T = p.T, band = p.band, half = p.band >> 1;
    // flat parameter layout (state-dict order, src/core/policy.py:33-35): W1 [h1][obs], b1, W2 [h2][h1], b2, W3 [act][h2], b3
    const int off_b1 = h1 * obs, off_w2 = off_b1 + h1, off_b2 = off_w2 + h2 * h1, off_w3 = off_b2 + h2, off_b3 = off_w3 + act * h2;

    for (int i = tid; i < band * obs; i += CL_THREADS) envA[i] = p.env_a[i];
    for (int i = tid; i < act * obs; i += CL_THREADS) envB[i] = p.env_b[i];
    for (int i = tid; i < 8 * J; i += CL_THREADS) x2[i] = make_float2(0.f, 0.f);
    const double my_mean = tid < obs ? p.ob_mean[tid] : 0.0, my_std = tid < obs ? 1.0 / p.ob_std[tid] : 1.0;      // (reciprocal)
    const float my_obs0 = tid < obs ? p.obs0[tid] : 0.f;
    const float ps = p.pos_scale;
    __syncthreads();

    for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x) {
        const long long base = es_checked_slice(p.idx[pair], p.n_params, p.table_len, p.err);
        const float* __restrict__ eps = p.table + base;
        const float* __restrict__ th = p.theta;
        const float sg = p.sigma;
        // theta +- sigma * eps: the product and the sum are rounded separately (numpy: flat + std * noise)
        auto wpm = [&](int at, float& wp, float& wm) {
            const float t = th[at], se = __fmul_rn(sg, eps[at]);
            wp = __fadd_rn(t, se);
            wm = __fsub_rn(t, se);
        };
        float wp1[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = 8 * j + s;
            float a = 0.f, b = 0.f;
            if (o < h1 && k < obs) wpm(o * obs + k, a, b);
            wp1[j] = a;
            Wm1[j * CL_THREADS + tid] = b;
        }
        for (int j = 0; j < 8; ++j) {
            const int k = 8 * j + s;
            float a = 0.f, b = 0.f;
            if (o < h2 && k < h1) wpm(off_w2 + o * h1 + k, a, b);
            W2p[j * CL_THREADS + tid] = a; W2m[j * CL_THREADS + tid] = b;
            a = 0.f; b = 0.f;
            if (o < act && k < h2) wpm(off_w3 + o * h2 + k, a, b);
            W3p[j * CL_THREADS + tid] = a; W3m[j * CL_THREADS + tid] = b;
        }
        if (tid < CL_H) {
            float a = 0.f, b = 0.f;
            if (tid < h1) wpm(off_b1 + tid, a, b);
            bias[tid] = a; bias[CL_H + tid] = b;
            a = 0.f; b = 0.f;
            if (tid < h2) wpm(off_b2 + tid, a, b);
            bias[2 * CL_H + tid] = a; bias[3 * CL_H + tid] = b;
            a = 0.f; b = 0.f;
            if (tid < act) wpm(off_b3 + tid, a, b);
            bias[4 * CL_H + tid] = a; bias[5 * CL_H + tid] = b;
        }
        if (tid < obs) {
            o2[tid] = make_float2(my_obs0, my_obs0);
            const float xn = cl_normalise(my_obs0, my_mean, my_std, p.ob_clip);
            x2[tid] = make_float2(xn, xn);
        }
        // the save_obs coins of the pair's two evaluations (legacy random_sample: (a >> 5, b >> 6) / 2^53 < chance)
        bool save_p = false, save_m = false;
        if (p.coins) {
            const uint32_t* c = p.coins + (size_t)pair * 4;
            save_p = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) / 9007199254740992.0 < p.chance;
            save_m = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6)) / 9007199254740992.0 < p.chance;
        }
        float sum_p = 0.f, sq_p = 0.f, sum_m = 0.f, sq_m = 0.f;    // column sums of the post-step observations (float32, step order)
        double fit = 0.0;                                           // threads obs_thr0 (+) and obs_thr0 + 1 (-)
        float pos0 = 0.f, pos1 = 0.f, pos2 = 0.f;
        const int rew_thr = (obs + 31) & ~31;                       // first thread of the warp after the observation threads
        __syncthreads();

        for (int t = 0; t < T; ++t) {
            const int cur = t & 1;
            float crow0 = 0.f, crow1 = 0.f;
            if (tid >= rew_thr && tid < rew_thr + 32) {
                const int ln = tid - rew_thr;
                const float* __restrict__ c = p.crew + (size_t)t * act;
                if (ln < act) crow0 = __ldg(c + ln);
                if (ln + 32 < act) crow1 = __ldg(c + ln + 32);
            }
            // ---- layer 1 ----
            {
                float zp = 0.f, zm = 0.f, zp1 = 0.f, zm1 = 0.f;
#pragma unroll
                for (int j = 0; j < J; j += 2) {
                    const float2 xv = x2[8 * j + s], xw = x2[8 * j + 8 + s];
                    zp = fmaf(wp1[j], xv.x, zp);
                    zm = fmaf(Wm1[j * CL_THREADS + tid], xv.y, zm);
                    zp1 = fmaf(wp1[j + 1], xw.x, zp1);
                    zm1 = fmaf(Wm1[(j + 1) * CL_THREADS + tid], xw.y, zm1);
                }
                zp = cl_group_sum(zp + zp1); zm = cl_group_sum(zm + zm1);
                if (s == 0 && o < CL_H) h1v[o] = (o < h1) ? make_float2(cl_tanh(zp + bias[o]), cl_tanh(zm + bias[CL_H + o])) : make_float2(0.f, 0.f);
            }
            __syncthreads();
            // ---- layer 2 ----
            {
                float zp = 0.f, zm = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 hv = h1v[8 * j + s];
                    zp = fmaf(W2p[j * CL_THREADS + tid], hv.x, zp);
                    zm = fmaf(W2m[j * CL_THREADS + tid], hv.y, zm);
                }
                zp = cl_group_sum(zp); zm = cl_group_sum(zm);
                if (s == 0 && o < CL_H) h2v[o] = (o < h2) ? make_float2(cl_tanh(zp + bias[2 * CL_H + o]), cl_tanh(zm + bias[3 * CL_H + o])) : make_float2(0.f, 0.f);
            }
            __syncthreads();
            // ---- layer 3 (rows >= act hold zero weights; every thread takes part in the shuffles) ----
            {
                float zp = 0.f, zm = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 hv = h2v[8 * j + s];
                    zp = fmaf(W3p[j * CL_THREADS + tid], hv.x, zp);
                    zm = fmaf(W3m[j * CL_THREADS + tid], hv.y, zm);
                }
                zp = cl_group_sum(zp); zm = cl_group_sum(zm);
                if (s == 0 && o < act) a2[o] = make_float2(cl_tanh(zp + bias[4 * CL_H + o]), cl_tanh(zm + bias[5 * CL_H + o]));
            }
            __syncthreads();
            // ---- env step (threads < obs), reward and position (two threads of the next warp) ----
            if (tid < obs) {
                const float2* __restrict__ oc = o2 + cur * obs;
                float ap = 0.f, am = 0.f;
                int k = tid - half;
                if (k < 0) k += obs;
#pragma unroll 8
                for (int d = 0; d < band; ++d) {
                    const float w = envA[d * obs + tid];
                    const float2 ov = oc[k];
                    ap = fmaf(w, ov.x, ap); am = fmaf(w, ov.y, am);
                    if (++k == obs) k = 0;
                }
#pragma unroll 6
                for (int j = 0; j < act; ++j) {
                    const float w = envB[j * obs + tid];
                    const float2 av = a2[j];
                    ap = fmaf(w, av.x, ap); am = fmaf(w, av.y, am);
                }
                const float np_ = cl_tanh(ap), nm = cl_tanh(am);
                o2[(cur ^ 1) * obs + tid] = make_float2(np_, nm);
                x2[tid] = make_float2(cl_normalise(np_, my_mean, my_std, p.ob_clip), cl_normalise(nm, my_mean, my_std, p.ob_clip));
                sum_p = __fadd_rn(sum_p, np_); sq_p = __fadd_rn(sq_p, __fmul_rn(np_, np_));
                sum_m = __fadd_rn(sum_m, nm);  sq_m = __fadd_rn(sq_m, __fmul_rn(nm, nm));
            } else if (tid >= rew_thr && tid < rew_thr + 32) {
                // the reward warp: its lanes hold the step's reward coefficients (requested before layer 1) and form the
                // products; lanes 0 / 1 add them in index order (the env's float32 dot) for the + / - evaluation
                const int ln = tid - rew_thr;
                if (ln < act) { const float2 av = a2[ln]; prod[ln] = make_float2(__fmul_rn(av.x, crow0), __fmul_rn(av.y, crow0)); }
                if (ln + 32 < act) { const float2 av = a2[ln + 32]; prod[ln + 32] = make_float2(__fmul_rn(av.x, crow1), __fmul_rn(av.y, crow1)); }
                __syncwarp();
                if (ln < 2) {
                    const int sgn = ln;
                    float acc = 0.f;
                    for (int j = 0; j < act; ++j) {
                        const float2 pv = prod[j];
                        acc = __fadd_rn(acc, sgn ? pv.y : pv.x);
                    }
                    fit += (double)acc;
                    const float2 q0 = a2[0], q1 = a2[1 % act], q2 = a2[2 % act];
                    pos0 = __fadd_rn(pos0, __fmul_rn(ps, sgn ? q0.y : q0.x));
                    pos1 = __fadd_rn(pos1, __fmul_rn(ps, sgn ? q1.y : q1.x));
                    pos2 = __fadd_rn(pos2, __fmul_rn(ps, sgn ? q2.y : q2.x));
                }
            }
            __syncthreads();
        }
        if (tid == rew_thr || tid == rew_thr + 1) {
            const int sgn = tid - rew_thr;
            (sgn ? p.fit_neg : p.fit_pos)[(size_t)pair * p.fit_stride] = fit;
            float* bv = sgn ? p.behv_neg : p.behv_pos;
            if (bv) { bv[(size_t)pair * 3 + 0] = pos0; bv[(size_t)pair * 3 + 1] = pos1; bv[(size_t)pair * 3 + 2] = pos2; }
        }
        if (p.ob_sum && (save_p || save_m)) {
            // ObStat.inc of the saved rollouts (src/core/es.py:73-74, src/gym/training_result.py:17-21): float32 column sums
            // added in float64 (the order over rollouts is the atomics' -- float64 sums of a handful of terms)
            if (tid < obs) {
                if (save_p) { atomicAdd(p.ob_sum + tid, (double)sum_p); atomicAdd(p.ob_sumsq + tid, (double)sq_p); }
                if (save_m) { atomicAdd(p.ob_sum + tid, (double)sum_m); atomicAdd(p.ob_sumsq + tid, (double)sq_m); }
            }
            if (tid == 0) {
                const int n = (save_p ? 1 : 0) + (save_m ? 1 : 0);
                atomicAdd(p.ob_count, (double)(n * T));
                atomicAdd(p.ob_count + 1, (double)n);
            }
        }
        __syncthreads();                                            // the next pair overwrites the shared weights
    }
}

template <int J>
int cl_launch(es_ctx* ctx, const ClParams& p, cudaStream_t stream) {
    const ClLayout L = cl_layout(J, p.obs, p.act, p.band);
    const size_t smem = (size_t)L.total * sizeof(float);
    if (smem > 227 * 1024) {
        es_set_error("es_rollout_closedloop: %zu bytes of shared memory needed (obs %d, act %d, band %d), 227 KB available", smem, p.obs,
                     p.act, p.band);
        return ES_ERR_UNSUPPORTED;
    }
    ES_CHECK_CUDA(cudaFuncSetAttribute(rollout_closed_kernel<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.n_pairs < ctx->sm_count ? p.n_pairs : ctx->sm_count;
    rollout_closed_kernel<J><<<grid, CL_THREADS, smem, stream>>>(p);
    ES_LAUNCHED(ctx);
    return ES_OK;
}

}  // namespace

int es_impl_rollout_closed(es_ctx* ctx, const float* table, int64_t table_len, const int64_t* idx, int n_pairs, const float* theta,
                           int n_params, float sigma, const int* dims, const double* ob_mean, const double* ob_std, double ob_clip,
                           const float* obs0, const float* env_a, int band, const float* env_b, const float* rew_vec, int T,
                           float pos_scale, const uint32_t* coins, double chance, double* fit_pos, double* fit_neg, int fit_stride,
                           float* behv_pos, float* behv_neg, double* ob_sum, double* ob_sumsq, double* ob_count,
                           cudaStream_t stream) {
    ClParams p;
    p.table = table; p.table_len = table_len; p.idx = idx; p.n_pairs = n_pairs; p.theta = theta; p.sigma = sigma;
    p.obs = dims[0]; p.h1 = dims[1]; p.h2 = dims[2]; p.act = dims[3]; p.T = T; p.n_params = n_params;
    p.ob_mean = ob_mean; p.ob_std = ob_std; p.ob_clip = ob_clip;
    p.obs0 = obs0; p.env_a = env_a; p.env_b = env_b; p.band = band; p.crew = rew_vec; p.pos_scale = pos_scale;
    p.coins = coins; p.chance = chance;
    p.fit_pos = fit_pos; p.fit_neg = fit_neg; p.fit_stride = fit_stride; p.behv_pos = behv_pos; p.behv_neg = behv_neg;
    p.ob_sum = ob_sum; p.ob_sumsq = ob_sumsq; p.ob_count = ob_count;
    p.err = ctx->err_dev;
    if (p.h1 > CL_H || p.h2 > CL_H || p.act > CL_A || p.obs > 384 || p.obs > CL_THREADS - 32) {
        es_set_error("es_rollout_closedloop: supports obs <= 384, hidden <= %d, act <= %d (got %d-%d-%d-%d)", CL_H, CL_A, p.obs, p.h1,
                     p.h2, p.act);
        return ES_ERR_UNSUPPORTED;
    }
    if (p.obs <= 32) return cl_launch<4>(ctx, p, stream);
    if (p.obs <= 128) return cl_launch<16>(ctx, p, stream);
    return cl_launch<48>(ctx, p, stream);
}
