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
//   * layer 1 (82 % of the weights): 512 threads x 128 registers; thread (o = tid / 8, s = tid % 8) owns the elements
//     k = 8 j + s of row o and keeps BOTH signs' weights in registers (2 x 48), so layer 1 reads only the observations from
//     shared memory: float2 (x+, x-) in [s][j] order, one 16-byte load per two elements, rows padded so that the eight rows a
//     warp reads at once sit in different banks; the 8 partial sums of a row meet by three shuffles;
//   * layers 2 / 3: both signs' weights in shared memory as float4 (w+[j], w-[j], w+[j+1], w-[j+1]), same thread <-> element
//     map, activations in the same [s][j] float2 layout;
//   * the env step: thread i owns observation i of both signs (A's diagonals and B transposed in shared memory; the raw
//     observations carry a wrap-around halo so the band is a linear read), normalises the new observation with its mean / std
//     (float64 like the reference: a float32 tensor minus a float64 ndarray) and keeps the float32 column sums of the
//     post-step observations for the ObStat of a rollout whose save_obs coin fell;
//   * reward (float32 dot in index order, summed in float64 like python's sum) and position by the last warp.
// Four barriers per step.  Measured (K = 10 000, T = 1 000, Humanoid shapes): 188 ms per generation = 2.8 us per step and SM.
// On the way (ncu, profiles/README.md): 512 threads with the - weights in shared memory were bound by the shared-memory pipe
// (4 550 wavefronts per step, mio_throttle, 215 ms); 256 threads x 255 registers removed the wavefronts but left two warps per
// scheduler waiting on every shared-memory load (218 ms); a row stride of the observation vector that is a multiple of 32
// words made every layer-1 load a 4-way bank conflict (371 ms); packed fma.rn.f32x2 on the (+, -) pairs (this version) cut
// layer 1 from 244 to 186 instructions per warp and step for 1 % of time.  What bounds it (source-level samples): layer 1 is
// 39 % of a step and 3/4 of that is short_scoreboard on the FMA that follows each 16-byte observation load -- the layer-1
// weights take 96 of the 128 registers, so the compiler keeps no load in flight and 4 warps per scheduler do not cover the
// shared-memory latency.  The register file is the constraint (2 x 24 064 weights are 73 % of it).  Parking the weights in
// tensor memory (both signs fit in 384 of the 512 columns) does not help: tcgen05.ld reads 64 B per cycle, i.e. ~3 000 cycles
// for the 196 KB a step needs, more than layer 1 takes now.
#include <math.h>
#include "common.cuh"

namespace {

constexpr int CL_G = 8;                    // lanes per row of a layer (its elements k = CL_G j + s)
constexpr int CL_THREADS = 64 * CL_G;
constexpr int CL_H = 64;                   // hidden units per layer at most (thread rows)
constexpr int CL_A = 64;                   // action units at most (layer-3 rows)
constexpr int CL_J23 = CL_H / CL_G;        // elements per thread in layers 2 / 3

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

constexpr int CL_HS = CL_J23 + 2;         // row stride (float2) of the [s][j] hidden-activation vectors: 16-byte aligned rows, conflict-free
constexpr int CL_HALO = 16;                // the raw observations carry a wrap-around halo of band entries (band <= 16)

// row stride (float2) of the [s][j] observation vector: >= J + 2 (the layer-1 loads run two elements past the row), 16-byte
// aligned rows, and the four rows a warp reads at once in different banks: stride = 2 (mod 4)
__host__ __device__ constexpr int cl_xs(int J) { return (J % 4 == 0) ? J + 2 : J + 4; }

struct ClLayout {                          // offsets in floats into dynamic shared memory
    int norm, w2, w3, bias, env_a, env_b, x2, o2, h1, h2, a2, prod, stat, racc, total, o2_stride, act_pad;
};
__host__ __device__ inline ClLayout cl_layout(int J, int obs, int act, int band) {
    ClLayout L;
    int at = 0;
    L.act_pad = (act + 3) & ~3;
    L.o2_stride = obs + CL_HALO;
    L.norm = at; at += 4 * obs;             // double mean[obs], double 1/std[obs]  (first: 8-byte aligned)
    L.racc = at; at += 16;                  // fitness (double x 2) and position (float x 6) of the pair, last warp
    L.x2 = at; at += 2 * CL_G * cl_xs(J);   // float2 (x+, x-) normalised observations, [s][cl_xs(J)]
    L.w2 = at; at += 4 * (CL_J23 / 2) * CL_THREADS;      // float4 (w+[j], w-[j], w+[j+1], w-[j+1]) at [j / 2][tid]
    L.w3 = at; at += 4 * (CL_J23 / 2) * CL_THREADS;
    L.h1 = at; at += 2 * CL_G * CL_HS;      // float2 (h+, h-), [s][CL_HS]
    L.h2 = at; at += 2 * CL_G * CL_HS;
    L.stat = at; at += 4 * obs;             // float4 (sum+, sumsq+, sum-, sumsq-) of the post-step observations
    L.bias = at; at += 6 * CL_H;            // b1+ b1- b2+ b2- b3+ b3-
    L.o2 = at; at += 2 * 2 * L.o2_stride;   // [2 buffers] float2 raw observations with halo
    L.a2 = at; at += 2 * CL_A;
    L.prod = at; at += 2 * CL_A;
    L.env_a = at; at += band * obs;
    L.env_b = at; at += L.act_pad * obs;
    L.total = at;
    return L;
}

__device__ __forceinline__ float cl_group_sum(float v) {       // sum over the CL_G lanes that share a row
#pragma unroll
    for (int m = 1; m < CL_G; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
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
// tanh(x) = 1 - 2 / (1 + e^2x) with the fast exponential and division: absolute error ~1e-7 (tanhf is ~40 dependent
// instructions per call, and every phase of a step ends in one)
__device__ __forceinline__ float cl_tanh(float x) {
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, 1.f + e);
}

// packed float32 pairs: (+, -) of a weight times (x+, x-) of an input is ONE fma.rn.f32x2 (two independent, identically
// rounded FMAs per instruction)
typedef unsigned long long cl_u64;
__device__ __forceinline__ cl_u64 cl_pk(float lo, float hi) {
    cl_u64 v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ void cl_unpk(cl_u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ cl_u64 cl_fma2(cl_u64 a, cl_u64 b, cl_u64 c) {
    cl_u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ cl_u64 cl_add2(cl_u64 a, cl_u64 b) {
    cl_u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int J>
__global__ void __launch_bounds__(CL_THREADS, 1) rollout_closed_kernel(const ClParams p) {
    extern __shared__ __align__(16) float cl_smem[];
    const ClLayout L = cl_layout(J, p.obs, p.act, p.band);
    double* __restrict__ nmean = reinterpret_cast<double*>(cl_smem + L.norm);
    double* __restrict__ nrstd = nmean + p.obs;
    double* __restrict__ rfit = reinterpret_cast<double*>(cl_smem + L.racc);        // [2]
    float* __restrict__ rpos = cl_smem + L.racc + 4;                                  // [2][3]
    float4* __restrict__ W2 = reinterpret_cast<float4*>(cl_smem + L.w2);
    float4* __restrict__ W3 = reinterpret_cast<float4*>(cl_smem + L.w3);
    float* __restrict__ bias = cl_smem + L.bias;
    float* __restrict__ envA = cl_smem + L.env_a;
    float* __restrict__ envB = cl_smem + L.env_b;
    float2* __restrict__ x2 = reinterpret_cast<float2*>(cl_smem + L.x2);
    float2* __restrict__ o2 = reinterpret_cast<float2*>(cl_smem + L.o2);
    float2* __restrict__ h1v = reinterpret_cast<float2*>(cl_smem + L.h1);
    float2* __restrict__ h2v = reinterpret_cast<float2*>(cl_smem + L.h2);
    float2* __restrict__ a2 = reinterpret_cast<float2*>(cl_smem + L.a2);
    float2* __restrict__ prod = reinterpret_cast<float2*>(cl_smem + L.prod);
    float4* __restrict__ stat = reinterpret_cast<float4*>(cl_smem + L.stat);
    constexpr int XS = cl_xs(J);

    const int tid = threadIdx.x, o = tid / CL_G, s = tid % CL_G, warp = tid >> 5, lane = tid & 31;
    const int obs = p.obs, h1 = p.h1, h2 = p.h2, act = p.act, T = p.T, band = p.band, half = p.band >> 1;
    const int act_pad = L.act_pad, o2s = L.o2_stride;
    // flat parameter layout (state-dict order, src/core/policy.py:33-35): W1 [h1][obs], b1, W2 [h2][h1], b2, W3 [act][h2], b3
    const int off_b1 = h1 * obs, off_w2 = off_b1 + h1, off_b2 = off_w2 + h2 * h1, off_w3 = off_b2 + h2, off_b3 = off_w3 + act * h2;
    const bool l3_warp = (tid & ~31) < CL_G * act;              // warps that hold rows of layer 3
    const bool rew_warp = warp == CL_THREADS / 32 - 1;
    const int hslot = (o % CL_G) * CL_HS + (o / CL_G);          // where row o's activation goes in a [s][j] vector

    for (int i = tid; i < band * obs; i += CL_THREADS) envA[i] = p.env_a[i];
    for (int i = tid; i < act_pad * obs; i += CL_THREADS) envB[i] = i < act * obs ? p.env_b[i] : 0.f;
    for (int i = tid; i < CL_G * XS; i += CL_THREADS) x2[i] = make_float2(0.f, 0.f);
    for (int i = tid; i < CL_G * CL_HS; i += CL_THREADS) { h1v[i] = make_float2(0.f, 0.f); h2v[i] = make_float2(0.f, 0.f); }
    for (int i = tid; i < CL_A; i += CL_THREADS) a2[i] = make_float2(0.f, 0.f);
    for (int i = tid; i < obs; i += CL_THREADS) { nmean[i] = p.ob_mean[i]; nrstd[i] = 1.0 / p.ob_std[i]; }
    __syncthreads();

    // a raw observation and its halo copies: slot q = k + half for k in [-half, obs + band - half)
    auto put_obs = [&](float2* __restrict__ buf, int i, float2 v) {
        buf[i + half] = v;
        if (i >= obs - half) buf[i - obs + half] = v;
        if (i < band - half) buf[i + obs + half] = v;
    };

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
        cl_u64 w1[J];                                              // (w+, w-) of this thread's layer-1 elements
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = CL_G * j + s;
            float a = 0.f, b = 0.f;
            if (o < h1 && k < obs) wpm(o * obs + k, a, b);
            w1[j] = cl_pk(a, b);
        }
        for (int j = 0; j < CL_J23; j += 2) {
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o < h2 && CL_G * j + s < h1) wpm(off_w2 + o * h1 + CL_G * j + s, w.x, w.y);
            if (o < h2 && CL_G * (j + 1) + s < h1) wpm(off_w2 + o * h1 + CL_G * (j + 1) + s, w.z, w.w);
            W2[(j >> 1) * CL_THREADS + tid] = w;
            w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o < act && CL_G * j + s < h2) wpm(off_w3 + o * h2 + CL_G * j + s, w.x, w.y);
            if (o < act && CL_G * (j + 1) + s < h2) wpm(off_w3 + o * h2 + CL_G * (j + 1) + s, w.z, w.w);
            W3[(j >> 1) * CL_THREADS + tid] = w;
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
        for (int i = tid; i < obs; i += CL_THREADS) {
            const float v = p.obs0[i];
            put_obs(o2, i, make_float2(v, v));
            const float xn = cl_normalise(v, nmean[i], nrstd[i], p.ob_clip);
            x2[(i % CL_G) * XS + (i / CL_G)] = make_float2(xn, xn);
            stat[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid < 2) rfit[tid] = 0.0;
        if (tid < 6) rpos[tid] = 0.f;
        // the save_obs coins of the pair's two evaluations (legacy random_sample: (a >> 5, b >> 6) / 2^53 < chance)
        bool save_p = false, save_m = false;
        if (p.coins) {
            const uint32_t* c = p.coins + (size_t)pair * 4;
            save_p = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) / 9007199254740992.0 < p.chance;
            save_m = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6)) / 9007199254740992.0 < p.chance;
        }
        const bool keep_stat = p.ob_sum && (save_p || save_m);
        __syncthreads();

        for (int t = 0; t < T; ++t) {
            const int cur = t & 1;
            float crow0 = 0.f, crow1 = 0.f;
            if (rew_warp) {                                         // this step's reward coefficients: in flight under the layers
                const float* __restrict__ c = p.crew + (size_t)t * act;
                if (lane < act) crow0 = __ldg(c + lane);
                if (lane + 32 < act) crow1 = __ldg(c + lane + 32);
            }
            // ---- layer 1: weights in registers; observations by 16-byte loads (two elements of both signs), three loads in
            //      flight (with 8 warps nothing else hides the shared-memory latency) ----
            {
                cl_u64 za = 0ull, zb = 0ull;                        // (z+, z-) over the even / the odd elements
                const ulonglong2* __restrict__ xr = reinterpret_cast<const ulonglong2*>(x2 + s * XS);
                ulonglong2 xa = xr[0], xb = xr[1], xc = xr[2];      // ((x+, x-)[j], (x+, x-)[j+1]); the row is padded
#pragma unroll
                for (int j = 0; j < J; j += 2) {
                    const ulonglong2 xv = xa;
                    xa = xb; xb = xc;
                    if (j + 6 < J + 2) xc = xr[(j >> 1) + 3];
                    za = cl_fma2(w1[j], xv.x, za);
                    zb = cl_fma2(w1[j + 1], xv.y, zb);
                }
                float zp, zm;
                cl_unpk(cl_add2(za, zb), zp, zm);
                zp = cl_group_sum(zp); zm = cl_group_sum(zm);
                if (s == 0) h1v[hslot] = (o < h1) ? make_float2(cl_tanh(zp + bias[o]), cl_tanh(zm + bias[CL_H + o])) : make_float2(0.f, 0.f);
            }
            __syncthreads();
            // ---- layer 2: one 16-byte load brings both signs' weights of two elements, another the two activations ----
            {
                cl_u64 za = 0ull, zb = 0ull;
                const ulonglong2* __restrict__ hr = reinterpret_cast<const ulonglong2*>(h1v + s * CL_HS);
                const ulonglong2* __restrict__ wr = reinterpret_cast<const ulonglong2*>(W2) + tid;
#pragma unroll
                for (int j = 0; j < CL_J23 / 2; ++j) {
                    const ulonglong2 w = wr[j * CL_THREADS], hv = hr[j];
                    za = cl_fma2(w.x, hv.x, za);
                    zb = cl_fma2(w.y, hv.y, zb);
                }
                float zp, zm;
                cl_unpk(cl_add2(za, zb), zp, zm);
                zp = cl_group_sum(zp); zm = cl_group_sum(zm);
                if (s == 0) h2v[hslot] = (o < h2) ? make_float2(cl_tanh(zp + bias[2 * CL_H + o]), cl_tanh(zm + bias[3 * CL_H + o])) : make_float2(0.f, 0.f);
            }
            __syncthreads();
            // ---- layer 3 (only the warps that hold its rows; rows >= act of the last such warp hold zero weights) ----
            if (l3_warp) {
                cl_u64 za = 0ull, zb = 0ull;
                const ulonglong2* __restrict__ hr = reinterpret_cast<const ulonglong2*>(h2v + s * CL_HS);
                const ulonglong2* __restrict__ wr = reinterpret_cast<const ulonglong2*>(W3) + tid;
#pragma unroll
                for (int j = 0; j < CL_J23 / 2; ++j) {
                    const ulonglong2 w = wr[j * CL_THREADS], hv = hr[j];
                    za = cl_fma2(w.x, hv.x, za);
                    zb = cl_fma2(w.y, hv.y, zb);
                }
                float zp, zm;
                cl_unpk(cl_add2(za, zb), zp, zm);
                zp = cl_group_sum(zp); zm = cl_group_sum(zm);
                if (s == 0 && o < act) a2[o] = make_float2(cl_tanh(zp + bias[4 * CL_H + o]), cl_tanh(zm + bias[5 * CL_H + o]));
            }
            __syncthreads();
            // ---- env step: thread i owns observation i; the raw observations carry a wrap-around halo, so the band is a
            //      linear read ----
            if (tid < obs) {
                const float2* __restrict__ oc = o2 + cur * o2s;
                const int i = tid;
                float ap = 0.f, am = 0.f, ap1 = 0.f, am1 = 0.f;
#pragma unroll 4
                for (int d = 0; d < band; d += 2) {                 // (band is even, checked on the host)
                    const float w0 = envA[d * obs + i], w1 = envA[(d + 1) * obs + i];
                    const float2 u0 = oc[i + d], u1 = oc[i + d + 1];
                    ap = fmaf(w0, u0.x, ap); am = fmaf(w0, u0.y, am);
                    ap1 = fmaf(w1, u1.x, ap1); am1 = fmaf(w1, u1.y, am1);
                }
#pragma unroll 2
                for (int j = 0; j < act_pad; j += 2) {
                    const float w0 = envB[j * obs + i], w1 = envB[(j + 1) * obs + i];
                    const float4 av = *reinterpret_cast<const float4*>(a2 + j);
                    ap = fmaf(w0, av.x, ap); am = fmaf(w0, av.y, am);
                    ap1 = fmaf(w1, av.z, ap1); am1 = fmaf(w1, av.w, am1);
                }
                const float np_ = cl_tanh(ap + ap1), nm = cl_tanh(am + am1);
                put_obs(o2 + (cur ^ 1) * o2s, i, make_float2(np_, nm));
                const double mu = nmean[i], rs_ = nrstd[i];
                x2[(i % CL_G) * XS + (i / CL_G)] = make_float2(cl_normalise(np_, mu, rs_, p.ob_clip), cl_normalise(nm, mu, rs_, p.ob_clip));
                if (keep_stat) {                                    // float32 column sums in step order (numpy's axis-0 reduction)
                    float4 st = stat[i];
                    st.x = __fadd_rn(st.x, np_); st.y = __fadd_rn(st.y, __fmul_rn(np_, np_));
                    st.z = __fadd_rn(st.z, nm);  st.w = __fadd_rn(st.w, __fmul_rn(nm, nm));
                    stat[i] = st;
                }
            }
            // ---- reward and position: the last warp's lanes form the products, lanes 0 / 1 add them in index order (the env's
            //      float32 dot) for the + / - evaluation ----
            if (rew_warp) {
                if (lane < act) { const float2 av = a2[lane]; prod[lane] = make_float2(__fmul_rn(av.x, crow0), __fmul_rn(av.y, crow0)); }
                if (lane + 32 < act) { const float2 av = a2[lane + 32]; prod[lane + 32] = make_float2(__fmul_rn(av.x, crow1), __fmul_rn(av.y, crow1)); }
                __syncwarp();
                if (lane < 2) {
                    float acc = 0.f;
                    for (int j = 0; j < act; ++j) {
                        const float2 pv = prod[j];
                        acc = __fadd_rn(acc, lane ? pv.y : pv.x);
                    }
                    rfit[lane] += (double)acc;
                    const float ps = p.pos_scale;
                    const float2 q0 = a2[0], q1 = a2[1 % act], q2 = a2[2 % act];
                    rpos[lane * 3 + 0] = __fadd_rn(rpos[lane * 3 + 0], __fmul_rn(ps, lane ? q0.y : q0.x));
                    rpos[lane * 3 + 1] = __fadd_rn(rpos[lane * 3 + 1], __fmul_rn(ps, lane ? q1.y : q1.x));
                    rpos[lane * 3 + 2] = __fadd_rn(rpos[lane * 3 + 2], __fmul_rn(ps, lane ? q2.y : q2.x));
                }
                __syncwarp();
            }
            __syncthreads();
        }
        if (rew_warp && lane < 2) {
            (lane ? p.fit_neg : p.fit_pos)[(size_t)pair * p.fit_stride] = rfit[lane];
            float* bv = lane ? p.behv_neg : p.behv_pos;
            if (bv) { bv[(size_t)pair * 3 + 0] = rpos[lane * 3 + 0]; bv[(size_t)pair * 3 + 1] = rpos[lane * 3 + 1]; bv[(size_t)pair * 3 + 2] = rpos[lane * 3 + 2]; }
        }
        if (keep_stat) {
            // ObStat.inc of the saved rollouts (src/core/es.py:73-74, src/gym/training_result.py:17-21): float32 column sums
            // added in float64 (the order over rollouts is the atomics' -- float64 sums of a handful of terms)
            for (int i = tid; i < obs; i += CL_THREADS) {
                const float4 st = stat[i];
                if (save_p) { atomicAdd(p.ob_sum + i, (double)st.x); atomicAdd(p.ob_sumsq + i, (double)st.y); }
                if (save_m) { atomicAdd(p.ob_sum + i, (double)st.z); atomicAdd(p.ob_sumsq + i, (double)st.w); }
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
    if (p.h1 > CL_H || p.h2 > CL_H || p.act > CL_A || p.obs > 384 || p.band > CL_HALO || p.obs < p.band || (p.band & 1)) {
        es_set_error("es_rollout_closedloop: supports even band <= obs <= 384, hidden <= %d, act <= %d, band <= %d (got %d-%d-%d-%d, band %d)",
                     CL_H, CL_A, CL_HALO, p.obs, p.h1, p.h2, p.act, p.band);
        return ES_ERR_UNSUPPORTED;
    }
    if (p.obs <= 32) return cl_launch<4>(ctx, p, stream);
    if (p.obs <= 128) return cl_launch<16>(ctx, p, stream);
    return cl_launch<48>(ctx, p, stream);
}
