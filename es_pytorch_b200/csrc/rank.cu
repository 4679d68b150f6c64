// rank.cu -- centered-rank transform of the 2K antithetic fitnesses -> K weights
// (reference: rank / Ranker / CenteredRanker / MultiObjectiveRanker,
//  src/utils/rankers.py:9-17,37-58,106-120).
//
// rank(x_i) = #{j : x_j < x_i} + #{j < i : x_j == x_i}  over x = concat(pos, neg)
// which equals ``ranks[argsort(x, kind='stable')] = arange`` -- integer-exact, and the
// count form shards: a GPU ranks only its own 2*k_count elements against all 2K keys.
// Keys are the float64 fitnesses mapped to order-preserving uint64 (-0.0 == +0.0 and
// NaNs last, like numpy's sort), so the O(n * n_local) inner loop is integer compares.
//   1. rank_keys_kernel      fitness -> keys[n_obj][2K]
//   2. rank_count_kernel     counts (int atomics: order-independent -> deterministic)
//   3. rank_finalize_kernel  y = float32(rank)/(2K-1) - 0.5, blend, weight = y+ - y-
// The float32 ops are the reference's, one IEEE operation each (no contraction).
#include "common.cuh"

constexpr int RK_THREADS = 256;
constexpr int RK_EPT = 4;          // elements ranked per thread
constexpr int RK_JTILE = 1024;     // keys staged in shared memory per step
constexpr int RK_JCHUNK = 1024;    // keys per CTA along j (grid.y = 2K / RK_JCHUNK)

__device__ __forceinline__ unsigned long long rk_key(double x) {
    if (x != x) return 0xFFFFFFFFFFFFFFFFull;     // NaN sorts last
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    if ((b << 1) == 0ull) b = 0ull;               // -0.0 -> +0.0 (they compare equal)
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void rank_keys_kernel(const double* __restrict__ fpos, const double* __restrict__ fneg, int K, int n_obj,
                                 unsigned long long* __restrict__ keys) {
    const int n = 2 * K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * n_obj; i += gridDim.x * blockDim.x) {
        const int c = i / n, e = i - c * n;
        const double x = (e < K) ? fpos[(size_t)e * n_obj + c] : fneg[(size_t)(e - K) * n_obj + c];
        keys[i] = rk_key(x);
    }
}

// local element e in [0, 2*k_count): e < k_count -> global index k_begin+e (pos part),
// else K + k_begin + (e - k_count) (neg part).
__device__ __forceinline__ int rk_global_index(int e, int K, int k_begin, int k_count) {
    return (e < k_count) ? (k_begin + e) : (K + k_begin + (e - k_count));
}

__global__ void __launch_bounds__(RK_THREADS)
rank_count_kernel(const unsigned long long* __restrict__ keys, int K, int k_begin, int k_count,
                  int* __restrict__ ranks /*[n_obj][2*k_count]*/) {
    __shared__ unsigned long long s_keys[RK_JTILE];
    const int n = 2 * K;
    const int n_local = 2 * k_count;
    const int c = blockIdx.z;
    const unsigned long long* kc = keys + (size_t)c * n;

    unsigned long long mykey[RK_EPT];
    int myidx[RK_EPT];
    int cnt[RK_EPT];
    const int e0 = (blockIdx.x * RK_THREADS + threadIdx.x) * RK_EPT;
#pragma unroll
    for (int r = 0; r < RK_EPT; ++r) {
        const int e = e0 + r;
        cnt[r] = 0;
        if (e < n_local) {
            myidx[r] = rk_global_index(e, K, k_begin, k_count);
            mykey[r] = kc[myidx[r]];
        } else {
            myidx[r] = -1;          // counts nothing: no j satisfies j < -1, key 0 has nothing below
            mykey[r] = 0ull;
        }
    }

    const int j_begin = blockIdx.y * RK_JCHUNK;
    const int j_end = min(n, j_begin + RK_JCHUNK);
    for (int jt = j_begin; jt < j_end; jt += RK_JTILE) {
        const int m = min(RK_JTILE, j_end - jt);
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += RK_THREADS) s_keys[t] = kc[jt + t];
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < m; ++t) {
            const unsigned long long kj = s_keys[t];
            const int j = jt + t;
#pragma unroll
            for (int r = 0; r < RK_EPT; ++r) cnt[r] += (kj < mykey[r]) || (kj == mykey[r] && j < myidx[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < RK_EPT; ++r) {
        const int e = e0 + r;
        if (e < n_local && cnt[r]) atomicAdd(&ranks[(size_t)c * n_local + e], cnt[r]);
    }
}

__global__ void rank_finalize_kernel(const int* __restrict__ ranks, int K, int n_obj, float w0, float w1, int k_count,
                                     float* __restrict__ weights_out, int32_t* __restrict__ ranks_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_count) return;
    const int n_local = 2 * k_count;
    const float denom = (float)(2 * K - 1);             // y /= (x.size - 1), rankers.py:56
    float yp, yn;
    {
        const int rp = ranks[k], rn = ranks[k_count + k];
        yp = __fsub_rn(__fdiv_rn((float)rp, denom), 0.5f);   // rankers.py:55-57
        yn = __fsub_rn(__fdiv_rn((float)rn, denom), 0.5f);
        if (ranks_out) { ranks_out[k] = rp; ranks_out[k_count + k] = rn; }
    }
    if (n_obj == 2) {
        const int rp = ranks[n_local + k], rn = ranks[n_local + k_count + k];
        const float yp1 = __fsub_rn(__fdiv_rn((float)rp, denom), 0.5f);
        const float yn1 = __fsub_rn(__fdiv_rn((float)rn, denom), 0.5f);
        yp = __fadd_rn(__fmul_rn(yp, w0), __fmul_rn(yp1, w1));  // rankers.py:120
        yn = __fadd_rn(__fmul_rn(yn, w0), __fmul_rn(yn1, w1));
        if (ranks_out) { ranks_out[n_local + k] = rp; ranks_out[n_local + k_count + k] = rn; }
    }
    weights_out[k] = __fsub_rn(yp, yn);                  // rankers.py:44
}

int es_impl_centered_rank(es_ctx* ctx, const double* fpos, const double* fneg, int K, int n_obj, float w0, float w1,
                          int k_begin, int k_count, float* weights_out, int32_t* ranks_out, cudaStream_t stream) {
    const size_t n = 2 * (size_t)K, n_local = 2 * (size_t)k_count;
    const size_t key_bytes = n * n_obj * sizeof(unsigned long long);
    const size_t rank_bytes = n_local * n_obj * sizeof(int);
    void* scratch = nullptr;
    int rc = es_ctx_scratch(ctx, key_bytes + rank_bytes, &scratch);
    if (rc) return rc;
    unsigned long long* keys = (unsigned long long*)scratch;
    int* ranks = (int*)((char*)scratch + key_bytes);

    ES_CHECK_CUDA(cudaMemsetAsync(ranks, 0, rank_bytes, stream));
    {
        int blocks = es_div_up((int64_t)n * n_obj, RK_THREADS);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        rank_keys_kernel<<<blocks, RK_THREADS, 0, stream>>>(fpos, fneg, K, n_obj, keys);
        ES_LAUNCHED(ctx);
    }
    {
        dim3 grid(es_div_up((int64_t)n_local, RK_THREADS * RK_EPT), es_div_up((int64_t)n, RK_JCHUNK), n_obj);
        ES_REQUIRE(grid.y <= 65535, "es_centered_rank: K too large for this kernel");
        rank_count_kernel<<<grid, RK_THREADS, 0, stream>>>(keys, K, k_begin, k_count, ranks);
        ES_LAUNCHED(ctx);
    }
    {
        rank_finalize_kernel<<<es_div_up(k_count, RK_THREADS), RK_THREADS, 0, stream>>>(ranks, K, n_obj, w0, w1, k_count,
                                                                                       weights_out, ranks_out);
        ES_LAUNCHED(ctx);
    }
    return ES_OK;
}
