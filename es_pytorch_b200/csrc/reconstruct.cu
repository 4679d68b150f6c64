// reconstruct.cu -- gradient reconstruction  out[p] = sum_k w[k] * table[idx[k] + p]
// (reference: scale_noise / batch_noise, src/utils/utils.py:14-39).
//
// HBM-bound streaming kernel: the n_idx slices (P floats each, starting at arbitrary
// 4-byte-aligned offsets of the table) are read exactly once; everything else is noise.
//
// Decomposition: grid = (column tiles) x (slice chunks).
//   * a CTA owns RC_TILE_P = 1024 consecutive columns and `k_per_chunk` slices;
//   * a warp owns 128 consecutive columns; lane l accumulates columns l, l+32, l+64, l+96,
//     so each warp request is one fully coalesced 128-byte run and a slice row costs the
//     warp 512 contiguous bytes (17 sectors for an unaligned start instead of 16: 6 %
//     over-fetch, independent of what neighbouring warps do);
//   * slices are unrolled by RC_UNROLL so every thread keeps 4*RC_UNROLL independent
//     4-byte loads in flight (memory-level parallelism is what saturates HBM here);
//   * per-chunk partial sums go to a scratch [n_chunks][P]; the last CTA to finish a
//     column tile (atomic ticket) adds the partials in chunk order -> deterministic.
#include <stdlib.h>
#include "common.cuh"

constexpr int RC_THREADS = 256;
constexpr int RC_WARPS = RC_THREADS / 32;
constexpr int RC_COLS_PER_LANE = 4;
constexpr int RC_TILE_P = RC_WARPS * 32 * RC_COLS_PER_LANE;  // 1024
constexpr int RC_UNROLL = 4;
constexpr int RC_MAX_CHUNK = 1024;  // slices staged in shared memory per CTA

__global__ void __launch_bounds__(RC_THREADS, 4)
reconstruct_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, const float* __restrict__ w,
                   int n_idx, int P, long long table_len, int* __restrict__ err, int k_per_chunk, int n_chunks,
                   float* __restrict__ partials, unsigned* __restrict__ tickets, float* __restrict__ out) {
    __shared__ long long s_off[RC_MAX_CHUNK];
    __shared__ float s_w[RC_MAX_CHUNK];
    __shared__ bool s_last;

    const int tile = blockIdx.x;
    const int chunk = blockIdx.y;
    const int k0 = chunk * k_per_chunk;
    const int kn = min(k_per_chunk, n_idx - k0);                      // >= 1 by construction
    const int kn_pad = (kn + RC_UNROLL - 1) / RC_UNROLL * RC_UNROLL;  // <= RC_MAX_CHUNK

    for (int i = threadIdx.x; i < kn_pad; i += RC_THREADS) {
        // padding entries read slice 0 with weight 0 (a valid address, contributes +0)
        s_off[i] = (i < kn) ? es_checked_slice((long long)idx[k0 + i], P, table_len, err) : 0ll;   // noisetable.py:34
        s_w[i] = (i < kn) ? w[k0 + i] : 0.0f;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col0 = tile * RC_TILE_P + warp * (32 * RC_COLS_PER_LANE) + lane;
    bool ok[RC_COLS_PER_LANE];
    int col[RC_COLS_PER_LANE];
#pragma unroll
    for (int j = 0; j < RC_COLS_PER_LANE; ++j) {
        col[j] = col0 + 32 * j;
        ok[j] = col[j] < P;
        if (!ok[j]) col[j] = 0;  // keep the address valid; result discarded
    }

    float acc[RC_COLS_PER_LANE] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < kn_pad; k += RC_UNROLL) {
        float v[RC_UNROLL][RC_COLS_PER_LANE];
        float wk[RC_UNROLL];
#pragma unroll
        for (int u = 0; u < RC_UNROLL; ++u) {
            const float* __restrict__ row = table + s_off[k + u];
            wk[u] = s_w[k + u];
#pragma unroll
            for (int j = 0; j < RC_COLS_PER_LANE; ++j) v[u][j] = __ldg(row + col[j]);
        }
#pragma unroll
        for (int u = 0; u < RC_UNROLL; ++u)
#pragma unroll
            for (int j = 0; j < RC_COLS_PER_LANE; ++j) acc[j] = fmaf(wk[u], v[u][j], acc[j]);
    }

    if (n_chunks == 1) {
#pragma unroll
        for (int j = 0; j < RC_COLS_PER_LANE; ++j)
            if (ok[j]) out[col[j]] = acc[j];
        return;
    }

    float* my = partials + (size_t)chunk * P;
#pragma unroll
    for (int j = 0; j < RC_COLS_PER_LANE; ++j)
        if (ok[j]) __stcg(my + col[j], acc[j]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&tickets[tile], 1u);
        s_last = (t == (unsigned)(n_chunks - 1));
        if (s_last) tickets[tile] = 0;  // re-arm for the next launch (stream-ordered)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA of this column tile: reduce the chunk partials in fixed order
    float sum[RC_COLS_PER_LANE] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < n_chunks; ++c) {
        const float* src = partials + (size_t)c * P;
#pragma unroll
        for (int j = 0; j < RC_COLS_PER_LANE; ++j) sum[j] += __ldcg(src + col[j]);
    }
#pragma unroll
    for (int j = 0; j < RC_COLS_PER_LANE; ++j)
        if (ok[j]) out[col[j]] = sum[j];
}

int es_impl_grad_reconstruct(es_ctx* ctx, const float* table, int64_t table_len, const int64_t* idx,
                             const float* weights, int n_idx, int P, float* out, cudaStream_t stream) {
    const int n_tiles = es_div_up(P, RC_TILE_P);
    // one wave: n_tiles * n_chunks <= sm_count * resident CTAs per SM (launch bound: 4), so no
    // tail wave; never stage more than RC_MAX_CHUNK slices per CTA.
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        const char* e = getenv("ES_RC_CTAS_PER_SM");  // tuning knob for profiling runs
        ctas_per_sm = e ? atoi(e) : 4;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int target_chunks = (int)(((int64_t)ctx->sm_count * ctas_per_sm) / n_tiles);
    if (target_chunks < 1) target_chunks = 1;
    int k_per_chunk = es_div_up(n_idx, target_chunks);
    if (k_per_chunk < 8) k_per_chunk = 8;
    k_per_chunk = es_div_up(k_per_chunk, RC_UNROLL) * RC_UNROLL;
    if (k_per_chunk > RC_MAX_CHUNK) k_per_chunk = RC_MAX_CHUNK;
    const int n_chunks = es_div_up(n_idx, k_per_chunk);
    ES_REQUIRE(n_chunks <= 65535, "es_grad_reconstruct: too many slice chunks (%d)", n_chunks);

    float* partials = nullptr;
    unsigned* tickets = nullptr;
    if (n_chunks > 1) {
        void* p = nullptr;
        int rc = es_ctx_scratch(ctx, (size_t)n_chunks * P * sizeof(float), &p);
        if (rc) return rc;
        partials = (float*)p;
        ES_REQUIRE(n_tiles < 4000, "es_grad_reconstruct: P too large for the ticket array");
        rc = es_ctx_counters(ctx, 4096, &tickets);
        if (rc) return rc;
    }
    dim3 grid(n_tiles, n_chunks);
    reconstruct_kernel<<<grid, RC_THREADS, 0, stream>>>(table, idx, weights, n_idx, P, (long long)table_len, ctx->err_dev,
                                                        k_per_chunk, n_chunks, partials, tickets, out);
    ES_LAUNCHED(ctx);
    return ES_OK;
}
