// mt_draw.cu -- draw the K noise indices on the device, bit-exact with numpy's legacy
// RandomState.randint(0, upper_bound) (reference: NoiseTable.sample_idx,
// src/core/noisetable.py:37-40, called once per antithetic pair from es.py:67-68).
//
// numpy 1.18 legacy path for ranges < 2^32 (numpy/random/src/distributions,
// buffered_bounded_masked_uint32): mask = next_pow2(rng)-1 with rng = upper_bound-1;
// repeat v = mt19937_next32() & mask until v <= rng.  After each accepted index the
// stream optionally consumes `extra` further 32-bit outputs (the rs.random() save_obs
// coins drawn by the fit_fn of simple_example.py:38, two words per double).
//
// One warp per virtual-rank stream.  Per 624-word block the warp (a) regenerates the
// MT state in three dependency-free phases, (b) tempers all words in parallel, (c) runs
// the accept/skip state machine as a parallel scan: every lane composes the transition
// function of its contiguous chunk for all `extra`+1 entry states, a Hillis-Steele scan
// composes the 32 functions, and each lane then replays its chunk from its true entry
// state and writes indices/extras at its true output offset.
#include "common.cuh"

constexpr int MT_N = 624, MT_M = 397;
constexpr int MT_MAXS = 8;  // machine states: 0 = seeking an index, j>0 = j extra words still to consume

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
    const uint32_t y = (u & 0x80000000u) | (v & 0x7FFFFFFFu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return y;
}

// transition function of a run of words: entry state s -> (exit state, indices emitted)
struct MtFn {
    unsigned char st[MT_MAXS];
    unsigned short cn[MT_MAXS];
};

__global__ void __launch_bounds__(32)
mt_draw_kernel(uint32_t* __restrict__ mt_key, int32_t* __restrict__ mt_pos, int n_per_stream, uint32_t rng,
               uint32_t mask, int extra, int64_t* __restrict__ idx_out, uint32_t* __restrict__ extra_out) {
    __shared__ uint32_t mt[MT_N];
    __shared__ uint32_t tw[MT_N];
    __shared__ MtFn fnA[32], fnB[32];

    const int lane = threadIdx.x;
    const int stream_id = blockIdx.x;
    uint32_t* key = mt_key + (size_t)stream_id * MT_N;
    int64_t* out = idx_out + (size_t)stream_id * n_per_stream;
    uint32_t* xout = extra_out ? extra_out + (size_t)stream_id * n_per_stream * extra : nullptr;
    const int S = extra + 1;

    for (int i = lane; i < MT_N; i += 32) mt[i] = key[i];
    int pos = mt_pos[stream_id];
    __syncwarp();

    int produced = 0;  // indices emitted so far
    int state = 0;     // machine state at `pos`
    while (produced < n_per_stream || state > 0) {
        if (pos >= MT_N) {
            // regenerate: new[i] = x[i+397] ^ twist(old[i], old[i+1]); x is old for i < 227, new after
            uint32_t y[8];
            // phase A: i in [0,227)
            for (int c = 0, i = lane; i < MT_N - MT_M; i += 32, ++c) y[c] = mt[i + MT_M] ^ mt_twist(mt[i], mt[i + 1]);
            __syncwarp();
            for (int c = 0, i = lane; i < MT_N - MT_M; i += 32, ++c) mt[i] = y[c];
            __syncwarp();
            // phase B: i in [227,454) uses new[i-227], old[i], old[i+1]
            for (int c = 0, i = (MT_N - MT_M) + lane; i < 2 * (MT_N - MT_M); i += 32, ++c)
                y[c] = mt[i - (MT_N - MT_M)] ^ mt_twist(mt[i], mt[i + 1]);
            __syncwarp();
            for (int c = 0, i = (MT_N - MT_M) + lane; i < 2 * (MT_N - MT_M); i += 32, ++c) mt[i] = y[c];
            __syncwarp();
            // phase C: i in [454,623) uses new[i-227] (phase B), old[i], old[i+1]
            for (int c = 0, i = 2 * (MT_N - MT_M) + lane; i < MT_N - 1; i += 32, ++c)
                y[c] = mt[i - (MT_N - MT_M)] ^ mt_twist(mt[i], mt[i + 1]);
            __syncwarp();
            for (int c = 0, i = 2 * (MT_N - MT_M) + lane; i < MT_N - 1; i += 32, ++c) mt[i] = y[c];
            __syncwarp();
            if (lane == 0) mt[MT_N - 1] = mt[MT_M - 1] ^ mt_twist(mt[MT_N - 1], mt[0]);
            __syncwarp();
            pos = 0;
        }
        // temper the available words
        for (int i = pos + lane; i < MT_N; i += 32) tw[i] = mt_temper(mt[i]);
        __syncwarp();

        const int avail = MT_N - pos;
        const int C = (avail + 31) >> 5;  // words per lane (<= 20)
        const int b = pos + lane * C;
        const int e = min(MT_N, b + C);
        const int len = max(0, e - b);

        // accept bits of this lane's chunk, in a register: bit i <=> word b+i passes the masked rejection test
        uint32_t am = 0;
        for (int i = 0; i < len; ++i) am |= (uint32_t)((tw[b + i] & mask) <= rng) << i;

        // (1) this lane's transition function, from the bit mask alone: skip s words, then repeatedly jump to the
        //     next accept bit and skip `extra` words after it
        MtFn f;
#pragma unroll
        for (int s = 0; s < MT_MAXS; ++s) {
            int cur = s, cn = 0, st = 0;
            if (s < S) {
                while (cur < len) {
                    const uint32_t m = am >> cur;
                    if (m == 0) { cur = len; break; }
                    cur += __ffs(m) + extra;          // accepted word at cur + ffs - 1, then `extra` words to skip
                    ++cn;
                }
                st = cur - len;                       // words still to skip in the next chunk (0 = seeking)
            } else {
                st = s;
            }
            f.st[s] = (unsigned char)st;
            f.cn[s] = (unsigned short)cn;
        }
        fnA[lane] = f;
        __syncwarp();
        // (2) inclusive scan of function composition (earlier lanes first)
        MtFn* src = fnA;
        MtFn* dst = fnB;
        for (int d = 1; d < 32; d <<= 1) {
            MtFn mine = src[lane];
            if (lane >= d) {
                const MtFn prev = src[lane - d];
#pragma unroll
                for (int s = 0; s < MT_MAXS; ++s) {
                    const int mid = prev.st[s];                      // run `prev` first, then this lane's run
                    mine.st[s] = src[lane].st[mid];
                    mine.cn[s] = (unsigned short)(prev.cn[s] + src[lane].cn[mid]);
                }
            }
            dst[lane] = mine;
            __syncwarp();
            MtFn* t = src; src = dst; dst = t;
        }
        // (3) entry state / offset of this lane = exclusive prefix applied to the block entry state
        int st = state, base = produced;
        if (lane > 0) {
            st = src[lane - 1].st[state];
            base = produced + src[lane - 1].cn[state];
        }
        const int blk_state = src[31].st[state];
        const int blk_count = src[31].cn[state];
        // (4) replay with the mask; stop where the machine is seeking and everything requested is out
        int stop = MT_N;  // first unconsumed position if the stream ends inside this lane's chunk
        {
            int cur = 0;
            // leading extra words of the previous lane's last draw
            while (st > 0 && cur < len) {
                if (xout && base - 1 < n_per_stream) xout[(size_t)(base - 1) * extra + (extra - st)] = tw[b + cur];
                --st; ++cur;
            }
            while (cur < len) {                       // st == 0 here
                if (base >= n_per_stream) { stop = b + cur; break; }
                const uint32_t m = am >> cur;
                if (m == 0) { cur = len; break; }
                cur += __ffs(m) - 1;                  // position of the accepted word
                out[base] = (int64_t)(tw[b + cur] & mask);
                ++base; ++cur;
                st = extra;
                while (st > 0 && cur < len) {
                    if (xout && base - 1 < n_per_stream) xout[(size_t)(base - 1) * extra + (extra - st)] = tw[b + cur];
                    --st; ++cur;
                }
            }
        }
        // a lane whose chunk starts after the end also reports its start
        if (b < e && stop == MT_N && st == 0 && base >= n_per_stream) stop = e;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) stop = min(stop, __shfl_xor_sync(0xffffffffu, stop, o));

        if (produced + blk_count >= n_per_stream && stop < MT_N) {
            // the stream finished inside this block at `stop` (state 0 there)
            pos = stop;
            produced = n_per_stream;
            state = 0;
        } else if (produced + blk_count >= n_per_stream && blk_state == 0) {
            pos = MT_N;  // finished exactly at the block end
            produced = n_per_stream;
            state = 0;
        } else {
            pos = MT_N;
            produced += blk_count;
            state = blk_state;
        }
        __syncwarp();
    }

    for (int i = lane; i < MT_N; i += 32) key[i] = mt[i];
    if (lane == 0) mt_pos[stream_id] = pos;
}

int es_impl_draw_indices(es_ctx* ctx, uint32_t* mt_key, int32_t* mt_pos, int n_streams, int n_per_stream,
                         uint64_t upper_bound, int extra_words, int64_t* idx_out, uint32_t* extra_out,
                         cudaStream_t stream) {
    const uint32_t rng = (uint32_t)(upper_bound - 1);
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    if (rng == 0) {
        // numpy returns `low` without consuming any random word when the range is empty
        ES_CHECK_CUDA(cudaMemsetAsync(idx_out, 0, sizeof(int64_t) * (size_t)n_streams * n_per_stream, stream));
        if (extra_words == 0) return ES_OK;
        es_set_error("es_draw_indices: upper_bound == 1 with extra_words > 0 is not supported");
        return ES_ERR_UNSUPPORTED;
    }
    mt_draw_kernel<<<n_streams, 32, 0, stream>>>(mt_key, mt_pos, n_per_stream, rng, mask, extra_words, idx_out, extra_out);
    ES_LAUNCHED(ctx);
    return ES_OK;
}
