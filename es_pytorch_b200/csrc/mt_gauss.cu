// mt_gauss.cu -- one generation's draws of a virtual-rank stream when the policy adds action noise (ac_std != 0), bit-exact
// in the MT19937 word stream with numpy's legacy RandomState.
//
// Reference program order per antithetic pair and per rank (src/core/es.py:66-72 with the fit_fn of simple_example.py:37-40 /
// obj.py:53-57, and FeedForward.forward, src/nn/nn.py:47-48):
//     idx  = rs.randint(0, len(table) - P)                      masked rejection on 32-bit words (mt_draw.cu)
//     + :   rs.random() save_obs coin (`coins` doubles, 2 words each),  then T steps x rs.randn(act) * ac_std
//     - :   the same
// rs.randn is the legacy polar method (numpy/random/src/legacy/legacy-distributions.c, legacy_gauss): attempts of two doubles
// (4 words) x1, x2 = 2 u - 1 until r2 = x1^2 + x2^2 is in (0, 1); f = sqrt(-2 log(r2) / r2); returns f*x2 and caches f*x1 for
// the next call.  The number of words a rollout consumes therefore depends on the words themselves, and the indices of all
// later pairs depend on it: the stream has to be walked in order.  Attempts inside one rollout are independent of each other,
// which is the parallelism used here: the CTA of a stream tests MG_THREADS attempts at once, ranks the accepted ones with a
// ballot scan and lets every accepting thread write its two gaussians at their place in the rollout's noise array.
//
// One CTA per stream.  The state lives in a ring of MG_RING 624-word blocks in shared memory (raw words for the recurrence and
// for the state handed back, tempered words for the consumers), regenerated ahead of the cursor.
//
// Output noise[((stream * n_pairs + pair) * 2 + sign) * N + t * act + j] = float32(gauss * scale): the float64 product rounded
// once; the rollout kernels add it to the float32 action (the reference forms the float64 sum and the env rounds it to
// float32: identical except when the float32 rounding of the noise term moves the sum across a rounding boundary, 2^-24
// relative on a term that is ~1e-2 of the action).
// log() here is CUDA's (<= 1 ulp), the reference's is glibc's: a gaussian can differ in its last float64 bit, which the
// float32 noise array does not see; the accept/reject arithmetic (the only part that steers the stream) is exact.
#include "common.cuh"

namespace {

constexpr int MG_N = 624, MG_M = 397, MG_D = MG_N - MG_M;      // 227
constexpr int MG_THREADS = 512, MG_WARPS = MG_THREADS / 32;
constexpr int MG_RING = 8;                                     // blocks in the ring (4992 words)
constexpr int MG_WIN = 4 * MG_THREADS;                         // words tested per step

__device__ __forceinline__ uint32_t mg_twist(uint32_t u, uint32_t v) {
    const uint32_t y = (u & 0x80000000u) | (v & 0x7FFFFFFFu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}
__device__ __forceinline__ uint32_t mg_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return y;
}
// legacy_double: (a >> 5, b >> 6) -> [0, 1) with 53 bits
__device__ __forceinline__ double mg_double(uint32_t a, uint32_t b) {
    return __dmul_rn(__dadd_rn(__dmul_rn((double)(a >> 5), 67108864.0), (double)(b >> 6)), 1.0 / 9007199254740992.0);
}

struct MgStream {
    uint32_t* raw;      // [MG_RING][624]
    uint32_t* tw;       // [MG_RING][624]
    int gen_b;          // newest generated block (block 0 = the incoming state)
};

// next block of the recurrence into ring slot (gen_b + 1) % MG_RING (separate source and destination: 3 phases, 3 barriers)
__device__ __forceinline__ void mg_regenerate(MgStream& s, int tid) {
    const uint32_t* __restrict__ src = s.raw + (s.gen_b % MG_RING) * MG_N;
    uint32_t* __restrict__ dst = s.raw + ((s.gen_b + 1) % MG_RING) * MG_N;
    uint32_t* __restrict__ twd = s.tw + ((s.gen_b + 1) % MG_RING) * MG_N;
    if (tid < MG_D) { const uint32_t y = src[tid + MG_M] ^ mg_twist(src[tid], src[tid + 1]); dst[tid] = y; twd[tid] = mg_temper(y); }
    __syncthreads();
    if (tid < MG_D) { const int i = MG_D + tid; const uint32_t y = dst[i - MG_D] ^ mg_twist(src[i], src[i + 1]); dst[i] = y; twd[i] = mg_temper(y); }
    __syncthreads();
    if (tid < MG_N - 2 * MG_D) {
        const int i = 2 * MG_D + tid;
        const uint32_t nxt = (i == MG_N - 1) ? dst[0] : src[i + 1];
        const uint32_t y = dst[i - MG_D] ^ mg_twist(src[i], nxt);
        dst[i] = y; twd[i] = mg_temper(y);
    }
    __syncthreads();
    ++s.gen_b;
}
// make words [.., upto) available (upto is uniform over the CTA)
__device__ __forceinline__ void mg_ensure(MgStream& s, long long upto, int tid) {
    while ((long long)(s.gen_b + 1) * MG_N < upto) mg_regenerate(s, tid);
}
__device__ __forceinline__ uint32_t mg_word(const MgStream& s, long long at) {
    return s.tw[(int)(at % (MG_RING * MG_N))];
}

__global__ void __launch_bounds__(MG_THREADS)
mt_gauss_kernel(uint32_t* __restrict__ mt_key, int32_t* __restrict__ mt_pos, int32_t* __restrict__ has_gauss_io,
                double* __restrict__ gauss_io, int n_pairs, uint32_t rng, uint32_t mask, int coins, int N, double scale,
                int64_t* __restrict__ idx_out, uint32_t* __restrict__ extra_out, float* __restrict__ noise_out) {
    __shared__ uint32_t s_raw[MG_RING * MG_N];
    __shared__ uint32_t s_tw[MG_RING * MG_N];
    __shared__ int s_wtot[2][MG_WARPS];
    __shared__ int s_end;
    __shared__ int s_has;
    __shared__ double s_spare;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sid = blockIdx.x;
    MgStream st{s_raw, s_tw, 0};
    for (int i = tid; i < MG_N; i += MG_THREADS) {
        const uint32_t y = mt_key[(size_t)sid * MG_N + i];
        s_raw[i] = y; s_tw[i] = mg_temper(y);
    }
    if (tid == 0) { s_has = has_gauss_io[sid]; s_spare = gauss_io[sid]; }
    __syncthreads();
    long long cur = mt_pos[sid];                                   // absolute word position (block 0 = the incoming state)
    int has = s_has;
    double spare = s_spare;
    int64_t* idx_o = idx_out + (size_t)sid * n_pairs;
    uint32_t* ext_o = extra_out ? extra_out + (size_t)sid * n_pairs * 4 * coins : nullptr;
    unsigned step = 0;

    for (int pair = 0; pair < n_pairs; ++pair) {
        // ---- rs.randint: masked rejection, every thread walks the same few words ----
        uint32_t w;
        do {
            mg_ensure(st, cur + 1, tid);
            w = mg_word(st, cur) & mask;
            ++cur;
        } while (w > rng);
        if (tid == 0) idx_o[pair] = (int64_t)w;
        for (int sgn = 0; sgn < 2; ++sgn) {
            // ---- the fit_fn's save_obs coin(s) ----
            mg_ensure(st, cur + 2 * coins, tid);
            if (ext_o && tid < 2 * coins) ext_o[(size_t)pair * 4 * coins + sgn * 2 * coins + tid] = mg_word(st, cur + tid);
            cur += 2 * coins;
            // ---- T x randn(act): N gaussians ----
            float* out = noise_out + ((size_t)((size_t)sid * n_pairs + pair) * 2 + sgn) * N;
            int o = 0;
            if (has && N > 0) {                                    // the cached second gaussian of an earlier attempt
                if (tid == 0) { out[0] = (float)(spare * scale); s_has = 0; }
                o = 1; has = 0; spare = 0.0;
            }
            while (o < N) {
                const int need = (N - o + 1) >> 1;                 // accepted attempts still to find
                mg_ensure(st, cur + MG_WIN, tid);
                int at = (int)(cur % (MG_RING * MG_N)) + 4 * tid;  // ring offsets of this thread's four words
                uint32_t wd[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int a = at + q;
                    if (a >= MG_RING * MG_N) a -= MG_RING * MG_N;
                    wd[q] = s_tw[a];
                }
                const double x1 = __dadd_rn(__dmul_rn(2.0, mg_double(wd[0], wd[1])), -1.0);
                const double x2 = __dadd_rn(__dmul_rn(2.0, mg_double(wd[2], wd[3])), -1.0);
                const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
                const bool acc = r2 < 1.0 && r2 != 0.0;
                const unsigned bal = __ballot_sync(0xffffffffu, acc);
                const unsigned buf = step & 1;
                ++step;
                if (lane == 0) s_wtot[buf][warp] = __popc(bal);
                __syncthreads();
                int before = 0, total = 0;
#pragma unroll
                for (int q = 0; q < MG_WARPS; ++q) {
                    const int c = s_wtot[buf][q];
                    before += (q < warp) ? c : 0;
                    total += c;
                }
                const int rank = before + __popc(bal & ((1u << lane) - 1u));
                if (acc && rank < need) {
                    const double f = sqrt(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
                    const int p = o + 2 * rank;
                    out[p] = (float)(__dmul_rn(__dmul_rn(f, x2), scale));
                    const double g2 = __dmul_rn(f, x1);
                    if (p + 1 < N) out[p + 1] = (float)(__dmul_rn(g2, scale));
                    else { s_spare = g2; s_has = 1; }              // the last attempt's second value stays cached
                    if (rank == need - 1) s_end = tid + 1;
                }
                if (total >= need) {
                    __syncthreads();                               // s_end (and s_has / s_spare) are written
                    cur += 4 * s_end;
                    o = N;
                } else {
                    cur += MG_WIN;
                    o += 2 * total;
                }
            }
            // the cache after this rollout (s_has: cleared by the consumer above, set by the thread that produced an (N+1)-th value)
            __syncthreads();
            has = s_has; spare = s_spare;
            __syncthreads();
        }
    }
    // ---- hand the state back: the block the cursor is in, and its position ----
    const int b_last = (int)(cur / MG_N) - ((cur % MG_N == 0 && cur > 0) ? 1 : 0);      // position 624 = block exhausted
    mg_ensure(st, (long long)(b_last + 1) * MG_N, tid);
    const uint32_t* fin = s_raw + (b_last % MG_RING) * MG_N;
    for (int i = tid; i < MG_N; i += MG_THREADS) mt_key[(size_t)sid * MG_N + i] = fin[i];
    if (tid == 0) {
        mt_pos[sid] = (int32_t)(cur - (long long)b_last * MG_N);
        has_gauss_io[sid] = has;
        gauss_io[sid] = has ? spare : 0.0;
    }
}

}  // namespace

int es_impl_draw_noisy(es_ctx* ctx, uint32_t* mt_key, int32_t* mt_pos, int32_t* has_gauss, double* gauss, int n_streams,
                       int n_per_stream, uint64_t upper_bound, int coins, int normals_per_eval, double scale, int64_t* idx_out,
                       uint32_t* extra_out, float* noise_out, cudaStream_t stream) {
    const uint32_t rng = (uint32_t)(upper_bound - 1);
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    if (rng == 0) {
        es_set_error("es_draw_noisy: upper_bound == 1 is not supported");
        return ES_ERR_UNSUPPORTED;
    }
    mt_gauss_kernel<<<n_streams, MG_THREADS, 0, stream>>>(mt_key, mt_pos, has_gauss, gauss, n_per_stream, rng, mask, coins,
                                                          normals_per_eval, scale, idx_out, extra_out, noise_out);
    ES_LAUNCHED(ctx);
    return ES_OK;
}
