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
// which is the parallelism used here: the CTA of a stream (736 threads) tests 1 472 attempts per step, ranks the accepted ones
// with a ballot scan and records their four words in order; mt_gauss_finish_kernel turns the records into gaussians on the
// whole GPU (float64 log / sqrt off the sequential path).
//
// One CTA per stream.  The state lives in a ring of MG_RING 624-word blocks in shared memory (raw words for the recurrence and
// for the state handed back, tempered words for the attempt windows), regenerated ahead of the cursor by the whole CTA with two
// barriers per block (see `ensure`).  What the time goes to, measured with cycle counters (MG_PROFILE) and ncu on the way from
// 109 ms to 72 ms per K = 10 000, T*act = 17 000 generation with 8 streams: the regeneration (69 blocks per rollout) is the
// limiter and it is instruction-issue bound (issue slots 61 %): the textbook three dependent 227-word phases cost ~1 100 cycles
// per block with 4 or 10 warps; ONE pass that recomputes up to three twists per word is as slow (~100 instructions per word);
// sharing the 623 twists through shared memory, warp-aligned ranges, branch-free word expressions and hoisted index arithmetic
// bring a block to ~450 cycles.  Splitting the CTA into producer and consumer warps did not help (the producer alone sets the
// pace), neither did moving the float64 math out (it was not the limiter).  A polynomial jump-ahead that spreads one stream's
// regeneration over many SMs is the known way beyond this; not built.
//
// Output noise[((stream * n_pairs + pair) * 2 + sign) * N + t * act + j] = float32(gauss * scale): the float64 product rounded
// once; the rollout kernels add it to the float32 action (the reference forms the float64 sum and the env rounds it to
// float32: identical except when the float32 rounding of the noise term moves the sum across a rounding boundary, 2^-24
// relative on a term that is ~1e-2 of the action).
// log() here is CUDA's (<= 1 ulp), the reference's is glibc's: a gaussian can differ in its last float64 bit, which the
// float32 noise array does not see; the accept/reject arithmetic (the only part that steers the stream) is exact.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "mt19937.cuh"

namespace {

constexpr int MG_N = 624, MG_M = 397, MG_D = MG_N - MG_M;      // 227
constexpr int MG_THREADS = 736, MG_WARPS = MG_THREADS / 32;    // 23 warps: see the word <-> thread map of the regeneration
static_assert(MG_WARPS <= 32, "one shuffle scan over the warp totals");
constexpr int MG_RING = 20;                                    // blocks in the ring (raw + tempered + one block of twists: 102 KB of shared memory)
constexpr int MG_RW = MG_RING * MG_N, MG_RQ = MG_RW / 4;       // ring words; the tempered ring is stored as 4 quarter rings (see mg_tw_slot)
static_assert(MG_RW % 4 == 0, "quarter rings");
constexpr int MG_APT = 2;                                      // attempts per thread and step (consecutive attempts)
constexpr int MG_WIN = 4 * MG_APT * MG_THREADS;                // words tested per step (5888 = 9.4 blocks)
static_assert((MG_RING - 2) * MG_N >= MG_WIN + MG_N, "the ring holds the cursor's block and a whole window ahead of it");

__device__ __forceinline__ uint32_t mg_twist(uint32_t u, uint32_t v) { return mt19937_twist(u, v); }
__device__ __forceinline__ uint32_t mg_temper(uint32_t y) { return mt19937_temper(y); }
// Where ring word k (0 <= k < MG_RW) of the TEMPERED stream lives: word k sits in quarter ring k % 4 at k / 4, so that the four
// words of consecutive attempts (k = s + 4a + q) are consecutive in shared memory for a fixed q: the attempt windows read
// without bank conflicts (stride-4 or stride-8 word reads were 4- / 8-way conflicts and half of a window step's time)
__device__ __forceinline__ int mg_tw_slot(int k) { return (k & 3) * MG_RQ + (k >> 2); }
// legacy_double: (a >> 5, b >> 6) -> [0, 1) with 53 bits
__device__ __forceinline__ double mg_double(uint32_t a, uint32_t b) {
    return __dmul_rn(__dadd_rn(__dmul_rn((double)(a >> 5), 67108864.0), (double)(b >> 6)), 1.0 / 9007199254740992.0);
}
struct MgShared {
    int wtot[2][MG_WARPS];
    int end;
};

// is a gaussian cached at the start of evaluation e of a stream?  Every evaluation draws N values; an accepted attempt makes
// two, so the cache state only depends on the parity of N and of e (c0 = the stream's incoming has_gauss)
__device__ __forceinline__ int mg_cached(int c0, int N, int e) { return (N & 1) ? (c0 ^ (e & 1)) : c0; }

__global__ void __launch_bounds__(MG_THREADS, 1)
mt_gauss_kernel(uint32_t* __restrict__ mt_key, int32_t* __restrict__ mt_pos, int32_t* __restrict__ has_gauss_io,
                const double* __restrict__ gauss_io, int n_pairs, uint32_t rng, uint32_t mask, int coins, int N,
                int64_t* __restrict__ idx_out, uint32_t* __restrict__ extra_out, uint4* __restrict__ acc4, int a_max,
                int32_t* __restrict__ c0_out, double* __restrict__ gauss0_out) {
    extern __shared__ uint32_t mg_smem[];
    uint32_t* s_raw = mg_smem;                                 // [MG_RING][624] raw state words (the recurrence, the state handed back)
    uint32_t* s_tw = mg_smem + MG_RING * MG_N;                 // [MG_RING][624] tempered words (what the consumers read)
    uint32_t* s_T = mg_smem + 2 * MG_RING * MG_N;              // [624] twists of the block being regenerated
    __shared__ MgShared sh;

    const int tid = threadIdx.x, lane = tid & 31;
    const int sid = blockIdx.x;
    for (int i = tid; i < MG_N; i += MG_THREADS) {
        const uint32_t y = mt_key[(size_t)sid * MG_N + i];
        s_raw[i] = y; s_tw[mg_tw_slot(i)] = mg_temper(y);
    }
    const long long cur0 = mt_pos[sid];
    const int c0 = has_gauss_io[sid] ? 1 : 0;
    if (tid == 0) { c0_out[sid] = c0; gauss0_out[sid] = gauss_io[sid]; }     // what the finishing kernel needs of the incoming cache
    __syncthreads();

    // the cursor as (block, offset in the block, offset in the ring): 32-bit bookkeeping, no 64-bit divisions per step
    int gen_b = 0;                                             // newest generated block (block 0 = the incoming state)
#ifdef MG_PROFILE
    long long pr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = clock64();
#endif
    int cblk = (int)(cur0 / MG_N), coff = (int)(cur0 % MG_N), cring = (int)(cur0 % (MG_RING * MG_N));
    auto advance = [&](int n) {
        coff += n;
        while (coff >= MG_N) { coff -= MG_N; ++cblk; }
        cring += n;
        if (cring >= MG_RING * MG_N) cring -= MG_RING * MG_N;
    };
    // Next block of the recurrence into ring slot (b + 1) % MG_RING by the whole CTA.  The recurrence
    // N[i] = N[i-227] ^ tw(O[i], O[i+1]) is XOR-linear in its first term, so with the 623 twists T[k] = tw(O[k], O[k+1]) of the
    // OLD block (one per thread, exchanged through shared memory) every new word is a few XORs of old words and twists -- two
    // barriers per block instead of the three dependent 227-word phases of the textbook form:
    //   i <  227:  N[i] = O[i+397] ^ T[i]
    //   i <  454:  N[i] = O[i+170] ^ T[i-227] ^ T[i]
    //   i <  623:  N[i] = O[i-57]  ^ T[i-454] ^ T[i-227] ^ T[i]
    //   N[623] = N[396] ^ tw(O[623], N[0])
    // (recomputing the twists instead of sharing them -- one barrier, ~100 instructions per word -- measured slower.)
    // word of this thread in the regeneration: every warp stays inside one of the three ranges and all three ranges run the SAME
    // code (an unused twist index points at s_T[623], which holds 0); the odd word out (N[623], the longest expression) has a
    // warp of its own: threads 0-226 | 256-482 | 512-680 | 704.  Everything that does not depend on the block is computed once.
    const int my_i = (tid < 256) ? (tid < MG_D ? tid : -1) : (tid < 512) ? (tid - 256 < MG_D ? tid - 256 + MG_D : -1)
                   : (tid < 704) ? (tid - 512 < MG_N - 1 - 2 * MG_D ? tid - 512 + 2 * MG_D : -1) : (tid == 704 ? MG_N - 1 : -1);
    const bool plain = my_i >= 0 && my_i < MG_N - 1;
    const int r_o = !plain ? 0 : (my_i < MG_D ? my_i + MG_M : my_i < 2 * MG_D ? my_i + MG_M - MG_D : my_i + MG_M - 2 * MG_D);
    const int r_a = plain ? my_i : MG_N - 1;
    const int r_b = (plain && my_i >= MG_D) ? my_i - MG_D : MG_N - 1;
    const int r_c = (plain && my_i >= 2 * MG_D) ? my_i - 2 * MG_D : MG_N - 1;
    const int r_tw = my_i >= 0 ? (my_i & 3) * MG_RQ + (my_i >> 2) : 0;       // mg_tw_slot(nslot * 624 + i) = r_tw + nslot * 156
    if (tid == 0) s_T[MG_N - 1] = 0;
    auto ensure = [&](int n) {                                 // the next n words are in the ring (n uniform over the CTA)
        const int blocks = cblk + (coff + n + MG_N - 1) / MG_N;               // blocks [0, blocks) are needed
        while (gen_b + 1 < blocks) {
            const uint32_t* __restrict__ O = s_raw + (gen_b % MG_RING) * MG_N;
            const int nslot = (gen_b + 1) % MG_RING;
            uint32_t* __restrict__ dst = s_raw + nslot * MG_N;
            if (tid < MG_N - 1) s_T[tid] = mg_twist(O[tid], O[tid + 1]);
            __syncthreads();
            if (plain) {
                const uint32_t y = O[r_o] ^ s_T[r_a] ^ s_T[r_b] ^ s_T[r_c];
                dst[my_i] = y; s_tw[r_tw + nslot * (MG_N / 4)] = mg_temper(y);
            } else if (my_i == MG_N - 1) {
                const uint32_t n0 = O[MG_M] ^ s_T[0];
                const uint32_t n396 = O[396 + MG_M - MG_D] ^ s_T[396 - MG_D] ^ s_T[396];
                const uint32_t y = n396 ^ mg_twist(O[MG_N - 1], n0);
                dst[my_i] = y; s_tw[r_tw + nslot * (MG_N / 4)] = mg_temper(y);
            }
            __syncthreads();
            ++gen_b;
        }
    };
    auto word = [&](int ahead) -> uint32_t {                   // the word `ahead` positions after the cursor
        int a = cring + ahead;
        if (a >= MG_RW) a -= MG_RW;
        return s_tw[mg_tw_slot(a)];
    };
    const int warp = tid >> 5;
    int64_t* idx_o = idx_out + (size_t)sid * n_pairs;
    uint32_t* ext_o = extra_out ? extra_out + (size_t)sid * n_pairs * 4 * coins : nullptr;
    unsigned step = 0;

    for (int pair = 0; pair < n_pairs; ++pair) {
        // ---- rs.randint: masked rejection, every thread walks the same few words ----
        uint32_t w;
        do {
            ensure(1);
            w = word(0) & mask;
            advance(1);
        } while (w > rng);
        if (tid == 0) idx_o[pair] = (int64_t)w;
        for (int sgn = 0; sgn < 2; ++sgn) {
            // ---- the fit_fn's save_obs coin(s) ----
            ensure(2 * coins);
            if (ext_o && tid < 2 * coins) ext_o[(size_t)pair * 4 * coins + sgn * 2 * coins + tid] = word(tid);
            advance(2 * coins);
            // ---- T x randn(act): N gaussians = the cached one (if any) + accepted attempts, two values each.  Only the accept /
            //      reject arithmetic steers the stream: the accepted attempts' words are recorded in order and turned into
            //      gaussians by mt_gauss_finish_kernel on the whole GPU ----
            const int e = pair * 2 + sgn;
            uint4* rec = acc4 + ((size_t)sid * 2 * n_pairs + e) * a_max;
            int need = (N - mg_cached(c0, N, e) + 1) >> 1;            // accepted attempts still to find
            int found = 0;
            while (need > 0) {
                // only as many words as the remaining attempts can use at the usual acceptance rate are regenerated up front
                ensure(MG_WIN);
                // this thread's MG_APT consecutive attempts (4 words each).  x = 2 u - 1 with u = (a >> 5, b >> 6) / 2^53: the
                // 53-bit integer converts exactly, and one fused multiply-add rounds the exact value of 2 u - 1 once -- the
                // reference's (2.0 * u) - 1.0 rounds the same exact value once (2 u is exact)
                uint32_t wd[MG_APT][4];
                bool acc[MG_APT];
                unsigned bal[MG_APT];
                const int ph = cring & 3;                      // the cursor's phase: word q of every attempt sits in quarter ring (ph + q) & 3
#pragma unroll
                for (int j = 0; j < MG_APT; ++j) {
                    const int b4 = (cring >> 2) + MG_APT * tid + j;            // (ring offset of the attempt's first word) / 4, before wrapping
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int k4 = b4 + ((ph + q) >> 2);
                        if (k4 >= MG_RQ) k4 -= MG_RQ;
                        wd[j][q] = s_tw[((ph + q) & 3) * MG_RQ + k4];
                    }
                    // u = (a >> 5, b >> 6) as a 53-bit integer, exactly, in float64; x = 2 u / 2^53 - 1: one fused multiply-add rounds
                    // the exact value once, like the reference's (2.0 * u) - 1.0 (2 u is exact)
                    const double v1 = fma((double)(wd[j][0] >> 5), 67108864.0, (double)(wd[j][1] >> 6));
                    const double v2 = fma((double)(wd[j][2] >> 5), 67108864.0, (double)(wd[j][3] >> 6));
                    const double x1 = fma(v1, 1.0 / 4503599627370496.0, -1.0), x2 = fma(v2, 1.0 / 4503599627370496.0, -1.0);
                    const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
                    acc[j] = r2 < 1.0 && r2 != 0.0;
                    bal[j] = __ballot_sync(0xffffffffu, acc[j]);
                }
                const unsigned buf = step & 1;
                ++step;
                int wsum = 0, below = 0;                       // accepted in this warp; accepted by lower lanes
#pragma unroll
                for (int j = 0; j < MG_APT; ++j) { wsum += __popc(bal[j]); below += __popc(bal[j] & ((1u << lane) - 1u)); }
                if (lane == 0) sh.wtot[buf][warp] = wsum;
                __syncthreads();
                int scan = (lane < MG_WARPS) ? sh.wtot[buf][lane] : 0;      // inclusive scan of the warp totals
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, scan, d);
                    if (lane >= d) scan += up;
                }
                const int total = __shfl_sync(0xffffffffu, scan, MG_WARPS - 1);
                const int before = warp ? __shfl_sync(0xffffffffu, scan, warp - 1) : 0;
                int rank = before + below;                     // rank of this thread's first attempt among the accepted ones
#pragma unroll
                for (int j = 0; j < MG_APT; ++j) {
                    if (acc[j] && rank < need) {
                        rec[found + rank] = make_uint4(wd[j][0], wd[j][1], wd[j][2], wd[j][3]);
                        if (rank == need - 1) sh.end = MG_APT * tid + j + 1;
                    }
                    rank += acc[j] ? 1 : 0;
                }
                if (total >= need) {
                    __syncthreads();                           // sh.end is written
                    advance(4 * sh.end);
                    need = 0;
                    __syncthreads();                           // ... and read by everyone before the next rollout can write it
                } else {
                    advance(MG_WIN);
                    need -= total;
                    found += total;
                }
            }
        }
    }
#ifdef MG_PROFILE
    if (sid == 0 && (tid == 0 || tid == 300 || tid == 623))
        printf("tid %d: %lld blocks: twist %lld, bar1 %lld, xor+temper %lld, bar2 %lld cycles per block; total kernel %lld cycles, %u steps\n", tid, pr[4],
               pr[0] / max(pr[4], 1LL), pr[1] / max(pr[4], 1LL), pr[2] / max(pr[4], 1LL), pr[3] / max(pr[4], 1LL), clock64() - t_start, step);
#endif
    // ---- hand the state back: the block the cursor is in (position 624 = block exhausted), and its position ----
    const bool at_end = coff == 0 && (cblk > 0);
    const int b_last = at_end ? cblk - 1 : cblk;
    ensure(0);                                                 // (block b_last is in the ring: the cursor has been there)
    const uint32_t* last = s_raw + (b_last % MG_RING) * MG_N;
    for (int i = tid; i < MG_N; i += MG_THREADS) mt_key[(size_t)sid * MG_N + i] = last[i];
    if (tid == 0) {
        mt_pos[sid] = at_end ? MG_N : coff;
        has_gauss_io[sid] = mg_cached(c0, N, 2 * n_pairs);     // (the cached VALUE is written by the finishing kernel)
    }
}

// The recorded attempts -> gaussians, on the whole GPU: attempt r of evaluation e gives values c_e + 2r (f*x2) and c_e + 2r + 1
// (f*x1) of the evaluation's N; a value N (the second of the last attempt) is the cache: the first value of the stream's next
// evaluation, or the cached gaussian the stream hands back.
__global__ void __launch_bounds__(256)
mt_gauss_finish_kernel(const uint4* __restrict__ acc4, int a_max, const int32_t* __restrict__ c0_in, const double* __restrict__ gauss0,
                       int n_streams, int n_pairs, int N, double scale, float* __restrict__ noise_out, double* __restrict__ gauss_io) {
    const long long per_eval = a_max;
    const long long total = (long long)n_streams * 2 * n_pairs * per_eval;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i % per_eval);
        const long long ge = i / per_eval;                     // evaluation, stream-major
        const int e = (int)(ge % (2 * n_pairs)), sid = (int)(ge / (2 * n_pairs));
        const int c0 = c0_in[sid];
        const int ce = mg_cached(c0, N, e);
        const int need = (N - ce + 1) >> 1;
        float* out = noise_out + (size_t)ge * N;
        if (r == 0) {
            if (e == 0 && ce && N > 0) out[0] = (float)__dmul_rn(gauss0[sid], scale);     // the stream's incoming cached gaussian
            if (e == 2 * n_pairs - 1 && !mg_cached(c0, N, 2 * n_pairs)) gauss_io[sid] = 0.0;   // nothing cached afterwards
            if (N == 0 && e == 0 && c0) gauss_io[sid] = gauss0[sid];                       // no draws at all: the cache is untouched
        }
        if (r >= need) continue;
        const uint4 w = acc4[i];
        const double x1 = __dadd_rn(__dmul_rn(2.0, mg_double(w.x, w.y)), -1.0);
        const double x2 = __dadd_rn(__dmul_rn(2.0, mg_double(w.z, w.w)), -1.0);
        const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
        const double f = sqrt(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
        const int p = ce + 2 * r;
        out[p] = (float)__dmul_rn(__dmul_rn(f, x2), scale);
        const double g2 = __dmul_rn(f, x1);
        if (p + 1 < N) out[p + 1] = (float)__dmul_rn(g2, scale);
        else if (e + 1 < 2 * n_pairs) out[N] = (float)__dmul_rn(g2, scale);              // = value 0 of the stream's next evaluation
        else gauss_io[sid] = g2;                                                          // the cache the stream hands back
    }
}

// =====================================================================================================================
// Jump-ahead: ONE stream regenerated by many CTAs.
//
// The state transition of MT19937 is linear over GF(2); with g_r(x) = x^(624 * 2^r) mod phi(x) (phi = its characteristic
// polynomial, tools/mt_jump/make_jump_polys.py) every word of the raw sequence obeys x[n + 624 * 2^r] = XOR_{i : g_r,i = 1}
// x[n + i].  So the state 2^r blocks ahead is, word by word, an XOR of ~10 000 words out of the next 20 560: independent per
// word, no recurrence to follow.  mt_fill_kernel gives every CTA one segment of 2^lb blocks of one stream: it jumps from the
// stream's current state to its segment's first block (one jump per set bit of the block index), regenerates the segment and
// writes the TEMPERED words to global memory.  What follows no longer touches the recurrence: mt_flags_kernel computes the accept
// bit of every possible attempt (all four phases) with per-chunk counts, mt_scan_kernel their prefixes; mt_walk_kernel -- the only
// sequential step left, one warp per stream -- follows the reference's order with two table look-ups per rollout ("the next
// `need` accepted attempts of this phase end at word ..."); mt_emit_kernel turns every rollout's accepted attempts into its
// gaussians, one CTA per rollout.
// =====================================================================================================================
constexpr int MJ_NPOLY = 18;
constexpr int MJ_WIN_BLOCKS = 33;                              // the state block + 32 more: 20 592 words >= 19 937 + 624
// Layout in global memory (per stream): the tempered words in stream order; per phase ph = 0..3 (an attempt starts at a word
// index = ph mod 4) one accept bit per attempt (attempt a of phase ph = words 4 a + ph .. 4 a + ph + 3), in chunks of
// MJ_CA = 1024 attempts (32 mask words) with the number of accepted attempts of every chunk and its exclusive prefix.
constexpr int MJ_CA = 1024, MJ_CW = 4 * MJ_CA;                  // attempts / words per chunk
__device__ const uint32_t mj_polys[MJ_NPOLY][MT_NW] = {
#include "mt_jump_polys.inc"
};

__global__ void __launch_bounds__(MG_THREADS, 1)
mt_fill_kernel(const uint32_t* __restrict__ mt_key, int n_seg, int lb_log2, uint32_t* __restrict__ words, size_t stride_words) {
    extern __shared__ uint32_t mj_smem[];
    uint32_t* xs = mj_smem;                                    // [33][624] raw words: the jump window; blocks 0 / 1: ping-pong of the fill
    uint32_t* s_T = mj_smem + MJ_WIN_BLOCKS * MT_NW;           // [624] twists
    uint32_t* s_g = s_T + MT_NW;                               // [624] the polynomial of the current jump
    const int tid = threadIdx.x;
    const int sid = blockIdx.x / n_seg, k = blockIdx.x % n_seg;
    Mt19937Regen rg;
    rg.init(tid);
    uint32_t* __restrict__ out = words + (size_t)sid * stride_words;
    for (int i = tid; i < MT_NW; i += MG_THREADS) {
        const uint32_t y = mt_key[(size_t)sid * MT_NW + i];
        xs[i] = y;
        if (k == 0) out[i] = mt19937_temper(y);               // block 0 = the incoming state: its unread words belong to the stream
    }
    if (tid == 0) s_T[MT_NW - 1] = 0;
    __syncthreads();
    // (blk >= 0: the block's index in the stream: its tempered words go to global memory)
    auto regen = [&](const uint32_t* __restrict__ O, uint32_t* __restrict__ dst, long long blk) {
        if (tid < MT_NW - 1) s_T[tid] = mt19937_twist(O[tid], O[tid + 1]);
        __syncthreads();
        if (rg.my_i >= 0) {
            const uint32_t y = rg.word(O, s_T);
            dst[rg.my_i] = y;
            if (blk >= 0) out[(size_t)blk * MT_NW + rg.my_i] = mt19937_temper(y);
        }
        __syncthreads();
    };
    // ---- jump to block k << lb_log2: one jump per set bit ----
    const unsigned target = (unsigned)k << lb_log2;
    for (int r = MJ_NPOLY - 1; r >= 0; --r) {
        if (!((target >> r) & 1u)) continue;
        for (int b = 1; b < MJ_WIN_BLOCKS; ++b) regen(xs + (b - 1) * MT_NW, xs + b * MT_NW, -1);
        for (int i = tid; i < MT_NW; i += MG_THREADS) s_g[i] = mj_polys[r][i];
        __syncthreads();
        uint32_t acc = 0;
        if (tid < MT_NW) {
            for (int w = 0; w < MT_NW; ++w) {
                uint32_t gw = s_g[w];
                const uint32_t* __restrict__ xw = xs + tid + 32 * w;
                while (gw) {
                    const int bit = __ffs(gw) - 1;
                    acc ^= xw[bit];
                    gw &= gw - 1;
                }
            }
        }
        __syncthreads();
        if (tid < MT_NW) xs[tid] = acc;                        // (word 0: only its top bit is state, and that bit is right)
        __syncthreads();
    }
    // ---- regenerate the segment: blocks target + 1 .. target + 2^lb ----
    const int Lb = 1 << lb_log2;
    int cur = 0;
    for (int b = 0; b < Lb; ++b) {
        regen(xs + cur * MT_NW, xs + (cur ^ 1) * MT_NW, (long long)1 + target + b);
        cur ^= 1;
    }
}

// ---- the accept bit of every possible attempt, in parallel: one CTA per (stream, chunk of 4096 words), thread q looks at the
//      four attempts that start at words 4 q .. 4 q + 3 of the chunk (one per phase) ----
__device__ __forceinline__ bool mj_accept(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    // x = 2 u - 1 with u = (a >> 5, b >> 6) / 2^53: the 53-bit integer is exact in float64 and one fused multiply-add rounds
    // the exact value of 2 u - 1 once, like the reference's (2.0 * u) - 1.0 (2 u is exact); r2 with two roundings as in C
    const double v1 = fma((double)(w0 >> 5), 67108864.0, (double)(w1 >> 6));
    const double v2 = fma((double)(w2 >> 5), 67108864.0, (double)(w3 >> 6));
    const double x1 = fma(v1, 1.0 / 4503599627370496.0, -1.0), x2 = fma(v2, 1.0 / 4503599627370496.0, -1.0);
    const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
    return r2 < 1.0 && r2 != 0.0;
}

__global__ void __launch_bounds__(MJ_CA)
mt_flags_kernel(const uint32_t* __restrict__ words, size_t stride_words, int n_chunks, uint32_t* __restrict__ masks,
                uint32_t* __restrict__ counts) {
    __shared__ int s_cnt[4];
    const int q = threadIdx.x, c = blockIdx.x, sid = blockIdx.y;
    if (q < 4) s_cnt[q] = 0;
    __syncthreads();
    const uint4* __restrict__ w4 = reinterpret_cast<const uint4*>(words + (size_t)sid * stride_words + (size_t)c * MJ_CW) + q;
    const uint4 a = __ldg(w4), b = __ldg(w4 + 1);                 // words 4 q .. 4 q + 7 (the buffer is padded past the last chunk)
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t* mk = masks + ((size_t)sid * 4 * n_chunks + c) * 32 + (q >> 5);          // + ph * n_chunks * 32
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        const unsigned bal = __ballot_sync(0xffffffffu, mj_accept(w[ph], w[ph + 1], w[ph + 2], w[ph + 3]));
        if ((q & 31) == 0) {
            mk[(size_t)ph * n_chunks * 32] = bal;
            atomicAdd(&s_cnt[ph], __popc(bal));
        }
    }
    __syncthreads();
    if (q < 4) counts[((size_t)sid * 4 + q) * n_chunks + c] = (uint32_t)s_cnt[q];
}

// exclusive prefix of the chunk counts of every (stream, phase): one CTA each
__global__ void __launch_bounds__(1024) mt_scan_kernel(uint32_t* __restrict__ counts, int n_chunks) {
    __shared__ uint32_t s_part[1024];
    uint32_t* v = counts + (size_t)blockIdx.x * n_chunks;
    const int per = (n_chunks + 1023) / 1024, lo = threadIdx.x * per, hi = min(n_chunks, lo + per);
    uint32_t sum = 0;
    for (int i = lo; i < hi; ++i) sum += v[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {                          // inclusive scan of the partial sums
        const uint32_t add = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
        __syncthreads();
        s_part[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? s_part[threadIdx.x - 1] : 0;
    for (int i = lo; i < hi; ++i) { const uint32_t t = v[i]; v[i] = run; run += t; }
}

// ---- the walk: ONE WARP per stream follows the reference's order; a rollout's N gaussians are "the next `need` accepted
//      attempts of phase ph from word s on", i.e. two look-ups in the prefix tables instead of a pass over the words ----
__global__ void __launch_bounds__(32)
mt_walk_kernel(uint32_t* __restrict__ mt_key, int32_t* __restrict__ mt_pos, int32_t* __restrict__ has_gauss_io,
               const double* __restrict__ gauss_io, int n_pairs, uint32_t rng, uint32_t mask, int coins, int N,
               int64_t* __restrict__ idx_out, uint32_t* __restrict__ extra_out, uint32_t* __restrict__ reg_start,
               int32_t* __restrict__ c0_out, double* __restrict__ gauss0_out, const uint32_t* __restrict__ words,
               size_t stride_words, int n_chunks, const uint32_t* __restrict__ masks, const uint32_t* __restrict__ cumul,
               uint32_t limit, int* __restrict__ err) {
    const int lane = threadIdx.x, sid = blockIdx.x;
    const uint32_t* __restrict__ gw = words + (size_t)sid * stride_words;
    const uint32_t* __restrict__ mk = masks + (size_t)sid * 4 * n_chunks * 32;
    const uint32_t* __restrict__ cu = cumul + (size_t)sid * 4 * n_chunks;
    uint32_t cpos = (uint32_t)mt_pos[sid];
    const int c0 = has_gauss_io[sid] ? 1 : 0;
    if (lane == 0) { c0_out[sid] = c0; gauss0_out[sid] = gauss_io[sid]; }
    int64_t* idx_o = idx_out + (size_t)sid * n_pairs;
    uint32_t* ext_o = extra_out ? extra_out + (size_t)sid * n_pairs * 4 * coins : nullptr;
    uint32_t* rs = reg_start + (size_t)sid * 2 * n_pairs;
    bool overflow = false;
    for (int pair = 0; pair < n_pairs && !overflow; ++pair) {
        uint32_t w;
        do {
            if (cpos + 1 > limit) { overflow = true; break; }
            w = __ldg(gw + cpos) & mask;
            ++cpos;
        } while (w > rng);
        if (overflow) break;
        if (lane == 0) idx_o[pair] = (int64_t)w;
        for (int sgn = 0; sgn < 2; ++sgn) {
            if (cpos + 2 * coins + 8 > limit) { overflow = true; break; }
            if (ext_o && lane < 2 * coins) ext_o[(size_t)pair * 4 * coins + sgn * 2 * coins + lane] = __ldg(gw + cpos + lane);
            cpos += 2 * coins;
            const int e = pair * 2 + sgn;
            if (lane == 0) rs[e] = cpos;
            const int need = (N - mg_cached(c0, N, e) + 1) >> 1;
            if (need == 0) continue;
            const int ph = cpos & 3;
            const uint32_t a0 = cpos >> 2;                        // first attempt of the rollout, in phase ph's numbering
            const uint32_t* __restrict__ mph = mk + (size_t)ph * n_chunks * 32;
            const uint32_t* __restrict__ cph = cu + (size_t)ph * n_chunks;
            // accepted attempts of the phase before a0
            const uint32_t ch0 = a0 >> 10, r0 = a0 & 1023;
            const uint32_t m0 = __ldg(mph + (size_t)ch0 * 32 + lane);
            const uint32_t base0 = __ldg(cph + ch0);
            int below = (lane < (int)(r0 >> 5)) ? __popc(m0) : (lane == (int)(r0 >> 5) ? __popc(m0 & ((1u << (r0 & 31)) - 1u)) : 0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
            const uint32_t target = base0 + (uint32_t)below + (uint32_t)need;      // accepted attempts before the END of the rollout
            // the chunk that contains the target-th accepted attempt: the last one whose prefix is < target, looked for in a
            // window of 32 chunks around the expected place (acceptance pi / 4; 32 chunks = 32 768 attempts >> its spread)
            long long est = ((long long)a0 + (long long)(need * 1.2732395447351628)) >> 10;
            long long wlo = est - 12;
            if (wlo < (long long)ch0) wlo = ch0;
            if (wlo + 32 > n_chunks) wlo = (long long)n_chunks - 32 > 0 ? n_chunks - 32 : 0;
            long long ci = wlo + lane;
            uint32_t pre = (ci < n_chunks) ? __ldg(cph + ci) : 0xFFFFFFFFu;
            unsigned lt = __ballot_sync(0xffffffffu, pre < target);
            if (lt == 0u || (lt == 0xFFFFFFFFu && wlo + 32 < n_chunks)) {
                // outside the window (tens of sigma away from the estimate): bisect the prefixes, then look again from there
                long long lo = ch0, hi = (long long)n_chunks - 1;          // invariant: prefix[lo] < target
                while (lo < hi) {
                    const long long mid = (lo + hi + 1) >> 1;
                    if (__ldg(cph + mid) < target) lo = mid; else hi = mid - 1;
                }
                wlo = lo;
                ci = wlo + lane;
                pre = (ci < n_chunks) ? __ldg(cph + ci) : 0xFFFFFFFFu;
                lt = __ballot_sync(0xffffffffu, pre < target);
            }
            const int sel = 31 - __clz((int)lt);                  // last lane with prefix < target (prefixes are non-decreasing)
            const uint32_t c1 = (uint32_t)(wlo + sel), pre1 = __shfl_sync(0xffffffffu, pre, sel);
            const uint32_t m1 = __ldg(mph + (size_t)c1 * 32 + lane);
            int inc = __popc(m1);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += up;
            }
            const uint32_t want = target - pre1;                  // the want-th (1-based) accepted attempt of chunk c1 ends the rollout
            const unsigned ge = __ballot_sync(0xffffffffu, (uint32_t)inc >= want);
            if (ge == 0u) { overflow = true; break; }
            const int L = __ffs((int)ge) - 1;
            const uint32_t mL = __shfl_sync(0xffffffffu, m1, L);
            const int incL = __shfl_sync(0xffffffffu, inc, L);
            const int kth = (int)want - (incL - __popc(mL));      // 1-based among the set bits of lane L's word
            const int bit = __fns(mL, 0, kth);
            const uint32_t a1 = (c1 << 10) + 32u * (uint32_t)L + (uint32_t)bit;
            cpos = 4u * (a1 + 1u) + (uint32_t)ph;
            if (cpos > limit) { overflow = true; break; }
        }
    }
    if (overflow) {
        if (lane == 0 && err) *(volatile int*)err = ES_ASYNC_RNG_OVERFLOW;
        return;
    }
    // ---- hand the state back: the raw words behind the tempered words of the cursor's block ----
    const uint32_t blk = cpos / MT_NW, off = cpos % MT_NW;
    const bool at_end = off == 0 && blk > 0;
    const uint32_t b_last = at_end ? blk - 1 : blk;
    for (int i = lane; i < MT_NW; i += 32) mt_key[(size_t)sid * MT_NW + i] = mt19937_untemper(__ldg(gw + (size_t)b_last * MT_NW + i));
    if (lane == 0) {
        mt_pos[sid] = at_end ? MT_NW : (int32_t)off;
        has_gauss_io[sid] = mg_cached(c0, N, 2 * n_pairs);
    }
}

// ---- the gaussians of one rollout per CTA: attempts from the rollout's first word on, accepted ones ranked by a block scan ----
__global__ void __launch_bounds__(1024)
mt_emit_kernel(const uint32_t* __restrict__ words, size_t stride_words, const uint32_t* __restrict__ reg_start,
               const int32_t* __restrict__ c0_in, const double* __restrict__ gauss0, int n_pairs, int N, double scale,
               float* __restrict__ noise_out, double* __restrict__ gauss_io) {
    __shared__ int s_wt[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ge = blockIdx.x;                                    // evaluation, stream-major
    const int e = ge % (2 * n_pairs), sid = ge / (2 * n_pairs);
    const int c0 = c0_in[sid], ce = mg_cached(c0, N, e);
    const int need = (N - ce + 1) >> 1;
    float* out = noise_out + (size_t)ge * N;
    if (tid == 0) {
        if (e == 0 && ce && N > 0) out[0] = (float)__dmul_rn(gauss0[sid], scale);          // the stream's incoming cached gaussian
        if (e == 2 * n_pairs - 1 && !mg_cached(c0, N, 2 * n_pairs)) gauss_io[sid] = 0.0;    // nothing cached afterwards
        if (N == 0 && e == 0 && c0) gauss_io[sid] = gauss0[sid];
    }
    const uint32_t* __restrict__ gw = words + (size_t)sid * stride_words + reg_start[ge];
    int found = 0;
    for (unsigned it = 0; found < need; ++it) {
        const uint32_t* __restrict__ wp = gw + (size_t)4 * (it * 1024u + tid);
        const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
        const bool acc = mj_accept(w0, w1, w2, w3);
        const unsigned bal = __ballot_sync(0xffffffffu, acc);
        if (lane == 0) s_wt[it & 1][warp] = __popc(bal);
        __syncthreads();
        int scan = s_wt[it & 1][lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, scan, d);
            if (lane >= d) scan += up;
        }
        const int total = __shfl_sync(0xffffffffu, scan, 31);
        const int rank = found + (warp ? __shfl_sync(0xffffffffu, scan, warp - 1) : 0) + __popc(bal & ((1u << lane) - 1u));
        if (acc && rank < need) {
            const double v1 = fma((double)(w0 >> 5), 67108864.0, (double)(w1 >> 6));
            const double v2 = fma((double)(w2 >> 5), 67108864.0, (double)(w3 >> 6));
            const double x1 = fma(v1, 1.0 / 4503599627370496.0, -1.0), x2 = fma(v2, 1.0 / 4503599627370496.0, -1.0);
            const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
            const double f = sqrt(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
            const int p = ce + 2 * rank;
            out[p] = (float)__dmul_rn(__dmul_rn(f, x2), scale);
            const double g2 = __dmul_rn(f, x1);
            if (p + 1 < N) out[p + 1] = (float)__dmul_rn(g2, scale);
            else if (e + 1 < 2 * n_pairs) out[N] = (float)__dmul_rn(g2, scale);            // = value 0 of the stream's next evaluation
            else gauss_io[sid] = g2;                                                        // the cache the stream hands back
        }
        found += total;
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
    const int N = normals_per_eval;
    // ---- how many words can a stream consume?  Per rollout ceil(N / 2) accepted attempts at acceptance pi / 4 (a negative
    //      binomial count), 4 words each; per pair one index draw (~1.08 words with rejections, bounded generously) and the
    //      coins.  The jump-ahead pass generates the mean + 12 sigma of the total (+ slack); the walk flags an overflow. ----
    const double p_acc = 0.78539816339744830962, n_acc = (N + 1) / 2;
    const double att_mean = n_acc / p_acc, att_sd = sqrt(n_acc * (1.0 - p_acc)) / p_acc;
    const double evals = 2.0 * n_per_stream;
    const double words_max = 624.0 + n_per_stream * (8.0 + 4.0 * coins) + 4.0 * (evals * att_mean + 12.0 * sqrt(evals) * att_sd + 64.0) +
                             2.0 * MG_WIN + 16.0 * MT_NW;
    const long long blocks_needed = (long long)(words_max / MT_NW) + 1;
    // jump-ahead when a stream is long enough to be worth splitting (ES_MT_JUMP=0 / 1 overrides; ES_MT_JUMP_LB: log2 of the
    // segment length in blocks, for tests)
    const char* ej = getenv("ES_MT_JUMP");
    const bool jump = ej ? atoi(ej) != 0 : blocks_needed >= 2048;
    int lb_log2 = 3;
    if (jump) {
        const char* el = getenv("ES_MT_JUMP_LB");
        if (el) lb_log2 = atoi(el);
        else
            while (lb_log2 < MJ_NPOLY - 1 && (long long)n_streams * ((blocks_needed + (1LL << lb_log2) - 1) >> lb_log2) > 4LL * ctx->sm_count) ++lb_log2;
        if (lb_log2 < 0) lb_log2 = 0;
        if (lb_log2 > MJ_NPOLY - 1) lb_log2 = MJ_NPOLY - 1;
    }
    const long long n_seg = jump ? (blocks_needed + (1LL << lb_log2) - 1) >> lb_log2 : 0;
    if (jump && (((n_seg << lb_log2) >> MJ_NPOLY) != 0 || (double)(n_seg << lb_log2) * MT_NW > 4.0e9)) {       // the block index of a segment start must fit the available jumps
        es_set_error("es_draw_noisy: %lld blocks per stream exceed the jump-ahead range (2^%d blocks)", n_seg << lb_log2, MJ_NPOLY);
        return ES_ERR_UNSUPPORTED;
    }
    // (whole chunks of MJ_CW words, + one chunk of padding: the flags kernel reads 4 words past every attempt start)
    const size_t gen_words = jump ? (size_t)(1 + (n_seg << lb_log2)) * MT_NW : 0;
    const long long n_chunks = jump ? (long long)((gen_words + MJ_CW - 1) / MJ_CW) : 0;
    const size_t stride_words = jump ? (size_t)(n_chunks + 1) * MJ_CW : 0;
    const int a_max = (N + 1) / 2 > 0 ? (N + 1) / 2 : 1;
    const size_t n_eval = (size_t)n_streams * 2 * n_per_stream;
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t c0_bytes = pad((size_t)n_streams * sizeof(int32_t)), g0_bytes = pad((size_t)n_streams * sizeof(double));
    if (jump) {
        // scratch: the streams' words | accept masks [stream][phase][chunk][32] | chunk counts -> prefixes [stream][phase][chunk]
        //          | first word of every rollout [stream][2 n] | incoming cache per stream
        const size_t words_bytes = pad((size_t)n_streams * stride_words * sizeof(uint32_t));
        const size_t mask_bytes = pad((size_t)n_streams * 4 * n_chunks * 32 * sizeof(uint32_t));
        const size_t cnt_bytes = pad((size_t)n_streams * 4 * n_chunks * sizeof(uint32_t));
        const size_t reg_bytes = pad(n_eval * sizeof(uint32_t));
        void* scratch = nullptr;
        int rc = es_ctx_scratch(ctx, words_bytes + mask_bytes + cnt_bytes + reg_bytes + c0_bytes + g0_bytes, &scratch);
        if (rc) return rc;
        char* at = (char*)scratch;
        uint32_t* words = (uint32_t*)at; at += words_bytes;
        uint32_t* masks = (uint32_t*)at; at += mask_bytes;
        uint32_t* counts = (uint32_t*)at; at += cnt_bytes;
        uint32_t* reg_start = (uint32_t*)at; at += reg_bytes;
        int32_t* c0 = (int32_t*)at; at += c0_bytes;
        double* gauss0 = (double*)at;
        const size_t smem = (size_t)(MJ_WIN_BLOCKS + 2) * MT_NW * sizeof(uint32_t);
        ES_CHECK_CUDA(cudaFuncSetAttribute(mt_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mt_fill_kernel<<<(unsigned)(n_streams * n_seg), MG_THREADS, smem, stream>>>(mt_key, (int)n_seg, lb_log2, words, stride_words);
        ES_LAUNCHED(ctx);
        mt_flags_kernel<<<dim3((unsigned)n_chunks, (unsigned)n_streams), MJ_CA, 0, stream>>>(words, stride_words, (int)n_chunks, masks, counts);
        ES_LAUNCHED(ctx);
        mt_scan_kernel<<<n_streams * 4, 1024, 0, stream>>>(counts, (int)n_chunks);
        ES_LAUNCHED(ctx);
        mt_walk_kernel<<<n_streams, 32, 0, stream>>>(mt_key, mt_pos, has_gauss, gauss, n_per_stream, rng, mask, coins, N, idx_out, extra_out,
                                                     reg_start, c0, gauss0, words, stride_words, (int)n_chunks, masks, counts,
                                                     (uint32_t)(gen_words - 8), ctx->err_dev);
        ES_LAUNCHED(ctx);
        mt_emit_kernel<<<(unsigned)n_eval, 1024, 0, stream>>>(words, stride_words, reg_start, c0, gauss0, n_per_stream, N, scale,
                                                              noise_out, gauss);
        ES_LAUNCHED(ctx);
        return ES_OK;
    }
    // sequential kernel.  scratch: the accepted attempts' words [stream][evaluation][a_max] (16 bytes per two gaussians), the
    // incoming cache per stream
    const size_t acc_bytes = pad(n_eval * a_max * sizeof(uint4));
    void* scratch = nullptr;
    int rc = es_ctx_scratch(ctx, acc_bytes + c0_bytes + g0_bytes, &scratch);
    if (rc) return rc;
    uint4* acc4 = (uint4*)scratch;
    int32_t* c0 = (int32_t*)((char*)scratch + acc_bytes);
    double* gauss0 = (double*)((char*)scratch + acc_bytes + c0_bytes);
    {
        const size_t smem = (size_t)(2 * MG_RING + 1) * MG_N * sizeof(uint32_t);
        ES_CHECK_CUDA(cudaFuncSetAttribute(mt_gauss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mt_gauss_kernel<<<n_streams, MG_THREADS, smem, stream>>>(mt_key, mt_pos, has_gauss, gauss, n_per_stream, rng, mask, coins, N,
                                                                 idx_out, extra_out, acc4, a_max, c0, gauss0);
        ES_LAUNCHED(ctx);
    }
    const long long total = (long long)n_eval * a_max;
    int blocks = es_div_up(total, 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    mt_gauss_finish_kernel<<<blocks, 256, 0, stream>>>(acc4, a_max, c0, gauss0, n_streams, n_per_stream, N, scale, noise_out, gauss);
    ES_LAUNCHED(ctx);
    return ES_OK;
}
