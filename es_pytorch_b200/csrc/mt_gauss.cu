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
// pace), neither did moving the float64 math out (it was not the limiter).  Long streams therefore take the polynomial
// jump-ahead path in the second half of this file, which spreads one stream's regeneration over the whole GPU (8.4 ms for the
// same draw); this kernel stays the path for short streams.
//
// Output noise[((stream * n_pairs + pair) * 2 + sign) * N + t * act + j] = float32(gauss * scale): the float64 product rounded
// once; the rollout kernels add it to the float32 action (the reference forms the float64 sum and the env rounds it to
// float32: identical except when the float32 rounding of the noise term moves the sum across a rounding boundary, 2^-24
// relative on a term that is ~1e-2 of the action).
// log() here is CUDA's (<= 1 ulp), the reference's is glibc's: a gaussian can differ in its last float64 bit, which the
// float32 noise array does not see; the accept/reject arithmetic (the only part that steers the stream) is exact.
#include <math.h>
#include <stdlib.h>
#include <type_traits>
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
// The state transition of MT19937 is linear over GF(2); with g(x) = x^(624 * B) mod phi(x) (phi = its characteristic
// polynomial, tools/mt_jump/make_jump_polys.py) every word of the raw sequence obeys x[n + 624 * B] = XOR_{i : g_i = 1}
// x[n + i].  So the state 2^r blocks ahead is, word by word, an XOR of ~10 000 words out of the next 20 560: independent per
// word, no recurrence to follow.  mt_fill_kernel gives every CTA one segment of 2^lb blocks of one stream: it jumps from the
// stream's current state to its segment's first block (one jump per non-zero hexadecimal digit of the block index: the polynomials of B = m * 16^q, m = 1 .. 15, q = 0 .. 4, are precomputed), regenerates the segment and
// writes the TEMPERED words to global memory.  What follows no longer touches the recurrence: mt_flags_kernel computes the accept
// bit of every possible attempt (all four phases) with per-chunk counts, mt_scan_kernel their prefixes; mt_walk_kernel -- the only
// sequential step left, one warp per stream -- follows the reference's order with two table look-ups per rollout ("the next
// `need` accepted attempts of this phase end at word ..."); mt_emit_kernel turns every rollout's accepted attempts into its
// gaussians, one CTA per rollout.
// =====================================================================================================================
constexpr int MJ_NQ = 5, MJ_BITS_RANGE = 4 * MJ_NQ;              // hex digits of a block index that have polynomials: 2^20 blocks
constexpr int MJ_NPOLY = 15 * MJ_NQ;                            // x^(624 * m * 16^q) mod phi at [q * 15 + m - 1], m = 1 .. 15
constexpr int MJ_WIN_BLOCKS = 33;                              // the state block + 32 more: 20 592 words >= 19 937 + 624
// Layout in global memory (per stream): the tempered words in stream order; per phase ph = 0..3 (an attempt starts at a word
// index = ph mod 4) one accept bit per attempt (attempt a of phase ph = words 4 a + ph .. 4 a + ph + 3), in chunks of
// MJ_CA = 1024 attempts (32 mask words) with the number of accepted attempts of every chunk and its exclusive prefix.
constexpr int MJ_CA = 1024, MJ_CW = 4 * MJ_CA;                  // attempts / words per chunk
__device__ const uint32_t mj_polys[MJ_NPOLY][MT_NW] = {
#include "mt_jump_polys.inc"
};

// the set bits of every jump polynomial as a list of word offsets (built once per device by mj_lists_kernel): applying a jump
// is then "XOR the window words at these offsets", 8 offsets per 16-byte load, with independent loads in flight
constexpr int MJ_BITS = MT_NW * 32;
__device__ __align__(16) uint16_t mj_idx[MJ_NPOLY][MJ_BITS];
__device__ int mj_cnt[MJ_NPOLY];

__global__ void __launch_bounds__(640) mj_lists_kernel() {
    __shared__ int s_w[20];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t g = tid < MT_NW ? mj_polys[r][tid] : 0u;
    int inc = __popc(g);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += up;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    int at = before + inc - __popc(g);
    while (g) {
        mj_idx[r][at++] = (uint16_t)(tid * 32 + __ffs(g) - 1);
        g &= g - 1;
    }
    if (tid == 639) mj_cnt[r] = before + inc;
}

// Thread <-> word map of the fill: warp w, lane l looks at word 31 w + l and OWNS it when l < 31 (the last lane of a warp only
// supplies its word to its neighbour, so no lane has to form a second word); word 623, whose expression differs, sits right
// after word 622 in warp 20.  One barrier per block: a thread forms its new word N[i] from the old block O and O's twists T
// (see mt19937.cuh), takes N[i + 1] from the next lane and stores N[i] AND the new block's twist T'[i] = tw(N[i], N[i + 1]) --
// so the next block can start right after the barrier.  T[623] stays 0 in both twist buffers.  The main loop is unrolled over
// the two block / twist buffers so that every shared-memory address is a per-thread register plus an immediate (the first
// version recomputed them and issued ~110 instructions per warp and block; the kernel is issue-bound).
constexpr int MF_WARPS = 21, MF_THREADS = 32 * MF_WARPS;
struct MjTaps { int o, a, b, c; };
__device__ __forceinline__ MjTaps mj_taps(int i) {            // N[i] = O[o] ^ T[a] ^ T[b] ^ T[c], 0 <= i <= 622
    MjTaps m;
    m.o = i < MT_DW ? i + MT_MW : (i < 2 * MT_DW ? i + MT_MW - MT_DW : i + MT_MW - 2 * MT_DW);
    m.a = i;
    m.b = i >= MT_DW ? i - MT_DW : MT_NW - 1;
    m.c = i >= 2 * MT_DW ? i - 2 * MT_DW : MT_NW - 1;
    return m;
}
__device__ __forceinline__ uint32_t mj_last_word(const uint32_t* __restrict__ O, const uint32_t* __restrict__ T) {     // N[623]
    const uint32_t n0 = O[MT_MW] ^ T[0];
    const uint32_t n396 = O[396 + MT_MW - MT_DW] ^ T[396 - MT_DW] ^ T[396];
    return n396 ^ mt19937_twist(O[MT_NW - 1], n0);
}

// segments in the order the CTAs should start: most jumps first, so that the long CTAs do not end up in the last wave
__device__ __forceinline__ int mj_jumps(unsigned block) {       // non-zero hexadecimal digits = jumps to reach the block
    int n = 0;
    for (; block; block >>= 4) n += (block & 15u) ? 1 : 0;
    return n;
}
__global__ void __launch_bounds__(1024) mj_order_kernel(int n_seg, int lb_log2, uint16_t* __restrict__ order) {
    for (int k = threadIdx.x; k < n_seg; k += 1024) {
        const int pc = mj_jumps((unsigned)k << lb_log2);
        int rank = 0;
        for (int j = 0; j < n_seg; ++j) {
            const int pj = mj_jumps((unsigned)j << lb_log2);
            rank += (pj > pc || (pj == pc && j < k)) ? 1 : 0;
        }
        order[rank] = (uint16_t)k;
    }
}

__global__ void __launch_bounds__(MF_THREADS, 2)
mt_fill_kernel(const uint32_t* __restrict__ mt_key, int n_streams, int n_seg, int lb_log2, const uint16_t* __restrict__ order,
               uint32_t* __restrict__ words, size_t stride_words) {
    extern __shared__ uint32_t mj_smem[];
    uint32_t* xs = mj_smem;                                    // [33][624] raw words: the jump window; blocks 0 / 1: ping-pong of the fill
    uint32_t* s_T = mj_smem + MJ_WIN_BLOCKS * MT_NW;           // [2][624] twists of the block being read / being written
    uint32_t* s_P = s_T + 2 * MT_NW;                           // [4][624] partial sums of a jump (one per quarter of the polynomial)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sid = blockIdx.x % n_streams, k = order[blockIdx.x / n_streams];
    const int iw = 31 * warp + lane;                           // the word this thread looks at
    const bool last = iw == MT_NW - 1;
    const bool owner = (lane < 31 && iw < MT_NW - 1) || last;
    const bool tail_warp = warp == MF_WARPS - 1;
    const int i = iw < MT_NW - 1 ? iw : MT_NW - 2;             // (clamped: whole warps take part in the shuffle)
    const MjTaps tp = mj_taps(i);
    uint32_t* __restrict__ out = words + (size_t)sid * stride_words;
    for (int j = tid; j < MT_NW; j += MF_THREADS) {
        const uint32_t y = mt_key[(size_t)sid * MT_NW + j];
        xs[j] = y;
        if (k == 0) out[j] = mt19937_temper(y);               // block 0 = the incoming state: its unread words belong to the stream
    }
    if (tid < 2) s_T[tid * MT_NW + MT_NW - 1] = 0;
    __syncthreads();
    // one block: O, T -> D (raw words), Tn (its twists), opw (this thread's tempered word in global memory, when OUT)
    const bool owner_t = owner && !last;
    auto regen = [&](auto out_tag, const uint32_t* __restrict__ O, const uint32_t* __restrict__ T, uint32_t* __restrict__ D,
                     uint32_t* __restrict__ Tn, uint32_t* __restrict__ opw) {
        constexpr bool OUT = decltype(out_tag)::value;
        uint32_t n = O[tp.o] ^ T[tp.a] ^ T[tp.b] ^ T[tp.c];
        if (tail_warp) {
            if (last) n = mj_last_word(O, T);
        }
        const uint32_t nx = __shfl_down_sync(0xffffffffu, n, 1);
        const uint32_t tw = mt19937_twist(n, nx);
        if (owner) D[iw] = n;
        if (owner_t) Tn[iw] = tw;
        if (OUT) {
            const uint32_t y = mt19937_temper(n);
            if (owner) *opw = y;
        }
        __syncthreads();
    };
    using Yes = std::true_type;
    using No = std::false_type;
    auto twists_of = [&](const uint32_t* __restrict__ O, uint32_t* __restrict__ T) {
        if (tid < MT_NW - 1) T[tid] = mt19937_twist(O[tid], O[tid + 1]);
        __syncthreads();
    };
    uint32_t* const T0 = s_T;
    uint32_t* const T1 = s_T + MT_NW;
    // ---- jump to block k << lb_log2 ----
    const unsigned target = (unsigned)k << lb_log2;
    for (int q = MJ_NQ - 1; q >= 0; --q) {                     // (one jump per non-zero hex digit of the first block's index)
        const unsigned m = (target >> (4 * q)) & 15u;
        if (m == 0u) continue;
        const int r = q * 15 + (int)m - 1;
        twists_of(xs, T0);
        for (int b = 1; b < MJ_WIN_BLOCKS; b += 2) {           // 32 more blocks: the window of the jump
            regen(No(), xs + (b - 1) * MT_NW, T0, xs + b * MT_NW, T1, nullptr);
            regen(No(), xs + b * MT_NW, T1, xs + (b + 1) * MT_NW, T0, nullptr);
        }
        // four output words per thread (j, j + 156, j + 312, j + 468: the offsets are decoded once for four loads), and the
        // polynomial's set bits in four quarters, one per group of 156 threads; the partial sums meet in shared memory
        if (tid < MT_NW) {
            const int grp = tid / (MT_NW / 4), j = tid - grp * (MT_NW / 4);
            const uint16_t* __restrict__ L = mj_idx[r];
            const int n = mj_cnt[r];
            const int per = ((n + 3) / 4 + 7) & ~7;            // a quarter of the list, whole 16-byte loads
            const int lo = min(n, grp * per), hi = min(n, lo + per);
            const uint32_t* __restrict__ xb = xs + j;
            uint32_t acc[4] = {0u, 0u, 0u, 0u}, alt[4] = {0u, 0u, 0u, 0u};
            auto tap = [&](uint32_t off, uint32_t* a) {
                const uint32_t* __restrict__ q = xb + off;
                a[0] ^= q[0]; a[1] ^= q[MT_NW / 4]; a[2] ^= q[2 * (MT_NW / 4)]; a[3] ^= q[3 * (MT_NW / 4)];
            };
            int e = lo;
            for (; e + 8 <= hi; e += 8) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(L + e));
                tap(q.x & 0xFFFFu, acc); tap(q.x >> 16, alt);
                tap(q.y & 0xFFFFu, acc); tap(q.y >> 16, alt);
                tap(q.z & 0xFFFFu, acc); tap(q.z >> 16, alt);
                tap(q.w & 0xFFFFu, acc); tap(q.w >> 16, alt);
            }
            for (; e < hi; ++e) tap(L[e], acc);
#pragma unroll
            for (int c = 0; c < 4; ++c) s_P[grp * MT_NW + j + c * (MT_NW / 4)] = acc[c] ^ alt[c];
        }
        __syncthreads();
        if (tid < MT_NW)                                       // (word 0: only its top bit is state, and that bit is right)
            xs[tid] = s_P[tid] ^ s_P[MT_NW + tid] ^ s_P[2 * MT_NW + tid] ^ s_P[3 * MT_NW + tid];
        __syncthreads();
    }
    twists_of(xs, T0);
    // ---- regenerate the segment: blocks target + 1 .. target + 2^lb, two per iteration (buffers 0 -> 1 -> 0) ----
    const int Lb = 1 << lb_log2;
    uint32_t* __restrict__ opw = out + (size_t)(1 + target) * MT_NW + iw;
    uint32_t* const X0 = xs;
    uint32_t* const X1 = xs + MT_NW;
    int b = 0;
    for (; b + 2 <= Lb; b += 2) {
        regen(Yes(), X0, T0, X1, T1, opw);
        regen(Yes(), X1, T1, X0, T0, opw + MT_NW);
        opw += 2 * MT_NW;
    }
    if (b < Lb) regen(Yes(), X0, T0, X1, T1, opw);
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

// One WARP per chunk (4096 words = 1024 attempts of each phase): 32 steps of 128 words; lane l holds words 4 l .. 4 l + 3 of a step
// (one 16-byte load), takes the next three from lane l + 1 (lane 31: from the next step's first load, which is already in
// flight), and tests the four attempts that start at its words; lane s keeps the ballots of step s, so the 32 mask words of a
// phase leave as one 128-byte store and the chunk's counts need no shared memory and no barrier.
constexpr int MFL_WARPS = 8;
__global__ void __launch_bounds__(32 * MFL_WARPS)
mt_flags_kernel(const uint32_t* __restrict__ words, size_t stride_words, int n_chunks, uint32_t* __restrict__ masks,
                uint32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31, sid = blockIdx.y;
    const int c = blockIdx.x * MFL_WARPS + (threadIdx.x >> 5);
    if (c >= n_chunks) return;
    const uint4* __restrict__ w4 = reinterpret_cast<const uint4*>(words + (size_t)sid * stride_words + (size_t)c * MJ_CW) + lane;
    uint32_t keep[4] = {0u, 0u, 0u, 0u};
    uint4 cur = __ldg(w4), nxt1 = __ldg(w4 + 32), nxt2 = __ldg(w4 + 64), nxt3 = __ldg(w4 + 96);   // (the buffer is padded by a chunk)
#pragma unroll 4
    for (int s = 0; s < 32; ++s) {
        const uint4 far = __ldg(w4 + (s + 4) * 32);           // (reads up to 4 steps past the chunk: inside the next chunk / the padding)
        uint4 b;
        b.x = __shfl_down_sync(0xffffffffu, cur.x, 1);
        b.y = __shfl_down_sync(0xffffffffu, cur.y, 1);
        b.z = __shfl_down_sync(0xffffffffu, cur.z, 1);
        const uint32_t n0 = __shfl_sync(0xffffffffu, nxt1.x, 0), n1 = __shfl_sync(0xffffffffu, nxt1.y, 0), n2 = __shfl_sync(0xffffffffu, nxt1.z, 0);
        if (lane == 31) { b.x = n0; b.y = n1; b.z = n2; }
        const unsigned b0 = __ballot_sync(0xffffffffu, mj_accept(cur.x, cur.y, cur.z, cur.w));
        const unsigned b1 = __ballot_sync(0xffffffffu, mj_accept(cur.y, cur.z, cur.w, b.x));
        const unsigned b2 = __ballot_sync(0xffffffffu, mj_accept(cur.z, cur.w, b.x, b.y));
        const unsigned b3 = __ballot_sync(0xffffffffu, mj_accept(cur.w, b.x, b.y, b.z));
        if (lane == s) { keep[0] = b0; keep[1] = b1; keep[2] = b2; keep[3] = b3; }
        cur = nxt1; nxt1 = nxt2; nxt2 = nxt3; nxt3 = far;
    }
    uint32_t* mk = masks + ((size_t)sid * 4 * n_chunks + c) * 32 + lane;              // + ph * n_chunks * 32
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        mk[(size_t)ph * n_chunks * 32] = keep[ph];
        const int cnt = __reduce_add_sync(0xffffffffu, __popc(keep[ph]));
        if (lane == 0) counts[((size_t)sid * 4 + ph) * n_chunks + c] = (uint32_t)cnt;
    }
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
    const unsigned FULL = 0xffffffffu;
    const uint32_t nc_u = (uint32_t)n_chunks;                     // (host: n_chunks * 128 < 2^32, every table index fits 32 bits)
    bool overflow = false;
    for (int pair = 0; pair < n_pairs && !overflow; ++pair) {
        // the index: the first of the next words that passes the masked rejection (32 candidates at a time)
        uint32_t w = 0;
        for (;;) {
            if (cpos + 1 > limit) { overflow = true; break; }
            const uint32_t v = __ldg(gw + cpos + lane) & mask;          // (the buffer is padded by a chunk past `limit`)
            const unsigned ok = __ballot_sync(FULL, v <= rng);
            if (ok) {
                const int f = __ffs((int)ok) - 1;
                w = __shfl_sync(FULL, v, f);
                cpos += (uint32_t)f + 1u;
                break;
            }
            cpos += 32u;
        }
        if (overflow || cpos > limit) { overflow = true; break; }
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
            // everything the rollout needs is loaded at once (one memory latency per rollout): the mask row and prefix of the
            // chunk it starts in, the prefixes of a window of 32 chunks around the expected end (acceptance pi / 4; the spread
            // of the end is ~sqrt(need) attempts, a window is 32 768), and the mask rows of the three likeliest end chunks.
            // (32-bit arithmetic throughout: one warp per stream leaves every instruction's latency exposed)
            const uint32_t ch0 = a0 >> 10, r0 = a0 & 1023;
            const uint32_t span = (uint32_t)need + (uint32_t)(((uint64_t)(uint32_t)need * 1173554908u) >> 32);      // need * 4 / pi
            const uint32_t a_est = a0 + span;
            uint32_t est = a_est >> 10;
            if (est > nc_u - 1u) est = nc_u - 1u;
            uint32_t wlo = est > ch0 + 12u ? est - 12u : ch0;
            if (wlo + 32u > nc_u) wlo = nc_u > 32u ? nc_u - 32u : 0u;
            uint32_t ci = wlo + lane;
            const uint32_t m0 = __ldg(mph + ch0 * 32u + lane);
            const uint32_t base0 = __ldg(cph + ch0);
            uint32_t pre = (ci < nc_u) ? __ldg(cph + ci) : 0xFFFFFFFFu;
            const uint32_t e0 = est > 0u ? est - 1u : 0u, e2 = est + 1u < nc_u ? est + 1u : est;
            const uint32_t mc0 = __ldg(mph + e0 * 32u + lane), mc1 = __ldg(mph + est * 32u + lane), mc2 = __ldg(mph + e2 * 32u + lane);
            // pull what the NEXT draws will touch towards L2 while this rollout is resolved: the words around the expected end
            // (the next index / coins); for the next rollout, whose phase is not known yet, in all four phases: the mask rows
            // around its start (= this end) and around its expected end, and the prefix windows around its expected end
            {
                const uint32_t a_pf = a_est + (uint32_t)(lane - 12) * 8u;
                if (a_pf < (limit >> 2)) asm volatile("prefetch.global.L2 [%0];" ::"l"(gw + 4u * a_pf));
                const uint32_t est2 = (a_est + span) >> 10;
                if (lane < 24) {
                    const uint32_t l12 = lane < 12 ? lane : lane - 12;
                    const uint32_t nc = (lane < 12 ? est : est2) + (l12 % 3u) - 1u;
                    if (nc < nc_u) asm volatile("prefetch.global.L2 [%0];" ::"l"(mk + ((size_t)(l12 / 3u) * nc_u + nc) * 32));
                } else {
                    const uint32_t nc = (est2 > 12u ? est2 - 12u : 0u) + ((lane & 1) ? 31u : 0u);
                    if (nc < nc_u) asm volatile("prefetch.global.L2 [%0];" ::"l"(cu + (size_t)((lane - 24) >> 1) * nc_u + nc));
                }
            }
            // accepted attempts of the phase before a0
            const int below = __reduce_add_sync(FULL, (lane < (int)(r0 >> 5)) ? __popc(m0)
                                                      : (lane == (int)(r0 >> 5) ? __popc(m0 & ((1u << (r0 & 31)) - 1u)) : 0));
            const uint32_t target = base0 + (uint32_t)below + (uint32_t)need;      // accepted attempts before the END of the rollout
            // the chunk that contains the target-th accepted attempt: the last one whose prefix is < target
            unsigned lt = __ballot_sync(FULL, pre < target);
            if (lt == 0u || (lt == 0xFFFFFFFFu && wlo + 32u < nc_u)) {
                // outside the window (tens of sigma away from the estimate): bisect the prefixes, then look again from there
                uint32_t lo = ch0, hi = nc_u - 1u;                          // invariant: prefix[lo] < target
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1u) >> 1;
                    if (__ldg(cph + mid) < target) lo = mid; else hi = mid - 1u;
                }
                wlo = lo;
                ci = wlo + lane;
                pre = (ci < nc_u) ? __ldg(cph + ci) : 0xFFFFFFFFu;
                lt = __ballot_sync(FULL, pre < target);
            }
            const int sel = 31 - __clz((int)lt);                  // last lane with prefix < target (prefixes are non-decreasing)
            const uint32_t c1 = wlo + (uint32_t)sel, pre1 = __shfl_sync(FULL, pre, sel);
            const uint32_t m1 = c1 == est ? mc1 : (c1 == e0 ? mc0 : (c1 == e2 ? mc2 : __ldg(mph + c1 * 32u + lane)));
            int inc = __popc(m1);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += up;
            }
            const uint32_t want = target - pre1;                  // the want-th (1-based) accepted attempt of chunk c1 ends the rollout
            const unsigned ge = __ballot_sync(FULL, (uint32_t)inc >= want);
            if (ge == 0u) { overflow = true; break; }
            const int L = __ffs((int)ge) - 1;
            const uint32_t mL = __shfl_sync(FULL, m1, L);
            const int incL = __shfl_sync(FULL, inc, L);
            const int kth = (int)want - (incL - __popc(mL));      // 1-based among the set bits of lane L's word
            // the kth set bit of mL: the lane whose bit is set with kth - 1 set bits below it
            const unsigned hit = __ballot_sync(FULL, ((mL >> lane) & 1u) && __popc(mL & ((1u << lane) - 1u)) == kth - 1);
            const int bit = __ffs((int)hit) - 1;
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
    bool jump = ej ? atoi(ej) != 0 : blocks_needed >= 2048;
    int lb_log2 = 3;
    if (jump) {
        const char* el = getenv("ES_MT_JUMP_LB");
        if (el) lb_log2 = atoi(el);
        else
            while (lb_log2 < MJ_BITS_RANGE - 1 && (long long)n_streams * ((blocks_needed + (1LL << lb_log2) - 1) >> lb_log2) > 4LL * ctx->sm_count) ++lb_log2;
        if (lb_log2 < 0) lb_log2 = 0;
        if (lb_log2 > MJ_BITS_RANGE - 1) lb_log2 = MJ_BITS_RANGE - 1;
    }
    long long n_seg = jump ? (blocks_needed + (1LL << lb_log2) - 1) >> lb_log2 : 0;
    if (jump && (((n_seg << lb_log2) >> MJ_BITS_RANGE) != 0 || (double)(n_seg << lb_log2) * MT_NW > 4.0e9 || n_seg > 65535)) {
        // the block index of a segment start must fit the available jumps (2^20 blocks = 654 M words per stream) and a
        // 32-bit word index: longer streams take the sequential kernel unless the jump-ahead path was asked for explicitly
        if (ej && atoi(ej) != 0) {
            es_set_error("es_draw_noisy: %lld blocks per stream exceed the jump-ahead range (2^%d blocks)", n_seg << lb_log2, MJ_BITS_RANGE);
            return ES_ERR_UNSUPPORTED;
        }
        jump = false;
        n_seg = 0;
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
        const size_t ord_bytes = pad((size_t)n_seg * sizeof(uint16_t));
        void* scratch = nullptr;
        int rc = es_ctx_scratch(ctx, words_bytes + mask_bytes + cnt_bytes + reg_bytes + c0_bytes + g0_bytes + ord_bytes, &scratch);
        if (rc) return rc;
        char* at = (char*)scratch;
        uint32_t* words = (uint32_t*)at; at += words_bytes;
        uint32_t* masks = (uint32_t*)at; at += mask_bytes;
        uint32_t* counts = (uint32_t*)at; at += cnt_bytes;
        uint32_t* reg_start = (uint32_t*)at; at += reg_bytes;
        int32_t* c0 = (int32_t*)at; at += c0_bytes;
        double* gauss0 = (double*)at; at += g0_bytes;
        uint16_t* order = (uint16_t*)at;
        if (!ctx->mj_lists_ready) {
            mj_lists_kernel<<<MJ_NPOLY, 640, 0, stream>>>();
            ES_LAUNCHED(ctx);
            ctx->mj_lists_ready = 1;
        }
        mj_order_kernel<<<1, 1024, 0, stream>>>((int)n_seg, lb_log2, order);
        ES_LAUNCHED(ctx);
        const size_t smem = (size_t)(MJ_WIN_BLOCKS + 6) * MT_NW * sizeof(uint32_t);
        ES_CHECK_CUDA(cudaFuncSetAttribute(mt_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mt_fill_kernel<<<(unsigned)(n_streams * n_seg), MF_THREADS, smem, stream>>>(mt_key, n_streams, (int)n_seg, lb_log2, order, words,
                                                                                    stride_words);
        ES_LAUNCHED(ctx);
        mt_flags_kernel<<<dim3((unsigned)((n_chunks + MFL_WARPS - 1) / MFL_WARPS), (unsigned)n_streams), 32 * MFL_WARPS, 0, stream>>>(
            words, stride_words, (int)n_chunks, masks, counts);
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
