// rollout_tc.cu -- fused perturb + MLP rollout + fitness on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM, operands staged in shared memory by cp.async.bulk).
//
// Same contract as rollout_f32.cu (reference: src/core/policy.py:61-64, src/nn/nn.py:35-46,
// src/gym/gym_runner.py:50-54, src/gym/training_result.py:28) for the policy family the
// BASELINE configs name: obs -> 64 -> 64 -> act (act <= 32, obs <= 1024), tanh after every layer.
//
// One CTA = one antithetic pair at a time (persistent over pairs), time on the MMA M dimension:
//   per 128-step tile of the episode
//     L1:  [U | V] (128 x 128, fp32, TMEM) = Xn_tile (128 x obs, bf16)  x  [theta1 ; sigma*eps1]^T
//          z1+- = U +- V + b1+-            (U is common to both signs: its bf16 rounding cancels
//                                            to first order in f+ - f-, V carries the perturbation)
//     epi1: h1+- = tanh(z1+-) -> bf16 -> shared (K-major, 128B swizzle) = A operand of L2
//     L2:  D2+- (128 x 64) = h1+- x (theta2 +- sigma*eps2)^T ;  epi2: h2+- = tanh(D2+- + b2+-)
//     L3:  D3+- (128 x 32) = h2+- x (theta3 +- sigma*eps3)^T ;  epi3: a = tanh(D3 + b3),
//          r_t = <a_t, c_t>, fitness += r_t, pos += a_t[0..2]
// Warp roles (448 threads, warp-specialised, mbarrier pipelines only -- no CTA-wide barrier in the loop):
//   warp 0      producer: cp.async.bulk of the observation stages (ring of TC_NST x 16 KB)
//   warp 1      L1 MMA issuer (one elected lane) and TMEM owner.  D1 is double-buffered in TMEM, so
//               L1 of tile m+1 runs on the tensor pipe while the epilogue works on tile m.
//   warp 22     L2/L3 MMA issuer: the short per-sign MMAs are issued by their own thread the moment
//               the epilogue publishes H (blocking mbarrier waits, no polling); issuing costs the
//               thread ~10^2 cycles per instruction, so one thread for everything was the bottleneck.
//   warps 2-17  epilogue (TMEM lane quarter = warp % 4, column quarter = (warp-2)/4; four warps per
//               scheduler hide the MUFU / dependent-ALU latency): the + and - sign
//               chains are interleaved, so waiting for L2+/L3+ is covered by the other sign's work.
//   warps 18-21 builders: convert the NEXT pair's noise slice (float32, arbitrary 4-byte alignment in
//               the table) into the bf16 / swizzled B operands as soon as the current pair's last L1
//               (resp. last L3) has retired, one pair ahead of the MMA issuer.
//
// The observation stream is pre-tiled once per generation by rollout_tc_prep_kernel into the
// exact shared-memory image of each (M-tile, K-chunk) stage, so a stage is ONE contiguous
// 16 KB cp.async.bulk (no tensor map needed).
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int TC_THREADS = 736;
constexpr int TC_EPI_WARP0 = 2, TC_EPI_WARPS = 16, TC_BLD_WARP0 = 18, TC_BLD_WARPS = 4, TC_MMA2_WARP = 22;
constexpr int TC_H = 64;            // hidden width (both hidden layers)
constexpr int TC_MT = 128;          // time steps per M tile
constexpr int TC_KC = 64;           // K elements per chunk (= 128 bytes of bf16 = one swizzle row)
constexpr int TC_NST = 4;           // A-operand stages
constexpr int TC_ACT_PAD = 32;      // L3 N (act_dim padded)
constexpr int TC_STAGE_BYTES = TC_MT * 128;          // 16 KB
constexpr int TC_TMEM_COLS = 512;
constexpr uint32_t TC_SPIN_LIMIT = 1u << 27;        // watchdog: trap instead of hanging the GPU

// ---- raw PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time when the phase is not complete,
// which would stall the MMA issuer's polling loop)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) {
        if (++spins > TC_SPIN_LIMIT) { printf("rollout_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* result_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(result_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// packed float32x2 add / subtract (sm_100: one instruction for two lanes)
__device__ __forceinline__ void add2(float& o0, float& o1, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "r"(b0), "r"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
    uint32_t c0, c1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(c0), "=r"(c1) : "l"(c));
    o0 = __uint_as_float(c0); o1 = __uint_as_float(c1);
}
__device__ __forceinline__ void sub2(float& o0, float& o1, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "r"(b0 ^ 0x80000000u), "r"(b1 ^ 0x80000000u));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
    uint32_t c0, c1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(c0), "=r"(c1) : "l"(c));
    o0 = __uint_as_float(c0); o1 = __uint_as_float(c1);
}
// warp-uniform leader election: keeps the surrounding control flow (and the descriptors) uniform so the
// MMA operands stay in uniform registers instead of being re-materialised per instruction
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// two tanh per MUFU op: round the pre-activations to bf16 (the result is stored as bf16 anyway)
__device__ __forceinline__ uint32_t tanh_bf16x2(float lo, float hi) {
    uint32_t packed, y;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(packed) : "f"(hi), "f"(lo));      // upper half <- hi, lower half <- lo
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(packed));
    return y;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float2 lds64f(uint32_t saddr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float lds32f(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
// tanh of two values -> packed bf16x2 (lower half = lo).  tanh.approx.bf16x2 is NOT a packed MUFU op on sm_100
// (SASS: two MUFU.TANH.BF16 + PRMTs), so the float32 approximation + one pack is both cheaper and more accurate.
__device__ __forceinline__ uint32_t tanh2_pack(float lo, float hi) {
    uint32_t y;
    const float a = tanh_fast(lo), b = tanh_fast(hi);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(b), "f"(a));
    return y;
}
// one arrival per warp: every lane orders its own shared writes towards the async proxy first
__device__ __forceinline__ void warp_arrive_after_smem_writes(uint64_t* bar, int lane) {
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);      // .x = lo (lower address), .y = hi
    return *reinterpret_cast<uint32_t*>(&v);
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (row, k) inside one K-chunk block (rows x 128 B) with the 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int k /*0..63*/) {
    return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1)));
}

// Fill rows [row0, row0+64) of every K-chunk block of B1 with bf16(scale * w[n][k]) (k < obs), bf16(scale * bias[n])
// in column `obs`, zero beyond.  A warp takes whole rows; a lane owns two adjacent columns per chunk, so global reads
// are coalesced 256-byte runs (12 independent loads in flight per row) and every shared store is one conflict-free
// 4-byte word of a swizzled 128-byte row.
__device__ __forceinline__ void tc_build_l1_rows(uint8_t* b1_base, int row0, const float* __restrict__ w,
                                                 const float* __restrict__ bvec, float scale, int obs, int nkc, int warp,
                                                 int nwarps, int lane) {
    constexpr int KB = 4;                                  // K chunks per batch; 2 rows per batch -> 16 loads in flight per lane
    const int c0 = 2 * lane;                               // this lane's two columns inside a chunk
    for (int n = 2 * warp; n < TC_H; n += 2 * nwarps) {    // two rows per iteration
        const float* __restrict__ wa = w + (size_t)n * obs;
        const float* __restrict__ wb = wa + obs;
        const float ba = __ldg(bvec + n), bb = __ldg(bvec + n + 1);
        for (int kc0 = 0; kc0 < nkc; kc0 += KB) {
            float xa0[KB], xa1[KB], xb0[KB], xb1[KB];
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const int k = (kc0 + j) * TC_KC + c0;
                // clamp the address instead of branching: out-of-range columns are replaced after the load
                const int k0 = min(k, obs - 1), k1 = min(k + 1, obs - 1);
                xa0[j] = __ldg(wa + k0); xa1[j] = __ldg(wa + k1);
                xb0[j] = __ldg(wb + k0); xb1[j] = __ldg(wb + k1);
            }
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const int kc = kc0 + j, k = kc * TC_KC + c0;
                if (kc < nkc) {
                    const float a0 = (k < obs) ? xa0[j] : ((k == obs) ? ba : 0.f), a1 = (k + 1 < obs) ? xa1[j] : ((k + 1 == obs) ? ba : 0.f);
                    const float b0 = (k < obs) ? xb0[j] : ((k == obs) ? bb : 0.f), b1 = (k + 1 < obs) ? xb1[j] : ((k + 1 == obs) ? bb : 0.f);
                    const uint32_t off = kc * TC_STAGE_BYTES;
                    *(uint32_t*)(b1_base + off + sw128_off(row0 + n, c0)) = pack_bf16x2(__fmul_rn(scale, a0), __fmul_rn(scale, a1));
                    *(uint32_t*)(b1_base + off + sw128_off(row0 + n + 1, c0)) = pack_bf16x2(__fmul_rn(scale, b0), __fmul_rn(scale, b1));
                }
            }
        }
    }
}

struct TcParams {
    const float* table;
    const int64_t* idx;
    const float* theta;
    const __nv_bfloat16* xnt;     // [n_mtiles][nkc][16 KB stage image]
    const float* rew_vec;         // [T][act]
    double* fit_pos;
    double* fit_neg;
    float* behv_pos;
    float* behv_neg;
    int n_pairs, obs, act, T, nkc, n_mtiles, fit_stride;
    float sigma, pos_scale;
    // flat parameter offsets
    int w1, b1, w2, b2, w3, b3;
    long long* trace;             // optional cycle-stamp trace of CTA 0 (ES_TC_TRACE env), NULL in production
    int dev_noload;               // dev experiment: skip the observation-tile copies (results are garbage)
};

#define TC_TRACE(role, slot) do { if (p.trace && blockIdx.x == 0 && (slot) < 512) p.trace[(role) * 512 + (slot)] = clock64(); } while (0)

struct TcSmemLayout {   // byte offsets from the 1024-aligned dynamic smem base
    uint32_t b1, a_stage, w2p, w2n, w3p, w3n, hp, hn, bias, red, bars, total;
};

constexpr int TC_BIAS_FLOATS = 2 * TC_H + 2 * TC_ACT_PAD;        // b2+, b2-, b3+, b3- (one buffer)

__host__ __device__ inline TcSmemLayout tc_layout(int nkc) {
    TcSmemLayout L;
    uint32_t o = 0;
    L.b1 = o;       o += (uint32_t)nkc * TC_STAGE_BYTES;         // [nkc][128 rows x 128 B]: rows 0-63 theta1, 64-127 sigma*eps1
    L.a_stage = o;  o += TC_NST * TC_STAGE_BYTES;
    L.w2p = o;      o += TC_H * 128;
    L.w2n = o;      o += TC_H * 128;
    L.w3p = o;      o += TC_ACT_PAD * 128;
    L.w3n = o;      o += TC_ACT_PAD * 128;
    L.hp = o;       o += TC_MT * 128;
    L.hn = o;       o += TC_MT * 128;
    L.bias = o;     o += 2 * TC_BIAS_FLOATS * 4;                 // double-buffered by pair parity
    L.red = o;      o += 2 * TC_EPI_WARPS * 8 * 4;               // per-pair reduction scratch, double-buffered
    L.bars = o;     o += 256;
    L.total = o;
    return L;
}

enum { BAR_FULL = 0, BAR_EMPTY = TC_NST, BAR_D1_FULL = 2 * TC_NST, BAR_D1_FREE = BAR_D1_FULL + 2,
       BAR_H1P = BAR_D1_FREE + 2, BAR_H1N, BAR_D2P, BAR_D2N, BAR_H2P, BAR_H2N, BAR_D3P, BAR_D3N,
       BAR_EPS_READY, BAR_EPS_FREE, BAR_W_READY, BAR_W_FREE, BAR_COUNT };
static_assert(BAR_COUNT * 8 + 16 <= 256, "barrier block too small");

// Descriptors are precomputed once (64-bit); stepping 16 bf16 (32 B) along K inside a 128B-swizzled row is +2 in the
// 16-byte-unit address field, stepping a whole 16 KB block is +1024.
// one K chunk (64 columns) of an L1 tile: 4 UMMA k-steps
__device__ __forceinline__ void tc_issue_l1_chunk(uint32_t d1, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int kc) {
    umma_bf16(d1, a_desc, b_desc, idesc, kc != 0);
    umma_bf16(d1, a_desc + 2, b_desc + 2, idesc, 1);
    umma_bf16(d1, a_desc + 4, b_desc + 4, idesc, 1);
    umma_bf16(d1, a_desc + 6, b_desc + 6, idesc, 1);
}
__device__ __forceinline__ void tc_issue_small(uint32_t d, uint64_t h_desc, uint64_t w_desc, uint32_t idesc) {
    umma_bf16(d, h_desc, w_desc, idesc, 0);
    umma_bf16(d, h_desc + 2, w_desc + 2, idesc, 1);
    umma_bf16(d, h_desc + 4, w_desc + 4, idesc, 1);
    umma_bf16(d, h_desc + 6, w_desc + 6, idesc, 1);
}

__global__ void __launch_bounds__(TC_THREADS, 1) rollout_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const TcSmemLayout L = tc_layout(p.nkc);
    uint64_t* bars = (uint64_t*)(smem + L.bars);
    uint32_t* tmem_slot = (uint32_t*)(smem + L.bars + BAR_COUNT * 8);
    float* bias_all = (float*)(smem + L.bias);
    float* red_all = (float*)(smem + L.red);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NMT = p.n_mtiles, NKC = p.nkc;

    // ---- one-time setup -------------------------------------------------------------------------
    for (uint32_t i = tid * 16; i < L.bars; i += TC_THREADS * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < TC_NST; ++s) { mbar_init(&bars[BAR_FULL + s], 1); mbar_init(&bars[BAR_EMPTY + s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars[BAR_D1_FULL + b], 1); mbar_init(&bars[BAR_D1_FREE + b], TC_EPI_WARPS); }
        mbar_init(&bars[BAR_H1P], TC_EPI_WARPS); mbar_init(&bars[BAR_H1N], TC_EPI_WARPS);
        mbar_init(&bars[BAR_H2P], TC_EPI_WARPS); mbar_init(&bars[BAR_H2N], TC_EPI_WARPS);
        mbar_init(&bars[BAR_D2P], 1); mbar_init(&bars[BAR_D2N], 1); mbar_init(&bars[BAR_D3P], 1); mbar_init(&bars[BAR_D3N], 1);
        mbar_init(&bars[BAR_EPS_READY], TC_BLD_WARPS); mbar_init(&bars[BAR_W_READY], TC_BLD_WARPS);
        mbar_init(&bars[BAR_EPS_FREE], 1); mbar_init(&bars[BAR_W_FREE], 1);
        fence_barrier_init();
    }
    __syncthreads();                    // zero fill + barrier init visible to everyone
    if (warp == 1) tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_d1[2] = {tmem, tmem + 128};
    const uint32_t tm_d2p = tmem + 256, tm_d2n = tmem + 320, tm_d3p = tmem + 384, tm_d3n = tmem + 416;

    if (warp == 0) {
        // ===================== producer: observation tiles =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x)
                for (int m = 0; m < NMT; ++m)
                    for (int kc = 0; kc < NKC; ++kc) {
                        mbar_wait(&bars[BAR_EMPTY + stage], phase ^ 1);
                        if (p.dev_noload) { mbar_arrive(&bars[BAR_FULL + stage]); }
                        else {
                        mbar_expect_tx(&bars[BAR_FULL + stage], TC_STAGE_BYTES);
                        bulk_g2s(smem + L.a_stage + stage * TC_STAGE_BYTES,
                                 (const uint8_t*)p.xnt + ((size_t)m * NKC + kc) * TC_STAGE_BYTES, TC_STAGE_BYTES,
                                 &bars[BAR_FULL + stage]);
                        }
                        if (++stage == TC_NST) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ===================== L1 MMA issuer (whole warp runs the loop; one elected lane issues) =====================
        {
            const uint32_t id_l1 = umma_idesc_bf16(TC_MT, 2 * TC_H);
            const uint64_t a_desc0 = umma_desc_sw128(smem_u32(smem + L.a_stage)), b_desc0 = umma_desc_sw128(smem_u32(smem + L.b1));
            uint32_t stage = 0, phase = 0, g = 0, i = 0;
            for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x, ++i) {
                mbar_wait(&bars[BAR_EPS_READY], i & 1);
                for (int m = 0; m < NMT; ++m, ++g) {
                    const uint32_t buf = g & 1, use = g >> 1;
                    mbar_wait(&bars[BAR_D1_FREE + buf], (use & 1) ^ 1);       // epilogue has drained this D1 buffer
                    tc_fence_after();
                    if (lane == 0) TC_TRACE(0, 4 * g + 0);
                    for (int kc = 0; kc < NKC; ++kc) {
                        mbar_wait(&bars[BAR_FULL + stage], phase);
                        tc_fence_after();
                        const uint64_t ad = a_desc0 + (uint64_t)stage * (TC_STAGE_BYTES >> 4);
                        const uint64_t bd = b_desc0 + (uint64_t)kc * (TC_STAGE_BYTES >> 4);
                        if (elect_one()) {
                            tc_issue_l1_chunk(tm_d1[buf], ad, bd, id_l1, kc);
                            umma_commit(&bars[BAR_EMPTY + stage]);           // stage reusable once these MMAs retire
                            if (kc == NKC - 1) umma_commit(&bars[BAR_D1_FULL + buf]);
                        }
                        __syncwarp();
                        // pace the stream: the tensor pipe runs MMAs in issue order, so a queue of L1 chunks would delay
                        // the short L2/L3 MMAs (issued by warp TC_MMA2_WARP) that the epilogue is waiting for.  Keep at
                        // most one chunk (4 MMAs) in flight: wait until this chunk has retired (its EMPTY commit fired).
                        mbar_wait(&bars[BAR_EMPTY + stage], phase);
                        if (++stage == TC_NST) { stage = 0; phase ^= 1; }
                    }
                    if (lane == 0) TC_TRACE(0, 4 * g + 1);
                }
                if (elect_one()) umma_commit(&bars[BAR_EPS_FREE]);           // the pair's last L1 is in flight
                __syncwarp();
            }
        }
    } else if (warp == TC_MMA2_WARP) {
        // ===================== L2 / L3 MMA issuer (warp-uniform loop, elected lane issues) =====================
        {
            const uint32_t id_l2 = umma_idesc_bf16(TC_MT, TC_H), id_l3 = umma_idesc_bf16(TC_MT, TC_ACT_PAD);
            const uint64_t hp = umma_desc_sw128(smem_u32(smem + L.hp)), hn = umma_desc_sw128(smem_u32(smem + L.hn));
            const uint64_t w2p = umma_desc_sw128(smem_u32(smem + L.w2p)), w2n = umma_desc_sw128(smem_u32(smem + L.w2n));
            const uint64_t w3p = umma_desc_sw128(smem_u32(smem + L.w3p)), w3n = umma_desc_sw128(smem_u32(smem + L.w3n));
            uint32_t g = 0, i = 0;
            for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x, ++i) {
                mbar_wait(&bars[BAR_W_READY], i & 1);
                for (int m = 0; m < NMT; ++m, ++g) {
                    const uint32_t par = g & 1;
                    mbar_wait(&bars[BAR_H1P], par); tc_fence_after();
                    if (elect_one()) { tc_issue_small(tm_d2p, hp, w2p, id_l2); umma_commit(&bars[BAR_D2P]); }
                    __syncwarp();
                    if (lane == 0) TC_TRACE(0, 4 * g + 2);
                    mbar_wait(&bars[BAR_H1N], par); tc_fence_after();
                    if (elect_one()) { tc_issue_small(tm_d2n, hn, w2n, id_l2); umma_commit(&bars[BAR_D2N]); }
                    __syncwarp();
                    mbar_wait(&bars[BAR_H2P], par); tc_fence_after();
                    if (elect_one()) { tc_issue_small(tm_d3p, hp, w3p, id_l3); umma_commit(&bars[BAR_D3P]); }
                    __syncwarp();
                    if (lane == 0) TC_TRACE(0, 4 * g + 3);
                    mbar_wait(&bars[BAR_H2N], par); tc_fence_after();
                    if (elect_one()) { tc_issue_small(tm_d3n, hn, w3n, id_l3); umma_commit(&bars[BAR_D3N]); }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(&bars[BAR_W_FREE]);             // fires when the pair's last L3 retires
                __syncwarp();
            }
        }
    } else if (warp < TC_BLD_WARP0) {
        // ===================== epilogue warps =====================
        const int we = warp - TC_EPI_WARP0;
        const int q = warp & 3, hq = we >> 2;                 // TMEM lane quarter, column quarter (16 columns)
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t hp_row = smem_u32(smem + L.hp + row * 128);
        const uint32_t hn_row = smem_u32(smem + L.hn + row * 128);
        const int sw = row & 7;
        const uint32_t ch0 = (uint32_t)(((2 * hq) ^ sw) << 4), ch1 = (uint32_t)(((2 * hq + 1) ^ sw) << 4);   // this warp's two 16-byte chunks
        const bool has_act = hq * 8 < p.act;                  // L3: this warp owns action columns 8hq..8hq+7
        uint32_t g = 0, i = 0;
        for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x, ++i) {
            const uint32_t bias = smem_u32(bias_all + (i & 1) * TC_BIAS_FLOATS);
            const uint32_t b2p = bias + hq * 64, b2n = bias + TC_H * 4 + hq * 64;
            const uint32_t b3p = bias + 2 * TC_H * 4 + hq * 32, b3n = bias + (2 * TC_H + TC_ACT_PAD) * 4 + hq * 32;
            float fitp = 0.f, fitn = 0.f, pp0 = 0.f, pp1 = 0.f, pp2 = 0.f, pn0 = 0.f, pn1 = 0.f, pn2 = 0.f;
            for (int m = 0; m < NMT; ++m, ++g) {
                const uint32_t par = g & 1, buf = g & 1, use = g >> 1;
                const int t = m * TC_MT + row;
                // reward coefficients of this warp's action columns: issue the loads before any waiting
                float cr[8];
                {
                    const float* __restrict__ crow = p.rew_vec + (size_t)(t < p.T ? t : 0) * p.act;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) cr[jj] = (hq * 8 + jj < p.act) ? __ldg(crow + hq * 8 + jj) : 0.f;
                }
                // ---- epi1: z1+- = U +- V (bias folded into the MMA) ----
                if (we == 0 && lane == 0) TC_TRACE(1, 8 * g + 0);
                mbar_wait(&bars[BAR_D1_FULL + buf], use & 1);
                if (we == 0 && lane == 0) TC_TRACE(1, 8 * g + 1);
                tc_fence_after();
                uint32_t u[16], v[16];
                tmem_ld16(tm_d1[buf] + lane_base + hq * 16, u);
                tmem_ld16(tm_d1[buf] + lane_base + TC_H + hq * 16, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[BAR_D1_FREE + buf]);   // U, V are in registers: the buffer may be refilled
                {
                    uint32_t w[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float z0, z1;
                        add2(z0, z1, u[2 * e], u[2 * e + 1], v[2 * e], v[2 * e + 1]);
                        w[e] = tanh2_pack(z0, z1);
                    }
                    sts128(hp_row + ch0, w[0], w[1], w[2], w[3]);
                    sts128(hp_row + ch1, w[4], w[5], w[6], w[7]);
                    warp_arrive_after_smem_writes(&bars[BAR_H1P], lane);
                    if (lane == 0) TC_TRACE(2, 16 * g + we);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float z0, z1;
                        sub2(z0, z1, u[2 * e], u[2 * e + 1], v[2 * e], v[2 * e + 1]);
                        w[e] = tanh2_pack(z0, z1);
                    }
                    sts128(hn_row + ch0, w[0], w[1], w[2], w[3]);
                    sts128(hn_row + ch1, w[4], w[5], w[6], w[7]);
                    warp_arrive_after_smem_writes(&bars[BAR_H1N], lane);
                }
                if (we == 0 && lane == 0) TC_TRACE(1, 8 * g + 2);
                if (m == 0) mbar_wait(&bars[BAR_W_READY], i & 1);      // biases of this pair are in place
                // ---- epi2 (+ then -): h2 = tanh(D2 + b2), overwrites this warp's part of H ----
#pragma unroll
                for (int sgn = 0; sgn < 2; ++sgn) {
                    mbar_wait(&bars[sgn ? BAR_D2N : BAR_D2P], par);
                    if (we == 0 && lane == 0) TC_TRACE(1, 8 * g + 3 + sgn);
                    if (lane == 0 && sgn == 0) TC_TRACE(3, 16 * g + we);
                    tc_fence_after();
                    uint32_t d[16];
                    tmem_ld16((sgn ? tm_d2n : tm_d2p) + lane_base + hq * 16, d);
                    tmem_ld_wait();
                    const uint32_t b2 = sgn ? b2n : b2p;
                    const uint32_t hrow = sgn ? hn_row : hp_row;
                    uint32_t w[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 bb = lds64f(b2 + e * 8);
                        float z0, z1;
                        add2(z0, z1, d[2 * e], d[2 * e + 1], __float_as_uint(bb.x), __float_as_uint(bb.y));
                        w[e] = tanh2_pack(z0, z1);
                    }
                    sts128(hrow + ch0, w[0], w[1], w[2], w[3]);
                    sts128(hrow + ch1, w[4], w[5], w[6], w[7]);
                    tc_fence_before();
                    warp_arrive_after_smem_writes(&bars[sgn ? BAR_H2N : BAR_H2P], lane);
                }
                // ---- epi3 (+ then -): a = tanh(D3 + b3); reward and position (columns 8hq..8hq+7) ----
#pragma unroll
                for (int sgn = 0; sgn < 2; ++sgn) {
                    if (we == 0 && lane == 0 && sgn == 0) TC_TRACE(1, 8 * g + 5);
                    mbar_wait(&bars[sgn ? BAR_D3N : BAR_D3P], par);
                    if (we == 0 && lane == 0) TC_TRACE(1, 8 * g + 6 + sgn);
                    if (has_act) {
                        tc_fence_after();
                        uint32_t d[8];
                        tmem_ld8((sgn ? tm_d3n : tm_d3p) + lane_base + hq * 8, d);
                        tmem_ld_wait();
                        const uint32_t b3 = sgn ? b3n : b3p;
                        float r = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const int j = hq * 8 + jj;
                            if (j < p.act) {
                                const float a = tanh_fast(__uint_as_float(d[jj]) + lds32f(b3 + jj * 4));
                                r = fmaf(a, cr[jj], r);
                                if (j == 0) q0 += a;
                                if (j == 1 % p.act) q1 += a;
                                if (j == 2 % p.act) q2 += a;
                            }
                        }
                        if (t < p.T) {
                            if (sgn) { fitn += r; pn0 += q0; pn1 += q1; pn2 += q2; }
                            else     { fitp += r; pp0 += q0; pp1 += q1; pp2 += q2; }
                        }
                        tc_fence_before();
                    }
                }
            }
            // ---- per-pair reduction over the epilogue threads ----
            float vals[8] = {fitp, fitn, pp0, pp1, pp2, pn0, pn1, pn2};
#pragma unroll
            for (int k = 0; k < 8; ++k) vals[k] = es_warp_sum(vals[k]);
            float* red = red_all + (i & 1) * TC_EPI_WARPS * 8;
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) red[we * 8 + k] = vals[k];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");     // epilogue warps only
            if (we == 0 && lane == 0) {
                float tot[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float sacc = 0.f;
#pragma unroll
                    for (int w = 0; w < TC_EPI_WARPS; ++w) sacc += red[w * 8 + k];
                    tot[k] = sacc;
                }
                p.fit_pos[(size_t)pair * p.fit_stride] = (double)tot[0];
                p.fit_neg[(size_t)pair * p.fit_stride] = (double)tot[1];
                if (p.behv_pos) {
                    p.behv_pos[pair * 3 + 0] = p.pos_scale * tot[2]; p.behv_pos[pair * 3 + 1] = p.pos_scale * tot[3];
                    p.behv_pos[pair * 3 + 2] = p.pos_scale * tot[4];
                    p.behv_neg[pair * 3 + 0] = p.pos_scale * tot[5]; p.behv_neg[pair * 3 + 1] = p.pos_scale * tot[6];
                    p.behv_neg[pair * 3 + 2] = p.pos_scale * tot[7];
                }
            }
        }
    } else if (warp < TC_MMA2_WARP) {
        // ===================== builder warps: next pair's B operands =====================
        const int bw = warp - TC_BLD_WARP0, btid = tid - TC_BLD_WARP0 * 32;
        constexpr int BT = TC_BLD_WARPS * 32;
        // theta1 half of B1 (rows 0..63 of every K chunk) is the same for every pair; column `obs` carries the bias
        tc_build_l1_rows(smem + L.b1, 0, p.theta + p.w1, p.theta + p.b1, 1.0f, p.obs, NKC, bw, TC_BLD_WARPS, lane);
        uint32_t i = 0;
        const float sg = p.sigma;
        for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x, ++i) {
            const float* __restrict__ eps = p.table + p.idx[pair];
            if (i > 0) mbar_wait(&bars[BAR_EPS_FREE], (i - 1) & 1);      // previous pair's last L1 has retired
            tc_build_l1_rows(smem + L.b1, TC_H, eps + p.w1, eps + p.b1, sg, p.obs, NKC, bw, TC_BLD_WARPS, lane);
            warp_arrive_after_smem_writes(&bars[BAR_EPS_READY], lane);
            if (pair + (int)gridDim.x < p.n_pairs) {                     // pull the next slice into L2 early
                const char* nxt = (const char*)(p.table + p.idx[pair + gridDim.x]);
                const int lines = (p.b3 + p.act) * 4 / 128 + 2;
                for (int l = btid; l < lines; l += BT) prefetch_l2(nxt + (size_t)l * 128);
            }
            // W2+-, W3+-, biases: loads first (registers), shared stores after the previous pair's last L3 retired
            constexpr int NB2 = (TC_H * TC_H / 2 + BT - 1) / BT;          // 16 column pairs per thread
            float e0[NB2], e1[NB2], t0[NB2], t1[NB2];
#pragma unroll
            for (int b = 0; b < NB2; ++b) {
                const int k2 = 2 * (btid + b * BT);
                e0[b] = __ldg(eps + p.w2 + k2); e1[b] = __ldg(eps + p.w2 + k2 + 1);
                t0[b] = __ldg(p.theta + p.w2 + k2); t1[b] = __ldg(p.theta + p.w2 + k2 + 1);
            }
            if (i > 0) mbar_wait(&bars[BAR_W_FREE], (i - 1) & 1);
#pragma unroll
            for (int b = 0; b < NB2; ++b) {
                const int k2 = 2 * (btid + b * BT);
                const int n = k2 >> 6, k = k2 & 63;
                const float d0 = __fmul_rn(sg, e0[b]), d1 = __fmul_rn(sg, e1[b]);
                *(uint32_t*)(smem + L.w2p + sw128_off(n, k)) = pack_bf16x2(__fadd_rn(t0[b], d0), __fadd_rn(t1[b], d1));
                *(uint32_t*)(smem + L.w2n + sw128_off(n, k)) = pack_bf16x2(__fadd_rn(t0[b], -d0), __fadd_rn(t1[b], -d1));
            }
            for (int k2 = 2 * btid; k2 < p.act * TC_H; k2 += 2 * BT) {    // W3+- (rows >= act stay zero)
                const int n = k2 >> 6, k = k2 & 63;
                const float d0 = __fmul_rn(sg, __ldg(eps + p.w3 + k2)), d1 = __fmul_rn(sg, __ldg(eps + p.w3 + k2 + 1));
                const float x0 = __ldg(p.theta + p.w3 + k2), x1 = __ldg(p.theta + p.w3 + k2 + 1);
                *(uint32_t*)(smem + L.w3p + sw128_off(n, k)) = pack_bf16x2(__fadd_rn(x0, d0), __fadd_rn(x1, d1));
                *(uint32_t*)(smem + L.w3n + sw128_off(n, k)) = pack_bf16x2(__fadd_rn(x0, -d0), __fadd_rn(x1, -d1));
            }
            {
                float* bias = bias_all + (i & 1) * TC_BIAS_FLOATS;
                if (btid < TC_H) {
                    const float d = __fmul_rn(sg, __ldg(eps + p.b2 + btid)), t = __ldg(p.theta + p.b2 + btid);
                    bias[btid] = __fadd_rn(t, d); bias[TC_H + btid] = __fadd_rn(t, -d);
                } else if (btid - TC_H < p.act) {
                    const int j = btid - TC_H;
                    const float d = __fmul_rn(sg, __ldg(eps + p.b3 + j)), t = __ldg(p.theta + p.b3 + j);
                    bias[2 * TC_H + j] = __fadd_rn(t, d); bias[2 * TC_H + TC_ACT_PAD + j] = __fadd_rn(t, -d);
                }
            }
            warp_arrive_after_smem_writes(&bars[BAR_W_READY], lane);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TC_TMEM_COLS);
}

// observation stream -> bf16, tiled into the shared-memory image of each (M tile, K chunk) stage
__global__ void rollout_tc_prep_kernel(const float* __restrict__ obsn, int T, int obs, int nkc, int n_mtiles,
                                       __nv_bfloat16* __restrict__ xnt) {
    const size_t total = (size_t)n_mtiles * nkc * TC_MT * TC_KC;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % TC_KC);
        const int row = (int)((i / TC_KC) % TC_MT);
        const int kc = (int)((i / (TC_KC * TC_MT)) % nkc);
        const int m = (int)(i / ((size_t)TC_KC * TC_MT * nkc));
        const int t = m * TC_MT + row, kk = kc * TC_KC + k;
        const float v = (kk < obs) ? ((t < T) ? obsn[(size_t)t * obs + kk] : 0.f) : ((kk == obs) ? 1.0f : 0.f);   // col `obs` = 1: bias
        uint8_t* stage = (uint8_t*)xnt + ((size_t)m * nkc + kc) * TC_STAGE_BYTES;
        *(__nv_bfloat16*)(stage + sw128_off(row, k)) = __float2bfloat16_rn(v);
    }
}

}  // namespace

int es_impl_rollout_tc(es_ctx* ctx, const float* table, int64_t table_len, const int64_t* idx, int n_pairs,
                       const float* theta, int P, float sigma, const int* layer_sizes, int n_layers, const float* obsn,
                       const float* rew_vec, int T, float pos_scale, double* fit_pos, double* fit_neg, int fit_stride,
                       float* behv_pos, float* behv_neg, cudaStream_t stream) {
    (void)table_len; (void)P;
    if (n_layers != 3 || layer_sizes[1] != TC_H || layer_sizes[2] != TC_H || layer_sizes[3] > TC_ACT_PAD ||
        layer_sizes[0] > 1023) {
        es_set_error("es_rollout_openloop(TC): the tensor-core path covers obs(<=1023)-64-64-act(<=32) tanh MLPs; "
                     "use ES_ROLLOUT_F32 for other shapes");
        return ES_ERR_UNSUPPORTED;
    }
    TcParams p;
    p.table = table; p.idx = idx; p.theta = theta; p.rew_vec = rew_vec;
    p.fit_pos = fit_pos; p.fit_neg = fit_neg; p.behv_pos = behv_pos; p.behv_neg = behv_neg;
    p.n_pairs = n_pairs; p.obs = layer_sizes[0]; p.act = layer_sizes[3]; p.T = T; p.fit_stride = fit_stride;
    p.sigma = sigma; p.pos_scale = pos_scale;
    p.nkc = es_div_up(p.obs + 1, TC_KC);                 // + the constant-1 column that carries the L1 bias
    p.n_mtiles = es_div_up(T, TC_MT);
    p.w1 = 0; p.b1 = p.obs * TC_H; p.w2 = p.b1 + TC_H; p.b2 = p.w2 + TC_H * TC_H; p.w3 = p.b2 + TC_H;
    p.b3 = p.w3 + TC_H * p.act;

    p.trace = nullptr;
    p.dev_noload = getenv("ES_TC_NOLOAD") ? 1 : 0;
    if (const char* e = getenv("ES_TC_TRACE")) p.trace = (long long*)strtoull(e, nullptr, 0);   // device pointer, dev tooling only
    const TcSmemLayout L = tc_layout(p.nkc);
    const size_t smem = (size_t)L.total + 1024;       // + alignment slack
    if (smem > 227 * 1024) {
        es_set_error("es_rollout_openloop(TC): obs_dim %d needs %zu bytes of shared memory (> 227 KB)", p.obs, smem);
        return ES_ERR_UNSUPPORTED;
    }
    // pre-tiled bf16 observation stream lives in the ctx scratch (768 KB for T=1000, obs=376)
    const size_t xnt_bytes = (size_t)p.n_mtiles * p.nkc * TC_STAGE_BYTES;
    void* scratch = nullptr;
    int rc = es_ctx_scratch(ctx, xnt_bytes, &scratch);
    if (rc) return rc;
    p.xnt = (const __nv_bfloat16*)scratch;
    {
        const size_t total = xnt_bytes / 2;
        int blocks = es_div_up((int64_t)total, 256);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        rollout_tc_prep_kernel<<<blocks, 256, 0, stream>>>(obsn, T, p.obs, p.nkc, p.n_mtiles, (__nv_bfloat16*)scratch);
        ES_LAUNCHED(ctx);
    }
    ES_CHECK_CUDA(cudaFuncSetAttribute(rollout_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = n_pairs < ctx->sm_count ? n_pairs : ctx->sm_count;
    rollout_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(p);
    ES_LAUNCHED(ctx);
    return ES_OK;
}
