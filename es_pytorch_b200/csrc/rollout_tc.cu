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
// Warp roles: warp 0 = bulk-copy producer of the observation tiles, warp 1 = MMA issuer
// (one elected lane) and TMEM owner, warps 2-5 = epilogue (TMEM -> registers -> tanh -> shared).
// All warps cooperate, once per pair, in converting the pair's noise slice (float32, arbitrary
// 4-byte alignment in the table) into the bf16 / swizzled B operands in shared memory.
//
// The observation stream is pre-tiled once per generation by rollout_tc_prep_kernel into the
// exact shared-memory image of each (M-tile, K-chunk) stage, so a stage is ONE contiguous
// 16 KB cp.async.bulk (no tensor map needed).
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int TC_THREADS = 192;
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
                                                 int lane) {
    constexpr int KB = 8;                                  // K chunks per batch (16 loads in flight per lane)
    for (int n = warp; n < TC_H; n += TC_THREADS / 32) {
        const float* __restrict__ wr = w + (size_t)n * obs;
        for (int kc0 = 0; kc0 < nkc; kc0 += KB) {
            float x0[KB], x1[KB];
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const int k = (kc0 + j) * TC_KC + 2 * lane;
                x0[j] = x1[j] = 0.f;
                if (kc0 + j < nkc) {
                    x0[j] = (k < obs) ? __ldg(wr + k) : ((k == obs) ? __ldg(bvec + n) : 0.f);
                    x1[j] = (k + 1 < obs) ? __ldg(wr + k + 1) : ((k + 1 == obs) ? __ldg(bvec + n) : 0.f);
                }
            }
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                if (kc0 + j < nkc)
                    *(uint32_t*)(b1_base + (kc0 + j) * TC_STAGE_BYTES + sw128_off(row0 + n, 2 * lane)) =
                        pack_bf16x2(__fmul_rn(scale, x0[j]), __fmul_rn(scale, x1[j]));
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
};

struct TcSmemLayout {   // byte offsets from the 1024-aligned dynamic smem base
    uint32_t b1, a_stage, w2p, w2n, w3p, w3n, hp, hn, bias, bars, total;
};

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
    L.bias = o;     o += 2 * (TC_H + TC_H + TC_ACT_PAD) * 4;     // b1+,b1-,b2+,b2-,b3+,b3-
    L.bars = o;     o += 256;
    L.total = o;
    return L;
}

enum { BAR_FULL = 0, BAR_EMPTY = TC_NST, BAR_D1_FULL = 2 * TC_NST, BAR_D1_FREE, BAR_H1_READY, BAR_D2_FULL, BAR_H2_READY,
       BAR_D3_FULL, BAR_COUNT };

__global__ void __launch_bounds__(TC_THREADS, 1) rollout_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const TcSmemLayout L = tc_layout(p.nkc);
    uint64_t* bars = (uint64_t*)(smem + L.bars);
    uint32_t* tmem_slot = (uint32_t*)(smem + L.bars + BAR_COUNT * 8);
    float* s_red = (float*)(smem + L.bars + BAR_COUNT * 8 + 16);     // 4 warps x 8 floats
    float* bias = (float*)(smem + L.bias);
    float* b2p = bias + 2 * TC_H, *b2n = bias + 3 * TC_H;
    float* b3p = bias + 4 * TC_H, *b3n = bias + 4 * TC_H + TC_ACT_PAD;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup -------------------------------------------------------------------------
    for (uint32_t i = tid * 16; i < L.bars; i += TC_THREADS * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < TC_NST; ++s) { mbar_init(&bars[BAR_FULL + s], 1); mbar_init(&bars[BAR_EMPTY + s], 1); }
        mbar_init(&bars[BAR_D1_FULL], 1);
        mbar_init(&bars[BAR_D1_FREE], 128);
        mbar_init(&bars[BAR_H1_READY], 128);
        mbar_init(&bars[BAR_D2_FULL], 1);
        mbar_init(&bars[BAR_H2_READY], 128);
        mbar_init(&bars[BAR_D3_FULL], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (warp == 1) tmem_alloc(tmem_slot, TC_TMEM_COLS);
    // theta1 half of B1 (rows 0..63 of every K chunk) is the same for every pair; column `obs` carries the
    // bias (the observation tiles hold a constant 1 there), so z1 = U +- V needs no bias add in the epilogue
    tc_build_l1_rows(smem + L.b1, 0, p.theta + p.w1, p.theta + p.b1, 1.0f, p.obs, p.nkc, warp, lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_d1 = tmem, tm_d2p = tmem + 128, tm_d2n = tmem + 192, tm_d3p = tmem + 256, tm_d3n = tmem + 288;

    // pipeline state (each role keeps only what it uses)
    uint32_t prod_stage = 0, prod_phase = 0;      // producer
    uint32_t cons_stage = 0, cons_phase = 0;      // MMA issuer
    uint32_t tile_parity = 0;                     // flips once per (pair, tile): d1_full, h1_ready, d2_full, h2_ready, d3_full
    uint32_t d1_free_parity = 0;

    for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x) {
        // ---- build this pair's B operands (all warps) ---------------------------------------------
        const float* __restrict__ eps = p.table + p.idx[pair];
        const float sg = p.sigma;
        tc_build_l1_rows(smem + L.b1, TC_H, eps + p.w1, eps + p.b1, sg, p.obs, p.nkc, warp, lane);   // sigma*eps1 (+ sigma*eps_b1)
        if (pair + (int)gridDim.x < p.n_pairs) {             // pull the next pair's slice into L2 while this one computes
            const char* nxt = (const char*)(p.table + p.idx[pair + gridDim.x]);
            const int lines = (p.b3 + p.act) * 4 / 128 + 2;
            for (int i = tid; i < lines; i += TC_THREADS) prefetch_l2(nxt + (size_t)i * 128);
        }
        {                                                                   // W2+- = theta2 +- sigma*eps2
            constexpr int NB = (TC_H * TC_H / 2 + TC_THREADS - 1) / TC_THREADS;     // pairs of columns per thread
            float e0[NB], e1[NB], t0[NB], t1[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int i = 2 * (tid + b * TC_THREADS);
                const bool ok = i < TC_H * TC_H;
                e0[b] = ok ? __ldg(eps + p.w2 + i) : 0.f; e1[b] = ok ? __ldg(eps + p.w2 + i + 1) : 0.f;
                t0[b] = ok ? __ldg(p.theta + p.w2 + i) : 0.f; t1[b] = ok ? __ldg(p.theta + p.w2 + i + 1) : 0.f;
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int i = 2 * (tid + b * TC_THREADS);
                if (i < TC_H * TC_H) {
                    const int n = i >> 6, k = i & 63;
                    const float d0 = __fmul_rn(sg, e0[b]), d1 = __fmul_rn(sg, e1[b]);
                    *(uint32_t*)(smem + L.w2p + sw128_off(n, k)) = pack_bf16x2(__fadd_rn(t0[b], d0), __fadd_rn(t1[b], d1));
                    *(uint32_t*)(smem + L.w2n + sw128_off(n, k)) = pack_bf16x2(__fadd_rn(t0[b], -d0), __fadd_rn(t1[b], -d1));
                }
            }
        }
        for (int i = tid; i < p.act * TC_H; i += TC_THREADS) {              // W3+- (rows >= act stay zero)
            const int n = i >> 6, k = i & 63;
            const float d = __fmul_rn(sg, __ldg(eps + p.w3 + i)), t = __ldg(p.theta + p.w3 + i);
            *(__nv_bfloat16*)(smem + L.w3p + sw128_off(n, k)) = __float2bfloat16_rn(__fadd_rn(t, d));
            *(__nv_bfloat16*)(smem + L.w3n + sw128_off(n, k)) = __float2bfloat16_rn(__fadd_rn(t, -d));
        }
        if (tid < TC_H) {
            float d = __fmul_rn(sg, __ldg(eps + p.b2 + tid)), t = __ldg(p.theta + p.b2 + tid);
            b2p[tid] = __fadd_rn(t, d); b2n[tid] = __fadd_rn(t, -d);
            if (tid < p.act) {
                d = __fmul_rn(sg, __ldg(eps + p.b3 + tid)); t = __ldg(p.theta + p.b3 + tid);
                b3p[tid] = __fadd_rn(t, d); b3n[tid] = __fadd_rn(t, -d);
            }
        }
        fence_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncthreads();

        if (warp == 0) {
            // ===== producer: observation tiles, one 16 KB bulk copy per (tile, K chunk) =====
            if (lane == 0) {
                for (int m = 0; m < p.n_mtiles; ++m)
                    for (int kc = 0; kc < p.nkc; ++kc) {
                        mbar_wait(&bars[BAR_EMPTY + prod_stage], prod_phase ^ 1);
                        mbar_expect_tx(&bars[BAR_FULL + prod_stage], TC_STAGE_BYTES);
                        bulk_g2s(smem + L.a_stage + prod_stage * TC_STAGE_BYTES,
                                 (const uint8_t*)p.xnt + ((size_t)m * p.nkc + kc) * TC_STAGE_BYTES, TC_STAGE_BYTES,
                                 &bars[BAR_FULL + prod_stage]);
                        if (++prod_stage == TC_NST) { prod_stage = 0; prod_phase ^= 1; }
                    }
            }
        } else if (warp == 1) {
            // ===== MMA issuer =====
            if (lane == 0) {
                const uint32_t id_l1 = umma_idesc_bf16(TC_MT, 2 * TC_H), id_l2 = umma_idesc_bf16(TC_MT, TC_H),
                               id_l3 = umma_idesc_bf16(TC_MT, TC_ACT_PAD);
                uint32_t tp = tile_parity, fp = d1_free_parity;
                for (int m = 0; m < p.n_mtiles; ++m) {
                    // D1 must have been drained by the epilogue of the previous tile
                    mbar_wait(&bars[BAR_D1_FREE], fp ^ 1);
                    fp ^= 1;
                    tc_fence_after();
                    for (int kc = 0; kc < p.nkc; ++kc) {
                        mbar_wait(&bars[BAR_FULL + cons_stage], cons_phase);
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(smem + L.a_stage + cons_stage * TC_STAGE_BYTES);
                        const uint32_t b0 = smem_u32(smem + L.b1 + kc * TC_STAGE_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < TC_KC / 16; ++k4)
                            umma_bf16(tm_d1, umma_desc_sw128(a0 + k4 * 32), umma_desc_sw128(b0 + k4 * 32), id_l1, (kc | k4) != 0);
                        umma_commit(&bars[BAR_EMPTY + cons_stage]);       // stage reusable once these MMAs retire
                        if (++cons_stage == TC_NST) { cons_stage = 0; cons_phase ^= 1; }
                    }
                    umma_commit(&bars[BAR_D1_FULL]);
                    // L2: D2+- = H1+- x W2+-^T
                    mbar_wait(&bars[BAR_H1_READY], tp);
                    tc_fence_after();
                    {
                        const uint32_t hp = smem_u32(smem + L.hp), hn = smem_u32(smem + L.hn);
                        const uint32_t wp = smem_u32(smem + L.w2p), wn = smem_u32(smem + L.w2n);
#pragma unroll
                        for (int k4 = 0; k4 < TC_H / 16; ++k4)
                            umma_bf16(tm_d2p, umma_desc_sw128(hp + k4 * 32), umma_desc_sw128(wp + k4 * 32), id_l2, k4 != 0);
#pragma unroll
                        for (int k4 = 0; k4 < TC_H / 16; ++k4)
                            umma_bf16(tm_d2n, umma_desc_sw128(hn + k4 * 32), umma_desc_sw128(wn + k4 * 32), id_l2, k4 != 0);
                        umma_commit(&bars[BAR_D2_FULL]);
                    }
                    // L3: D3+- = H2+- x W3+-^T
                    mbar_wait(&bars[BAR_H2_READY], tp);
                    tc_fence_after();
                    {
                        const uint32_t hp = smem_u32(smem + L.hp), hn = smem_u32(smem + L.hn);
                        const uint32_t wp = smem_u32(smem + L.w3p), wn = smem_u32(smem + L.w3n);
#pragma unroll
                        for (int k4 = 0; k4 < TC_H / 16; ++k4)
                            umma_bf16(tm_d3p, umma_desc_sw128(hp + k4 * 32), umma_desc_sw128(wp + k4 * 32), id_l3, k4 != 0);
#pragma unroll
                        for (int k4 = 0; k4 < TC_H / 16; ++k4)
                            umma_bf16(tm_d3n, umma_desc_sw128(hn + k4 * 32), umma_desc_sw128(wn + k4 * 32), id_l3, k4 != 0);
                        umma_commit(&bars[BAR_D3_FULL]);
                    }
                    tp ^= 1;
                }
            }
        } else {
            // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
            const int q = warp & 3;
            const int row = q * 32 + lane;                           // row of the tile == TMEM lane
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            uint8_t* hp_row = smem + L.hp + row * 128;
            uint8_t* hn_row = smem + L.hn + row * 128;
            const int sw = row & 7;
            float fitp = 0.f, fitn = 0.f, pp0 = 0.f, pp1 = 0.f, pp2 = 0.f, pn0 = 0.f, pn1 = 0.f, pn2 = 0.f;
            uint32_t tp = tile_parity;
            for (int m = 0; m < p.n_mtiles; ++m) {
                // ---- epi1: h1+- = tanh(U +- V + b1+-) ----
                mbar_wait(&bars[BAR_D1_FULL], tp);
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t u[32], v[32];
                    tmem_ld32(tm_d1 + lane_base + half * 32, u);
                    tmem_ld32(tm_d1 + lane_base + TC_H + half * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {                    // 16-byte chunk = 8 columns
                        uint32_t wp[4], wn[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = c * 8 + e * 2;
                            const float u0 = __uint_as_float(u[j]), u1 = __uint_as_float(u[j + 1]);
                            const float v0 = __uint_as_float(v[j]), v1 = __uint_as_float(v[j + 1]);
                            wp[e] = tanh_bf16x2(u0 + v0, u1 + v1);
                            wn[e] = tanh_bf16x2(u0 - v0, u1 - v1);
                        }
                        const int chunk = (half * 4 + c) ^ sw;
                        *(uint4*)(hp_row + chunk * 16) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
                        *(uint4*)(hn_row + chunk * 16) = make_uint4(wn[0], wn[1], wn[2], wn[3]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&bars[BAR_D1_FREE]);                     // next tile's L1 may overwrite D1
                fence_async_smem();
                mbar_arrive(&bars[BAR_H1_READY]);
                // ---- epi2: h2+- = tanh(D2+- + b2+-) (overwrites H1: the L2 MMAs have retired) ----
                mbar_wait(&bars[BAR_D2_FULL], tp);
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t dp[32], dn[32];
                    tmem_ld32(tm_d2p + lane_base + half * 32, dp);
                    tmem_ld32(tm_d2n + lane_base + half * 32, dn);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t wp[4], wn[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = c * 8 + e * 2, n = half * 32 + j;
                            const float2 bp = *(const float2*)(b2p + n), bn = *(const float2*)(b2n + n);
                            wp[e] = tanh_bf16x2(__uint_as_float(dp[j]) + bp.x, __uint_as_float(dp[j + 1]) + bp.y);
                            wn[e] = tanh_bf16x2(__uint_as_float(dn[j]) + bn.x, __uint_as_float(dn[j + 1]) + bn.y);
                        }
                        const int chunk = (half * 4 + c) ^ sw;
                        *(uint4*)(hp_row + chunk * 16) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
                        *(uint4*)(hn_row + chunk * 16) = make_uint4(wn[0], wn[1], wn[2], wn[3]);
                    }
                }
                tc_fence_before();
                fence_async_smem();
                mbar_arrive(&bars[BAR_H2_READY]);
                // ---- epi3: a = tanh(D3 + b3); reward and position ----
                mbar_wait(&bars[BAR_D3_FULL], tp);
                tc_fence_after();
                {
                    uint32_t dp[32], dn[32];
                    tmem_ld32(tm_d3p + lane_base, dp);
                    tmem_ld32(tm_d3n + lane_base, dn);
                    tmem_ld_wait();
                    const int t = m * TC_MT + row;
                    if (t < p.T) {
                        const float* __restrict__ c = p.rew_vec + (size_t)t * p.act;
                        float rp = 0.f, rn = 0.f;
#pragma unroll
                        for (int j = 0; j < TC_ACT_PAD; ++j) {
                            if (j < p.act) {
                                const float ap = tanh_fast(__uint_as_float(dp[j]) + b3p[j]);
                                const float an = tanh_fast(__uint_as_float(dn[j]) + b3n[j]);
                                const float cj = __ldg(c + j);
                                rp = fmaf(ap, cj, rp);
                                rn = fmaf(an, cj, rn);
                                if (j == 0) { pp0 += ap; pn0 += an; }
                                if (j == 1 % p.act) { pp1 += ap; pn1 += an; }
                                if (j == 2 % p.act) { pp2 += ap; pn2 += an; }
                            }
                        }
                        fitp += rp;
                        fitn += rn;
                    }
                }
                tc_fence_before();
                tp ^= 1;
            }
            // ---- per-pair reduction over the 128 epilogue threads ----
            float vals[8] = {fitp, fitn, pp0, pp1, pp2, pn0, pn1, pn2};
#pragma unroll
            for (int i = 0; i < 8; ++i) vals[i] = es_warp_sum(vals[i]);
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) s_red[q * 8 + i] = vals[i];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");           // epilogue warps only
            if (warp == 2 && lane == 0) {
                float tot[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) tot[i] = (s_red[i] + s_red[8 + i]) + (s_red[16 + i] + s_red[24 + i]);
                p.fit_pos[(size_t)pair * p.fit_stride] = (double)tot[0];
                p.fit_neg[(size_t)pair * p.fit_stride] = (double)tot[1];
                if (p.behv_pos) {
                    // act_dim < 3: components repeat (a[j % act]); handled by the j==k%act tests above
                    p.behv_pos[pair * 3 + 0] = p.pos_scale * tot[2]; p.behv_pos[pair * 3 + 1] = p.pos_scale * tot[3];
                    p.behv_pos[pair * 3 + 2] = p.pos_scale * tot[4];
                    p.behv_neg[pair * 3 + 0] = p.pos_scale * tot[5]; p.behv_neg[pair * 3 + 1] = p.pos_scale * tot[6];
                    p.behv_neg[pair * 3 + 2] = p.pos_scale * tot[7];
                }
            }
        }
        // every role advances the per-tile parities identically
        if (p.n_mtiles & 1) { tile_parity ^= 1; d1_free_parity ^= 1; }
        __syncthreads();             // all MMAs of this pair have retired (the epilogue saw the last D3) -> B buffers reusable
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
