// mt19937.cuh -- pieces of the MT19937 generator shared by mt_gauss.cu and mt_jump.cu (device code, sm_100a).
#pragma once
#include <stdint.h>

constexpr int MT_NW = 624, MT_MW = 397, MT_DW = MT_NW - MT_MW;      // state words, middle word, 227

__device__ __forceinline__ uint32_t mt19937_twist(uint32_t u, uint32_t v) {
    const uint32_t y = (u & 0x80000000u) | (v & 0x7FFFFFFFu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}
__device__ __forceinline__ uint32_t mt19937_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return y;
}
// the inverse of the tempering (a bijection of 32-bit words): the raw state word behind an output word
__device__ __forceinline__ uint32_t mt19937_untemper(uint32_t y) {
    y ^= y >> 18;
    y ^= (y << 15) & 0xEFC60000u;
    uint32_t t = y;                                   // undo y ^= (y << 7) & B: 7 bits are recovered per round
    t = y ^ ((t << 7) & 0x9D2C5680u);
    t = y ^ ((t << 7) & 0x9D2C5680u);
    t = y ^ ((t << 7) & 0x9D2C5680u);
    t = y ^ ((t << 7) & 0x9D2C5680u);
    y = t;
    t = y;                                            // undo y ^= y >> 11: 11 bits per round
    t = y ^ (t >> 11);
    t = y ^ (t >> 11);
    return t;
}

// Thread <-> word map of the two-barrier block regeneration (needs >= 705 threads).  The recurrence
// N[i] = N[i-227] ^ tw(O[i], O[i+1]) is XOR-linear in its first term, so with the 623 twists T[k] = tw(O[k], O[k+1]) of the OLD
// block (one per thread, exchanged through shared memory) every new word is a few XORs of old words and twists:
//   i <  227:  N[i] = O[i+397] ^ T[i]
//   i <  454:  N[i] = O[i+170] ^ T[i-227] ^ T[i]
//   i <  623:  N[i] = O[i-57]  ^ T[i-454] ^ T[i-227] ^ T[i]
//   N[623] = N[396] ^ tw(O[623], N[0])
// Every warp stays inside one range and all three ranges run the SAME code (an unused twist index points at T[623], which holds
// 0); the odd word out has a warp of its own: threads 0-226 | 256-482 | 512-680 | 704.
struct Mt19937Regen {
    int my_i, r_o, r_a, r_b, r_c;
    bool plain;
    __device__ __forceinline__ void init(int tid) {
        my_i = (tid < 256) ? (tid < MT_DW ? tid : -1) : (tid < 512) ? (tid - 256 < MT_DW ? tid - 256 + MT_DW : -1)
             : (tid < 704) ? (tid - 512 < MT_NW - 1 - 2 * MT_DW ? tid - 512 + 2 * MT_DW : -1) : (tid == 704 ? MT_NW - 1 : -1);
        plain = my_i >= 0 && my_i < MT_NW - 1;
        r_o = !plain ? 0 : (my_i < MT_DW ? my_i + MT_MW : my_i < 2 * MT_DW ? my_i + MT_MW - MT_DW : my_i + MT_MW - 2 * MT_DW);
        r_a = plain ? my_i : MT_NW - 1;
        r_b = (plain && my_i >= MT_DW) ? my_i - MT_DW : MT_NW - 1;
        r_c = (plain && my_i >= 2 * MT_DW) ? my_i - 2 * MT_DW : MT_NW - 1;
    }
    // the new word of this thread (valid when my_i >= 0) after the twists of O are in T (T[623] == 0); call between two barriers
    __device__ __forceinline__ uint32_t word(const uint32_t* __restrict__ O, const uint32_t* __restrict__ T) const {
        if (plain) return O[r_o] ^ T[r_a] ^ T[r_b] ^ T[r_c];
        const uint32_t n0 = O[MT_MW] ^ T[0];
        const uint32_t n396 = O[396 + MT_MW - MT_DW] ^ T[396 - MT_DW] ^ T[396];
        return n396 ^ mt19937_twist(O[MT_NW - 1], n0);
    }
};
