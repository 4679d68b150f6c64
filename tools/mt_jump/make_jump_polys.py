"""Jump-ahead polynomials of MT19937: g_r(x) = x^(624 * 2^r) mod phi(x) for r = 0 .. R-1, where phi is the characteristic
polynomial (degree 19937) of the generator's state transition over GF(2).

With g = sum_i g_i x^i, every word of the (raw, untempered) output sequence obeys  x[n + J] = XOR_{i : g_i = 1} x[n + i]
for J = 624 * 2^r -- a jump by 2^r state blocks is an XOR of 19937-term-window words, independent per output word, which is
what es_pytorch_b200/csrc/mt_jump.cu evaluates on the GPU to start many CTAs at different points of ONE stream.

  phi    Berlekamp-Massey on 2 * 19937 + 64 output bits (bit 31 of the raw word sequence) -> connection polynomial C of degree
         19937; phi(x) = x^19937 * C(1 / x)
  g_0    x^624 (624 < 19937: no reduction);  g_r = g_(r-1)^2 mod phi  (squaring over GF(2) = spreading the bits)

Writes es_pytorch_b200/mt_jump_polys.npy: uint32 [R][624] (bit i of the polynomial = bit i % 32 of word i // 32; bits >= 19937
are zero).  Pure Python big-int arithmetic, about a minute.  Checked by tests/test_host_logic.py against sequentially generated
words (numpy's own MT19937)."""
import os
import sys

import numpy as np

N, M, DEG = 624, 397, 19937
R = 18                                      # jumps of 1 .. 2^17 blocks (2^17 * 624 = 81.8 M words)


def raw_words(seed: int, count: int) -> np.ndarray:
    """`count` raw (untempered) state words following the seeded state: block after block of the recurrence."""
    key = np.random.RandomState(seed).get_state()[1].astype(np.uint64)
    out = []
    mt = [int(v) for v in key]
    while len(out) < count:
        for i in range(N):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % N] & 0x7FFFFFFF)
            mt[i] = mt[(i + M) % N] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        out.extend(mt)
    return np.array(out[:count], dtype=np.uint64)


def berlekamp_massey(bits) -> int:
    """Connection polynomial C (int, bit i = c_i, c_0 = 1) of the shortest LFSR generating `bits`: s_n = XOR_i>=1 c_i s_(n-i)."""
    C, B, L, m = 1, 1, 0, 1
    window = 0                               # bit i = s_(n-i)
    for n, s in enumerate(bits):
        window = (window << 1) | int(s)
        d = bin(C & window).count('1') & 1
        if d:
            T = C
            C ^= B << m
            if 2 * L <= n:
                L, B, m = n + 1 - L, T, 1
            else:
                m += 1
        else:
            m += 1
    return C, L


def gf2_square(a: int) -> int:
    """a(x)^2 over GF(2): bit i -> bit 2 i."""
    s = bin(a)[2:]
    return int('0'.join(s), 2)


def gf2_mod(a: int, phi: int, deg: int) -> int:
    while a.bit_length() > deg:
        a ^= phi << (a.bit_length() - 1 - deg)
    return a


def main():
    seq = raw_words(12345, 2 * DEG + 64 + N)
    bits = [(int(w) >> 31) & 1 for w in seq[N:]]           # (skip the first block: its words still depend on the seed words' low bits)
    C, L = berlekamp_massey(bits[:2 * DEG + 64])
    assert L == DEG, L
    phi = int(bin(C)[2:].zfill(DEG + 1)[::-1], 2) if False else sum(((C >> i) & 1) << (DEG - i) for i in range(DEG + 1))
    assert phi.bit_length() == DEG + 1 and phi & 1
    polys = []
    g = 1 << N                                             # x^624
    for r in range(R):
        polys.append(g)
        g = gf2_mod(gf2_square(g), phi, DEG)
    out = np.zeros((R, N), dtype=np.uint32)
    for r, g in enumerate(polys):
        for w in range(N):
            out[r, w] = (g >> (32 * w)) & 0xFFFFFFFF
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'es_pytorch_b200',
                        'mt_jump_polys.npy')
    np.save(path, out)
    print(path, out.shape, 'weights', [bin(p).count('1') for p in polys[:4]], '...')


if __name__ == '__main__':
    main()
