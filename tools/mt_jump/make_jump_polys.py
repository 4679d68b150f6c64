"""Jump-ahead polynomials of MT19937: g(q, m) = x^(624 * m * 16^q) mod phi(x) for the hexadecimal digits m = 1 .. 15 at the
positions q = 0 .. Q-1 of a block count, where phi is the characteristic polynomial (degree 19937) of the generator's state
transition over GF(2).

With g = sum_i g_i x^i, every word of the (raw, untempered) output sequence obeys  x[n + J] = XOR_{i : g_i = 1} x[n + i]
for J = 624 * m * 16^q -- a jump by m * 16^q state blocks is an XOR of 19937-term-window words, independent per output word,
which is what es_pytorch_b200/csrc/mt_gauss.cu (mt_fill_kernel) evaluates on the GPU to start many CTAs at different points
of ONE stream: one jump per non-zero hex digit of the segment's first block (the first version shipped the 18 powers of two
and paid one jump per set BIT).

  phi      Berlekamp-Massey on 2 * 19937 + 64 output bits (bit 31 of the raw word sequence) -> connection polynomial C of
           degree 19937; phi(x) = x^19937 * C(1 / x)
  2^r      x^624 (624 < 19937: no reduction), then repeated squaring mod phi (squaring over GF(2) = spreading the bits)
  g(q, m)  product mod phi of the powers of two in m * 16^q (carry-less multiplication with python integers)

Writes es_pytorch_b200/mt_jump_polys.npy: uint32 [Q][15][624] (bit i of the polynomial = bit i % 32 of word i // 32; bits >=
19937 are zero) and es_pytorch_b200/csrc/mt_jump_polys.inc (the same as C initialisers, [Q * 15][624]).  Checks on the way:
every polynomial with a jump of at most 2^22 words against numpy's own MT19937 (words generated sequentially), and every
g(q, m) with q >= 1 against four squarings of g(q - 1, m).  Pure python big-int arithmetic, about a minute; a sample is
checked again by tests/test_host_logic.py."""
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


def gf2_mul(a: int, b: int) -> int:
    """Carry-less product a(x) b(x) over GF(2)."""
    r = 0
    while b:
        low = b & -b
        r ^= a << (low.bit_length() - 1)
        b ^= low
    return r


def untemper(y: np.ndarray) -> np.ndarray:
    """Inverse of MT19937's output tempering (vectorised): the raw state words behind output words."""
    y = y.astype(np.uint64)
    y ^= y >> np.uint64(18)
    y ^= (y << np.uint64(15)) & np.uint64(0xEFC60000)
    t = y.copy()
    for _ in range(4):
        t = y ^ ((t << np.uint64(7)) & np.uint64(0x9D2C5680))
    y = t & np.uint64(0xFFFFFFFF)
    t = y.copy()
    for _ in range(2):
        t = y ^ (t >> np.uint64(11))
    return (t & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def numpy_raw_words(seed: int, count: int) -> np.ndarray:
    """`count` raw words of numpy's legacy RandomState(seed), from its first regenerated block on."""
    rs = np.random.RandomState(seed)
    return untemper(rs.randint(0, 2 ** 32, size=count, dtype=np.uint64))


def check_against_numpy(g: int, blocks: int, raw: np.ndarray, starts=(0, 1, 397, 623, 1000)) -> bool:
    """x[n + 624 * blocks] == XOR_{i : g_i = 1} x[n + i] on numpy's word sequence."""
    idx = np.array([i for i in range(DEG) if (g >> i) & 1], dtype=np.int64)
    J = N * blocks
    for n in starts:
        if n + J >= len(raw) or n + DEG >= len(raw):
            return False
        if int(np.bitwise_xor.reduce(raw[idx + n])) != int(raw[n + J]):
            return False
    return True


Q = 5                                       # hex digit positions: jumps of up to 16^5 = 2^20 blocks


def main():
    seq = raw_words(12345, 2 * DEG + 64 + N)
    bits = [(int(w) >> 31) & 1 for w in seq[N:]]           # (skip the first block: its words still depend on the seed words' low bits)
    C, L = berlekamp_massey(bits[:2 * DEG + 64])
    assert L == DEG, L
    phi = sum(((C >> i) & 1) << (DEG - i) for i in range(DEG + 1))
    assert phi.bit_length() == DEG + 1 and phi & 1
    pow2 = []
    g = 1 << N                                             # x^624
    for r in range(4 * Q):
        pow2.append(g)
        g = gf2_mod(gf2_square(g), phi, DEG)
    raw = numpy_raw_words(2024, (1 << 22) + 2 * DEG + 2048)
    polys, checked = {}, 0
    for q in range(Q):
        for m in range(1, 16):
            g = None
            for b in range(4):
                if (m >> b) & 1:
                    g = pow2[4 * q + b] if g is None else gf2_mod(gf2_mul(g, pow2[4 * q + b]), phi, DEG)
            polys[(q, m)] = g
            blocks = m * 16 ** q
            if N * blocks + DEG + 1024 < len(raw):
                assert check_against_numpy(g, blocks, raw), (q, m)
                checked += 1
            if q >= 1:
                h = polys[(q - 1, m)]
                for _ in range(4):
                    h = gf2_mod(gf2_square(h), phi, DEG)
                assert h == g, (q, m)
    out = np.zeros((Q, 15, N), dtype=np.uint32)
    for (q, m), g in polys.items():
        for w in range(N):
            out[q, m - 1, w] = (g >> (32 * w)) & 0xFFFFFFFF
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    path = os.path.join(root, 'es_pytorch_b200', 'mt_jump_polys.npy')
    np.save(path, out)
    inc = os.path.join(root, 'es_pytorch_b200', 'csrc', 'mt_jump_polys.inc')
    with open(inc, 'w') as f:
        f.write('// generated by tools/mt_jump/make_jump_polys.py: x^(624 * m * 16^q) mod phi, [q * 15 + m - 1][624], bit i = bit i % 32 of word i / 32\n')
        for q in range(Q):
            for m in range(1, 16):
                f.write('{' + ','.join('0x%08xu' % int(v) for v in out[q, m - 1]) + '},\n')
    print(path, out.shape, 'checked against numpy:', checked, 'weights', [bin(polys[(0, m)]).count('1') for m in (1, 2, 3)], '...')


if __name__ == '__main__':
    main()
