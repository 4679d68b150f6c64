"""``gym.utils.seeding`` subset.

``np_random(seed)`` seeds a legacy RandomState with the seed directly -- the behaviour the reference's own test asserts for
the noise table (test/es/noisetable_test.py:26).  gym 0.17.1 (the reference's pinned version, env.yml:57) instead hashes the
seed first; ``np_random(seed, hashed=True)`` restates that path as recalled from gym 0.17.1's ``gym/utils/seeding.py``
(sha512 of ``str(seed)``, first 8 bytes as little-endian uint32 words, ``RandomState.seed(list_of_words)``).  The package is
not available offline, so that variant is UNPINNED: no golden vector could be generated for it."""
import hashlib
import os
import struct

import numpy as np


def _bigint_from_bytes(data: bytes) -> int:
    sizeof_int = 4
    padding = sizeof_int - len(data) % sizeof_int            # (gym pads a full word when the length is already a multiple)
    data += b'\0' * padding
    words = struct.unpack('{}I'.format(len(data) // sizeof_int), data)
    return sum(2 ** (sizeof_int * 8 * i) * w for i, w in enumerate(words))


def _int_list_from_bigint(bigint):
    out = []
    while bigint > 0:
        bigint, mod = divmod(bigint, 2 ** 32)
        out.append(mod)
    return out or [0]


def create_seed(a=None, max_bytes=8):
    if a is None:
        return _bigint_from_bytes(os.urandom(max_bytes))
    if isinstance(a, str):
        a = a.encode('utf8') + hashlib.sha512(a.encode('utf8')).digest()
        return _bigint_from_bytes(a[:max_bytes])
    if isinstance(a, int):
        return a % 2 ** (8 * max_bytes)
    raise TypeError('Invalid type for seed: {} ({})'.format(type(a), a))


def hash_seed(seed=None, max_bytes=8):
    if seed is None:
        seed = create_seed(max_bytes=max_bytes)
    return _bigint_from_bytes(hashlib.sha512(str(seed).encode('utf8')).digest()[:max_bytes])


def np_random(seed=None, hashed: bool = False):
    if hashed:
        seed = create_seed(seed)
        rng = np.random.RandomState()
        rng.seed(_int_list_from_bigint(hash_seed(seed)))
        return rng, seed
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(1)[0] % (2 ** 31))
    return np.random.RandomState(seed), seed
