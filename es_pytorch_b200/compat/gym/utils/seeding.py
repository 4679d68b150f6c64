"""``gym.utils.seeding`` subset.  gym 0.17.1 hashes the seed (sha512) before seeding a legacy
RandomState; that package is not available to pin the hash, so ``np_random`` here seeds the
RandomState with the seed directly -- the behaviour the reference's own test asserts for the
noise table (test/es/noisetable_test.py:26)."""
import numpy as np


def _int_list_from_bigint(bigint):
    out = []
    while bigint > 0:
        bigint, mod = divmod(bigint, 2 ** 32)
        out.append(mod)
    return out or [0]


def np_random(seed=None):
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(1)[0] % (2 ** 31))
    return np.random.RandomState(seed), seed
