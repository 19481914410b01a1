"""NumPy restatement of the JAX 0.4.13 threefry PRNG pieces behind ConditionalVarianceInducingPointsSelector
(TEST INFRASTRUCTURE + host-side helper source of truth).

Third-party algorithm: jax==0.4.13 / jaxlib==0.4.13 (poetry.lock:1078-1079, 1134-1135) is not vendored under
/root/reference; this restates its published algorithm (jax/_src/prng.py threefry2x32, jax/_src/random.py
_shuffle/permutation/randint, default jax_threefry_partitionable=False) and is anchored on the reference's own
call sites (src/inducing_points_selection.py:115-119, :40-46) and pinned index vectors
(tests/test_inducing_points_selection.py:128,139,151,162).
"""
from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def threefry_2x32(key, counts):
    key = np.asarray(key, dtype=np.uint32)
    counts = np.asarray(counts, dtype=np.uint32).reshape(-1)
    n = counts.size
    odd = n % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, np.uint32)])
    half = counts.size // 2
    x0, x1 = counts[:half].copy(), counts[half:].copy()
    ks = [key[0], key[1], np.uint32(key[0] ^ key[1] ^ np.uint32(0x1BD11BDA))]
    with np.errstate(over="ignore"):
        x0 = (x0 + ks[0]).astype(np.uint32)
        x1 = (x1 + ks[1]).astype(np.uint32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(i + 2) % 3] + np.uint32(i + 1)).astype(np.uint32)
    out = np.concatenate([x0, x1])
    return out[:n] if odd else out


def prng_key(seed: int):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num: int = 2):
    return threefry_2x32(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)


def random_bits32(key, n: int):
    return threefry_2x32(key, np.arange(n, dtype=np.uint32))


def random_bits64(key, n: int):
    bits = threefry_2x32(key, np.arange(2 * n, dtype=np.uint32))
    return (bits[:n].astype(np.uint64) << np.uint64(32)) | bits[n:].astype(np.uint64)


def permutation(key, n: int):
    """jax.random.permutation(key, n): rounds of stable sort by fresh 32-bit keys."""
    x = np.arange(n)
    if n <= 1:
        return x
    num_rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    for _ in range(num_rounds):
        key, sub = split(key)
        sort_keys = random_bits32(sub, n)
        x = x[np.argsort(sort_keys, kind="stable")]
    return x


def randint(key, n: int, minval: int, maxval: int):
    """jax.random.randint under x64 (int64 dtype): two 64-bit draws combined modulo the span."""
    k1, k2 = split(key)
    hi = random_bits64(k1, n)
    lo = random_bits64(k2, n)
    span = np.uint64(maxval - minval)
    mult = np.uint64((int(2**32) % int(span)) ** 2 % int(span))
    with np.errstate(over="ignore"):
        off = ((hi % span) * mult + (lo % span)) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int64)


def choice_with_replacement(key, n_items: int, k: int):
    """jax.random.choice(key, arange(N), (k,)) with the default replace=True."""
    return randint(key, k, 0, n_items)
