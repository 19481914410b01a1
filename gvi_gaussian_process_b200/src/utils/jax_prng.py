"""Host-side JAX 0.4.13 threefry2x32 PRNG pieces used by the inducing-point selectors, so that a PRNGKey gives
the same permutation as `jax.random.permutation` / `jax.random.choice` in the reference
(src/inducing_points_selection.py:40-46, 115-119).  jax is an un-vendored third-party dependency (pinned at
0.4.13 in the reference's poetry.lock); this follows its published algorithm with the 0.4.13 default
jax_threefry_partitionable=False.  It is checked against the reference's pinned index vectors in tests/.
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
_ROTATIONS = np.array([[13, 15, 26, 6], [17, 29, 16, 24]], dtype=np.uint32)


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def _threefry2x32(key: np.ndarray, count: np.ndarray) -> np.ndarray:
    count = np.ascontiguousarray(count, dtype=_U32).ravel()
    n = count.size
    pad = n & 1
    if pad:
        count = np.append(count, _U32(0))
    lo, hi = np.split(count, 2)
    k0, k1 = _U32(key[0]), _U32(key[1])
    sched = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
    old = np.seterr(over="ignore")
    try:
        a = lo + sched[0]
        b = hi + sched[1]
        for group in range(5):
            for rot in _ROTATIONS[group & 1]:
                a = a + b
                b = (b << rot) | (b >> (_U32(32) - rot))
                b = b ^ a
            a = a + sched[(group + 1) % 3]
            b = b + sched[(group + 2) % 3] + _U32(group + 1)
    finally:
        np.seterr(**old)
    out = np.concatenate([a, b]).astype(_U32)
    return out[:-1] if pad else out


def split(key: np.ndarray, num: int = 2) -> np.ndarray:
    return _threefry2x32(np.asarray(key, dtype=_U32), np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def _bits32(key, n):
    return _threefry2x32(np.asarray(key, dtype=_U32), np.arange(n, dtype=_U32))


def _bits64(key, n):
    raw = _threefry2x32(np.asarray(key, dtype=_U32), np.arange(2 * n, dtype=_U32)).astype(np.uint64)
    return (raw[:n] << np.uint64(32)) | raw[n:]


def permutation(key: np.ndarray, n: int) -> np.ndarray:
    """jax.random.permutation(key, n): ceil(3 ln n / ln(2^32-1)) rounds of a stable sort by fresh 32-bit keys."""
    order = np.arange(n, dtype=np.int64)
    if n <= 1:
        return order
    rounds = int(np.ceil(3.0 * np.log(n) / np.log(float(np.iinfo(np.uint32).max))))
    key = np.asarray(key, dtype=_U32)
    for _ in range(rounds):
        key, sub = split(key)
        order = order[np.argsort(_bits32(sub, n), kind="stable")]
    return order


def choice(key: np.ndarray, n_items: int, k: int) -> np.ndarray:
    """jax.random.choice(key, arange(n_items), (k,)) with replacement (x64 randint recipe)."""
    k1, k2 = split(np.asarray(key, dtype=_U32))
    hi, lo = _bits64(k1, k), _bits64(k2, k)
    span = np.uint64(n_items)
    mult = np.uint64(pow(2**32 % n_items, 2, n_items))
    old = np.seterr(over="ignore")
    try:
        out = ((hi % span) * mult + (lo % span)) % span
    finally:
        np.seterr(**old)
    return out.astype(np.int64)


def permutation_device(key: np.ndarray, n: int, device):
    """`permutation(key, n)` computed on `device` (a CUDA torch device): the threefry hash is one CUDA launch per shuffle
    round (gvi_threefry_random_bits) and the stable key sort is the host framework's device sort; key splitting (a few
    hashes) stays on the host.  Bit-identical to `permutation` (tests/test_gpu_parity.py); at N = 1e7 the host version
    costs seconds of NumPy sorting, this one milliseconds."""
    import torch

    from .. import _lib_proxy as L

    order = torch.arange(n, dtype=torch.int64, device=device)
    if n <= 1:
        return order
    rounds = int(np.ceil(3.0 * np.log(n) / np.log(float(np.iinfo(np.uint32).max))))
    key = np.asarray(key, dtype=_U32)
    for _ in range(rounds):
        key, sub = split(key)
        bits = L.threefry_random_bits(sub, n, device)
        order = order[torch.sort(bits, stable=True).indices]
    return order
