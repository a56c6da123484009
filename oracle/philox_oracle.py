"""TEST INFRASTRUCTURE (oracle): numpy restatement of the device-side random draws of csrc/philox.cuh / csrc/rng.cu.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module; the product path never does.

The reference draws its random numbers from torch's global generator (`torch.randperm` U/cleanrl/ppo.py:295,
`Normal.sample` :112-114, `torch.bernoulli` U/mdp/commands.py:75,88 and U/mdp/events.py:71-75), whose stream no other
implementation can reproduce; what CAN be pinned is that the CUDA kernels implement the published Philox4x32-10
algorithm (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) and the documented
mapping from (seed, stream, counter) to every output.  The known-answer vectors of the Random123 distribution
(kat_vectors: philox4x32 10 rounds) pin `philox4x32_10` itself in tests/test_philox.py.
"""

from __future__ import annotations

import numpy as np

STREAM_ACTION_NOISE, STREAM_PERMUTATION, STREAM_BERNOULLI, STREAM_UNIFORM = 0, 1, 2, 3

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10_raw(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32 (10 rounds, Random123 word order)."""
    c = [np.asarray(counter[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[0].astype(np.uint64)
            p1 = _M1 * c[2].astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c[1] ^ k0
            n1 = (p1 & _MASK32).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c[3] ^ k1
            n3 = (p0 & _MASK32).astype(np.uint32)
            c = [n0, n1, n2, n3]
            k0 = k0 + _W0
            k1 = k1 + _W1
    return np.stack(c, axis=-1)


def philox(seed: int, stream: int, ctr) -> np.ndarray:
    """The device's addressing: counter words (ctr lo, ctr hi, stream, 0), key (seed lo, seed hi)."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    counter = np.stack(
        [(ctr & _MASK32).astype(np.uint32), (ctr >> np.uint64(32)).astype(np.uint32), np.full(ctr.shape, stream, np.uint32), np.zeros(ctr.shape, np.uint32)],
        axis=-1,
    )
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32), ctr.shape + (2,))
    return philox4x32_10_raw(counter, key)


def u24(w: np.ndarray) -> np.ndarray:
    """uniform in (0, 1) from the top 24 bits: exact in fp32."""
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24) + np.float32(2.0**-25)


def uniform(seed: int, stream: int, offset: int, n: int) -> np.ndarray:
    return u24(philox(seed, stream, np.uint64(offset) + np.arange(n, dtype=np.uint64))[..., 0])


def normal(seed: int, offset: int, n: int) -> np.ndarray:
    """Box-Muller on words 0, 1 (device: sqrtf/logf/cosf in fp32; here fp32 numpy -> agreement to ~1e-6)."""
    w = philox(seed, STREAM_ACTION_NOISE, np.uint64(offset) + np.arange(n, dtype=np.uint64))
    r = np.sqrt(np.float32(-2.0) * np.log(u24(w[..., 0])))
    return (r * np.cos(np.float32(6.283185307179586) * u24(w[..., 1]))).astype(np.float32)


def feistel_bijection(x: np.ndarray, bits: int, rk: np.ndarray) -> np.ndarray:
    wl, wr = bits // 2, bits - bits // 2
    x = x.astype(np.uint32)
    L, R = x >> np.uint32(wr), x & np.uint32((1 << wr) - 1)
    with np.errstate(over="ignore"):
        for r in range(6):
            f = (R ^ rk[r]) * np.uint32(0x9E3779B1)
            f ^= f >> np.uint32(15)
            f = f * np.uint32(0x85EBCA77)
            f ^= f >> np.uint32(13)
            L, R = R, (L ^ f) & np.uint32((1 << wl) - 1)
            wl, wr = wr, wl
    return (L << np.uint32(wr)) | R


def random_permutation(n: int, seed: int, offset: int) -> np.ndarray:
    """catb200_random_permutation: round keys = Philox blocks (offset, offset + 1) of the permutation stream."""
    bits = 1
    while (1 << bits) < n:
        bits += 1
    rk = philox(seed, STREAM_PERMUTATION, np.array([offset, offset + 1], dtype=np.uint64)).reshape(-1)
    x = np.arange(n, dtype=np.uint32)
    out = feistel_bijection(x, bits, rk)
    todo = out >= n
    while todo.any():  # cycle walking
        out[todo] = feistel_bijection(out[todo], bits, rk)
        todo = out >= n
    return out.astype(np.int64)


def bernoulli_mask(p: np.ndarray, seed: int, offset: int) -> np.ndarray:
    return uniform(seed, STREAM_BERNOULLI, offset, p.shape[0]) < p.astype(np.float32)
