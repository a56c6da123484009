"""Philox4x32-10 and the draws built on it (csrc/philox.cuh, csrc/rng.cu).

CPU part: the C++ block function (compiled host-side from the very header the kernels inline) and the numpy oracle
both reproduce the known-answer vectors published with Random123 (kat_vectors, "philox4x32 10"); the Feistel
permutation of the C++ host path equals the oracle's.  GPU part: kernels == oracle bit for bit for every integer
output (permutation, Bernoulli mask and its index list), 1e-5 for the Box-Muller normals (device libm vs numpy)."""

import ctypes

import numpy as np
import pytest
import torch

from oracle import philox_oracle as po

# counter words, key words, expected output (Random123 kat_vectors)
KATS = [
    ("00000000 00000000 00000000 00000000", "00000000 00000000", "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
    ("ffffffff ffffffff ffffffff ffffffff", "ffffffff ffffffff", "408f276d 41c83b0e a20bc7c6 6d5451fd"),
    ("243f6a88 85a308d3 13198a2e 03707344", "a4093822 299f31d0", "d16cfe09 94fdcceb 5001e420 24126ea1"),
]


def _words(text):
    return np.array([int(x, 16) for x in text.split()], dtype=np.uint32)


@pytest.mark.parametrize("ctr,key,want", KATS)
def test_philox_known_answers(lib, ctr, key, want):
    c, k, w = _words(ctr), _words(key), _words(want)
    assert np.array_equal(po.philox4x32_10_raw(c[None], k[None])[0], w)
    out = (ctypes.c_uint32 * 4)()
    assert lib.catb200_philox4x32_10(c.ctypes.data, k.ctypes.data, ctypes.addressof(out)) == 0
    assert np.array_equal(np.array(out[:], dtype=np.uint32), w)


@pytest.mark.parametrize("n", [1, 2, 3, 1000, 4096 * 24, 100003])
def test_permutation_host_path_equals_oracle_and_is_a_permutation(lib, n):
    out = np.empty(n, dtype=np.int64)
    assert lib.catb200_random_permutation_host(n, 77, 5, out.ctypes.data) == 0
    assert np.array_equal(out, po.random_permutation(n, 77, 5))
    assert np.array_equal(np.sort(out), np.arange(n))


def test_permutation_quality():
    """Not a sort of random keys, but close enough to uniform for minibatch shuffling: positions decorrelated from
    values, every minibatch-sized slice covers the index range evenly, different offsets give different shuffles."""
    n, mb = 4096 * 24, 16384
    a, b = po.random_permutation(n, 1, 0), po.random_permutation(n, 1, 2)
    assert abs(np.corrcoef(np.arange(n), a)[0, 1]) < 0.02
    assert (a == b).mean() < 0.001
    for s in range(0, n, mb):
        hist = np.bincount(a[s : s + mb] * 16 // n, minlength=16)
        assert hist.min() > 0.85 * mb / 16 and hist.max() < 1.15 * mb / 16
    assert abs(float(a[:mb].mean()) - n / 2) < 0.02 * n


def test_uniform_and_normal_moments():
    u = po.uniform(3, po.STREAM_UNIFORM, 0, 200000)
    assert 0.0 < u.min() and u.max() < 1.0 and abs(float(u.mean()) - 0.5) < 3e-3
    z = po.normal(3, 0, 200000)
    assert abs(float(z.mean())) < 1e-2 and abs(float(z.std()) - 1.0) < 1e-2


@pytest.mark.gpu
def test_device_draws_match_the_oracle():
    from constraints_as_terminations_b200 import ops

    dev = "cuda:0"
    seed = 987654321
    rng = ops.make_rng_state(seed, dev)
    n = 4096 * 24
    perm = ops.random_permutation(n, rng)
    assert torch.equal(perm.cpu(), torch.from_numpy(po.random_permutation(n, seed, 0)))
    assert rng.cpu().tolist() == [seed, 2]
    perm2 = ops.random_permutation(1000, rng)  # the offset advanced on the device: a different draw
    assert torch.equal(perm2.cpu(), torch.from_numpy(po.random_permutation(1000, seed, 2)))
    # Bernoulli mask + ascending index list
    g = torch.Generator().manual_seed(0)
    p = torch.rand(5000, generator=g) * 0.3
    p[::7] = 0.0
    p[3::11] = 1.0
    rng = ops.make_rng_state(seed, dev, offset=100)
    mask, ids, count = ops.bernoulli_mask(p.to(dev), rng, with_ids=True)
    want = po.bernoulli_mask(p.numpy(), seed, 100)
    assert np.array_equal(mask.cpu().numpy(), want)
    assert int(count) == int(want.sum())
    assert np.array_equal(ids[: int(count)].cpu().numpy(), np.nonzero(want)[0])
    assert rng.cpu().tolist() == [seed, 100 + 5000]
    assert not mask.cpu()[p == 0.0].any() and mask.cpu()[p == 1.0].all()
