"""Pins the oracle PRNG (oracle/prng.py) with external known answers.

The reference has no test that pins Threefry output bits; the generator lives in
jax>=0.7 which is absent here.  Pins used instead: the Random123 Threefry-2x32-20 KAT vectors and
the values JAX's own documentation prints."""
import numpy as np

from oracle import prng


def test_threefry_random123_kats():
    cases = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
             ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
             ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for k, c, want in cases:
        x0, x1 = prng.threefry2x32(k[0], k[1], c[0], c[1])
        assert (int(x0), int(x1)) == want


def test_split_matches_jax_docs():
    np.testing.assert_array_equal(prng.split(prng.key(0)),
                                  [[1797259609, 2579123966], [928981903, 3453687069]])
    np.testing.assert_array_equal(prng.split_legacy(prng.key(0)),
                                  [[4146024105, 967050713], [2718843009, 1272950319]])


def test_normal_matches_jax_docs():
    assert prng.normal(prng.key(42)) == np.float32(-0.028304616)


def test_key_layout():
    np.testing.assert_array_equal(prng.key(1), [0, 1])
    np.testing.assert_array_equal(prng.key((5 << 32) + 7), [5, 7])


def test_uniform_range_and_bernoulli():
    k = prng.key(3)
    u = prng.uniform(k, 4096)
    assert u.dtype == np.float32 and u.min() >= 0 and u.max() < 1
    assert abs(u.mean() - 0.5) < 0.02
    v = prng.uniform(k, 4096, -2, 2)
    assert v.min() >= -2 and v.max() < 2
    np.testing.assert_array_equal(prng.bernoulli(k, 0.5, 4096), u < 0.5)
    # scalar draw is element 0 of the flat stream
    assert prng.uniform(k) == u[0]


def test_normal_moments():
    x = prng.normal(prng.key(7), 20000)
    assert abs(x.mean()) < 0.03 and abs(x.std() - 1) < 0.03
