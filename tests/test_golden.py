"""Committed fixtures (tests/golden/, built by tests/golden/make_golden.py).

CPU part: the oracle against the reference's own known answers + external PRNG pins (reference_known_answers.json) and
against its own recorded streams (oracle_vectors.npz: guards the oracle against silent changes).
GPU part: the CUDA engine's PRNG and det-f32 math hooks (through the C ABI) against the recorded streams, bit for bit."""
import json
import os

import numpy as np
import pytest

from oracle import adapt, detmath as dm, diag, prng, tree

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F = np.float32


@pytest.fixture(scope="module")
def known():
    with open(os.path.join(HERE, "reference_known_answers.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def vec():
    return dict(np.load(os.path.join(HERE, "oracle_vectors.npz")))


def test_oracle_against_reference_known_answers(known):
    for c in known["threefry2x32_20_kat"]["cases"]:
        x0, x1 = prng.threefry2x32(c["key"][0], c["key"][1], c["ctr"][0], c["ctr"][1])
        assert [int(x0), int(x1)] == c["out"]
    np.testing.assert_array_equal(prng.split(prng.key(0)), known["split_key0"]["partitionable"])
    np.testing.assert_array_equal(prng.split_legacy(prng.key(0)), known["split_key0"]["legacy"])
    assert prng.normal(prng.key(42)) == F(known["normal_key42"]["partitionable"])
    for n, want in known["adaptation_schedule"]["cases"].items():
        assert [list(w) for w in adapt.build_adaptation_schedule(int(n))] == want
    for leaf, want in known["leaf_idx_to_ckpt_idxs"]["cases"].items():
        assert list(tree.leaf_idx_to_ckpt_idxs(int(leaf))) == want
    imm = np.ones(1, F)
    r_ckpts = np.array([[1.0], [2.0], [3.0], [-2.0]], F)
    r_sum_ckpts = np.array([[2.0], [4.0], [4.0], [-1.0]], F)
    for (lo, hi), want in known["is_iterative_turning"]["cases"]:
        assert tree.is_iterative_turning(imm, np.array([1.0], F), np.array([3.0], F), r_ckpts, r_sum_ckpts, lo, hi) == want
    d = known["diagnostics"]
    np.testing.assert_allclose(diag.autocorrelation(np.arange(10.0), bias=False), d["autocorrelation_arange10"], atol=d["atol"])
    y = np.stack([np.arange(10.0), np.arange(10.0) + 1])
    np.testing.assert_allclose(diag.gelman_rubin(y), d["gelman_rubin_two_shifted_aranges"], atol=d["atol"])
    np.testing.assert_allclose(diag.effective_sample_size(np.arange(1000.0).reshape(100, 10), bias=False),
                               d["ess_arange1000_100x10"], atol=d["atol"])


def test_oracle_reproduces_its_recorded_streams(vec):
    for i in range(3):
        k = vec[f"key{i}"]
        np.testing.assert_array_equal(prng.random_bits(k, 257), vec[f"bits{i}"])
        np.testing.assert_array_equal(prng.uniform(k, 257), vec[f"uniform{i}"])
        np.testing.assert_array_equal(prng.uniform(k, 257, -2.0, 2.0), vec[f"uniform_m2_2_{i}"])
        np.testing.assert_array_equal(prng.normal(k, 257), vec[f"normal{i}"])
        np.testing.assert_array_equal(prng.split(k, 5), vec[f"split5_{i}"])
    np.testing.assert_array_equal(np.array([dm.exp(v) for v in vec["dm_x"]], F), vec["dm_exp"])
    np.testing.assert_array_equal(np.array([dm.erfinv(v) for v in vec["dm_u"]], F), vec["dm_erfinv"])


@pytest.mark.gpu
def test_cuda_reproduces_the_recorded_streams(vec):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from numpyro_b200 import engine as eng
    for i in range(3):
        k = vec[f"key{i}"]
        np.testing.assert_array_equal(eng.prng_bits(k, 257), vec[f"bits{i}"])
        np.testing.assert_array_equal(eng.prng_uniform(k, 257, 0.0, 1.0), vec[f"uniform{i}"])
        np.testing.assert_array_equal(eng.prng_uniform(k, 257, -2.0, 2.0), vec[f"uniform_m2_2_{i}"])
        np.testing.assert_array_equal(eng.prng_normal(k, 257), vec[f"normal{i}"])
        np.testing.assert_array_equal(eng.prng_split(k[None], 5)[0], vec[f"split5_{i}"])
    for op, x, want in ((0, "dm_x", "dm_exp"), (3, "dm_x", "dm_expit"), (1, "dm_pos", "dm_log"), (2, "dm_pos", "dm_log1p"),
                        (4, "dm_u", "dm_erfinv")):
        np.testing.assert_array_equal(eng.detmath(op, vec[x]), vec[want])
