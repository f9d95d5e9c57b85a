"""Builds the committed fixtures of tests/golden/.

1. reference_known_answers.json -- the known answers the REFERENCE's own tests hold for the NUTS/HMC path (restated by hand
   from /root/reference, which cannot be imported in this image: jax is absent) plus the external pins of the PRNG that
   lives in the un-vendored jax dependency.  Every entry names its source.
2. oracle_vectors.npz -- streams produced by the CPU oracle AFTER it passed (1): Threefry bit / uniform / normal streams,
   key splits, det-f32 math tables and a deterministic warm-up adaptation script.  The CUDA path must reproduce them bit for
   bit (tests/test_golden.py, GPU part), and the oracle must keep reproducing them (CPU part: guards the oracle itself).

Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import adapt, detmath as dm, prng          # noqa: E402

F = np.float32

KNOWN = {
    "threefry2x32_20_kat": {
        "source": "Random123 kat_vectors (threefry2x32, 20 rounds); jax.random's generator (jax>=0.7, un-vendored)",
        "cases": [{"key": [0, 0], "ctr": [0, 0], "out": [0x6B200159, 0x99BA4EFE]},
                  {"key": [0xFFFFFFFF, 0xFFFFFFFF], "ctr": [0xFFFFFFFF, 0xFFFFFFFF], "out": [0x1CB996FC, 0xBB002BE7]},
                  {"key": [0x13198A2E, 0x03707344], "ctr": [0x243F6A88, 0x85A308D3], "out": [0xC4923A9C, 0x483DF7A0]}]},
    "split_key0": {"source": "jax.random documentation (partitionable threefry, the jax>=0.5 default) / legacy mode",
                   "partitionable": [[1797259609, 2579123966], [928981903, 3453687069]],
                   "legacy": [[4146024105, 967050713], [2718843009, 1272950319]]},
    "normal_key42": {"source": "jax.random documentation", "partitionable": -0.028304616},
    "adaptation_schedule": {
        "source": "test/infer/test_hmc_util.py:278-292 (build_adaptation_schedule) + SURVEY.md row a11 (1000 steps, hand-verified)",
        "cases": {"18": [[0, 17]], "50": [[0, 6], [7, 44], [45, 49]], "100": [[0, 14], [15, 89], [90, 99]],
                  "150": [[0, 74], [75, 99], [100, 149]], "200": [[0, 74], [75, 99], [100, 149], [150, 199]],
                  "280": [[0, 74], [75, 99], [100, 229], [230, 279]],
                  "1000": [[0, 74], [75, 99], [100, 149], [150, 249], [250, 449], [450, 949], [950, 999]]}},
    "leaf_idx_to_ckpt_idxs": {"source": "test/infer/test_hmc_util.py:381-386",
                              "cases": {"0": [1, 0], "6": [3, 2], "7": [0, 2], "13": [2, 2], "15": [0, 3]}},
    "is_iterative_turning": {
        "source": "test/infer/test_hmc_util.py:389-403 (inverse mass 1, r = 1, r_sum = 3, r_ckpts [1,2,3,-2], r_sum_ckpts [2,4,4,-1])",
        "cases": [[[3, 2], False], [[3, 3], True], [[0, 0], False], [[0, 1], True], [[1, 3], True]]},
    "diagnostics": {"source": "test/test_diagnostics.py:60-108",
                    "autocorrelation_arange10": [1, 0.78, 0.52, 0.21, -0.13, -0.52, -0.94, -1.4, -1.91, -2.45],
                    "gelman_rubin_two_shifted_aranges": 0.98, "ess_arange1000_100x10": 52.64, "atol": 0.01},
    "readme_eight_schools_noncentered": {
        "source": "README.md:118-143 (1 chain, 500 warm-up, 1000 samples): statistical pin, checked within Monte-Carlo error",
        "mu_mean": 4.08, "tau_mean": 3.96, "divergences": 0},
}


def oracle_vectors():
    out = {}
    keys = [prng.key(0), prng.key(42), prng.key((7 << 32) + 11)]
    for i, k in enumerate(keys):
        out[f"key{i}"] = np.asarray(k, np.uint32)
        out[f"bits{i}"] = prng.random_bits(k, 257)
        out[f"uniform{i}"] = prng.uniform(k, 257)
        out[f"uniform_m2_2_{i}"] = prng.uniform(k, 257, -2.0, 2.0)
        out[f"normal{i}"] = prng.normal(k, 257)
        out[f"split5_{i}"] = prng.split(k, 5)
    # det-f32 math tables (shared explicit algorithms, csrc/detmath.cuh <-> oracle/detmath.py)
    rng = np.random.default_rng(2024)
    x = np.concatenate([rng.normal(size=300) * 8, [0.0, -0.0, 1.0, -1.0, 88.0, -88.0, 1e-8, -1e-8]]).astype(F)
    out["dm_x"] = x
    out["dm_exp"] = np.array([dm.exp(v) for v in x], F)
    out["dm_expit"] = np.array([dm.expit(v) for v in x], F)
    pos = np.abs(x) + F(1e-6)
    out["dm_pos"] = pos
    out["dm_log"] = np.array([dm.log(v) for v in pos], F)
    out["dm_log1p"] = np.array([dm.log1p(v) for v in pos], F)
    u = np.clip(x / 100.0, -0.999999, 0.999999).astype(F)
    out["dm_u"] = u
    out["dm_erfinv"] = np.array([dm.erfinv(v) for v in u], F)
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
        json.dump(KNOWN, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **oracle_vectors())
    print("wrote", os.listdir(HERE))
