"""CPU: the oracle against the committed golden vectors (produced from the unmodified
reference by oracle/make_golden.py) and against properties of the operator."""
import ast
import glob
import os

import numpy as np
import pytest

from oracle import cosine_sim_attention_oracle as oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load_case(path):
    z = np.load(path)
    kw = ast.literal_eval(str(z["kwargs"]))
    mask = z["mask"] if z["mask"].size else None
    return z, kw, mask


def test_golden_files_present():
    assert len(GOLDEN) >= 11


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    z, kw, mask = load_case(path)
    bias = z["attn_bias"] if "attn_bias" in z.files else None
    res = oracle.attention(z["q"], z["k"], z["v"], mask=mask, attn_bias=bias, d_out=z["d_out"], **kw)
    names = ("o", "dq", "dk", "dv") + (("d_bias",) if bias is not None else ())
    assert len(res) == len(names)
    for name, got in zip(names, res):
        assert got.shape == z[name].shape
        assert np.abs(got - z[name]).max() < 1e-9, name


def _rand(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape)


def test_rows_are_convex_combinations_of_values():
    q, k = _rand((1, 2, 17, 64), 0), _rand((1, 2, 29, 64), 1)
    v = np.ones((1, 2, 29, 64))
    o = oracle.attention(q, k, v)
    assert np.allclose(o, 1.0, atol=1e-12)


def test_linear_in_values():
    q, k = _rand((1, 2, 20, 64), 2), _rand((1, 2, 33, 64), 3)
    v1, v2 = _rand((1, 2, 33, 64), 4), _rand((1, 2, 33, 64), 5)
    a = oracle.attention(q, k, 2 * v1 - 3 * v2, causal=True)
    b = 2 * oracle.attention(q, k, v1, causal=True) - 3 * oracle.attention(q, k, v2, causal=True)
    assert np.allclose(a, b, atol=1e-12)


def test_scale_invariance_of_normalised_inputs():
    q, k, v = _rand((1, 1, 9, 64), 6), _rand((1, 1, 11, 64), 7), _rand((1, 1, 11, 64), 8)
    assert np.allclose(oracle.attention(q, k, v, groups=2), oracle.attention(5 * q, 0.1 * k, v, groups=2), atol=1e-12)


def test_causal_is_bottom_right_aligned():
    # with Nq < Nk the first query already sees Nk - Nq + 1 keys (reference py:112-115)
    q, k, v = _rand((1, 1, 3, 64), 9), _rand((1, 1, 7, 64), 10), _rand((1, 1, 7, 64), 11)
    full = oracle.attention(q, k, v, causal=True)
    first = oracle.attention(q[:, :, :1], k[:, :, :5], v[:, :, :5])
    assert np.allclose(full[:, :, :1], first, atol=1e-12)


def test_key_mask_equals_dropping_keys():
    q, k, v = _rand((2, 2, 6, 64), 12), _rand((2, 2, 10, 64), 13), _rand((2, 2, 10, 64), 14)
    mask = np.ones((2, 10), dtype=bool)
    mask[:, 7:] = False
    assert np.allclose(oracle.attention(q, k, v, mask=mask), oracle.attention(q, k[:, :, :7], v[:, :, :7]), atol=1e-12)


def test_empty_rows_mean_vs_zero():
    q, k, v = _rand((1, 1, 4, 64), 15), _rand((1, 1, 5, 64), 16), _rand((1, 1, 5, 64), 17)
    mask = np.zeros((1, 5), dtype=bool)
    assert np.allclose(oracle.attention(q, k, v, mask=mask, empty_rows="zero"), 0.0)
    assert np.allclose(oracle.attention(q, k, v, mask=mask, empty_rows="mean"), v.mean(-2, keepdims=True))


def test_gradients_match_finite_differences():
    rng = np.random.default_rng(18)
    q, k, v = rng.standard_normal((1, 2, 5, 64)), rng.standard_normal((1, 7, 64)), rng.standard_normal((1, 7, 64))
    do = rng.standard_normal((1, 2, 5, 64))
    kw = dict(causal=True, groups=2, scale=3.0)
    _, dq, dk, dv = oracle.attention(q, k, v, d_out=do, **kw)
    f = lambda q_, k_, v_: (oracle.attention(q_, k_, v_, **kw) * do).sum()
    eps = 1e-6
    for arr, grad, idx in ((q, dq, (0, 1, 2, 5)), (k, dk, (0, 3, 9)), (v, dv, (0, 6, 1))):
        p, m = arr.copy(), arr.copy()
        p[idx] += eps
        m[idx] -= eps
        args_p = [p if a is arr else a for a in (q, k, v)]
        args_m = [m if a is arr else a for a in (q, k, v)]
        fd = (f(*args_p) - f(*args_m)) / (2 * eps)
        assert abs(fd - grad[idx]) < 1e-6


def test_round_to_matches_torch():
    import torch
    x = np.random.default_rng(19).standard_normal(4096) * np.exp(np.random.default_rng(20).uniform(-20, 20, 4096))
    assert np.array_equal(oracle.round_to(x, "bf16"), torch.from_numpy(x).float().bfloat16().double().numpy())
    assert np.array_equal(oracle.round_to(x, "f16"), torch.from_numpy(x).float().half().double().numpy())
