"""CPU: host-side logic of the Python operator API (no kernels run here)."""
import numpy as np
import pytest
import torch

from flash_cosine_sim_attention_b200 import (flash_cosine_sim_attention, l2norm_tensors,
                                             plain_cosine_sim_attention)
import importlib

op_module = importlib.import_module("flash_cosine_sim_attention_b200.flash_cosine_sim_attention")
from oracle import cosine_sim_attention_oracle as oracle


def test_public_surface_matches_reference_package():
    import flash_cosine_sim_attention_b200 as pkg
    for name in ("flash_cosine_sim_attention", "plain_cosine_sim_attention", "l2norm_tensors", "debug"):
        assert hasattr(pkg, name)


def test_signature_matches_reference():
    import inspect
    sig = inspect.signature(flash_cosine_sim_attention)
    names = list(sig.parameters)
    assert names[:10] == ["q", "k", "v", "mask", "attn_bias", "scale", "groups", "causal", "l2norm_qk",
                          "attn_bias_batch_dim"]
    assert sig.parameters["scale"].default == 8 and sig.parameters["groups"].default == 1
    assert sig.parameters["l2norm_qk"].default is True and sig.parameters["causal"].default is False


def test_cpu_tensors_fail_loudly():
    q = torch.randn(1, 2, 8, 64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        flash_cosine_sim_attention(q, q, q)


def test_mask_and_causal_are_exclusive():
    q = torch.randn(1, 2, 8, 64)
    with pytest.raises(AssertionError, match="mask should not be supplied"):
        flash_cosine_sim_attention(q, q, q, mask=torch.ones(1, 8, dtype=torch.bool), causal=True)


@pytest.mark.parametrize("kw", [dict(), dict(causal=True), dict(groups=4, scale=2.0), dict(l2norm_qk=False, scale=1.0)])
def test_plain_matches_oracle(kw):
    g = torch.Generator().manual_seed(3)
    q, k, v = (torch.randn(2, 3, 21, 64, generator=g, dtype=torch.float64) for _ in range(3))
    got = plain_cosine_sim_attention(q, k, v, **kw).numpy()
    assert np.abs(got - oracle.attention(q.numpy(), k.numpy(), v.numpy(), **kw)).max() < 1e-10


def test_plain_variants_match_oracle():
    g = torch.Generator().manual_seed(4)
    q = torch.randn(2, 4, 10, 64, generator=g, dtype=torch.float64)
    k, v = (torch.randn(2, 14, 64, generator=g, dtype=torch.float64) for _ in range(2))   # single-head kv
    mask = torch.rand(2, 14, generator=g) > 0.4
    mask[:, 0] = True
    got = plain_cosine_sim_attention(q, k, v, mask=mask).numpy()
    assert np.abs(got - oracle.attention(q.numpy(), k.numpy(), v.numpy(), mask=mask.numpy())).max() < 1e-10
    qm = torch.randn(6, 9, 64, generator=g, dtype=torch.float64)                           # merged batch-heads
    km, vm = (torch.randn(6, 12, 64, generator=g, dtype=torch.float64) for _ in range(2))
    got = plain_cosine_sim_attention(qm, km, vm, causal=True).numpy()
    assert got.shape == (6, 9, 64)
    assert np.abs(got - oracle.attention(qm.numpy(), km.numpy(), vm.numpy(), causal=True)).max() < 1e-10


def test_plain_with_bias_matches_oracle():
    g = torch.Generator().manual_seed(5)
    q, k, v = (torch.randn(2, 3, 7, 64, generator=g, dtype=torch.float64) for _ in range(3))
    bias_h = torch.randn(3, 7, 7, generator=g, dtype=torch.float64)
    bias_b = torch.randn(2, 7, 7, generator=g, dtype=torch.float64)
    a = plain_cosine_sim_attention(q, k, v, attn_bias=bias_h).numpy()
    assert np.abs(a - oracle.attention(q.numpy(), k.numpy(), v.numpy(), attn_bias=bias_h.numpy())).max() < 1e-10
    b = plain_cosine_sim_attention(q, k, v, attn_bias=bias_b, attn_bias_batch_dim=True).numpy()
    assert np.abs(b - oracle.attention(q.numpy(), k.numpy(), v.numpy(), attn_bias=bias_b.numpy(),
                                       attn_bias_batch_dim=True)).max() < 1e-10


def test_l2norm_tensors_cpu_matches_oracle_and_keeps_dtype():
    g = torch.Generator().manual_seed(6)
    a, b = torch.randn(2, 3, 5, 64, generator=g), torch.randn(2, 5, 64, generator=g)
    for groups in (1, 2, 8):
        ya, yb = l2norm_tensors(a, b, groups=groups)
        assert ya.dtype == a.dtype and ya.shape == a.shape and yb.shape == b.shape
        assert np.abs(ya.numpy() - oracle.l2norm(a.numpy(), groups)[0]).max() < 1e-6
    (h,) = l2norm_tensors(a.half(), groups=2)
    assert h.dtype == torch.float16


def test_shape_canonicalisation():
    S = op_module._Shapes
    s = S(torch.empty(2, 8, 10, 64), torch.empty(2, 8, 12, 64), torch.empty(2, 8, 12, 64))
    assert (s.B, s.H, s.kv_heads, s.Nq, s.Nk, s.D, s.merged) == (2, 8, 8, 10, 12, 64, False)
    s = S(torch.empty(2, 8, 10, 64), torch.empty(2, 12, 64), torch.empty(2, 12, 64))
    assert (s.H, s.kv_heads, s.kkind) == (8, 1, "bnd")
    s = S(torch.empty(16, 10, 64), torch.empty(16, 12, 64), torch.empty(16, 12, 64))
    assert (s.B, s.H, s.kv_heads, s.merged) == (16, 1, 1, True)
    with pytest.raises(AssertionError):
        S(torch.empty(16, 10, 64), torch.empty(2, 8, 12, 64), torch.empty(2, 8, 12, 64))


def test_tma_ready_copies_only_when_needed():
    f = op_module._tma_ready
    x = torch.empty(2, 10, 8, 64, dtype=torch.float16).permute(0, 2, 1, 3)    # transformer.py layout
    assert f(x) is x                                                           # strided but TMA-expressible
    e = torch.zeros((), dtype=torch.float16).expand(2, 8, 10, 64)             # grad of o.sum()
    assert f(e).is_contiguous() and f(e) is not e
    t = torch.empty(2, 8, 10, 128, dtype=torch.float16)[..., ::2]
    assert f(t).stride(-1) == 1


def test_extension_error_path_raises_runtime_error():
    """A C-ABI error must surface as a Python RuntimeError carrying the library's message (the module once
    crashed while unwinding: built with a toolchain whose exception tables did not match libtorch's)."""
    from flash_cosine_sim_attention_b200.flash_cosine_sim_attention import _ext
    with pytest.raises(RuntimeError, match="null problem"):
        _ext()._error_path_selftest()


def test_extension_module_has_the_reference_surface():
    """forward / backward / debug under the reference's module name (cu:1928-1933, version.py:3)."""
    import importlib
    from flash_cosine_sim_attention_b200.flash_cosine_sim_attention import _ext
    from flash_cosine_sim_attention_b200.version import __cuda_pkg_name__
    m = _ext()
    assert __cuda_pkg_name__ == "flash_cosine_sim_attention_cuda_0_1_40"
    assert importlib.import_module(__cuda_pkg_name__) is m
    for name in ("forward", "backward", "debug"):
        assert callable(getattr(m, name))
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        m.forward(torch.zeros(1, 2, 8, 64), torch.zeros(1, 2, 8, 64), torch.zeros(1, 2, 8, 64), None, None, False, 8.0, False)


def test_persistent_work_item_dealing_covers_every_item_once_and_balances():
    """The persistent kernels deal work items (numbered heaviest first) to G CTAs in snake order:
    round r -> item r*G + c on even rounds, r*G + G-1-c on odd ones (fwd_kernel.cuh / bwd_kernel.cuh `item_index`).
    Restated here: every item is visited exactly once, each CTA stops at its first missing item, and for the causal
    benchmark shapes the heaviest CTA stays within 5 % of a longest-first dynamic assignment (what one CTA per item
    under the hardware's block scheduler gives)."""
    import heapq

    def deal(n_items, G):
        per_cta = []
        for c in range(G):
            mine, r = [], 0
            while True:
                idx = r * G + ((G - 1 - c) if (r & 1) else c)
                if idx >= n_items:
                    break
                mine.append(idx)
                r += 1
            per_cta.append(mine)
        return per_cta

    for n_items, G in [(1, 148), (147, 148), (148, 148), (149, 148), (512, 148), (1024, 148), (66000, 148), (7, 3)]:
        per_cta = deal(n_items, G)
        seen = sorted(i for m in per_cta for i in m)
        assert seen == list(range(n_items))

    def makespans(weights, G):
        static = max(sum(weights[i] for i in m) for m in deal(len(weights), G))
        heap = [0] * G
        for w in weights:                       # items arrive heaviest first: longest-processing-time dispatch
            heapq.heappush(heap, heapq.heappop(heap) + w)
        return static, max(heap)

    # forward at (4,8,4096,64) causal: 16 query blocks x 32 (b,h), block b costs 2(b+1) iterations + a fixed part
    fwd = [2 * (16 - i // 32) * 2075 + 4800 for i in range(512)]
    # backward: 32 key tiles x 32 (b,h), key tile j is seen by 32 - j query tiles
    bwd = [(32 - i // 32) * 2400 + 5000 for i in range(1024)]
    for w in (fwd, bwd):
        static, dynamic = makespans(w, 148)
        assert static <= 1.05 * dynamic
