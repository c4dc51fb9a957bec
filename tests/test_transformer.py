"""The caller side (SURVEY.md par. 8f row 2): the re-authored transformer wrapper and the drop-in
import name.  CPU tests use the naive attention; the GPU tests run the fused kernels with fp16
autocast + GradScaler exactly as the reference's train.py does (train.py:53-64, 98-117)."""
import pytest
import torch


def test_drop_in_import_names():
    import flash_cosine_sim_attention as shim
    from flash_cosine_sim_attention.benchmark import benchmark
    from flash_cosine_sim_attention.transformer import Attention, CosineSimCausalTransformer
    for name in ("flash_cosine_sim_attention", "plain_cosine_sim_attention", "l2norm_tensors", "debug"):
        assert hasattr(shim, name)
    assert callable(benchmark) and Attention and CosineSimCausalTransformer


def _model(**kw):
    from flash_cosine_sim_attention_b200.transformer import CosineSimCausalTransformer
    torch.manual_seed(0)
    return CosineSimCausalTransformer(num_tokens=256, dim=128, max_seq_len=64, depth=2, heads=2, dim_head=64, **kw)


def test_state_dict_keys_match_reference_layout():
    keys = set(_model().state_dict().keys())
    for k in ("token_emb.weight", "pos_emb.weight", "layers.0.0.to_q.weight", "layers.0.0.to_out.weight",
              "layers.1.2.1.weight", "layers.1.2.3.weight", "to_logits.1.weight"):
        assert k in keys


@pytest.mark.parametrize("pre_norm", [False, True])
def test_forward_loss_generate_cpu(pre_norm):
    m = _model(pre_norm=pre_norm, attn_scale=1, attn_l2norm_groups=8)
    assert m.layers[0][0].l2norm_groups == 8            # the setting takes effect (reference quirk fixed)
    x = torch.randint(0, 256, (2, 33))
    logits = m(x)
    assert logits.shape == (2, 33, 256)
    loss = m(x, return_loss=True)
    loss.backward()
    assert torch.isfinite(loss) and all(p.grad is not None for p in m.parameters())
    out = m.generate(x[:, :5], 7)
    assert out.shape == (2, 7)


def test_causality_cpu():
    m = _model().eval()
    x = torch.randint(0, 256, (1, 20))
    y = x.clone()
    y[0, -1] = (y[0, -1] + 1) % 256
    a, b = m(x), m(y)
    assert torch.allclose(a[:, :-1], b[:, :-1], atol=1e-6) and not torch.allclose(a[:, -1], b[:, -1])


@pytest.mark.gpu
def test_train_step_fp16_autocast_matches_plain_attention():
    """train.py's configuration in miniature: autocast fp16, GradScaler, fused kernel vs naive path."""
    dev = "cuda"
    kw = dict(attn_scale=1, attn_l2norm_groups=8, pre_norm=True)
    fused, plain = _model(use_cuda_kernel=True, **kw).to(dev), _model(use_cuda_kernel=False, **kw).to(dev)
    plain.load_state_dict(fused.state_dict())
    x = torch.randint(0, 256, (2, 65), device=dev)
    scaler = torch.amp.GradScaler("cuda")
    losses = []
    for m in (fused, plain):
        with torch.autocast("cuda", dtype=torch.float16):
            loss = m(x, return_loss=True)
        scaler.scale(loss).backward()
        losses.append(loss.item())
    assert abs(losses[0] - losses[1]) < 2e-2
    g0 = fused.layers[0][0].to_q.weight.grad.float()
    g1 = plain.layers[0][0].to_q.weight.grad.float()
    assert torch.isfinite(g0).all() and (g0 - g1).abs().max() <= 5e-2 * g1.abs().max() + 1e-6


@pytest.mark.gpu
def test_generate_ragged_lengths_with_kernel():
    m = _model(use_cuda_kernel=True).cuda().half()
    start = torch.randint(0, 256, (2, 3), device="cuda")
    out = m.generate(start, 9)
    assert out.shape == (2, 9)


@pytest.mark.gpu
def test_reference_benchmark_helper_runs():
    from flash_cosine_sim_attention.benchmark import benchmark
    from flash_cosine_sim_attention import flash_cosine_sim_attention
    q, k, v = (torch.randn(1, 2, 256, 64, device="cuda", dtype=torch.float16).requires_grad_() for _ in range(3))
    ms = benchmark(flash_cosine_sim_attention, forwards=True, backwards=True, num_times=3, warmup_iters=2)(q, k, v, causal=True)
    assert ms > 0
