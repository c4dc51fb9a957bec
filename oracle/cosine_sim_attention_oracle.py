"""CPU oracle for fused cosine-similarity attention.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product (flash_cosine_sim_attention_b200) never does.

It restates, in numpy float64 and from the reference's source, what
`plain_cosine_sim_attention` computes (reference flash_cosine_sim_attention/
flash_cosine_sim_attention.py:75-126) and the closed-form backward the reference's CUDA kernel
implements (flash_cosine_sim_attention_cuda.cu:1487-1626).  Each function cites the lines it
follows.

Pinned: `oracle/make_golden.py` runs the UNMODIFIED reference module (loaded by path from
/root/reference) on seeded inputs, checks this restatement against it (forward and autograd
gradients) and writes tests/golden/*.npz; tests/test_oracle.py re-checks the restatement
against those committed vectors on every run.  The reference ships no golden vectors of its
own (its tests are differential on unseeded randn, tests/test.py:53-58).
"""
import numpy as np

F64 = np.float64


def l2norm(x, groups=1, eps=1e-12):
    """x / max(||x||_2, eps) over each of `groups` equal chunks of the last dim.
    reference py:38-55 (l2norm -> F.normalize eps 1e-12 on GPU; grouped_l2norm reshapes to
    (..., groups, dim // groups))."""
    x = np.asarray(x, dtype=F64)
    shape = x.shape
    g = x.reshape(*shape[:-1], groups, shape[-1] // groups)
    n = np.sqrt((g * g).sum(-1, keepdims=True))
    return (g / np.maximum(n, eps)).reshape(shape), n.reshape(*shape[:-1], groups)


def l2norm_backward(dy, x, groups=1, eps=1e-12):
    """Gradient of l2norm w.r.t. x: (dy - y <y, dy>_group) / ||x||_group.
    (autograd of F.normalize in the reference; restated analytically)."""
    x = np.asarray(x, dtype=F64)
    dy = np.asarray(dy, dtype=F64)
    shape = x.shape
    g = x.reshape(*shape[:-1], groups, shape[-1] // groups)
    d = dy.reshape(g.shape)
    n = np.maximum(np.sqrt((g * g).sum(-1, keepdims=True)), eps)
    y = g / n
    return ((d - y * (y * d).sum(-1, keepdims=True)) / n).reshape(shape)


def round_to(x, fmt):
    """Round float64 values to the nearest bf16 / fp16 value (ties to even), back as float64.
    The reference casts the normalised q, k back to the input dtype before they enter the
    attention (py:64); 16-bit parity runs have to model that rounding."""
    if fmt is None:
        return x
    if fmt in ("f16", "float16"):
        return np.asarray(x, dtype=np.float32).astype(np.float16).astype(F64)
    if fmt in ("bf16", "bfloat16"):
        u = np.ascontiguousarray(np.asarray(x, dtype=np.float32)).view(np.uint32).astype(np.uint64)
        u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
        return u.astype(np.uint32).view(np.float32).reshape(np.shape(x)).astype(F64)
    raise ValueError(fmt)


def _canon(q, k, v):
    """reference py:90-97: 3-D q = merged batch-heads (then k, v must be 3-D too);
    3-D k with 4-D q = one key/value head shared by all query heads."""
    merged = q.ndim == 3
    single_head_kv = k.ndim == 3
    if merged:
        assert k.ndim == 3 and v.ndim == 3
        q = q[:, None]
    if single_head_kv:
        k, v = k[:, None], v[:, None]
    return q, k, v, merged, single_head_kv


def _visibility(b_idx, Nq, Nk, mask, causal):
    """Boolean (Nq, Nk) matrix of attendable pairs for batch element b_idx.
    causal: reference py:112-115 masks triu(j - i + 1), i.e. key j is visible to query i iff
    j <= i + (Nk - Nq) (bottom-right aligned).  key mask: py:117-118, True = keep."""
    vis = np.ones((Nq, Nk), dtype=bool)
    if causal:
        i = np.arange(Nq)[:, None]
        j = np.arange(Nk)[None, :]
        vis &= j <= i + (Nk - Nq)
    if mask is not None:
        vis &= np.asarray(mask[b_idx], dtype=bool)[None, :]
    return vis


def attention(q, k, v, mask=None, attn_bias=None, scale=8, groups=1, causal=False, l2norm_qk=True,
              attn_bias_batch_dim=False, d_out=None, empty_rows="mean", round_qk=None):
    """Forward (and, when d_out is given, backward) of cosine-sim attention in float64.

    Forward: reference py:75-126 - normalise q, k (py:99-100); sim = scale * q k^T (py:102-104);
    + bias (py:106-108); masked_fill(-max) for causal / key mask (py:110-118); softmax (py:120);
    out = attn v (py:121).
    Backward: the closed form of cu:1487-1626 (dV = P^T dO, dP = dO V^T, dS = P*(dP - rowsum(dO*O)),
    dK = scale dS^T q, dQ = scale dS k), followed by the l2norm backward, so that the returned
    dq, dk are w.r.t. the RAW q, k exactly like autograd through the reference.

    empty_rows: what a query with no visible key yields - "mean" (plain reference: softmax of an
    all-(-max) row = uniform average of v) or "zero" (the fused kernels, cu:1239: 1/max(l, eps)).

    round_qk: None | "bf16" | "f16" - round the NORMALISED q, k to that format before the
    similarity (what the reference's l2norm_tensors does for 16-bit inputs, py:64); gradients flow
    through the rounding unchanged, as autograd does through a dtype cast.

    Returns o, or (o, dq, dk, dv) when d_out is given - plus d_bias (shape of attn_bias: dS summed
    over the batch unless the bias has a batch dimension, cu:1574-1576) when a bias is given too.
    Shapes follow the inputs."""
    assert not (causal and mask is not None), "mask should not be supplied if causality is needed"
    q0, k0, v0 = (np.asarray(t, dtype=F64) for t in (q, k, v))
    q4, k4, v4, merged, single_head_kv = _canon(q0, k0, v0)
    if merged:
        attn_bias_batch_dim = True
    B, H, Nq, D = q4.shape
    Hk, Nk = k4.shape[1], k4.shape[2]
    if l2norm_qk:
        qn, _ = l2norm(q4, groups)
        kn, _ = l2norm(k4, groups)
        qn, kn = round_to(qn, round_qk), round_to(kn, round_qk)
    else:
        qn, kn = q4, k4
    o = np.zeros((B, H, Nq, D), dtype=F64)
    want_grad = d_out is not None
    if want_grad:
        do4 = np.asarray(d_out, dtype=F64)
        if merged:
            do4 = do4[:, None]
        dqn = np.zeros_like(qn)
        dkn = np.zeros_like(kn)
        dv4 = np.zeros_like(v4)
        dbias = None if attn_bias is None else np.zeros(np.asarray(attn_bias).shape, dtype=F64)
    for b in range(B):
        vis = _visibility(b, Nq, Nk, mask, causal)
        any_vis = vis.any(-1)
        for h in range(H):
            hk = 0 if Hk == 1 else h
            s = scale * (qn[b, h] @ kn[b, hk].T)
            if attn_bias is not None:
                bias = np.asarray(attn_bias, dtype=F64)
                s = s + (bias[b] if attn_bias_batch_dim else bias[h])
            s = np.where(vis, s, -np.inf)
            m = np.where(any_vis, s.max(-1, initial=-np.inf), 0.0)
            e = np.exp(s - m[:, None])
            e = np.where(vis, e, 0.0)
            l = e.sum(-1)
            p = e / np.maximum(l, 1e-300)[:, None]
            if empty_rows == "mean":
                p[~any_vis] = 1.0 / Nk
            o[b, h] = p @ v4[b, hk]
            if want_grad:
                do_bh = do4[b, h]
                dv4[b, hk] += p.T @ do_bh
                dp = do_bh @ v4[b, hk].T
                delta = (do_bh * o[b, h]).sum(-1, keepdims=True)
                ds = p * (dp - delta)
                if empty_rows == "mean":
                    ds[~any_vis] = 0.0      # the logits of such rows are constants (-max)
                dqn[b, h] = scale * (ds @ kn[b, hk])
                dkn[b, hk] += scale * (ds.T @ qn[b, h])
                if dbias is not None:
                    dbias[b if attn_bias_batch_dim else h] += ds
    out = o[:, 0] if merged else o
    if not want_grad:
        return out
    if l2norm_qk:
        dq = l2norm_backward(dqn, q4, groups)
        dk = l2norm_backward(dkn, k4, groups)
    else:
        dq, dk = dqn, dkn
    if merged:
        dq = dq[:, 0]
    if single_head_kv:
        dk, dv4 = dk[:, 0], dv4[:, 0]
    if dbias is not None:
        return out, dq, dk, dv4, dbias
    return out, dq, dk, dv4


# ------------------------------------------------------------------------------------------------
# torch float32 port of the reference's naive path, used ONLY to time a CPU baseline
# (bench.py cpu_baseline / --impl reference): same op sequence as reference py:75-126
# (einsum -> scale -> masked_fill -> softmax -> einsum) with autograd for the backward.
# ------------------------------------------------------------------------------------------------

def torch_cpu_forward_backward(q, k, v, scale=8, groups=1, causal=False, mask=None, backward=True):
    import torch
    import torch.nn.functional as F

    def norm(t):
        shape = t.shape
        t = t.reshape(*shape[:-1], groups, shape[-1] // groups)
        return F.normalize(t, dim=-1).reshape(shape)

    qn, kn = norm(q), norm(k)
    sim = torch.einsum("bhid,bhjd->bhij", qn, kn) * scale
    neg = -torch.finfo(sim.dtype).max
    if causal:
        i, j = sim.shape[-2:]
        sim = sim.masked_fill(torch.ones((i, j), dtype=torch.bool).triu(j - i + 1), neg)
    if mask is not None:
        sim = sim.masked_fill(~mask[:, None, None, :], neg)
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
    if backward:
        out.sum().backward()
    return out
