"""Python operator API - the host-side mirror of the reference's
flash_cosine_sim_attention/flash_cosine_sim_attention.py, kept name-for-name so callers
(benchmark.py, train.py, transformer.py, the reference's tests) switch over unchanged:

    flash_cosine_sim_attention(q, k, v, mask, attn_bias, scale, groups, causal, l2norm_qk,
                               attn_bias_batch_dim)                     reference py:308-334
    l2norm_tensors(*tensors, groups)                                    reference py:57-65
    plain_cosine_sim_attention(...)   naive un-fused formulation        reference py:75-126
    FlashCosineSimAttention (autograd.Function) / forward / backward / debug   py:245-304, cu:1928-1933

What differs, on purpose:
  * the compute is libfcsa_b200.so (hand-written sm_100a kernels behind a C ABI), reached
    through ctypes with raw device pointers and the current CUDA stream;
  * with l2norm_qk=True the normalisation and its backward run as fused CUDA kernels inside
    one autograd node instead of PyTorch ops around it (reference py:320-321);
  * CUDA tensors only: the reference's tiled CPU forward (py:130-241) is not reproduced (it is
    wrong for causal N > 512, py:215) - CPU tensors raise;
  * the exponent shift is scale*groups (not scale): mathematically identical (softmax shift
    invariance), but safe in 16 bit when groups > 1 (q.k can reach `groups`).
"""
import math
import warnings

import torch
from torch.autograd import Function

from . import _abi
from ._abi import FcsaProblem, FcsaTensor

__all__ = [
    "flash_cosine_sim_attention",
    "plain_cosine_sim_attention",
    "l2norm_tensors",
    "FlashCosineSimAttention",
    "flash_cosine_sim_attention_cuda",
    "forward",
    "backward",
    "debug",
]

_KERNEL_DTYPES = {torch.float16: _abi.FCSA_F16, torch.bfloat16: _abi.FCSA_BF16}
_KERNEL_HEAD_DIMS = (64, 128)


def exists(val):
    return val is not None


# --------------------------------------------------------------------------------------------
# low level: tensors -> C ABI structs
# --------------------------------------------------------------------------------------------

def _tma_ready(t):
    """Tensor usable behind a TMA tensor map as is: feature dim contiguous, 16-byte aligned
    base and (batch, head, row) strides.  Otherwise a contiguous copy is made - e.g. for the
    stride-0 expanded grad that `o.sum().backward()` produces."""
    ok = t.stride(-1) == 1 and t.data_ptr() % 16 == 0 and all(s % 8 == 0 for s in t.stride()[:-1])
    return t if ok else t.contiguous()


def _view4(t, kind):
    """Canonical (batch, head, row, feature) addressing of a 3-D or 4-D tensor.
    kind: 'bhnd' 4-D as is | 'bnd' 3-D (batch, row, feature) -> one head."""
    if kind == "bhnd":
        sb, sh, sn = t.stride(0), t.stride(1), t.stride(2)
    else:
        sb, sh, sn = t.stride(0), 0, t.stride(1)
    return FcsaTensor(t.data_ptr(), sb, sh, sn)


class _Shapes:
    """Shape canonicalisation of the reference's host op (cu:1647-1660, cu:1679)."""

    def __init__(self, q, k, v):
        self.merged = q.ndim == 3
        if self.merged:
            assert k.ndim == 3 and v.ndim == 3, (
                "if batch and heads are merged for queries, keys and values must also similarly "
                "have only 3 dimensions")
            self.B, self.Nq, self.D = q.shape
            self.H = 1
            self.kv_heads = 1
            self.qkind = "bnd"
            self.kkind = "bnd"
        else:
            assert q.ndim == 4, "queries must be (batch, heads, seq, dim) or (batch*heads, seq, dim)"
            self.B, self.H, self.Nq, self.D = q.shape
            self.qkind = "bhnd"
            if k.ndim == 3:
                assert v.ndim == 3, "keys and values must both be single-headed"
                self.kv_heads = 1
                self.kkind = "bnd"
            else:
                self.kv_heads = self.H
                self.kkind = "bhnd"
        self.Nk = k.shape[-2]
        assert k.shape[-1] == self.D and v.shape[-1] == self.D, "head dimensions of q, k, v must match"
        assert k.shape[0] == self.B and v.shape[0] == self.B, "batch sizes of q, k, v must match"
        assert v.shape[-2] == self.Nk


def _problem(sh, dtype, scale, shift, causal, mask):
    p = FcsaProblem()
    p.dtype = _KERNEL_DTYPES[dtype]
    p.batch, p.heads, p.kv_heads = sh.B, sh.H, sh.kv_heads
    p.seq_q, p.seq_k, p.head_dim = sh.Nq, sh.Nk, sh.D
    p.causal = 1 if causal else 0
    p.scale = float(scale)
    p.shift = float(shift)
    if exists(mask):
        p.key_mask = mask.data_ptr()
        p.key_mask_stride = mask.stride(0)
    else:
        p.key_mask = None
        p.key_mask_stride = 0
    return p


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _prep_mask(mask, sh):
    if not exists(mask):
        return None
    assert mask.shape == (sh.B, sh.Nk), f"mask must be (batch, seq_k) = {(sh.B, sh.Nk)}, got {tuple(mask.shape)}"
    m = mask.to(torch.bool).contiguous()
    return m.view(torch.uint8)


def _kernel_supported(q, k, v, attn_bias):
    return (q.is_cuda and q.dtype in _KERNEL_DTYPES and k.dtype == q.dtype and v.dtype == q.dtype
            and q.shape[-1] in _KERNEL_HEAD_DIMS and not exists(attn_bias))


def _bias_ready(attn_bias, sh, dtype, batch_dim):
    """(heads, i, j) - or (batch, i, j) when batch_dim - bias -> tensor in q's dtype whose rows are 16-byte
    aligned (row length padded to a multiple of 8), plus the fcsa_bias struct addressing it as
    [batch][head][i][j] (a stride of 0 for the dimension the bias does not have)."""
    lead = sh.B if batch_dim else sh.H
    assert tuple(attn_bias.shape) == (lead, sh.Nq, sh.Nk), \
        f"attn_bias must be {(lead, sh.Nq, sh.Nk)} ({'batch' if batch_dim else 'heads'}, i, j), got {tuple(attn_bias.shape)}"
    t = attn_bias.detach().to(dtype)
    pad = (-sh.Nk) % 8
    if pad:
        t = torch.nn.functional.pad(t, (0, pad))
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    bs = _abi.FcsaBias()
    bs.ptr = t.data_ptr()
    plane = t.stride(0)
    bs.sb, bs.sh, bs.sn = (plane, 0, t.stride(1)) if batch_dim else (0, plane, t.stride(1))
    return t, bs


def _attn_forward(q, k, v, mask_u8, scale, shift, causal, need_inv_l=True, bias=None, bias_batch_dim=False):
    """q, k (already normalised if wanted), v -> o, inv_l through fcsa_forward (fcsa_forward_bias with a bias)."""
    lib = _abi.load()
    sh = _Shapes(q, k, v)
    q, k, v = _tma_ready(q), _tma_ready(k), _tma_ready(v)
    o = torch.empty(q.shape, dtype=q.dtype, device=q.device)
    inv_l = torch.empty((sh.B, sh.H, sh.Nq), dtype=torch.float32, device=q.device) if need_inv_l else None
    p = _problem(sh, q.dtype, scale, shift, causal, mask_u8)
    tq, tk, tv, to = _view4(q, sh.qkind), _view4(k, sh.kkind), _view4(v, sh.kkind), _view4(o, sh.qkind)
    with torch.cuda.device(q.device):
        if bias is None:
            _abi.check(lib.fcsa_forward(_abi.ref(p), _abi.ref(tq), _abi.ref(tk), _abi.ref(tv), _abi.ref(to),
                                        inv_l.data_ptr() if need_inv_l else None, _stream(q.device)))
        else:
            keep, bs = _bias_ready(bias, sh, q.dtype, bias_batch_dim)
            _abi.check(lib.fcsa_forward_bias(_abi.ref(p), _abi.ref(tq), _abi.ref(tk), _abi.ref(tv), _abi.ref(bs),
                                             _abi.ref(to), inv_l.data_ptr() if need_inv_l else None,
                                             _stream(q.device)))
            del keep
    return o, inv_l


def _attn_backward(do, o, inv_l, q, k, v, mask_u8, scale, shift, causal, bias=None, bias_batch_dim=False,
                   bias_grad=False):
    """-> dq, dk, dv (and d_bias, in the bias's dtype and shape, when bias_grad)."""
    lib = _abi.load()
    sh = _Shapes(q, k, v)
    q, k, v, o, do = _tma_ready(q), _tma_ready(k), _tma_ready(v), _tma_ready(o), _tma_ready(do)
    dq = torch.empty(q.shape, dtype=q.dtype, device=q.device)
    dk = torch.empty(k.shape, dtype=k.dtype, device=q.device)
    dv = torch.empty(v.shape, dtype=v.dtype, device=q.device)
    p = _problem(sh, q.dtype, scale, shift, causal, mask_u8)
    nbytes = lib.fcsa_backward_workspace_bytes(_abi.ref(p))
    if nbytes == 0:
        _abi.check(_abi.FCSA_ERR_INVALID)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    t = lambda x, kind: _view4(x, kind)
    with torch.cuda.device(q.device):
        if bias is None:
            _abi.check(lib.fcsa_backward(
                _abi.ref(p), _abi.ref(t(q, sh.qkind)), _abi.ref(t(k, sh.kkind)), _abi.ref(t(v, sh.kkind)),
                _abi.ref(t(o, sh.qkind)), _abi.ref(t(do, sh.qkind)), inv_l.data_ptr(),
                _abi.ref(t(dq, sh.qkind)), _abi.ref(t(dk, sh.kkind)), _abi.ref(t(dv, sh.kkind)),
                ws.data_ptr(), nbytes, _stream(q.device)))
            return dq, dk, dv
        keep, bs = _bias_ready(bias, sh, q.dtype, bias_batch_dim)
        db_acc = torch.zeros(bias.shape, dtype=torch.float32, device=q.device) if bias_grad else None
        plane = sh.Nq * sh.Nk
        dsb, dsh = (plane, 0) if bias_batch_dim else (0, plane)
        _abi.check(lib.fcsa_backward_bias(
            _abi.ref(p), _abi.ref(t(q, sh.qkind)), _abi.ref(t(k, sh.kkind)), _abi.ref(t(v, sh.kkind)),
            _abi.ref(t(o, sh.qkind)), _abi.ref(t(do, sh.qkind)), inv_l.data_ptr(), _abi.ref(bs),
            db_acc.data_ptr() if bias_grad else None, dsb, dsh,
            _abi.ref(t(dq, sh.qkind)), _abi.ref(t(dk, sh.kkind)), _abi.ref(t(dv, sh.kkind)),
            ws.data_ptr(), nbytes, _stream(q.device)))
        del keep
    return dq, dk, dv, (db_acc.to(bias.dtype) if bias_grad else None)


def _l2norm_struct(sh, groups, qn, kn, rq, rk):
    n = _abi.FcsaL2Norm()
    n.groups = groups
    n.q_hat, n.k_hat = _view4(qn, sh.qkind), _view4(kn, sh.kkind)
    n.q_rnorm, n.k_rnorm = rq.data_ptr(), rk.data_ptr()
    return n


def _attn_forward_fused(q, k, v, mask_u8, scale, shift, causal, groups, need_inv_l=True):
    """raw q, k -> (o, inv_l, q_hat, k_hat, q_rnorm, k_rnorm) through fcsa_forward_fused: one
    launch normalises q and k, one runs the attention."""
    lib = _abi.load()
    sh = _Shapes(q, k, v)
    q, k, v = _tma_ready(q), _tma_ready(k), _tma_ready(v)
    dev = q.device
    o = torch.empty(q.shape, dtype=q.dtype, device=dev)
    qn = torch.empty(q.shape, dtype=q.dtype, device=dev)
    kn = torch.empty(k.shape, dtype=k.dtype, device=dev)
    rq = torch.empty((sh.B, sh.H, sh.Nq, groups), dtype=torch.float32, device=dev)
    rk = torch.empty((sh.B, sh.kv_heads, sh.Nk, groups), dtype=torch.float32, device=dev)
    inv_l = torch.empty((sh.B, sh.H, sh.Nq), dtype=torch.float32, device=dev) if need_inv_l else None
    p = _problem(sh, q.dtype, scale, shift, causal, mask_u8)
    n = _l2norm_struct(sh, groups, qn, kn, rq, rk)
    with torch.cuda.device(dev):
        _abi.check(lib.fcsa_forward_fused(_abi.ref(p), _abi.ref(_view4(q, sh.qkind)), _abi.ref(_view4(k, sh.kkind)),
                                          _abi.ref(_view4(v, sh.kkind)), _abi.ref(n), _abi.ref(_view4(o, sh.qkind)),
                                          inv_l.data_ptr() if need_inv_l else None, _stream(dev)))
    return o, inv_l, qn, kn, rq, rk


def _attn_backward_fused(do, o, inv_l, qn, kn, v, rq, rk, mask_u8, scale, shift, causal, groups):
    """gradients w.r.t. the raw q, k and v through fcsa_backward_fused."""
    lib = _abi.load()
    sh = _Shapes(qn, kn, v)
    v, o, do = _tma_ready(v), _tma_ready(o), _tma_ready(do)
    dev = qn.device
    dq = torch.empty(qn.shape, dtype=qn.dtype, device=dev)
    dk = torch.empty(kn.shape, dtype=kn.dtype, device=dev)
    dv = torch.empty(v.shape, dtype=v.dtype, device=dev)
    p = _problem(sh, qn.dtype, scale, shift, causal, mask_u8)
    n = _l2norm_struct(sh, groups, qn, kn, rq, rk)
    nbytes = lib.fcsa_backward_workspace_bytes(_abi.ref(p))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _abi.check(lib.fcsa_backward_fused(
            _abi.ref(p), _abi.ref(n), _abi.ref(_view4(v, sh.kkind)), _abi.ref(_view4(o, sh.qkind)),
            _abi.ref(_view4(do, sh.qkind)), inv_l.data_ptr(), _abi.ref(_view4(dq, sh.qkind)),
            _abi.ref(_view4(dk, sh.kkind)), _abi.ref(_view4(dv, sh.kkind)), ws.data_ptr(), nbytes, _stream(dev)))
    return dq, dk, dv


def _l2norm_forward(x, groups):
    """CUDA kernel: x -> (x normalised per group, 1/norm) with x 3-D or 4-D."""
    lib = _abi.load()
    x = _tma_ready(x)
    kind = "bhnd" if x.ndim == 4 else "bnd"
    B = x.shape[0]
    H = x.shape[1] if x.ndim == 4 else 1
    N, D = x.shape[-2], x.shape[-1]
    y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    rnorm = torch.empty((B, H, N, groups), dtype=torch.float32, device=x.device)
    tx, ty = _view4(x, kind), _view4(y, kind)
    with torch.cuda.device(x.device):
        _abi.check(lib.fcsa_l2norm_forward(_KERNEL_DTYPES[x.dtype], B, H, N, D, groups, _abi.ref(tx),
                                           _abi.ref(ty), rnorm.data_ptr(), _stream(x.device)))
    return y, rnorm


def _l2norm_backward(dy, y, rnorm, groups):
    lib = _abi.load()
    dy, y = _tma_ready(dy), _tma_ready(y)
    kind = "bhnd" if y.ndim == 4 else "bnd"
    B = y.shape[0]
    H = y.shape[1] if y.ndim == 4 else 1
    N, D = y.shape[-2], y.shape[-1]
    dx = torch.empty(y.shape, dtype=y.dtype, device=y.device)
    with torch.cuda.device(y.device):
        _abi.check(lib.fcsa_l2norm_backward(_KERNEL_DTYPES[y.dtype], B, H, N, D, groups,
                                            _abi.ref(_view4(dy, kind)), _abi.ref(_view4(y, kind)),
                                            rnorm.data_ptr(), _abi.ref(_view4(dx, kind)), _stream(y.device)))
    return dx


# --------------------------------------------------------------------------------------------
# the reference's extension-module surface: forward / backward / debug (cu:1630, 1752, 1921)
# --------------------------------------------------------------------------------------------

def forward(q, k, v, mask, attn_bias, attn_bias_batch_dim, scale, causal, shift=None):
    """Same contract as the reference's pybind `forward`: returns (o, inv_l, should_backwards).
    `shift` (extra, optional): the constant subtracted from the logits; the reference's is `scale`."""
    assert not (causal and exists(mask)), "mask should not be supplied if causality is needed"
    sh = _Shapes(q, k, v)
    if sh.merged:
        attn_bias_batch_dim = True          # reference cu:1647-1654
    should_backwards = any(t.requires_grad for t in (q, k, v)) or (exists(attn_bias) and attn_bias.requires_grad)
    o, inv_l = _attn_forward(q, k, v, _prep_mask(mask, sh), scale, scale if shift is None else shift, causal,
                             need_inv_l=True, bias=attn_bias, bias_batch_dim=attn_bias_batch_dim)
    return o, inv_l, should_backwards


def backward(d_out, o, inv_l, q, k, v, mask, attn_bias, attn_bias_batch_dim, scale, causal, shift=None):
    """Same contract as the reference's pybind `backward`: returns (dq, dk, dv, db)."""
    sh = _Shapes(q, k, v)
    sft = scale if shift is None else shift
    if not exists(attn_bias):
        dq, dk, dv = _attn_backward(d_out, o, inv_l, q, k, v, _prep_mask(mask, sh), scale, sft, causal)
        return dq, dk, dv, None
    if sh.merged:
        attn_bias_batch_dim = True
    return _attn_backward(d_out, o, inv_l, q, k, v, _prep_mask(mask, sh), scale, sft, causal, bias=attn_bias,
                          bias_batch_dim=attn_bias_batch_dim, bias_grad=attn_bias.requires_grad)


def debug():
    """Reference: a no-op hook (cu:1921).  Here: number of kernels launched by the library."""
    return int(_abi.load().fcsa_debug())


class FlashCosineSimAttention(Function):
    """The reference's autograd.Function (py:245-304) on already-normalised q, k."""

    @staticmethod
    def forward(ctx, q, k, v, mask, attn_bias, scale, causal, attn_bias_batch_dim, shift=None):
        o, inv_l, should_backwards = forward(q, k, v, mask, attn_bias, attn_bias_batch_dim, scale, causal, shift)
        if not should_backwards:
            return o
        ctx.should_backwards = should_backwards
        ctx.save_for_backward(o, inv_l, q, k, v, mask, attn_bias)
        ctx.params = (scale, causal, attn_bias_batch_dim, shift)
        return o

    @staticmethod
    def backward(ctx, do):
        assert ctx.should_backwards
        o, inv_l, q, k, v, mask, attn_bias = ctx.saved_tensors
        scale, causal, attn_bias_batch_dim, shift = ctx.params
        dq, dk, dv, db = backward(do, o, inv_l, q, k, v, mask, attn_bias, attn_bias_batch_dim, scale, causal, shift)
        return dq, dk, dv, None, db, None, None, None, None


flash_cosine_sim_attention_cuda = FlashCosineSimAttention.apply


def _choose_shift(dtype, scale, groups, l2norm_qk):
    """Constant subtracted from the logits before exp (any constant gives the same attention;
    the reference hard-codes `scale`, cu:1216).  p = exp(logit - shift) is stored in 16 bit:
      bf16: shift = scale*groups - p <= 1 for every possible q.k (<= groups); bf16 has fp32's range.
      fp16, groups == 1: q.k <= 1, so shift = scale - 15 ln 2 puts p in (0, 2^15]: the top of the
            fp16 range instead of its subnormals.
      fp16, groups > 1: exp(scale*q.k) spans e^(2*scale*groups) - more than fp16 can hold without a
            row max.  Same choice as the reference (shift = scale); the kernels saturate p at
            65504 instead of producing inf.  Use bf16 for grouped l2norm with large scale*groups."""
    if not l2norm_qk:
        return float(scale)
    if dtype == torch.bfloat16:
        return float(scale * groups)
    if groups == 1:
        return float(scale) - 15.0 * math.log(2.0)
    return float(scale)


class _FusedCosineSimAttention(Function):
    """One autograd node for l2norm(q), l2norm(k) -> attention, all in CUDA kernels."""

    @staticmethod
    def forward(ctx, q, k, v, mask, scale, causal, groups, l2norm_qk, shift_groups=0):
        sh = _Shapes(q, k, v)
        mask_u8 = _prep_mask(mask, sh)
        # shift_groups > 0: q, k arrive already normalised over that many groups (padded head dims)
        shift = (_choose_shift(q.dtype, scale, shift_groups, True) if shift_groups > 0
                 else _choose_shift(q.dtype, scale, groups, l2norm_qk))
        needs_grad = any(ctx.needs_input_grad[:3])
        if l2norm_qk:
            o, inv_l, qn, kn, rq, rk = _attn_forward_fused(q, k, v, mask_u8, scale, shift, causal, groups,
                                                           need_inv_l=needs_grad)
        else:
            qn, kn, rq, rk = q, k, None, None
            o, inv_l = _attn_forward(qn, kn, v, mask_u8, scale, shift, causal, need_inv_l=needs_grad)
        if needs_grad:
            ctx.save_for_backward(o, inv_l, qn, kn, v, mask_u8, rq, rk)
            ctx.params = (scale, shift, causal, groups, l2norm_qk)
        return o

    @staticmethod
    def backward(ctx, do):
        o, inv_l, qn, kn, v, mask_u8, rq, rk = ctx.saved_tensors
        scale, shift, causal, groups, l2norm_qk = ctx.params
        if l2norm_qk:
            dq, dk, dv = _attn_backward_fused(do, o, inv_l, qn, kn, v, rq, rk, mask_u8, scale, shift, causal, groups)
        else:
            dq, dk, dv = _attn_backward(do, o, inv_l, qn, kn, v, mask_u8, scale, shift, causal)
        return dq, dk, dv, None, None, None, None, None, None


class _L2Norm(Function):
    """l2norm_tensors on CUDA 16-bit inputs: the fused kernel with its own backward."""

    @staticmethod
    def forward(ctx, x, groups):
        y, rnorm = _l2norm_forward(x, groups)
        ctx.save_for_backward(y, rnorm)
        ctx.groups = groups
        return y

    @staticmethod
    def backward(ctx, dy):
        y, rnorm = ctx.saved_tensors
        return _l2norm_backward(dy, y, rnorm, ctx.groups), None


# --------------------------------------------------------------------------------------------
# public API
# --------------------------------------------------------------------------------------------

def _l2norm_torch(t, groups):
    shape = t.shape
    g = t.reshape(*shape[:-1], groups, shape[-1] // groups)
    acc = g.float() if g.dtype in (torch.float16, torch.bfloat16) else g
    g = torch.nn.functional.normalize(acc, dim=-1).to(t.dtype)
    return g.reshape(shape)


def l2norm_tensors(*tensors, groups=1):
    """l2-normalise each tensor over `groups` chunks of the last dim; results keep the dtype of
    the first tensor (reference py:57-65)."""
    assert len(tensors) > 0
    dtype = tensors[0].dtype
    out = []
    for t in tensors:
        assert t.shape[-1] % groups == 0, "groups must divide the feature dimension"
        gs = t.shape[-1] // groups
        fused = (t.is_cuda and t.dtype in _KERNEL_DTYPES and t.ndim in (3, 4)
                 and t.shape[-1] in (32, 64, 128, 256) and (gs & (gs - 1)) == 0)
        y = _L2Norm.apply(t, groups) if fused else _l2norm_torch(t, groups)
        out.append(y.type(dtype))
    return tuple(out)


def plain_cosine_sim_attention(q, k, v, mask=None, attn_bias=None, scale=8, groups=1, causal=False,
                               l2norm_qk=True, attn_bias_batch_dim=False):
    """The naive, un-fused formulation (reference py:75-126): explicit similarity matrix,
    masked softmax, weighted sum.  Public API of the reference package (used by
    transformer.py when the fused kernel is switched off); any device, any float dtype."""
    assert not (causal and exists(mask)), "mask should not be supplied if causality is needed"
    merged = q.ndim == 3
    single_head_kv = k.ndim == 3
    if merged:
        assert k.ndim == 3 and v.ndim == 3, (
            "if batch and heads are merged for queries, keys and values must also similarly "
            "have only 3 dimensions")
        attn_bias_batch_dim = True
        q = q.unsqueeze(1)
    if l2norm_qk:
        q, k = (_l2norm_torch(t, groups) for t in (q, k))
    kk = k.unsqueeze(1) if single_head_kv else k
    vv = v.unsqueeze(1) if single_head_kv else v
    sim = torch.matmul(q, kk.transpose(-1, -2)) * scale
    if exists(attn_bias):
        sim = sim + attn_bias.unsqueeze(1 if attn_bias_batch_dim else 0)
    neg = -torch.finfo(sim.dtype).max
    if causal:
        i, j = sim.shape[-2:]
        future = torch.ones((i, j), device=q.device, dtype=torch.bool).triu(j - i + 1)
        sim = sim.masked_fill(future, neg)
    if exists(mask):
        sim = sim.masked_fill(~mask[:, None, None, :], neg)
    out = torch.matmul(sim.softmax(dim=-1), vv)
    return out.squeeze(1) if merged else out


_warned = set()


def _warn_once(key, msg):
    if key not in _warned:
        _warned.add(key)
        warnings.warn(msg, stacklevel=3)


def flash_cosine_sim_attention(q, k, v, mask=None, attn_bias=None, scale=8, groups=1, causal=False,
                               l2norm_qk=True, attn_bias_batch_dim=False, l2norm_groups=None):
    """Fused cosine-similarity attention (reference py:308-334, same arguments and meaning;
    `l2norm_groups` is accepted as an alias of `groups`)."""
    if exists(l2norm_groups):
        groups = l2norm_groups
    assert not (causal and exists(mask)), "mask should not be supplied if causality is needed"
    if not q.is_cuda:
        raise RuntimeError(
            "flash_cosine_sim_attention: CUDA tensors required - this build has no CPU path "
            "(use plain_cosine_sim_attention on CPU tensors)")
    D = q.shape[-1]
    if (q.dtype in _KERNEL_DTYPES and k.dtype == q.dtype and v.dtype == q.dtype and not exists(attn_bias)
            and D not in _KERNEL_HEAD_DIMS and D < 128 and D % 8 == 0 and D % groups == 0):
        # Head dims the kernels are not instantiated for (the reference's 32 and 96): zero-pad the
        # features to 64 / 128 and run the same tcgen05 kernels.  Zero features change neither q.k nor
        # the norms (the l2norm runs first, on the real features), the padded output columns are
        # zero and sliced off; autograd takes care of the slices.  Costs two pad copies per tensor.
        Dp = 64 if D < 64 else 128
        if l2norm_qk:
            q, k = _l2norm_torch(q, groups), _l2norm_torch(k, groups)
        pad = lambda t: torch.nn.functional.pad(t, (0, Dp - D))
        o = _FusedCosineSimAttention.apply(pad(q), pad(k), pad(v), mask, float(scale), bool(causal), 1, False,
                                           int(groups) if l2norm_qk else 0)
        return o[..., :D]
    if (exists(attn_bias) and q.dtype in _KERNEL_DTYPES and k.dtype == q.dtype and v.dtype == q.dtype
            and D in _KERNEL_HEAD_DIMS and D % groups == 0 and attn_bias.is_cuda):
        # additive bias: l2norm kernels (with their own backward), then the BIAS instantiations of the
        # attention kernels; d_bias is reduced in fp32 and returned in the bias's dtype
        if l2norm_qk:
            q, k = l2norm_tensors(q, k, groups=groups)
        shift = _choose_shift(q.dtype, scale, groups if l2norm_qk else 1, True)
        if q.dtype == torch.float16:
            # p = exp(logit - shift) is stored in fp16: keep its top below 2^15 whatever the bias adds
            shift += max(float(attn_bias.detach().amax()), 0.0)
        return FlashCosineSimAttention.apply(q, k, v, mask, attn_bias, float(scale), bool(causal),
                                             bool(attn_bias_batch_dim), float(shift))
    if not _kernel_supported(q, k, v, attn_bias):
        # float32 inputs and head dims above 128 / not a multiple of 8 have no sm_100a kernel yet:
        # they run the un-fused formulation on the same GPU (correct, slower), never silently wrong.
        why = ("attn_bias" if exists(attn_bias) else
               f"dtype {q.dtype}" if q.dtype not in _KERNEL_DTYPES else f"head_dim {q.shape[-1]}")
        _warn_once(why, f"flash_cosine_sim_attention: no fused sm_100a kernel for {why}; "
                        "running the un-fused GPU formulation")
        return plain_cosine_sim_attention(q, k, v, mask=mask, attn_bias=attn_bias, scale=scale,
                                          groups=groups, causal=causal, l2norm_qk=l2norm_qk,
                                          attn_bias_batch_dim=attn_bias_batch_dim)
    gs = q.shape[-1] // groups
    assert q.shape[-1] % groups == 0 and (gs & (gs - 1)) == 0, "groups must divide the head dim into power-of-two chunks"
    return _FusedCosineSimAttention.apply(q, k, v, mask, float(scale), bool(causal), int(groups), bool(l2norm_qk))
