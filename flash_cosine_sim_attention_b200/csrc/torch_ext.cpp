// PyTorch extension module over the C ABI of libfcsa_b200.so.
//
// This is the module the reference builds from its .cu with setup.py (setup.py:30-39) and imports
// by its versioned name (flash_cosine_sim_attention.py:15-20, version.py:3):
//     flash_cosine_sim_attention_cuda_0_1_40 . forward / backward / debug      (cu:1928-1933)
// with the reference's exact signatures (cu:1630-1639, 1752-1764, 1921), so that even the reference's
// own unmodified flash_cosine_sim_attention.py loads it.  All compute happens behind
// include/fcsa_b200.h; this file only does what the reference's host op did around its launches
// (cu:1640-1700, 1766-1830): shape canonicalisation, allocation of the outputs, the current stream.
// Extra entry points (forward_ex / backward_ex / l2norm_*) expose the fused-l2norm and explicit-shift
// variants the Python operator layer uses; they keep the per-call host cost at a few microseconds.
//
// No kernels, no CUDA code here: this file is compiled by the host C++ compiler only.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <map>
#include <mutex>
#include <tuple>

#include "../../include/fcsa_b200.h"

namespace {

using at::Tensor;
using c10::optional;

#define FCSA_CHECK(call)                                                                      \
  do {                                                                                        \
    const int rc_ = (call);                                                                   \
    TORCH_CHECK(rc_ == FCSA_OK, "libfcsa_b200 error ", rc_, ": ", fcsa_last_error());          \
  } while (0)

int32_t dtype_code(const Tensor& t) {
  if (t.scalar_type() == at::kHalf) return FCSA_F16;
  if (t.scalar_type() == at::kBFloat16) return FCSA_BF16;
  TORCH_CHECK(false, "flash_cosine_sim_attention (sm_100a): dtype ", t.scalar_type(),
              " has no fused kernel behind this entry point (float16 / bfloat16 only)");
}

// Tensor usable behind a TMA tensor map as is: feature dim contiguous, 16-byte aligned base and
// (batch, head, row) strides, no stride-0 (expanded) dimension of extent > 1.  Otherwise a contiguous
// copy (e.g. the expanded grad of `o.sum().backward()`, head-expanded keys).
Tensor tma_ready(const Tensor& t) {
  const int64_t nd = t.dim();
  bool ok = t.stride(nd - 1) == 1 && (reinterpret_cast<uintptr_t>(t.data_ptr()) % 16) == 0;
  for (int64_t d = 0; ok && d < nd - 1; ++d) {
    const int64_t s = t.stride(d);
    ok = (s % 8 == 0) && (s > 0 || t.size(d) == 1);
  }
  return ok ? t : t.contiguous();
}

// canonical (batch, head, row, feature) addressing: 4-D as is, 3-D (batch, row, feature) = one head
fcsa_tensor view4(const Tensor& t) {
  fcsa_tensor v;
  v.ptr = t.data_ptr();
  if (t.dim() == 4) {
    v.sb = t.stride(0); v.sh = t.stride(1); v.sn = t.stride(2);
  } else {
    v.sb = t.stride(0); v.sh = 0; v.sn = t.stride(1);
  }
  return v;
}

// shape canonicalisation of the reference's host op (cu:1647-1660, cu:1679)
struct Shapes {
  bool merged;
  int64_t B, H, kv_heads, Nq, Nk, D;
  Shapes(const Tensor& q, const Tensor& k, const Tensor& v) {
    merged = q.dim() == 3;
    if (merged) {
      TORCH_CHECK(k.dim() == 3 && v.dim() == 3,
                  "if batch and heads are merged for queries, keys and values must also similarly have only 3 dimensions");
      B = q.size(0); Nq = q.size(1); D = q.size(2); H = 1; kv_heads = 1;
    } else {
      TORCH_CHECK(q.dim() == 4, "queries must be (batch, heads, seq, dim) or (batch*heads, seq, dim)");
      B = q.size(0); H = q.size(1); Nq = q.size(2); D = q.size(3);
      if (k.dim() == 3) {
        TORCH_CHECK(v.dim() == 3, "keys and values must both be single-headed");
        kv_heads = 1;
      } else {
        TORCH_CHECK(k.dim() == 4 && v.dim() == 4, "keys and values must be 3- or 4-dimensional");
        kv_heads = H;
      }
    }
    Nk = k.size(-2);
    TORCH_CHECK(k.size(-1) == D && v.size(-1) == D, "head dimensions of q, k, v must match");
    TORCH_CHECK(k.size(0) == B && v.size(0) == B, "batch sizes of q, k, v must match");
    TORCH_CHECK(v.size(-2) == Nk, "keys and values must have the same length");
    TORCH_CHECK(q.is_cuda() && k.is_cuda() && v.is_cuda(),
                "flash_cosine_sim_attention: CUDA tensors required - this build has no CPU path");
    TORCH_CHECK(k.scalar_type() == q.scalar_type() && v.scalar_type() == q.scalar_type(),
                "q, k, v must share one dtype");
  }
};

fcsa_problem make_problem(const Shapes& sh, const Tensor& q, double scale, double shift, bool causal,
                          const optional<Tensor>& mask_u8, bool out_f32 = false) {
  fcsa_problem p;
  p.out_f32 = out_f32 ? 1 : 0;
  p.reserved_ = 0;
  p.dtype = dtype_code(q);
  p.batch = (int32_t)sh.B; p.heads = (int32_t)sh.H; p.kv_heads = (int32_t)sh.kv_heads;
  p.seq_q = (int32_t)sh.Nq; p.seq_k = (int32_t)sh.Nk; p.head_dim = (int32_t)sh.D;
  p.causal = causal ? 1 : 0;
  p.scale = (float)scale; p.shift = (float)shift;
  if (mask_u8.has_value()) {
    p.key_mask = reinterpret_cast<const uint8_t*>(mask_u8->data_ptr());
    p.key_mask_stride = mask_u8->stride(0);
  } else {
    p.key_mask = nullptr;
    p.key_mask_stride = 0;
  }
  return p;
}

// key-padding mask (batch, seq_k), True = attend -> contiguous bytes
optional<Tensor> prep_mask(const optional<Tensor>& mask, const Shapes& sh) {
  if (!mask.has_value() || !mask->defined()) return c10::nullopt;
  TORCH_CHECK(mask->dim() == 2 && mask->size(0) == sh.B && mask->size(1) == sh.Nk, "mask must be (batch, seq_k) = (",
              sh.B, ", ", sh.Nk, ")");
  return mask->to(at::kBool).contiguous();
}

// (heads, i, j) - or (batch, i, j) when batch_dim - bias -> tensor in q's dtype whose rows are 16-byte
// aligned (row length padded to a multiple of 8) + the fcsa_bias addressing it as [batch][head][i][j]
Tensor prep_bias(const Tensor& bias, const Shapes& sh, at::ScalarType dtype, bool batch_dim) {
  const int64_t lead = batch_dim ? sh.B : sh.H;
  TORCH_CHECK(bias.dim() == 3 && bias.size(0) == lead && bias.size(1) == sh.Nq && bias.size(2) == sh.Nk,
              "attn_bias must be (", lead, ", ", sh.Nq, ", ", sh.Nk, ") (", batch_dim ? "batch" : "heads", ", i, j)");
  TORCH_CHECK(bias.is_cuda(), "attn_bias must be a CUDA tensor");
  Tensor t = bias.detach().to(dtype);
  const int64_t pad = (8 - sh.Nk % 8) % 8;
  if (pad) t = at::constant_pad_nd(t, {0, pad});
  t = t.contiguous();
  if (reinterpret_cast<uintptr_t>(t.data_ptr()) % 16) t = t.clone();
  return t;
}

fcsa_bias bias_struct(const Tensor& t, bool batch_dim, const optional<Tensor>& amax) {
  fcsa_bias b;
  b.ptr = t.data_ptr();
  const int64_t plane = t.stride(0);
  if (batch_dim) { b.sb = plane; b.sh = 0; } else { b.sb = 0; b.sh = plane; }
  b.sn = t.stride(1);
  b.amax = (amax.has_value() && amax->defined()) ? amax->data_ptr<float>() : nullptr;
  return b;
}

// Backward workspaces, one pair per (device, stream): `scratch` (contents irrelevant) and `zeroed`
// (fp32 dq accumulator + tile counters; zero-filled once, left zero by every backward).  Only grow.
struct Workspaces {
  Tensor scratch, zeroed;
};
std::mutex g_ws_mutex;
std::map<std::pair<int, void*>, Workspaces> g_ws;

Workspaces backward_workspaces(const fcsa_problem& p, const Tensor& like, cudaStream_t stream) {
  const size_t need_s = fcsa_backward_workspace_bytes(&p), need_z = fcsa_backward_zeroed_bytes(&p);
  TORCH_CHECK(need_s > 0 && need_z > 0, "libfcsa_b200: ", fcsa_last_error());
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  Workspaces& w = g_ws[{(int)like.get_device(), (void*)stream}];
  const auto opts = like.options().dtype(at::kByte);
  if (!w.scratch.defined() || (size_t)w.scratch.numel() < need_s) w.scratch = at::empty({(int64_t)need_s}, opts);
  if (!w.zeroed.defined() || (size_t)w.zeroed.numel() < need_z) {
    w.zeroed = at::empty({(int64_t)need_z}, opts);
    FCSA_CHECK(fcsa_zeroed_init(w.zeroed.data_ptr(), need_z, stream));
  }
  return w;
}

// ------------------------------------------------------------------------------------------------
// l2norm (flash_cosine_sim_attention.py:38-65)
// ------------------------------------------------------------------------------------------------
std::tuple<Tensor, Tensor> l2norm_forward(const Tensor& x_in, int64_t groups) {
  TORCH_CHECK(x_in.is_cuda() && (x_in.dim() == 3 || x_in.dim() == 4), "l2norm: 3-D or 4-D CUDA tensor expected");
  const c10::cuda::CUDAGuard guard(x_in.device());
  const Tensor x = tma_ready(x_in);
  const int64_t B = x.size(0), H = x.dim() == 4 ? x.size(1) : 1, N = x.size(-2), D = x.size(-1);
  Tensor y = at::empty(x.sizes(), x.options());
  Tensor rnorm = at::empty({B, H, N, groups}, x.options().dtype(at::kFloat));
  const fcsa_tensor tx = view4(x), ty = view4(y);
  FCSA_CHECK(fcsa_l2norm_forward(dtype_code(x), (int32_t)B, (int32_t)H, (int32_t)N, (int32_t)D, (int32_t)groups, &tx,
                                 &ty, rnorm.data_ptr<float>(), at::cuda::getCurrentCUDAStream().stream()));
  return {y, rnorm};
}

Tensor l2norm_backward(const Tensor& dy_in, const Tensor& y_in, const Tensor& rnorm, int64_t groups) {
  const c10::cuda::CUDAGuard guard(y_in.device());
  const Tensor dy = tma_ready(dy_in), y = tma_ready(y_in);
  const int64_t B = y.size(0), H = y.dim() == 4 ? y.size(1) : 1, N = y.size(-2), D = y.size(-1);
  Tensor dx = at::empty(y.sizes(), y.options());
  const fcsa_tensor tdy = view4(dy), ty = view4(y), tdx = view4(dx);
  FCSA_CHECK(fcsa_l2norm_backward(dtype_code(y), (int32_t)B, (int32_t)H, (int32_t)N, (int32_t)D, (int32_t)groups,
                                  &tdy, &ty, rnorm.data_ptr<float>(), &tdx,
                                  at::cuda::getCurrentCUDAStream().stream()));
  return dx;
}

// ------------------------------------------------------------------------------------------------
// float32 front end: (l2norm +) cast of a float32 tensor to the 16-bit operand dtype, and its backward
// ------------------------------------------------------------------------------------------------
// x: float32 (b, h, n, d) or (b, n, d); returns y (16 bit, `pad_to` >= d features, zero-padded) and rnorm (groups > 0)
std::tuple<Tensor, Tensor> f32_cast(const Tensor& x_in, bool to_bf16, int64_t groups, const optional<Tensor>& mul,
                                    bool mul_reciprocal, int64_t pad_to) {
  TORCH_CHECK(x_in.is_cuda() && x_in.scalar_type() == at::kFloat && (x_in.dim() == 3 || x_in.dim() == 4),
              "f32_cast: 3-D or 4-D float32 CUDA tensor expected");
  const c10::cuda::CUDAGuard guard(x_in.device());
  const Tensor x = (x_in.stride(-1) == 1 && reinterpret_cast<uintptr_t>(x_in.data_ptr()) % 16 == 0) ? x_in : x_in.contiguous();
  const int64_t B = x.size(0), H = x.dim() == 4 ? x.size(1) : 1, N = x.size(-2), D = x.size(-1);
  std::vector<int64_t> shape = x.sizes().vec();
  shape.back() = std::max(pad_to, D);
  const auto hopt = x.options().dtype(to_bf16 ? at::kBFloat16 : at::kHalf);
  Tensor y = shape.back() == D ? at::empty(shape, hopt) : at::zeros(shape, hopt);
  Tensor rnorm;
  if (groups > 0) rnorm = at::empty({B, H, N, groups}, x.options());
  const fcsa_tensor tx = view4(x), ty = view4(y);
  FCSA_CHECK(fcsa_f32_cast(to_bf16 ? FCSA_BF16 : FCSA_F16, (int32_t)B, (int32_t)H, (int32_t)N, (int32_t)D, (int32_t)groups,
                           &tx, &ty, groups > 0 ? rnorm.data_ptr<float>() : nullptr,
                           (mul.has_value() && mul->defined()) ? mul->data_ptr<float>() : nullptr, mul_reciprocal ? 1 : 0,
                           at::cuda::getCurrentCUDAStream().stream()));
  return {y, rnorm};
}

// dy: float32 gradient w.r.t. y (may carry padded features: only the first d are read); y / rnorm: what f32_cast
// returned (groups > 0); returns the float32 gradient w.r.t. x with d features
Tensor f32_cast_backward(const Tensor& dy_in, const optional<Tensor>& y, const optional<Tensor>& rnorm, int64_t groups,
                         const optional<Tensor>& mul, bool mul_reciprocal, int64_t d) {
  TORCH_CHECK(dy_in.is_cuda() && dy_in.scalar_type() == at::kFloat, "f32_cast_backward: float32 CUDA gradient expected");
  const c10::cuda::CUDAGuard guard(dy_in.device());
  const Tensor dy = (dy_in.stride(-1) == 1 && reinterpret_cast<uintptr_t>(dy_in.data_ptr()) % 16 == 0) ? dy_in : dy_in.contiguous();
  const int64_t B = dy.size(0), H = dy.dim() == 4 ? dy.size(1) : 1, N = dy.size(-2);
  std::vector<int64_t> shape = dy.sizes().vec();
  shape.back() = d;
  Tensor dx = at::empty(shape, dy.options());
  const fcsa_tensor tdy = view4(dy), tdx = view4(dx);
  fcsa_tensor ty = {nullptr, 0, 0, 0};
  bool bf = false;
  if (groups > 0) {
    TORCH_CHECK(y.has_value() && y->defined() && rnorm.has_value() && rnorm->defined(), "f32_cast_backward: y and rnorm needed");
    ty = view4(*y);
    bf = y->scalar_type() == at::kBFloat16;
  }
  FCSA_CHECK(fcsa_f32_cast_backward(bf ? FCSA_BF16 : FCSA_F16, (int32_t)B, (int32_t)H, (int32_t)N, (int32_t)d,
                                    (int32_t)groups, &tdy, &ty, groups > 0 ? rnorm->data_ptr<float>() : nullptr, &tdx,
                                    (mul.has_value() && mul->defined()) ? mul->data_ptr<float>() : nullptr,
                                    mul_reciprocal ? 1 : 0, at::cuda::getCurrentCUDAStream().stream()));
  return dx;
}

// ------------------------------------------------------------------------------------------------
// forward_ex: the general forward.
//   l2norm_groups > 0 : q, k are RAW; they are normalised over that many groups by the fused pre-pass
//                       (returns q_hat, k_hat, q_rnorm, k_rnorm for the backward); no bias on this path
//   l2norm_groups == 0: q, k are used as given (already normalised, or l2norm_qk=False)
//   bias_prepared     : the tensor prep_bias() made (saved by the caller for the backward), optional
//   bias_amax         : optional fp32 device scalar; the kernels add max(amax, 0) to `shift` (fp16 range)
//   out_f32           : o is returned as float32 (fp32 accumulator, no rounding to 16 bit) - float32 callers
// returns (o, inv_l, q_hat, k_hat, q_rnorm, k_rnorm) - the last four undefined when l2norm_groups == 0
// ------------------------------------------------------------------------------------------------
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> forward_ex(
    const Tensor& q_in, const Tensor& k_in, const Tensor& v_in, const optional<Tensor>& mask,
    const optional<Tensor>& bias_prepared, bool bias_batch_dim, const optional<Tensor>& bias_amax, double scale,
    double shift, bool causal, int64_t l2norm_groups, bool need_inv_l, bool out_f32) {
  const Shapes sh(q_in, k_in, v_in);
  TORCH_CHECK(!(causal && mask.has_value() && mask->defined()), "mask should not be supplied if causality is needed");
  const c10::cuda::CUDAGuard guard(q_in.device());
  const Tensor q = tma_ready(q_in), k = tma_ready(k_in), v = tma_ready(v_in);
  const optional<Tensor> mask_u8 = prep_mask(mask, sh);
  const fcsa_problem p = make_problem(sh, q, scale, shift, causal, mask_u8, out_f32);
  Tensor o = at::empty(q.sizes(), out_f32 ? q.options().dtype(at::kFloat) : q.options());   // out_f32: no final rounding
  Tensor inv_l;
  if (need_inv_l) inv_l = at::empty({sh.B, sh.H, sh.Nq}, q.options().dtype(at::kFloat));
  float* inv_l_ptr = need_inv_l ? inv_l.data_ptr<float>() : nullptr;
  const fcsa_tensor tq = view4(q), tk = view4(k), tv = view4(v), to = view4(o);
  cudaStream_t stream = at::cuda::getCurrentCUDAStream().stream();
  const bool has_bias = bias_prepared.has_value() && bias_prepared->defined();
  Tensor qn, kn, rq, rk;
  if (l2norm_groups > 0) {
    TORCH_CHECK(!has_bias, "forward_ex: the fused-l2norm path takes no attn_bias (normalise first)");
    qn = at::empty(q.sizes(), q.options());
    kn = at::empty(k.sizes(), k.options());
    rq = at::empty({sh.B, sh.H, sh.Nq, l2norm_groups}, q.options().dtype(at::kFloat));
    rk = at::empty({sh.B, sh.kv_heads, sh.Nk, l2norm_groups}, q.options().dtype(at::kFloat));
    fcsa_l2norm n;
    n.groups = (int32_t)l2norm_groups;
    n.q_hat = view4(qn); n.k_hat = view4(kn);
    n.q_rnorm = rq.data_ptr<float>(); n.k_rnorm = rk.data_ptr<float>();
    FCSA_CHECK(fcsa_forward_fused(&p, &tq, &tk, &tv, &n, &to, inv_l_ptr, stream));
  } else if (has_bias) {
    const fcsa_bias b = bias_struct(*bias_prepared, bias_batch_dim, bias_amax);
    FCSA_CHECK(fcsa_forward_bias(&p, &tq, &tk, &tv, &b, &to, inv_l_ptr, stream));
  } else {
    FCSA_CHECK(fcsa_forward(&p, &tq, &tk, &tv, &to, inv_l_ptr, stream));
  }
  return {o, inv_l, qn, kn, rq, rk};
}

// ------------------------------------------------------------------------------------------------
// backward_ex: gradients of forward_ex.
//   groups > 0 (with q_rnorm, k_rnorm): q, k are the NORMALISED tensors forward_ex returned and dq, dk
//                       are gradients w.r.t. the raw ones (l2norm backward fused)
//   bias_prepared + bias_grad: d_bias accumulated in fp32, returned in `bias_dtype_like`'s dtype / shape
// returns (dq, dk, dv, d_bias or undefined)
// ------------------------------------------------------------------------------------------------
std::tuple<Tensor, Tensor, Tensor, Tensor> backward_ex(
    const Tensor& d_out, const Tensor& o_in, const Tensor& inv_l, const Tensor& q_in, const Tensor& k_in,
    const Tensor& v_in, const optional<Tensor>& q_rnorm, const optional<Tensor>& k_rnorm,
    const optional<Tensor>& mask, const optional<Tensor>& bias_prepared, bool bias_batch_dim,
    const optional<Tensor>& bias_amax, bool bias_grad, const optional<Tensor>& bias_like, double scale, double shift,
    bool causal, int64_t groups, bool out_f32) {
  const Shapes sh(q_in, k_in, v_in);
  const c10::cuda::CUDAGuard guard(q_in.device());
  const Tensor q = tma_ready(q_in), k = tma_ready(k_in), v = tma_ready(v_in), o = tma_ready(o_in),
               d_o = tma_ready(d_out);
  TORCH_CHECK(d_o.sizes() == q.sizes() && d_o.scalar_type() == q.scalar_type(),
              "d_out must have the shape and dtype of the queries");
  const optional<Tensor> mask_u8 = prep_mask(mask, sh);
  const fcsa_problem p = make_problem(sh, q, scale, shift, causal, mask_u8, out_f32);
  TORCH_CHECK(o.scalar_type() == (out_f32 ? at::kFloat : q.scalar_type()), "backward_ex: `o` must be ",
              out_f32 ? "float32 (out_f32)" : "in the operand dtype");
  const auto gopt = out_f32 ? q.options().dtype(at::kFloat) : q.options();     // out_f32: float32 gradients
  Tensor dq = at::empty(q.sizes(), gopt);
  Tensor dk = at::empty(k.sizes(), gopt);
  Tensor dv = at::empty(v.sizes(), gopt);
  cudaStream_t stream = at::cuda::getCurrentCUDAStream().stream();
  const Workspaces ws = backward_workspaces(p, q, stream);
  const fcsa_tensor tq = view4(q), tk = view4(k), tv = view4(v), to = view4(o), tdo = view4(d_o), tdq = view4(dq),
                    tdk = view4(dk), tdv = view4(dv);
  void* wsp = ws.scratch.data_ptr();
  void* zsp = ws.zeroed.data_ptr();
  const size_t wsn = (size_t)ws.scratch.numel(), zsn = (size_t)ws.zeroed.numel();
  const bool has_bias = bias_prepared.has_value() && bias_prepared->defined();
  const bool fused = q_rnorm.has_value() && q_rnorm->defined();
  Tensor db;
  if (fused) {
    TORCH_CHECK(!has_bias && k_rnorm.has_value() && k_rnorm->defined() && groups > 0, "backward_ex: bad fused-l2norm arguments");
    fcsa_l2norm n;
    n.groups = (int32_t)groups;
    n.q_hat = tq; n.k_hat = tk;
    n.q_rnorm = q_rnorm->data_ptr<float>(); n.k_rnorm = k_rnorm->data_ptr<float>();
    FCSA_CHECK(fcsa_backward_fused(&p, &n, &tv, &to, &tdo, inv_l.data_ptr<float>(), &tdq, &tdk, &tdv, wsp, wsn, zsp,
                                   zsn, stream));
  } else if (has_bias) {
    const fcsa_bias b = bias_struct(*bias_prepared, bias_batch_dim, bias_amax);
    Tensor db_acc;
    float* db_ptr = nullptr;
    const int64_t plane = sh.Nq * sh.Nk;
    if (bias_grad) {
      db_acc = at::zeros({bias_batch_dim ? sh.B : sh.H, sh.Nq, sh.Nk}, q.options().dtype(at::kFloat));
      db_ptr = db_acc.data_ptr<float>();
    }
    FCSA_CHECK(fcsa_backward_bias(&p, &tq, &tk, &tv, &to, &tdo, inv_l.data_ptr<float>(), &b, db_ptr,
                                  bias_batch_dim ? plane : 0, bias_batch_dim ? 0 : plane, &tdq, &tdk, &tdv, wsp, wsn,
                                  zsp, zsn, stream));
    if (bias_grad)
      db = (bias_like.has_value() && bias_like->defined()) ? db_acc.to(bias_like->scalar_type()) : db_acc;
  } else {
    FCSA_CHECK(fcsa_backward(&p, &tq, &tk, &tv, &to, &tdo, inv_l.data_ptr<float>(), &tdq, &tdk, &tdv, wsp, wsn, zsp,
                             zsn, stream));
  }
  return {dq, dk, dv, db};
}

// ------------------------------------------------------------------------------------------------
// the reference's surface (cu:1630-1639, 1752-1764, 1921): q, k already normalised, shift = scale
// ------------------------------------------------------------------------------------------------
std::tuple<Tensor, Tensor, bool> forward(const Tensor& q, const Tensor& k, const Tensor& v,
                                         const optional<Tensor>& mask, const optional<Tensor>& attn_bias,
                                         bool attn_bias_batch_dim, double scale, bool causal) {
  const Shapes sh(q, k, v);
  if (sh.merged) attn_bias_batch_dim = true;   // cu:1647-1654
  const bool has_bias = attn_bias.has_value() && attn_bias->defined();
  const bool should_backwards =
      q.requires_grad() || k.requires_grad() || v.requires_grad() || (has_bias && attn_bias->requires_grad());
  optional<Tensor> bias_prepared;
  if (has_bias) bias_prepared = prep_bias(*attn_bias, sh, q.scalar_type(), attn_bias_batch_dim);
  auto r = forward_ex(q, k, v, mask, bias_prepared, attn_bias_batch_dim, c10::nullopt, scale, scale, causal, 0, true,
                      false);
  return {std::get<0>(r), std::get<1>(r), should_backwards};
}

std::tuple<Tensor, Tensor, Tensor, optional<Tensor>> backward(
    const Tensor& d_out, const Tensor& o, const Tensor& l, const Tensor& q, const Tensor& k, const Tensor& v,
    const optional<Tensor>& mask, const optional<Tensor>& attn_bias, bool attn_bias_batch_dim, double scale,
    bool causal) {
  const Shapes sh(q, k, v);
  if (sh.merged) attn_bias_batch_dim = true;
  const bool has_bias = attn_bias.has_value() && attn_bias->defined();
  optional<Tensor> bias_prepared;
  if (has_bias) bias_prepared = prep_bias(*attn_bias, sh, q.scalar_type(), attn_bias_batch_dim);
  auto r = backward_ex(d_out, o, l, q, k, v, c10::nullopt, c10::nullopt, mask, bias_prepared, attn_bias_batch_dim,
                       c10::nullopt, has_bias && attn_bias->requires_grad(), attn_bias, scale, scale, causal, 0, false);
  optional<Tensor> db;
  if (std::get<3>(r).defined()) db = std::get<3>(r);
  return {std::get<0>(r), std::get<1>(r), std::get<2>(r), db};
}

// Drops the cached backward workspaces of every (device, stream) (they only ever grow: a long-context call leaves a
// 1 GB accumulator behind).  Returns the number of bytes released to the caching allocator.  Safe at any time between
// calls: buffers still referenced by enqueued kernels stay alive through the allocator's stream semantics.
int64_t release_workspaces() {
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  int64_t bytes = 0;
  for (auto& kv : g_ws) {
    if (kv.second.scratch.defined()) bytes += kv.second.scratch.numel();
    if (kv.second.zeroed.defined()) bytes += kv.second.zeroed.numel();
  }
  g_ws.clear();
  return bytes;
}

// the reference's debug() is an empty hook (cu:1921); this one reports the library's launch counter
int64_t debug() { return fcsa_debug(); }

Tensor prepare_bias(const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& attn_bias, bool batch_dim) {
  const Shapes sh(q, k, v);
  return prep_bias(attn_bias, sh, q.scalar_type(), batch_dim || sh.merged);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "B200 (sm_100a) fused cosine-similarity attention: torch binding of libfcsa_b200.so";
  m.def("forward", &forward, "Flash Cosine-Sim Attention Forward");
  m.def("backward", &backward, "Flash Cosine-Sim Attention Backward");
  m.def("debug", &debug, "Debug");
  m.def("forward_ex", &forward_ex, "forward with explicit shift / fused l2norm / prepared bias");
  m.def("backward_ex", &backward_ex, "backward of forward_ex");
  m.def("prepare_bias", &prepare_bias, "attn_bias -> padded, aligned tensor in the problem dtype");
  m.def("f32_cast", &f32_cast, "float32 -> (l2norm +) scaled cast to the 16-bit operand dtype");
  m.def("f32_cast_backward", &f32_cast_backward, "backward of f32_cast: float32 gradient w.r.t. the float32 input");
  m.def("l2norm_forward", &l2norm_forward, "grouped l2norm: x -> (y, 1/norm)");
  m.def("l2norm_backward", &l2norm_backward, "grouped l2norm backward from the normalised y");
  m.def("release_workspaces", &release_workspaces, "free the cached backward workspaces; returns the bytes released");
  m.def("abi_version", []() { return fcsa_version(); });
  m.def("_error_path_selftest", []() { FCSA_CHECK(fcsa_forward(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr)); },
        "raises the library's error for a null problem (exercises the C ABI error -> Python exception path)");
}
