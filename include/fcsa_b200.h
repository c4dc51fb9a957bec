/*
 * fcsa_b200.h - C ABI of libfcsa_b200.so, the B200 (sm_100a) fused cosine-similarity
 * attention library.
 *
 * This is the drop-in boundary for the reference's CUDA extension
 * (lucidrains/flash-cosine-sim-attention @ v0.1.40).  Each entry point names the reference
 * interface it replaces (file:line, paths relative to the reference repository):
 *
 *   fcsa_forward            <- flash_cosine_sim_attention_forward   flash_cosine_sim_attention_cuda.cu:1630-1748
 *                              (pybind `forward`, cu:1928-1933; called from flash_cosine_sim_attention.py:247-256)
 *   fcsa_backward           <- flash_cosine_sim_attention_backward  cu:1752-1917
 *                              (pybind `backward`; called from flash_cosine_sim_attention.py:281-302)
 *   fcsa_debug              <- debug()                              cu:1921
 *   fcsa_l2norm_forward /   <- l2norm / grouped_l2norm / l2norm_tensors  flash_cosine_sim_attention.py:38-65
 *   fcsa_l2norm_backward       (PyTorch F.normalize + autograd in the reference; fused kernels here)
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every tensor is a DEVICE pointer to 16-bit data (FCSA_F16 / FCSA_BF16) addressed as
 *     [batch][head][row][feature] with element strides (sb, sh, sn) and a contiguous feature
 *     dimension; pointers and strides*2 must be multiples of 16 bytes (TMA requirement)
 *   - `stream` is a cudaStream_t; all work is enqueued on it and nothing synchronises
 *     (the reference device-synchronised after every call, cu:1745/1889 - deliberately not kept)
 *   - inputs are borrowed and never written; outputs are caller-allocated
 *   - return value: FCSA_OK or an error code; fcsa_last_error() returns a thread-local message
 *     (the reference only printed CUDA errors to stderr, cu:17-28)
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 */
#ifndef FCSA_B200_H
#define FCSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FCSA_F16 = 0, FCSA_BF16 = 1 };

enum {
  FCSA_OK = 0,
  FCSA_ERR_INVALID = 1,      /* bad argument (shape, alignment, null pointer)           */
  FCSA_ERR_UNSUPPORTED = 2,  /* valid request this build has no kernel for              */
  FCSA_ERR_CUDA = 3,         /* a CUDA runtime / driver call failed                     */
  FCSA_ERR_WORKSPACE = 4     /* workspace too small                                     */
};

/* A (batch, head, row, feature) view with a contiguous feature dimension. */
typedef struct fcsa_tensor {
  void* ptr;
  int64_t sb, sh, sn; /* element strides of batch, head, row */
} fcsa_tensor;

/*
 * One attention problem.  Mirrors the argument list of the reference's forward/backward
 * (cu:1630-1639, cu:1752-1764) after its shape canonicalisation (cu:1647-1660):
 *   - merged batch-heads (3-D q) is expressed as heads = 1
 *   - single-head keys/values (3-D k, v with 4-D q; cu:1679) is kv_heads = 1
 */
typedef struct fcsa_problem {
  int32_t dtype;     /* FCSA_F16 | FCSA_BF16 (q, k, v, o, do, dq, dk, dv all share it)        */
  int32_t batch;
  int32_t heads;
  int32_t kv_heads;  /* == heads, or 1 for keys/values shared across heads                   */
  int32_t seq_q;
  int32_t seq_k;
  int32_t head_dim;  /* 64 or 128                                                             */
  int32_t causal;    /* bottom-right aligned (cu:1097, cu:1210): key j visible iff j <= i + seq_k - seq_q */
  float scale;       /* logits = scale * <q, k>                                               */
  float shift;       /* p = exp(logit - shift); the reference uses shift = scale (cu:1216)    */
  const uint8_t* key_mask;   /* (batch, seq_k) bytes, nonzero = attend; NULL = no mask (cu:1198-1212) */
  int64_t key_mask_stride;   /* bytes between batches                                         */
  int32_t out_f32;   /* 0: o (forward; also the `o` the backward reads), dq, dk, dv have the problem dtype.
                        1: they are float32 tensors (strides then count float32 elements) - the accumulators are
                        fp32 anyway, this only skips the final rounding to 16 bit.  q, k, v, d_o stay 16-bit
                        operands.  Used for float32 callers (reference: Float dispatch, cu:1702-1703)        */
  int32_t reserved_;
} fcsa_problem;

/* Library / ABI version: major*10000 + minor*100 + patch. */
int fcsa_version(void);

/* Message describing the last error on the calling thread ("" if none). */
const char* fcsa_last_error(void);

/* Reference `debug()` (cu:1921): a no-op hook.  Here it returns the number of kernels this
 * library has launched in the current process (bench.py reports it as gpu_launches). */
int64_t fcsa_debug(void);

/*
 * o = softmax-like(q k^T) v with the fixed-shift formulation; inv_l[b][h][i] = 1 / max(l_i, 1e-37)
 * is the saved normaliser the backward needs (the reference's `l` output, cu:1698/1239).
 * inv_l: (batch, heads, seq_q) fp32 contiguous, may be NULL when no backward will follow.
 * mask and causal are mutually exclusive (flash_cosine_sim_attention.py:88, cu:1675).
 * Rows with no visible key produce o = 0 (cu:1239, reference behaviour of the fused kernel).
 */
int fcsa_forward(const fcsa_problem* p, const fcsa_tensor* q, const fcsa_tensor* k,
                 const fcsa_tensor* v, const fcsa_tensor* o, float* inv_l, void* stream);

/* Bytes of SCRATCH fcsa_backward needs for this problem (per-row constants, shared-kv accumulators):
 * contents irrelevant on entry, garbage on exit. */
size_t fcsa_backward_workspace_bytes(const fcsa_problem* p);

/* Bytes of the ZEROED workspace of fcsa_backward (the fp32 dq accumulator tiles).
 * Contract: every byte is zero when a backward call starts and zero again when its kernels have
 * finished - the pass that converts an accumulator tile to dq also clears it - so one buffer,
 * zero-filled ONCE (fcsa_zeroed_init), serves any number of calls of any shape enqueued on the same
 * stream.  This is what replaces the reference's per-call 33.5 MB memset of its fp32 dq (cu:1818).
 * Do not share one buffer between streams that may run concurrently. */
size_t fcsa_backward_zeroed_bytes(const fcsa_problem* p);

/* Zero-fill a freshly allocated (or grown) zeroed workspace: cudaMemsetAsync on `stream`. */
int fcsa_zeroed_init(void* zeroed, size_t bytes, void* stream);

/*
 * Gradients of fcsa_forward w.r.t. q, k, v (the q, k given are the already-normalised ones,
 * exactly as in the reference: cu:1487-1626).  dq/dk/dv are caller-allocated in the problem
 * dtype; with kv_heads == 1, dk/dv have a single head and receive the sum over heads
 * (cu:1613-1619).  `workspace` must hold fcsa_backward_workspace_bytes(p) bytes, `zeroed`
 * fcsa_backward_zeroed_bytes(p) bytes that are all zero (see above).
 */
int fcsa_backward(const fcsa_problem* p, const fcsa_tensor* q, const fcsa_tensor* k,
                  const fcsa_tensor* v, const fcsa_tensor* o, const fcsa_tensor* d_o,
                  const float* inv_l, const fcsa_tensor* dq, const fcsa_tensor* dk,
                  const fcsa_tensor* dv, void* workspace, size_t workspace_bytes, void* zeroed,
                  size_t zeroed_bytes, void* stream);

/*
 * y = x / max(||x||_2, 1e-12) over `groups` equal chunks of the feature dimension
 * (flash_cosine_sim_attention.py:38-55).  x: strided view; y: same shape (any strides);
 * rnorm: (batch, heads, rows, groups) fp32 contiguous = 1 / max(||x||, eps), may be NULL.
 */
int fcsa_l2norm_forward(int32_t dtype, int32_t batch, int32_t heads, int32_t rows, int32_t head_dim,
                        int32_t groups, const fcsa_tensor* x, const fcsa_tensor* y, float* rnorm,
                        void* stream);

/*
 * dx = (dy - y <y, dy>_group) * rnorm_group : the l2norm backward given the NORMALISED y.
 * dy, y, dx: 16-bit strided views.
 */
int fcsa_l2norm_backward(int32_t dtype, int32_t batch, int32_t heads, int32_t rows,
                         int32_t head_dim, int32_t groups, const fcsa_tensor* dy,
                         const fcsa_tensor* y, const float* rnorm, const fcsa_tensor* dx,
                         void* stream);

/*
 * Fused variants: the grouped l2-normalisation of q and k that the reference performs with PyTorch
 * ops around its kernels (flash_cosine_sim_attention.py:320-321, l2norm_tensors py:57-65) and its
 * autograd backward run inside the same call.
 *   q_hat, k_hat : normalised q / k in the problem dtype - outputs of fcsa_forward_fused, inputs of
 *                  fcsa_backward_fused (the reference saves exactly these for its backward, py:270)
 *   q_rnorm      : (batch, heads, seq_q, groups)    fp32 = 1 / max(||q||_group, 1e-12)
 *   k_rnorm      : (batch, kv_heads, seq_k, groups) fp32
 * head_dim / groups must be a power of two.
 */
typedef struct fcsa_l2norm {
  int32_t groups;
  fcsa_tensor q_hat, k_hat;
  float* q_rnorm;
  float* k_rnorm;
} fcsa_l2norm;

/* q_hat, k_hat = l2norm(q), l2norm(k); o, inv_l = fcsa_forward(q_hat, k_hat, v). */
int fcsa_forward_fused(const fcsa_problem* p, const fcsa_tensor* q, const fcsa_tensor* k,
                       const fcsa_tensor* v, const fcsa_l2norm* n, const fcsa_tensor* o, float* inv_l,
                       void* stream);

/* Gradients w.r.t. the RAW q, k (and v): fcsa_backward followed by the l2norm backward, the latter
 * folded into the dq conversion and the dk epilogue.  Same workspaces as fcsa_backward. */
int fcsa_backward_fused(const fcsa_problem* p, const fcsa_l2norm* n, const fcsa_tensor* v,
                        const fcsa_tensor* o, const fcsa_tensor* d_o, const float* inv_l,
                        const fcsa_tensor* dq, const fcsa_tensor* dk, const fcsa_tensor* dv,
                        void* workspace, size_t workspace_bytes, void* zeroed, size_t zeroed_bytes,
                        void* stream);

/* ---- float32 front end ---------------------------------------------------------------------------
 * The reference dispatches Float in forward and backward (cu:1702-1703, 1832-1834).  Here float32 callers run the
 * 16-bit tensor-core kernels (fp32 accumulation, `out_f32` results); these two passes are all that surrounds them.
 * `x` / `dy` / `dx` are FLOAT32 tensors (fcsa_tensor with float32 element strides, feature dim contiguous, rows
 * 16-byte aligned), `y` is the 16-bit tensor (`dtype`), which may be wider than head_dim (zero-padded features are
 * the caller's business: only head_dim features per row are written / read).
 *   fcsa_f32_cast          y = round16( l2norm_groups(x) * m )   groups > 0 ;  round16( x * m )   groups == 0
 *   fcsa_f32_cast_backward dx = (dy - y <y,dy>_group) * rnorm_group * m   groups > 0 ;  dy * m   groups == 0
 * m = 1 if `mul` is NULL, else *mul (a DEVICE scalar - e.g. a power-of-two range scale chosen on the device, the
 * host never reads it), or 1 / *mul when mul_reciprocal != 0.  rnorm: (batch, heads, rows, groups) fp32.
 * head_dim must be 16, 32, 64 or 128 and head_dim / groups a power of two. */
int fcsa_f32_cast(int32_t dtype, int32_t batch, int32_t heads, int32_t rows, int32_t head_dim, int32_t groups,
                  const fcsa_tensor* x, const fcsa_tensor* y, float* rnorm, const float* mul, int32_t mul_reciprocal,
                  void* stream);
int fcsa_f32_cast_backward(int32_t dtype, int32_t batch, int32_t heads, int32_t rows, int32_t head_dim,
                           int32_t groups, const fcsa_tensor* dy, const fcsa_tensor* y, const float* rnorm,
                           const fcsa_tensor* dx, const float* mul, int32_t mul_reciprocal, void* stream);

/* ---- additive attention bias -------------------------------------------------------------------
 * Replaces the `attn_bias` / `attn_bias_batch_dim` arguments of the reference's forward / backward
 * (flash_cosine_sim_attention_cuda.cu:1630-1639, 1752-1764; added to the logits at cu:1168,1214 and
 * cu:1474-1476, gradient cu:1574-1576).  Elements have the problem's dtype; the bias is addressed as
 * [batch][head][query][key] with ELEMENT strides sb (0 when the bias has no batch dimension -
 * attn_bias_batch_dim == false), sh, sn and contiguous keys.  Rows must be 16-byte aligned: sb, sh, sn
 * multiples of 8 and sn >= seq_k rounded up to 8 (pad the rows; the padding is never used). */
typedef struct fcsa_bias {
  const void* ptr;
  int64_t sb, sh, sn;
  /* Optional DEVICE pointer to one fp32: an upper bound of the bias values (e.g. its max).  The kernels
   * add max(*amax, 0) to the problem's `shift`, which keeps p = exp(logit - shift) inside the fp16 range
   * whatever the bias adds - without the host ever reading the bias (no device synchronisation).
   * Pass the same pointer to the forward and the backward; NULL = use `shift` as is. */
  const float* amax;
} fcsa_bias;

/* fcsa_forward with the bias added to scale * q.k before the exponential. */
int fcsa_forward_bias(const fcsa_problem* p, const fcsa_tensor* q, const fcsa_tensor* k,
                      const fcsa_tensor* v, const fcsa_bias* bias, const fcsa_tensor* o, float* inv_l,
                      void* stream);

/* fcsa_backward with the bias.  d_bias_acc: fp32, zero-filled by the caller, same index space as the bias
 * with element strides (dsb, dsh) over batch / head and contiguous [seq_q][seq_k] planes (dsb = 0 sums
 * the gradient over the batch, as the reference does for a bias without batch dimension, cu:1574-1576);
 * NULL when the bias needs no gradient (reference: db is empty unless it requires grad, cu:1827). */
int fcsa_backward_bias(const fcsa_problem* p, const fcsa_tensor* q, const fcsa_tensor* k,
                       const fcsa_tensor* v, const fcsa_tensor* o, const fcsa_tensor* d_o,
                       const float* inv_l, const fcsa_bias* bias, float* d_bias_acc, int64_t dsb,
                       int64_t dsh, const fcsa_tensor* dq, const fcsa_tensor* dk, const fcsa_tensor* dv,
                       void* workspace, size_t workspace_bytes, void* zeroed, size_t zeroed_bytes,
                       void* stream);

/*
 * Measurement hook (bench.py roofline line; no reference counterpart - the reference only timed
 * whole calls, flash_cosine_sim_attention/benchmark.py:7-58).  While set, fcsa_forward (which = 0)
 * or fcsa_backward (which = 1) records the cudaEvent_t `start` / `stop` on the launch stream
 * immediately before / after its dominant kernel (the tcgen05 attention kernel, not the
 * pre/post passes); which = 2, 3, 4 do the same for the l2norm(q,k) pass of fcsa_forward_fused, the
 * backward preprocess and the dq finish pass.  An event between two launches suspends their programmatic
 * dependent launch, so a hooked run is for attribution, not for the headline time.  Pass NULL, NULL to clear.  State is process-wide (autograd calls the
 * backward from its own thread), so use it from single-stream measurement code only.
 */
int fcsa_set_kernel_events(int32_t which, void* start_event, void* stop_event);

#ifdef __cplusplus
}
#endif

#endif /* FCSA_B200_H */
