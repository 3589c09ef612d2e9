// Native launch sequencers: one C call enqueues every kernel of the transformer layers of a stack
// (ResidualAttentionBlock.forward, model/base/transformer.py:126-146, and its backward), instead of one
// Python -> ctypes round trip per kernel. At the reference's batch size (about 3 clips per step) the step is bound by
// the host's launch rate, not by the GPU; the kernels and their order are exactly those of
// titok_video_b200/engine.py::_layers and titok_video_b200/backward.py::_layers_train / _layers_backward, which remain
// as the per-kernel paths (used when a per-kernel profiler is installed) and are tested to agree.
#include "../../include/titok_b200.h"

#include <cuda_runtime.h>

namespace {

inline const void* P(int64_t v) { return reinterpret_cast<const void*>(static_cast<uintptr_t>(v)); }
inline const float* PF(int64_t v) { return reinterpret_cast<const float*>(static_cast<uintptr_t>(v)); }
inline float* PFm(int64_t v) { return reinterpret_cast<float*>(static_cast<uintptr_t>(v)); }

enum { W_QKV = 0, W_OUT, W_12, W_3, LN_PRE, LN_ATTN_POST, LN_FFN, LN_FFD_POST, LN_NEXT, W_COLS };
enum { F_QKV = 0, F_ATT, F_O, F_YA, F_XF, F_XNF, F_H12, F_H, F_YF, F_XN, F_XNN, F_COLS };
enum { B_DUF = 0, B_DH, B_DH12, B_DXN, B_GF, B_DUA, B_DATT, B_DQKV, B_DO, B_G0, B_G1, B_COLS };
enum { G_FFD_POST = 0, G_W3, G_W12, G_FFN, G_ATTN_POST, G_WOUT, G_WQKV, G_PRE, G_COLS };

#define TTK_TRY(expr)            \
  do {                           \
    int _e = (expr);             \
    if (_e != 0) return _e;      \
  } while (0)

}  // namespace

extern "C" {

// Inference layers (engine._layers): x, xn updated in place; qkv / att / h (and y when the fused residual kernel is not
// applicable, i.e. width != 256 or y != NULL) are scratch buffers.
int ttk_layers_fwd(const ttk_layers_desc* d, void* x, void* xn, void* qkv, void* att, void* h, void* y, ttk_stream_t st) {
  if (!d || !d->weights || !x || !xn || !qkv || !att || !h) return TTK_ERR_BAD_ARG;
  const int M = d->M, w = d->width, g = d->gqa, inner = d->inner, L = d->n_layers;
  const int64_t ldq = 2 * (int64_t)w + 2 * g;
  const bool fused = (y == nullptr);
  if (fused && w != 256) return TTK_ERR_BAD_SHAPE;
  for (int i = 0; i < L; ++i) {
    const int64_t* W = d->weights + (int64_t)i * W_COLS;
    const int mode = i == 0 ? 0 : 1;
    TTK_TRY(ttk_gemm_qkv_rope(xn, w, P(W[W_QKV]), w, M, w, w, g, d->rope, qkv, ldq, d->k_norm2, st));
    TTK_TRY(ttk_attn_varlen_fwd(qkv, ldq, M, w, g, d->attn_work, d->n_attn_work, d->softmax_scale, att, w, d->k_norm2, st));
    if (fused) {
      TTK_TRY(ttk_gemm_resid_norm256(att, w, P(W[W_OUT]), w, M, w, x, w, mode, d->alpha, PF(W[LN_ATTN_POST]), PF(W[LN_FFN]), x,
                                     xn, w, st));
    } else {
      TTK_TRY(ttk_gemm_bf16(att, w, P(W[W_OUT]), w, M, w, w, nullptr, y, w, nullptr, 0, st));
      TTK_TRY(ttk_resid_norm(x, y, x, xn, PF(W[LN_ATTN_POST]), PF(W[LN_FFN]), d->alpha, mode, M, w, w, st));
    }
    TTK_TRY(ttk_gemm_geglu(xn, w, P(W[W_12]), w, M, inner, w, h, inner, st));
    if (fused) {
      TTK_TRY(ttk_gemm_resid_norm256(h, inner, P(W[W_3]), inner, M, inner, x, w, mode, d->alpha, PF(W[LN_FFD_POST]),
                                     PF(W[LN_NEXT]), x, xn, w, st));
    } else {
      TTK_TRY(ttk_gemm_bf16(h, inner, P(W[W_3]), inner, M, w, inner, nullptr, y, w, nullptr, 0, st));
      TTK_TRY(ttk_resid_norm(x, y, x, xn, PF(W[LN_FFD_POST]), PF(W[LN_NEXT]), d->alpha, mode, M, w, w, st));
    }
  }
  return TTK_OK;
}

// Last encoder layer on the latent rows only (engine._layer_latent): see include/titok_b200.h.
int ttk_layer_fwd_latent(const ttk_layers_desc* d, int layer, const void* x, const void* xn, void* qkv, void* att,
                         const void* tail_work, int n_tail_work, const int32_t* latent_row, int T, void* xc, void* xnc,
                         void* attc, void* hc, void* yc, ttk_stream_t st) {
  if (!d || !d->weights || !x || !xn || !qkv || !att || !tail_work || !latent_row || !xc || !xnc || !attc || !hc)
    return TTK_ERR_BAD_ARG;
  const int M = d->M, w = d->width, g = d->gqa, inner = d->inner;
  if (layer < 0 || layer >= d->n_layers || T <= 0) return TTK_ERR_BAD_ARG;
  const int64_t ldq = 2 * (int64_t)w + 2 * g;
  const bool fused = (yc == nullptr);
  if (fused && w != 256) return TTK_ERR_BAD_SHAPE;
  const int64_t* W = d->weights + (int64_t)layer * W_COLS;
  const int mode = layer == 0 ? 0 : 1;
  TTK_TRY(ttk_gemm_qkv_rope(xn, w, P(W[W_QKV]), w, M, w, w, g, d->rope, qkv, ldq, d->k_norm2, st));
  TTK_TRY(ttk_attn_varlen_fwd(qkv, ldq, M, w, g, tail_work, n_tail_work, d->softmax_scale, att, w, d->k_norm2, st));
  TTK_TRY(ttk_gather_rows(att, w, latent_row, attc, w, T, w, st));
  TTK_TRY(ttk_gather_rows(x, w, latent_row, xc, w, T, w, st));
  if (fused) {
    TTK_TRY(ttk_gemm_resid_norm256(attc, w, P(W[W_OUT]), w, T, w, xc, w, mode, d->alpha, PF(W[LN_ATTN_POST]), PF(W[LN_FFN]), xc,
                                   xnc, w, st));
  } else {
    TTK_TRY(ttk_gemm_bf16(attc, w, P(W[W_OUT]), w, T, w, w, nullptr, yc, w, nullptr, 0, st));
    TTK_TRY(ttk_resid_norm(xc, yc, xc, xnc, PF(W[LN_ATTN_POST]), PF(W[LN_FFN]), d->alpha, mode, T, w, w, st));
  }
  TTK_TRY(ttk_gemm_geglu(xnc, w, P(W[W_12]), w, T, inner, w, hc, inner, st));
  if (fused) {
    TTK_TRY(ttk_gemm_resid_norm256(hc, inner, P(W[W_3]), inner, T, inner, xc, w, mode, d->alpha, PF(W[LN_FFD_POST]),
                                   PF(W[LN_NEXT]), xc, xnc, w, st));
  } else {
    TTK_TRY(ttk_gemm_bf16(hc, inner, P(W[W_3]), inner, T, w, inner, nullptr, yc, w, nullptr, 0, st));
    TTK_TRY(ttk_resid_norm(xc, yc, xc, xnc, PF(W[LN_FFD_POST]), PF(W[LN_NEXT]), d->alpha, mode, T, w, w, st));
  }
  return TTK_OK;
}

// Training forward (backward._layers_train). slab: bf16 [n_layers][per_layer]; offs: HOST int64 [11] element offsets of
// {qkv, att, o, y_a, x_f, xn_f, h12, h, y_f, x_n, xn_n} inside a layer's block; lse: fp32 [n_layers][width/64][M].
// Layer i reads x_n / xn_n of layer i-1 (x0 / xn0 for layer 0).
int ttk_layers_fwd_train(const ttk_layers_desc* d, const void* x0, const void* xn0, void* slab, int64_t per_layer,
                         const int64_t* offs, float* lse, ttk_stream_t st) {
  if (!d || !d->weights || !x0 || !xn0 || !slab || !offs || !lse) return TTK_ERR_BAD_ARG;
  const int M = d->M, w = d->width, g = d->gqa, inner = d->inner, L = d->n_layers;
  const int64_t ldq = 2 * (int64_t)w + 2 * g;
  const int hq = w / 64;
  char* base = static_cast<char*>(slab);
  auto B = [&](int layer, int col) -> void* { return base + ((int64_t)layer * per_layer + offs[col]) * 2; };
  const void* x = x0;
  const void* xn = xn0;
  for (int i = 0; i < L; ++i) {
    const int64_t* W = d->weights + (int64_t)i * W_COLS;
    const int mode = i == 0 ? 0 : 1;
    float* lse_i = lse + (int64_t)i * hq * M;
    TTK_TRY(ttk_gemm_qkv_rope(xn, w, P(W[W_QKV]), w, M, w, w, g, d->rope, B(i, F_QKV), ldq, d->k_norm2, st));
    TTK_TRY(ttk_attn_varlen_fwd_train(B(i, F_QKV), ldq, M, w, g, d->attn_work, d->n_attn_work, d->softmax_scale, B(i, F_ATT), w,
                                      B(i, F_O), lse_i, d->k_norm2, st));
    TTK_TRY(ttk_gemm_bf16(B(i, F_ATT), w, P(W[W_OUT]), w, M, w, w, nullptr, B(i, F_YA), w, nullptr, 0, st));
    TTK_TRY(ttk_resid_norm(x, B(i, F_YA), B(i, F_XF), B(i, F_XNF), PF(W[LN_ATTN_POST]), PF(W[LN_FFN]), d->alpha, mode, M, w, w, st));
    TTK_TRY(ttk_gemm_bf16(B(i, F_XNF), w, P(W[W_12]), w, M, 2 * inner, w, nullptr, B(i, F_H12), 2 * (int64_t)inner, nullptr, 0, st));
    TTK_TRY(ttk_geglu_fwd(B(i, F_H12), 2 * (int64_t)inner, inner, B(i, F_H), inner, M, st));
    TTK_TRY(ttk_gemm_bf16(B(i, F_H), inner, P(W[W_3]), inner, M, w, inner, nullptr, B(i, F_YF), w, nullptr, 0, st));
    TTK_TRY(ttk_resid_norm(B(i, F_XF), B(i, F_YF), B(i, F_XN), B(i, F_XNN), PF(W[LN_FFD_POST]), PF(W[LN_NEXT]), d->alpha, mode, M,
                           w, w, st));
    x = B(i, F_XN);
    xn = B(i, F_XNN);
  }
  return TTK_OK;
}

// Backward of the layers (backward._layers_backward). g_in: dL/dx after the last layer. work: bf16 scratch; woffs: HOST
// int64 [11] element offsets of {du_f, dh, dh12, dxn, g_f, du_a, d_att, dqkv, dO, g_a, g_b} (temporaries are reused by
// every layer, the layer outputs alternate between g_a and g_b); delta: fp32 [width/64][M] scratch; grads: HOST int64
// [n_layers][8] device pointers of the fp32 gradients {ffd_post_ln, w3, w12, ffn_norm, attn_post_ln, out_proj, to_qkv,
// pre_ln} (0 where the layer has no such parameter). *g_out receives the pointer (g_a or g_b) holding dL/dx0.
int ttk_layers_bwd(const ttk_layers_desc* d, const void* x0, const void* xn0, const void* slab, int64_t per_layer,
                   const int64_t* offs, const float* lse, const void* g_in, void* work, const int64_t* woffs, float* delta,
                   const int64_t* grads, void** g_out, ttk_stream_t st) {
  if (!d || !d->weights || !x0 || !xn0 || !slab || !offs || !lse || !g_in || !work || !woffs || !delta || !grads || !g_out)
    return TTK_ERR_BAD_ARG;
  const int M = d->M, w = d->width, gq = d->gqa, inner = d->inner, L = d->n_layers;
  const int64_t ldq = 2 * (int64_t)w + 2 * gq;
  const int hq = w / 64;
  const char* base = static_cast<const char*>(slab);
  char* wb = static_cast<char*>(work);
  auto T = [&](int layer, int col) -> const void* { return base + ((int64_t)layer * per_layer + offs[col]) * 2; };
  auto Wk = [&](int col) -> void* { return wb + woffs[col] * 2; };
  const void* g = g_in;
  for (int i = L - 1; i >= 0; --i) {
    const int64_t* W = d->weights + (int64_t)i * W_COLS;
    const int64_t* G = grads + (int64_t)i * G_COLS;
    const int mode = i == 0 ? 0 : 1;
    const float c = mode == 1 ? d->alpha : 1.0f;
    const void* x_a = i == 0 ? x0 : T(i - 1, F_XN);
    const void* xn_a = i == 0 ? xn0 : T(i - 1, F_XNN);
    const float* lse_i = lse + (int64_t)i * hq * M;
    // ---- GEGLU block: x_out = x_f + y_f | RMSNorm(alpha x_f + y_f)
    const void* du = g;
    if (mode == 1) {
      TTK_TRY(ttk_rmsnorm_bwd(T(i, F_XF), T(i, F_YF), d->alpha, PF(W[LN_FFD_POST]), nullptr, nullptr, g, nullptr, 0.f, Wk(B_DUF),
                              PFm(G[G_FFD_POST]), nullptr, M, w, w, st));
      du = Wk(B_DUF);
    }
    TTK_TRY(ttk_gemm_bf16(du, w, P(W[W_3]), inner, M, inner, w, nullptr, Wk(B_DH), inner, nullptr, 1, st));
    TTK_TRY(ttk_gemm_wgrad(du, w, T(i, F_H), inner, M, w, inner, PFm(G[G_W3]), inner, st));
    TTK_TRY(ttk_geglu_bwd(T(i, F_H12), 2 * (int64_t)inner, inner, Wk(B_DH), inner, Wk(B_DH12), 2 * (int64_t)inner, M, st));
    TTK_TRY(ttk_gemm_bf16(Wk(B_DH12), 2 * (int64_t)inner, P(W[W_12]), w, M, w, 2 * inner, nullptr, Wk(B_DXN), w, nullptr, 1, st));
    TTK_TRY(ttk_gemm_wgrad(Wk(B_DH12), 2 * (int64_t)inner, T(i, F_XNF), w, M, 2 * inner, w, PFm(G[G_W12]), w, st));
    TTK_TRY(ttk_rmsnorm_bwd(T(i, F_XF), nullptr, 1.f, PF(W[LN_FFN]), nullptr, nullptr, Wk(B_DXN), du, c, Wk(B_GF), PFm(G[G_FFN]),
                            nullptr, M, w, w, st));
    // ---- attention block: x_f = x_a + y_a | RMSNorm(alpha x_a + y_a)
    du = Wk(B_GF);
    if (mode == 1) {
      TTK_TRY(ttk_rmsnorm_bwd(x_a, T(i, F_YA), d->alpha, PF(W[LN_ATTN_POST]), nullptr, nullptr, Wk(B_GF), nullptr, 0.f, Wk(B_DUA),
                              PFm(G[G_ATTN_POST]), nullptr, M, w, w, st));
      du = Wk(B_DUA);
    }
    TTK_TRY(ttk_gemm_bf16(du, w, P(W[W_OUT]), w, M, w, w, nullptr, Wk(B_DATT), w, nullptr, 1, st));
    TTK_TRY(ttk_gemm_wgrad(du, w, T(i, F_ATT), w, M, w, w, PFm(G[G_WOUT]), w, st));
    TTK_TRY(ttk_attn_bwd_prep(Wk(B_DATT), w, T(i, F_O), w, T(i, F_QKV), ldq, M, w, Wk(B_DO), w, Wk(B_DQKV), ldq, delta, st));
    TTK_TRY(ttk_attn_bwd_dkv(T(i, F_QKV), ldq, Wk(B_DO), w, M, w, gq, d->dkv_work, d->n_dkv_work, lse_i, delta, d->rope,
                             d->softmax_scale, Wk(B_DQKV), ldq, st));
    TTK_TRY(ttk_attn_bwd_dq(T(i, F_QKV), ldq, Wk(B_DO), w, M, w, gq, d->dq_work, d->n_dq_work, lse_i, delta, d->rope,
                            d->softmax_scale, Wk(B_DQKV), ldq, st));
    TTK_TRY(ttk_gemm_bf16(Wk(B_DQKV), ldq, P(W[W_QKV]), w, M, w, (int)ldq, nullptr, Wk(B_DXN), w, nullptr, 1, st));
    TTK_TRY(ttk_gemm_wgrad(Wk(B_DQKV), ldq, xn_a, w, M, (int)ldq, w, PFm(G[G_WQKV]), w, st));
    void* g_next = Wk(((L - 1 - i) & 1) ? B_G1 : B_G0);
    TTK_TRY(ttk_rmsnorm_bwd(x_a, nullptr, 1.f, PF(W[LN_PRE]), nullptr, nullptr, Wk(B_DXN), du, c, g_next, PFm(G[G_PRE]), nullptr, M,
                            w, w, st));
    g = g_next;
  }
  *g_out = const_cast<void*>(g);
  return TTK_OK;
}

}  // extern "C"
