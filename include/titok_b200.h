/*
 * titok_b200.h -- C ABI of libtitok_b200.so: the sm_100a kernels behind the TiTok-Video tokenizer
 * hot path (encode -> quantize -> decode).
 *
 * The reference (NilanEkanayake/TiTok-Video) is pure Python/PyTorch and has NO native interface;
 * each entry point below replaces the chain of library kernels that one stretch of the reference's
 * Python launches. The "replaces" notes cite reference files (relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked "host".
 *   - bf16 tensors are row-major with an explicit leading dimension in ELEMENTS (ld*), 16-byte aligned.
 *   - the caller owns every buffer; nothing is allocated or freed; all work is enqueued on `stream`
 *     (no internal synchronisation) so calls can be captured into a CUDA graph.
 *   - return value: 0 on success, negative ttk_status otherwise; never throws. There is no CPU or
 *     non-sm_100 fallback: a device that is not compute capability 10.x returns TTK_ERR_ARCH.
 */
#ifndef TITOK_B200_H_
#define TITOK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* ttk_stream_t; /* == cudaStream_t */

enum ttk_status {
  TTK_OK = 0,
  TTK_ERR_BAD_ARG = -1,
  TTK_ERR_BAD_SHAPE = -2,
  TTK_ERR_ALIGNMENT = -3,
  TTK_ERR_ARCH = -4,
  TTK_ERR_CUDA = -5,
  TTK_ERR_DRIVER = -6,
  TTK_ERR_WORKSPACE = -7
};

const char* ttk_strerror(int status);
int ttk_version(void);

/* ---------------------------------------------------------------------------------------------
 * Quantizer: model/quantizer/fsq.py
 * FSQ constants are HOST arrays of D (<= 8) entries, computed by the caller exactly as
 * FSQ.bound / FSQ.quantize compute them (fsq.py:78-90): half_l, offset, shift, half_width; basis and
 * levels are FSQ._basis / FSQ._levels (fsq.py:63-67).
 * ------------------------------------------------------------------------------------------- */

/* FSQ.forward (fsq.py:123-135). dtype 0 = bf16, 1 = fp32. codes has z's dtype, indices are int32. */
int ttk_fsq_fwd(const void* z, void* codes, int32_t* indices, int64_t n, int dtype, int D, const float* half_l,
                const float* offset, const float* shift, const float* half_width, const int32_t* basis,
                const int32_t* levels, ttk_stream_t stream);

/* Backward of FSQ.forward through round_ste (fsq.py:48-51): dz = dcodes*half_l*(1-tanh^2(z+shift))/half_width */
int ttk_fsq_bwd(const void* z, const void* dcodes, void* dz, int64_t n, int dtype, int D, const float* half_l,
                const float* offset, const float* shift, const float* half_width, const int32_t* basis,
                const int32_t* levels, ttk_stream_t stream);

/* FSQ.indices_to_codes (fsq.py:96-103,111-121). idx_dtype 0 = int32, 1 = int64; out_dtype 0 = bf16, 1 = fp32. */
int ttk_fsq_indices_to_codes(const void* idx, int idx_dtype, void* codes, int out_dtype, int64_t n, int D,
                             const float* half_width, const int32_t* basis, const int32_t* levels,
                             ttk_stream_t stream);

/* torch.bincount of CodebookLogger.get_scores (train_utils/codebook_logging.py:20-24), on device:
 * counts[K] (uint32) += histogram(idx[n]). */
int ttk_hist_u32(const int32_t* idx, int64_t n, int K, uint32_t* counts, ttk_stream_t stream);

/* usage / entropy of CodebookLogger.get_scores (codebook_logging.py:26-29), on device.
 * out[3] doubles: #non-zero bins, entropy in nats, total count. */
int ttk_codebook_stats(const uint32_t* counts, int K, double* out, ttk_stream_t stream);

/* Generic VQ (BASELINE.json north_star): idx[n] = argmin_k ||z_n - c_k||^2 over a [K,D] bf16 codebook, the
 * oracle being torch.cdist(z, C).argmin(-1) (for FSQ: C = FSQ.implicit_codebook, fsq.py:75-76).
 * The codebook is augmented once: cb_aug [ttk_vq_aug_rows(K,D), ttk_vq_aug_dim(D)]: K rows [-2c | 3-term bf16 split
 * of |c|^2 | 0] followed by the fp32 squared norms of ceil256(K) codes (+inf past K; used when D % 64 == 0, where the
 * kernel adds |c|^2 in its epilogue instead of spending a k block on the norm columns).
 * z is [N,D] bf16 with row pitch ldz (multiple of 8 elements). best (optional) = |c|^2 - 2 z.c of the winner. */
int ttk_vq_aug_dim(int D);
int ttk_vq_aug_rows(int K, int D);
int ttk_vq_prepare_codebook(const void* codebook, int64_t ldc, int K, int D, void* cb_aug, int64_t lda,
                            ttk_stream_t stream);
int ttk_vq_argmin(const void* z, int64_t ldz, const void* cb_aug, int64_t lda, int64_t N, int K, int D, int32_t* idx,
                  float* best, ttk_stream_t stream);
/* zq[n] = C[idx[n]] (16-byte gather, D % 8 == 0) and, if loss_sum != NULL, loss_sum[0] += sum ||zq - z||^2
 * (numerator of the commitment / codebook loss of a learned-codebook VQ; the reference's FSQ has no such loss). */
int ttk_vq_gather_loss(const void* z, int64_t ldz, const void* codebook, int64_t ldc, const int32_t* idx, int64_t N,
                       int D, void* zq, int64_t ldq, float* loss_sum, ttk_stream_t stream);
/* Backward of the learned-codebook quantizer z_q = z + sg(C[idx] - z) with commitment = mean((z - sg(c))^2) and
 * codebook = mean((sg(z) - c)^2):  dz = dzq + commit_scale * (z - c),  dC[idx] += codebook_scale * (c - z)
 * (commit_scale = dL/dcommitment * 2/(N D), codebook_scale = dL/dcodebook * 2/(N D)). bf16 activations, dC fp32
 * [K, lddc], ACCUMULATED (zero it first). dzq / dz / dC may be NULL. dev_scales (device float[2], may be NULL)
 * multiplies the two scales on the device, so the loss gradients never visit the host. (The reference's FSQ has only the
 * straight-through part, fsq.py:48-51; this is north_star's generic VQ.) */
int ttk_vq_bwd(const void* dzq, int64_t lddq, const void* z, int64_t ldz, const void* codebook, int64_t ldc,
               const int32_t* idx, int64_t N, int D, float commit_scale, float codebook_scale, const float* dev_scales,
               void* dz, int64_t lddz, float* dC, int64_t lddc, ttk_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Linear layers (tcgen05 GEMMs).  C[M,N] = A[M,K] @ W[N,K]^T, bf16 in, fp32 accumulate, bf16 out.
 * ------------------------------------------------------------------------------------------- */

/* nn.Linear (blocks.py:49,67,125,143; transformer.py:45,83). bias bf16 [N] or NULL. out_row_map int32 [M]
 * or NULL (scatter rows; negative = skip). w_is_kn != 0: W given as [K,N] row-major. */
int ttk_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, const void* bias,
                  void* out, int64_t ldo, const int32_t* out_row_map, int w_is_kn, ttk_stream_t stream);

/* Development aid (not used by the product path): buf = device int64 [148][64], zeroed by the caller, or NULL to
 * switch off. GEMM launches after this call record clock64 stamps of their producer / MMA / epilogue pipelines. */
int ttk_debug_set_trace(void* buf);

/* Attn.to_qkv + split + apply_rotary_emb(q), (k) (transformer.py:85-98, rope.py:19-27).
 * rope: fp32 [M,60] (cos,sin) of the 30 rotated complex lanes per token (RoPE.forward, rope.py:57-71).
 * out [M, 2*width+2*gqa] = [rope(q) | gate | rope(k) | v]. k_norm2 (optional, fp32 [gqa/64][M]): |k|^2 of every row and
 * kv head, a by-product of the epilogue that ttk_attn_varlen_fwd turns into its score bound. */
int ttk_gemm_qkv_rope(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, int width, int gqa,
                      const float* rope, void* out, int64_t ldo, float* k_norm2, ttk_stream_t stream);

/* GEGLU.w12 + chunk + gelu(gate)*x (transformer.py:47-52). W12 [2*inner, K]; out [M, inner]. */
int ttk_gemm_geglu(const void* A, int64_t lda, const void* W12, int64_t ldw, int M, int inner, int K, void* out,
                   int64_t ldo, ttk_stream_t stream);

/* out_proj / w3 (N = 256) fused with the residual or KEEL update and the next pre-norm
 * (transformer.py:128-130,141-145). mode 0: x' = x + y; mode 1: x' = RMSNorm(alpha*x + y)*w_post.
 * x_out = x'; xn_out (optional) = RMSNorm(x')*w_next. Norm weights fp32 [256]. */
int ttk_gemm_resid_norm256(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, const void* x,
                           int64_t ldx, int mode, float alpha, const float* w_post, const float* w_next, void* x_out,
                           void* xn_out, int64_t ldo, ttk_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Attention: flash_attn_varlen_func(...) * sigmoid(gate) (transformer.py:100-103).
 * qkv is the output of ttk_gemm_qkv_rope. work: device array of n_work 48-byte records
 * {q_row0[2], q_valid[2], q_head[2], kv_head, kv_row0, kv_len, kmax2, leader, kmax2b} (int32) built by the host planner
 * from cu_seqlens (blocks.py:81-83); leader = index of the first record of the same (clip, kv head); kmax2 / kmax2b are SCRATCH
 * of the library (the call writes max_j |k_j|^2 of that clip / kv head into the leader records: the work list must be
 * writable device memory, and concurrent calls on different streams need their own copy). out [M, width].
 * k_norm2 (optional): the by-product of ttk_gemm_qkv_rope; with it the call enqueues ONE kernel and does not touch the
 * work list. Without it, it enqueues two kernels: the key-norm bound (reads K again), then the attention kernel proper.
 * out must be 32-byte aligned with ldo % 16 == 0 (256-bit row stores).
 * ------------------------------------------------------------------------------------------- */
int ttk_attn_varlen_fwd(const void* qkv, int64_t ld, int M, int width, int gqa, const void* work, int n_work,
                        float softmax_scale, void* out, int64_t ldo, const float* k_norm2, ttk_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Row kernels (width in {256,512,768,1024}; norm weights fp32 [width]; mask_token fp32 [1]).
 * ------------------------------------------------------------------------------------------- */

/* flash_attn.ops.triton.layer_norm.RMSNorm forward (eps 1e-5). */
int ttk_rmsnorm_fwd(const void* x, int64_t ldx, const float* w, void* y, int64_t ldy, int M, int width,
                    ttk_stream_t stream);

/* Residual / KEEL update + next pre-norm, unfused form of ttk_gemm_resid_norm256 (any width). */
int ttk_resid_norm(const void* x, const void* y, void* x_out, void* xn_out, const float* w_post, const float* w_next,
                   float alpha, int mode, int M, int width, int64_t ld, ttk_stream_t stream);

/* Packing metadata of a batch (TiTokEncoder.forward, blocks.py:72-89; RoPE ids, rope.py:57-71; patch geometry,
 * utils.py:26-51) expanded on the device from per-clip descriptors desc int64 [n_clips,12] =
 * {row_start, tok_start, pat_start, token_count, n_patches, g1, g2, clip_offset, W, H*W, T*H*W, 0}.
 * Outputs: enc_src_row / dec_src_row int32 [M], latent_row int32 [T], patch_row int32 [G], geom int64 [G,4],
 * rope_pos int32 [M,3]. */
int ttk_build_plan(const int64_t* desc, int n_clips, int64_t M, int P0, int P1, int P2, int32_t* enc_src_row,
                   int32_t* dec_src_row, int32_t* latent_row, int32_t* patch_row, int64_t* geom, int32_t* rope_pos,
                   ttk_stream_t stream);

/* ttk_build_plan for a BUCKET of batch compositions: the launch extents (M_max, T_max, G_max) are fixed -- they are baked
 * into one captured CUDA graph -- and the real sizes of the step live in device memory: hdr int64[8] = {n_clips, M, T, G,
 * element offset of a 768-element scratch patch behind the clips, 0, 0, 0}. Padded rows / tokens / patches get harmless
 * defaults (mask-token rows, token row 0, the scratch patch), see csrc/rowops.cu. Replaces, like ttk_build_plan, the
 * metadata code of blocks.py:72-89 / rope.py:57-71 for the reference's ragged batches (video_dataset.py:130-172). */
int ttk_build_plan_bucket(const int64_t* desc, const int64_t* hdr, int64_t M_max, int64_t T_max, int64_t G_max, int P0,
                          int P1, int P2, int32_t* enc_src_row, int32_t* dec_src_row, int32_t* latent_row,
                          int32_t* patch_row, int64_t* geom, int32_t* rope_pos, ttk_stream_t stream);

/* RoPE table of a packed batch (RoPE.forward + _get_freqs_cis, rope.py:48-71): rope [M,60] fp32 (cos, sin) pairs,
 * complex lane = freq*3 + axis, gathered from cs_table [n_ids,10,2] fp32 = (cos, sin)(inv_freq[f] * id), evaluated once
 * in float64 by the host planner, by the integer position ids pos [M,3] int32. */
int ttk_rope_table_gather(const int32_t* pos, const float* cs_table, int n_ids, float* rope, int64_t M,
                          ttk_stream_t stream);

/* Per-clip reconstruction error of two flat clip buffers (same layout): out[2i] += sum|a-b| (the L1 reconstruction
 * loss numerator, loss_module.py:118), out[2i+1] += sum (a-b)^2 (PSNR, eval_metrics.py). out: fp64 [2*n_clips], zeroed
 * by the caller; clip_offset / clip_numel: device int64 [n_clips] in elements (multiples of 8). */
int ttk_clip_error(const void* a, const void* b, const int64_t* clip_offset, const int64_t* clip_numel, int n_clips,
                   int64_t max_clip_numel, double* out, ttk_stream_t stream);

/* The step before the path: decoded uint8 frames -> clips in [-1, 1] (dataset/video_dataset.py:118-119:
 * `chunk.to(dtype) / 255; chunk * 2 - 1` on bf16 tensors), bit-identical to that torch expression, so that a job can
 * ship uint8 frames over PCIe (half the bytes of bf16 clips). src uint8 [n], dst bf16 [n], both 16-byte aligned. */
int ttk_normalize_u8(const void* src, void* dst, int64_t n, ttk_stream_t stream);

/* TiTokEncoder embed (blocks.py:95-97). src_row int32 [M]: >= 0 row of proj (patch), < 0 latent row. */
int ttk_enc_embed(const void* proj, int64_t ldp, const int32_t* src_row, const float* mask_token, const float* w_t,
                  const float* w_p, const float* w_next, void* x_out, void* xn_out, int M, int width, int64_t ld,
                  ttk_stream_t stream);

/* TiTokDecoder embed (blocks.py:164-167). src_row int32 [M]: >= 0 latent token index into codes, < 0 patch row.
 * w_in bf16 [width, token_size], b_in bf16 [width]. */
int ttk_dec_embed(const void* codes, int token_size, const int32_t* src_row, const void* w_in, const void* b_in,
                  const float* mask_token, const float* w_t, const float* w_p, const float* w_next, void* x_out,
                  void* xn_out, int M, int width, int64_t ld, ttk_stream_t stream);

/* Encoder head + quantizer (blocks.py:101-103 -> titok.py:49 -> fsq.py:123-135).
 * w_out bf16 [token_size, width], b_out bf16 [token_size]. Outputs z/codes bf16 [T, token_size], idx int32 [T].
 * latent_row int32 [T]: packed row of every latent token; NULL: x holds the latent rows only, in token order. */
int ttk_enc_head_fsq(const void* x, int64_t ld, const int32_t* latent_row, const float* w_post, int pre_normed,
                     const void* w_out, const void* b_out, int token_size, void* z_out, void* codes_out,
                     int32_t* idx_out, int T, int width, const float* half_l, const float* offset, const float* shift,
                     const float* half_width, const int32_t* basis, const int32_t* levels, ttk_stream_t stream);

/* patch_rearrange / unpatch_rearrange (model/base/utils.py:26-51) with feature order (c p0 p1 p2).
 * geom int64 [G,4] = {element offset of the patch origin in the flat clip buffer, W, H*W, T*H*W}. P2 must be 8. */
int ttk_patchify(const void* clips, const int64_t* geom, int C, int P0, int P1, int P2, void* patches, int64_t ldp,
                 int64_t G, ttk_stream_t stream);
/* ttk_patchify on decoded uint8 frames: patch gather + the dataset's normalisation `x / 255 * 2 - 1`
 * (dataset/video_dataset.py:118-119) in one pass; bit-identical to ttk_normalize_u8 followed by ttk_patchify. */
int ttk_patchify_u8(const void* clips, const int64_t* geom, int C, int P0, int P1, int P2, void* patches, int64_t ldp,
                    int64_t G, ttk_stream_t stream);
int ttk_unpatchify(const void* proj, int64_t ldp, const int32_t* patch_row, const int64_t* geom, int C, int P0,
                   int P1, int P2, void* clips, int64_t G, ttk_stream_t stream);

/* =============================================================================================
 * Training path: backward kernels. The reference has no hand-written backward; each entry point replaces what
 * torch.autograd runs for the cited forward lines under bf16 autocast (train.py:68-80: manual_backward of the
 * reconstruction loss through TiTokDecoder -> FSQ (straight-through) -> TiTokEncoder). Activation gradients are
 * bf16, parameter gradients fp32 and ACCUMULATED (+=) into caller-zeroed buffers.
 * =========================================================================================== */

/* ttk_attn_varlen_fwd that also saves what the backward needs: o_save [M, ldo] = attention output before the gate,
 * lse fp32 [width/64][M] = log2-domain log-sum-exp of the scaled scores. */
int ttk_attn_varlen_fwd_train(const void* qkv, int64_t ld, int M, int width, int gqa, const void* work, int n_work,
                              float softmax_scale, void* out, int64_t ldo, void* o_save, float* lse, const float* k_norm2,
                              ttk_stream_t stream);

/* Backward of `attn * sigmoid(gate)` (transformer.py:101-103) + the row sums flash-attn's backward needs:
 * dO = d_out * sigmoid(gate) [M,width]; dqkv[:, width:2*width] = d gate; delta fp32 [width/64][M] = rowsum(dO o O). */
int ttk_attn_bwd_prep(const void* d_out, int64_t ldd, const void* o, int64_t ldo, const void* qkv, int64_t ld, int M,
                      int width, void* dO, int64_t lddo, void* dqkv, int64_t ldq, float* delta, ttk_stream_t stream);

/* Backward of flash_attn_varlen_func (transformer.py:100) and of apply_rotary_emb on q, k (rope.py:19-27).
 * work: int32 [n_work, 8] records {st_row0, st_valid, st_head, o_head0, n_heads, clip_row0, clip_len, 0} (plan.py:
 * attn_bwd_work_lists). dkv writes dqkv[:, 2w:2w+g] (dk, rotated back) and dqkv[:, 2w+g:] (dv); dq writes dqkv[:, :w]. */
int ttk_attn_bwd_dkv(const void* qkv, int64_t ld, const void* dO, int64_t lddo, int M, int width, int gqa,
                     const void* work, int n_work, const float* lse, const float* delta, const float* rope,
                     float softmax_scale, void* dqkv, int64_t ldq, ttk_stream_t stream);
int ttk_attn_bwd_dq(const void* qkv, int64_t ld, const void* dO, int64_t lddo, int M, int width, int gqa,
                    const void* work, int n_work, const float* lse, const float* delta, const float* rope,
                    float softmax_scale, void* dqkv, int64_t ldq, ttk_stream_t stream);

/* nn.Linear weight gradient: dW[n_out, k_in] (fp32, row pitch ldw) += dY[M, n_out]^T @ X[M, k_in] (split-K tcgen05
 * GEMM over the token dimension). The input gradient dX = dY @ W is ttk_gemm_bf16(dY, W, w_is_kn = 1). */
int ttk_gemm_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, int M, int n_out, int k_in, float* dw,
                   int64_t ldw, ttk_stream_t stream);

/* RMSNorm backward (fa:ops/triton/layer_norm.py:1093-1126), optionally through the KEEL pre-sum (transformer.py:141-145):
 *   u = y ? bf16(bf16(alpha*x) + y) : x;  dx = d(RMSNorm(u)*w)/du . dy + add_scale * add;  dw += sum_rows dy * u * rstd.
 * sel (optional int32 [M]): rows with sel < 0 use w2 / dw2 (the two pre-norms of the embed, blocks.py:95-97,164-167). */
int ttk_rmsnorm_bwd(const void* x, const void* y, float alpha, const float* w, const float* w2, const int32_t* sel,
                    const void* dy, const void* add, float add_scale, void* dx, float* dw, float* dw2, int M, int width,
                    int64_t ld, ttk_stream_t stream);

/* GEGLU on a stored w12 output h12 [M, 2*inner] = [value | gate] (transformer.py:50-52) and its backward. */
int ttk_geglu_fwd(const void* h12, int64_t ld12, int inner, void* h, int64_t ldh, int64_t M, ttk_stream_t stream);
int ttk_geglu_bwd(const void* h12, int64_t ld12, int inner, const void* dh, int64_t ldh, void* dh12, int64_t ldd,
                  int64_t M, ttk_stream_t stream);

/* dst[i] = src[idx[i]] / dst[idx[i]] = src[i] for n rows of `width` bf16 (latent / patch row maps, blocks.py:85-86). */
int ttk_gather_rows(const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd, int64_t n, int width,
                    ttk_stream_t stream);
int ttk_scatter_rows(const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd, int64_t n, int width,
                     ttk_stream_t stream);

/* Bias / mask_token gradients: out[c] (optional) += sum_r x[r,c]; total[0] (optional) += sum of all of x.
 * x bf16 [M, N], N and ld multiples of 8, 16-byte aligned. */
int ttk_colsum(const void* x, int64_t ld, int64_t M, int N, float* out, float* total, ttk_stream_t stream);

/* Weight refresh after an optimizer step: the master parameters of a stack -> their kernel-layout copies in one launch.
 * table: DEVICE int64 [n][4] = {src pointer, dst pointer, numel, kind}; kind = 2 * (src is bf16, else fp32) + (dst is fp32,
 * else bf16). Replaces the `.to(bfloat16)` casts autocast performs on every nn.Linear weight in every forward. */
int ttk_multi_cast(const int64_t* table, int n, int64_t max_numel, ttk_stream_t stream);

/* Backward of the encoder head Linear(width -> token_size) on the latent rows (blocks.py:101-103). */
int ttk_head_bwd(const void* dz, int token_size, const void* xn, int64_t ld, const int32_t* latent_row, const void* w_out,
                 void* dxn, float* dw, float* db, int T, int width, ttk_stream_t stream);

/* Backward of the decoder's Linear(token_size -> width) on the latent rows (blocks.py:164-165). dcodes fp32 [T, TS]. */
int ttk_dec_in_bwd(const void* de, int64_t ld, const int32_t* latent_row, const void* codes, int token_size,
                   const void* w_in, float* dcodes, float* dw, float* db, int T, int width, ttk_stream_t stream);

/* ttk_enc_embed / ttk_dec_embed that also store the rows before ln_pre_t / ln_pre_p (e0_out [M, ld]). */
int ttk_enc_embed_train(const void* proj, int64_t ldp, const int32_t* src_row, const float* mask_token, const float* w_t,
                        const float* w_p, const float* w_next, void* x_out, void* xn_out, void* e0_out, int M, int width,
                        int64_t ld, ttk_stream_t stream);
int ttk_dec_embed_train(const void* codes, int token_size, const int32_t* src_row, const void* w_in, const void* b_in,
                        const float* mask_token, const float* w_t, const float* w_p, const float* w_next, void* x_out,
                        void* xn_out, void* e0_out, int M, int width, int64_t ld, ttk_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Native launch sequencers: one call enqueues every kernel of the transformer layers of a stack
 * (ResidualAttentionBlock.forward, transformer.py:126-146, and its backward) -- the same kernels in the same order as
 * the per-kernel entry points above; they exist because at the reference's batch size the step is bound by the host's
 * launch rate. weights: HOST int64 [n_layers][9] of device pointers {to_qkv, out_proj, w12, w3 (bf16), pre_ln,
 * attn_post_ln, ffn_norm, ffd_post_ln, next_ln (fp32; 0 = the layer has no such norm)}; next_ln = pre_ln of layer i+1,
 * ln_post after the last layer.
 * ------------------------------------------------------------------------------------------- */
typedef struct ttk_layers_desc {
  int32_t M, width, gqa, inner, n_layers, n_attn_work, n_dkv_work, n_dq_work;
  float alpha, softmax_scale;
  const float* rope;      /* [M,60] */
  const void* attn_work;  /* ttk_attn_varlen_fwd work list */
  const void* dkv_work;   /* ttk_attn_bwd_dkv work list (backward only) */
  const void* dq_work;    /* ttk_attn_bwd_dq work list (backward only) */
  const int64_t* weights; /* host */
  float* k_norm2;         /* scratch fp32 [gqa/64][M]: key norms from the qkv GEMM to the attention kernel (may be null) */
} ttk_layers_desc;

/* Inference: x, xn [M,width] updated in place; qkv [M,2w+2g], att [M,w], h [M,inner] scratch; y [M,w] scratch selects the
 * unfused residual path (required unless width == 256). */
int ttk_layers_fwd(const ttk_layers_desc* d, void* x, void* xn, void* qkv, void* att, void* h, void* y, ttk_stream_t stream);

/* The encoder's LAST layer, restricted to what the head reads (blocks.py:101 `x[latent_mask]`): every row-wise step of
 * ResidualAttentionBlock.forward (transformer.py:126-146) acts on a packed row alone and attention output row i depends on
 * query row i only, so after the last qkv projection (keys / values of ALL rows are still needed) only the latent rows have
 * to be carried on. Runs layer `layer` of the stack described by d: qkv GEMM + RoPE on all M rows, attention for the query
 * tiles of `tail_work` (planner: the tiles that hold latent rows), gathers the T latent rows (latent_row int32 [T]) of att
 * and x into the compact xc / attc [T,width], then out_proj, GEGLU and w3 (+ residual / KEEL / norms) on T rows. On return
 * xc = x[latent_row] after the layer and xnc = RMSNorm(xc) * next norm weight, bit-identical to those rows of
 * ttk_layers_fwd. hc [T,inner] scratch; yc [T,width] scratch selects the unfused residual path (required unless width == 256). */
int ttk_layer_fwd_latent(const ttk_layers_desc* d, int layer, const void* x, const void* xn, void* qkv, void* att,
                         const void* tail_work, int n_tail_work, const int32_t* latent_row, int T, void* xc, void* xnc,
                         void* attc, void* hc, void* yc, ttk_stream_t stream);

/* Training forward: slab bf16 [n_layers][per_layer]; offs HOST int64 [11] = element offsets of {qkv, att, o, y_a, x_f, xn_f,
 * h12, h, y_f, x_n, xn_n} in a layer's block; lse fp32 [n_layers][width/64][M]. */
int ttk_layers_fwd_train(const ttk_layers_desc* d, const void* x0, const void* xn0, void* slab, int64_t per_layer,
                         const int64_t* offs, float* lse, ttk_stream_t stream);

/* Backward: g_in = dL/dx after the last layer; work bf16 scratch with HOST int64 [11] offsets woffs = {du_f, dh, dh12, dxn,
 * g_f, du_a, d_att, dqkv, dO, g_a, g_b}; delta fp32 [width/64][M] scratch; grads HOST int64 [n_layers][8] device pointers of
 * the fp32 gradients {ffd_post_ln, w3, w12, ffn_norm, attn_post_ln, out_proj, to_qkv, pre_ln}. *g_out = dL/dx0 (g_a or g_b). */
int ttk_layers_bwd(const ttk_layers_desc* d, const void* x0, const void* xn0, const void* slab, int64_t per_layer,
                   const int64_t* offs, const float* lse, const void* g_in, void* work, const int64_t* woffs, float* delta,
                   const int64_t* grads, void** g_out, ttk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TITOK_B200_H_ */
