/* mvlt_b200.h — C ABI of libmvlt_b200.so: hand-written sm_100a kernels for the MVLT multimodal forward hot path.
 *
 * The reference (Control-xl/Medical-Vision-Langauge-Transformer) has no FFI of its own: its boundary is the Python
 * nn.Module surface.  Each entry point below replaces the ATen call sites named beside it; the
 * Python host (medical_vision_langauge_transformer_b200/) keeps the reference's class names, forward signatures and
 * state_dict keys and calls these through ctypes with raw device pointers.
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer is a DEVICE pointer unless noted;
 * functions are stream-ordered on `stream`, allocate nothing, keep no global mutable state beyond one-time
 * attribute setup, never throw or exit.  Return 0 on success, a negative MVLT_ERR_* for rejected arguments, or a
 * positive cudaError_t.  dtype codes: 0 = fp32, 1 = bf16.  Row strides (ld*) are in ELEMENTS.
 */
#ifndef MVLT_B200_H_
#define MVLT_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mvlt_stream_t; /* == cudaStream_t */

#define MVLT_OK 0
#define MVLT_ERR_INVALID (-1)
#define MVLT_ERR_UNSUPPORTED (-2)
#define MVLT_ERR_DRIVER (-3)
#define MVLT_F32 0
#define MVLT_BF16 1
#define MVLT_ACT_NONE 0
#define MVLT_ACT_GELU 1 /* exact erf GELU (nn.GELU default), vfe.py:126, HF modeling_bert.py:336 */
#define MVLT_ACT_TANH 2 /* BertPooler, HF modeling_bert.py:466 */
#define MVLT_ACT_RELU 3 /* applied AFTER the residual add: torchvision resnet.py Bottleneck.forward (out += identity; relu) */
#define MVLT_ACT_RELU_GELU 4 /* last bottleneck of the trunk: ReLU, then the nn.GELU of model.py:232-235 */

/* One-time setup (driver entry points, opt-in shared memory sizes).  Call once per process after the CUDA
 * context exists and before any stream capture.  Also reports the library ABI version. */
int mvlt_init(void);
int mvlt_abi_version(void);

/* C[M,N] = act(A[M,K] . W[N,K]^T + bias) + residual — tcgen05/TMEM/TMA, bf16 operands, fp32 accumulate.
 * Replaces nn.Linear at vfe.py:231 (qkv), :252 (proj), :136/:139 (fc1/fc2), :443 (reduction, bias=NULL) and
 * HF modeling_bert.py:179-181 (Q|K|V packed as one [2304,768] weight), :295, :338, :352, :463, :476.
 * A,W bf16; C fp32|bf16 (out_dtype); bias fp32 or NULL; residual (or NULL) of C's dtype: fp32 C takes the fp32 residual
 * stream (residual == C accumulates in place at the L2), bf16 C a bf16 residual (act NONE / RELU / RELU_GELU).
 * MVLT_ACT_RELU / RELU_GELU need a bf16 C.  K % 16 == 0, lda/ldw % 8 == 0.  block_n = 0 picks the tile width. */
int mvlt_gemm_bf16_tc(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc,
                      const float* bias, const void* residual, long long ldres, int res_dtype, int M, int N, int K,
                      int act, int out_dtype, int block_n, mvlt_stream_t stream);

/* Conv2d over an NHWC bf16 activation as an implicit GEMM on the same tcgen05 kernel: the A operand is fetched by
 * im2col-mode TMA (64 channels x 128 output pixels per filter tap and load; padding arrives as zeros), nothing is
 * materialised.  out[B*Ho*Wo, N] (bf16, row stride ldc) = act(patches(x) . w^T + bias (+ residual)); w bf16 [N, R*S*C]
 * packed TAP-MAJOR (k = (ky*S + kx)*C + c) with the eval-mode BatchNorm scale folded in, bias fp32 = the folded shift,
 * residual bf16 [B*Ho*Wo, N] or NULL.  C % 64 == 0.  Replaces the nn.Conv2d + BatchNorm2d (+ ReLU, + identity) call
 * sites of torchvision resnet.py Bottleneck.forward as used by vfe.py:7-24 / :27-44: conv2 (3x3, stride 1|2, pad 1) and
 * the stride-2 1x1 downsample; stride-1 1x1 convolutions are plain mvlt_gemm_bf16_tc calls on the NHWC matrix. */
int mvlt_conv2d_nhwc_bf16_tc(const void* x, int B, int H, int W, int C, const void* w, long long ldw, void* out,
                             long long ldc, const float* bias, const void* residual, long long ldres, int N, int R, int S,
                             int stride, int pad, int act, int block_n, mvlt_stream_t stream);

/* Explicit tap-major patch matrix out[B*Ho*Wo, R*S*C] (row stride ld_out) of an NHWC activation (dtype fp32|bf16): the A
 * operand of the fp32 parity-mode convolutions (mvlt_gemm_f32_simt); the bf16 path uses mvlt_conv2d_nhwc_bf16_tc. */
int mvlt_im2col_nhwc(const void* x, int dtype, void* out, long long ld_out, int B, int H, int W, int C, int R, int S,
                     int stride, int pad, mvlt_stream_t stream);

/* fp32 parity mode only: patch matrix of the ResNet stem (conv1 7x7/2 pad 3 on the 3-channel NCHW image, vfe.py:15):
 * out[B*Ho*Wo, kpad] (out_dtype), k = (c*R + ky)*S + kx = conv1.weight.view(64, -1) order, columns >= Cin*R*S zero. */
int mvlt_stem_im2col_nchw(const float* img, void* out, int out_dtype, long long ld_out, int B, int Cin, int H, int W, int R,
                          int S, int stride, int pad, int kpad, mvlt_stream_t stream);

/* The whole ResNet stem in one kernel (bf16 mode): conv1 7x7/2 pad 3 (3 -> 64 channels) + folded BatchNorm + ReLU +
 * MaxPool2d(3, 2, 1); NCHW fp32 image [B,3,H,W] -> NHWC bf16 out [B, Hp, Wp, 64] (H = W = 224 -> 56 x 56).  w bf16 [64, 160]
 * = conv1.weight.view(64, 147) * BatchNorm scale, zero padded; bias fp32 [64] = the folded shift.  mma.sync tiles over a
 * shared-memory input patch; neither the patch matrix nor the 112 x 112 convolution output touches HBM.  vfe.py:15-18. */
int mvlt_resnet_stem_tc(const float* img, const void* w, const float* bias, void* out, int B, int H, int W,
                        mvlt_stream_t stream);

/* nn.MaxPool2d(k, stride, pad) on NHWC (vfe.py:18: MaxPool2d(3, 2, 1) after the stem); dtype fp32|bf16. */
int mvlt_maxpool_nhwc(const void* x, void* out, int dtype, int B, int H, int W, int C, int k, int stride, int pad,
                      mvlt_stream_t stream);

/* Fused MLP half of a Swin block, in place on the fp32 residual stream:  x += fc2(GELU(fc1(LayerNorm(x)))).
 * One tcgen05 kernel per call: LayerNorm in the prologue (fp32 statistics), the [128, 4C] hidden tile stays in shared
 * memory / TMEM, fp32 accumulate, fp32 residual.  Replaces vfe.py:385 (x + drop_path(mlp(norm2(x))), eval) with
 * vfe.py:136-139 inside.  x fp32 [M, C] (row stride ldx); gamma/beta/b1/b2 fp32; w1 bf16 [4C, C]; w2 bf16 [C, 4C].
 * C in {96, 192, 384} (Swin-S stages 0-2; stage 3 uses the unfused kernels), hidden == 4C. */
int mvlt_swin_mlp_fused(float* x, long long ldx, const float* gamma, const float* beta, float eps, const void* w1,
                        const float* b1, const void* w2, const float* b2, long long M, int C, int hidden,
                        mvlt_stream_t stream);

/* LayerNorm + Linear in one tcgen05 kernel (A-stationary): out[M, N] (bf16, row stride ldc) = act(LayerNorm(x; gamma, beta,
 * eps) . w^T + bias).  x fp32 [M, C] (row stride ldx), w bf16 [N, C], bias fp32 [N] or NULL, act NONE | GELU.  The 128 normalised
 * rows of a CTA stay in shared memory as the MMA A operand while the N columns stream by.  Replaces vfe.py:356 + :231
 * (norm1 -> qkv) and vfe.py:385 + :136 (norm2 -> fc1 -> GELU).  C in {192, 384}, N % 32 == 0. */
int mvlt_ln_linear_bf16(const float* x, long long ldx, const float* gamma, const float* beta, float eps, const void* w,
                        const float* bias, void* out, long long ldc, long long M, int C, int N, int act, mvlt_stream_t stream);

/* Same contract as mvlt_gemm_bf16_tc in fp32 on the CUDA cores (parity mode, 1e-4 vs the reference); all five
 * activation codes, fp32 residual. */
int mvlt_gemm_f32_simt(const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc,
                       const float* bias, const float* residual, long long ldres, int M, int N, int K, int act,
                       mvlt_stream_t stream);

/* out[r,:] = LayerNorm(in[r,:]) * gamma + beta, optional erf-GELU after it.  out_bf16_copy (or NULL) receives the
 * same rows rounded to bf16: `out` fp32 stays the residual, the copy is the next GEMM's A operand.
 * vfe.py:356,:385,:685 (+ model.py:232-235 GELU), HF modeling_bert.py:298,:356,:483. */
int mvlt_layernorm_rows(const void* in, int in_dtype, long long ld_in, void* out, int out_dtype, long long ld_out,
                        const float* gamma, const float* beta, long long rows, int C, float eps, int gelu,
                        void* out_bf16_copy, long long ld_copy, mvlt_stream_t stream);

/* PatchEmbed: Conv2d(3,96,k=4,s=4) + LayerNorm(96); img fp32 NCHW [B,3,224,224] -> out fp32 [B,3136,96].
 * vfe.py:557-565. */
int mvlt_patch_embed_ln(const float* img, const float* weight, const float* bias, const float* gamma,
                        const float* beta, float* out, int B, int img_size, int patch, int embed_dim, float eps,
                        mvlt_stream_t stream);

/* Same contract on the tensor cores (mma.sync, one warp per 16 patches, A fragments read straight from the NCHW image);
 * both operands are split into bf16 hi + lo halves and all four cross products accumulate in fp32, so the result is
 * fp32-accurate (~1e-6 relative).  The bf16-mode stem: 5x faster than the FMA-bound kernel above. */
int mvlt_patch_embed_ln_tc(const float* img, const float* weight, const float* bias, const float* gamma,
                           const float* beta, float* out, int B, int img_size, int patch, int embed_dim, float eps,
                           const float* gamma2, const float* beta2, float eps2, void* out2_bf16, int out2_window,
                           mvlt_stream_t stream);
/* out2_bf16 (or NULL): additionally LayerNorm(out; gamma2, beta2, eps2) rounded to bf16 [B*3136, 96] — norm1 of the first
 * Swin block (vfe.py:356) computed on the rows while they are still in registers.  out2_window > 0: its rows are written
 * WINDOW-MAJOR (window_partition of vfe.py:144-156 with that window size, no shift) for mvlt_window_attention_tc. */

/* LayerNorm(x) -> bf16 [B*H*W, C] with the OUTPUT ROWS in window-major order of the image rolled by -shift: row
 * (b*nW + w)*window^2 + i = token i of window w — norm1 of a Swin block (vfe.py:356) fused with torch.roll (vfe.py:361) and
 * window_partition (vfe.py:144-156, :363-364).  The qkv GEMM keeps the row order, so the tcgen05 window-attention kernel
 * below fetches each window as one contiguous TMA box.  x fp32 [B*H*W, C] (row stride ld_in) in natural token order. */
int mvlt_layernorm_rows_winmajor(const float* in, long long ld_in, void* out_bf16, const float* gamma, const float* beta, int B,
                                 int H, int W, int C, int window, int shift, float eps, mvlt_stream_t stream);

/* PatchMerging gather + LayerNorm(4C): x fp32 [B,H,W,C] -> out [B*H/2*W/2, 4C] in quad order (0,0),(1,0),(0,1),(1,1).
 * vfe.py:433-442 (the Linear(4C,2C) that follows is mvlt_gemm_*). */
int mvlt_patch_merge_ln(const float* x, void* out, int out_dtype, const float* gamma, const float* beta, int B, int H,
                        int W, int C, float eps, mvlt_stream_t stream);

/* Shifted-window attention with tokens in natural order: qkv [B*H*W, 3C] -> out [B*H*W, C]; the roll / partition /
 * reverse of vfe.py:144-173,:361,:378 are folded into the row index map.  shift > 0 adds the -100 region mask of
 * vfe.py:318-344.  relbias, dtype fp32: [heads,64,64] = the gathered relative_position_bias (vfe.py:236-238), zero
 * padded.  relbias, dtype bf16: the same bias with the shift mask added and 1/scale folded in, repacked in
 * mma.m16n8 C-fragment order as fp32 [n_cls, heads, 4, 7, 32, 4] (n_cls = 4 window classes when shift > 0, else 1;
 * key columns 49..55 = -1e30), 16-byte aligned — built once per block by the host (ops.window_bias_fragments). */
int mvlt_window_attention(const void* qkv, void* out, int dtype, const float* relbias, int B, int H, int W, int C,
                          int heads, int window, int shift, float scale, mvlt_stream_t stream);

/* First half of a Swin block up to the attention input in ONE tcgen05 kernel on CTA pairs (cta_group::2):
 *   out (bf16 [B*H*W, N], rows WINDOW-MAJOR for the image rolled by -shift) = LayerNorm(x) . w^T + bias
 * — norm1 (vfe.py:356) + torch.roll (:361) + window_partition (:363-364) + the qkv Linear (:231).  The LayerNorm prologue gathers
 * the token rows of the natural-order fp32 residual stream x [B*H*W, C] (row stride ldx) that belong to 256 consecutive output
 * rows and keeps them in shared memory as the A operand while the N columns are walked in 256-wide chunks.  w bf16 [N, C],
 * bias fp32 [N] or NULL, N % 32 == 0, C in {96, 192, 384}. */
int mvlt_swin_ln_qkv(const float* x, long long ldx, const float* gamma, const float* beta, float eps, const void* w,
                     const float* bias, void* out, int B, int H, int W, int C, int N, int window, int shift, mvlt_stream_t stream);

/* Second half of a Swin block in ONE tcgen05 kernel on CTA pairs (cta_group::2), in place on the fp32 residual stream x [M, C]:
 *   x <- x + o . w_proj^T + b_proj (vfe.py:252, :384), then x <- x + fc2(GELU(fc1(LayerNorm(x)))) (vfe.py:385, :136-139).
 * o = window-attention output, bf16 [M, C] dense, or NULL (MLP half only; w_proj / b_proj ignored).  The new residual rows stay
 * in tensor memory between the two halves; x is read once and written once.  Weights bf16 in nn.Linear layout (w_proj [C, C],
 * w1 [4C, C], w2 [C, 4C]); biases and LayerNorm parameters fp32, 16-byte aligned.  C in {192, 384}, hidden == 4C. */
int mvlt_swin_block_tail(const void* o, float* x, long long ldx, const void* w_proj, const float* b_proj, const float* gamma,
                         const float* beta, float eps, const void* w1, const float* b1, const void* w2, const float* b2,
                         long long M, int C, int hidden, mvlt_stream_t stream);

/* nn.Linear + residual + LayerNorm of the BERT post-LN sites in ONE tcgen05 kernel (clusters of four CTAs, see csrc/gemm_ln.cu):
 *   y = LayerNorm(A . W^T + bias + residual) * gamma + beta
 * — BertSelfOutput (HF modeling_bert.py:295-297: dense, dropout (eval: identity), LayerNorm(hidden + input)) and BertOutput (:352-354).
 * A bf16 [M, K] (row stride lda), W bf16 [N, K] (nn.Linear layout), bias / gamma / beta fp32 [N], residual fp32 [M, N];
 * y is written as fp32 to out_f32 (may alias residual) and, when out_bf16 != NULL, as bf16 to out_bf16 (the A operand of the next
 * GEMM).  N must be 768 (MVLT_ERR_UNSUPPORTED otherwise: the caller keeps GEMM + mvlt_layernorm_rows), K % 16 == 0. */
int mvlt_linear_residual_layernorm(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                                   const float* residual, long long ldres, const float* gamma, const float* beta, float eps,
                                   float* out_f32, long long ld32, void* out_bf16, long long ld16, int M, int N, int K,
                                   mvlt_stream_t stream);

/* Number of 256-row tiles mvlt_linear_residual_layernorm processes concurrently on the current device (resident clusters of four
 * CTAs: 33 on a 148-SM B200), or a negative error code.  Host-side routing (ops.use_linear_ln) asks whether M fills a wave. */
int mvlt_linear_ln_resident_tiles(void);

/* The same attention on tcgen05 / TMEM / TMA (bf16): qkv [B*nW*49, 3C] with rows WINDOW-MAJOR for this block's shift (as
 * written by mvlt_layernorm_rows_winmajor + the qkv GEMM), out [B*H*W, C] in NATURAL token order (window_reverse + the
 * reverse roll of vfe.py:159-173, :373-381 are the output scatter).  Two windows per 128-lane accumulator tile, S = Q.K^T
 * and O = P.V as tcgen05.mma with S, P and O in tensor memory, one score row per thread.  bias_table: fp32
 * [n_cls][heads][49][52] = (relative-position bias + the -100 shift mask of vfe.py:318-344) * log2(e), n_cls = 4 window
 * classes when shift > 0 else 1 (ops.window_bias_table), 16-byte aligned.  window must be 7, C == heads * 32. */
int mvlt_window_attention_tc(const void* qkv, void* out, const float* bias_table, int B, int H, int W, int C, int heads,
                             int window, int shift, float scale, mvlt_stream_t stream);

/* Joint embedding assembly + additive key mask: model.py:110-160, :162-183.
 * feat [n_feat, n_obj, D]; img_index int32 [B] (row of feat per sample) or NULL for identity; ids int64 [B,L];
 * text_mask uint8 [B,L] or NULL (= ids > 0, model.py:337); image_mask uint8 [B,n_obj] or NULL (= all ones);
 * word_emb fp32 [vocab+1, D];
 * typepos fp32 [S, D] = token_type_emb[s <= n_obj+1] + position_emb[s]; out [B, S, D]; out_bf16_copy as in
 * mvlt_layernorm_rows (or NULL); kmask fp32 [B,S]. */
int mvlt_joint_embed(const void* feat, int feat_dtype, const int* img_index, const long long* ids,
                     const unsigned char* text_mask, const unsigned char* image_mask, const float* word_emb,
                     const float* typepos, void* out, int out_dtype, void* out_bf16_copy, float* kmask, int B, int n_obj,
                     int L, int D, int cls_id, int sep_id, int vocab_rows, mvlt_stream_t stream);

/* ViT token assembly (torchvision vision_transformer.py `_process_input` + class token + pos_embedding, vfe.py:94-107):
 * out fp32 [B, 1 + n_patch, D]; out[b,0] = cls + pos[0]; out[b,1+i] = patches[b*n_patch + i] + pos[1+i]. */
int mvlt_vit_embed(const float* patches, const float* cls, const float* pos, float* out, int B, int n_patch, int D,
                   mvlt_stream_t stream);

/* BERT self-attention over the joint sequence: qkv [B*S, 3*heads*64] -> out [B*S, heads*64].
 * HF modeling_bert.py:115-140; seq2seq != 0 applies the mask of model.py:118-123 instead of kmask.  S <= 288 in bf16 (131 for
 * Swin / ResNet features at L = 80, 180 for the two-view input, 197 inside the ViT trunk, 278 for ViT / linear-patch
 * features); the same kernel is the nn.MultiheadAttention of the ViT encoder blocks (kmask = 0). */
int mvlt_joint_attention(const void* qkv, void* out, int dtype, const float* kmask, int B, int S, int heads,
                         int head_dim, int seq2seq, int obj_end, float scale, mvlt_stream_t stream);

/* The same attention on tcgen05 / TMEM / TMA (bf16, head_dim 64): a tile is 128 consecutive rows of qkv for one head; the
 * samples it touches are multiplied under the MMA's disable-output-lane mask, so every accumulator lane is a real row.
 * S in [64, 96] or [128, 144] (the MVLT joint sequences 74 / 81 / 131); other lengths return MVLT_ERR_UNSUPPORTED and the
 * caller uses mvlt_joint_attention. */
int mvlt_joint_attention_tc(const void* qkv, void* out, const float* kmask, int B, int S, int heads, int head_dim, int seq2seq,
                            int obj_end, float scale, mvlt_stream_t stream);

/* Masked-LM loss WITHOUT the logits (model.py:396-410: BertOnlyMLMHead decoder, HF modeling_bert.py:502-512, + CrossEntropyLoss with
 * ignore_index): the vocabulary GEMM t[rows,K] . w[N,K]^T + bias runs on the tcgen05 kernel with an online-logsumexp epilogue that
 * writes one (max, sum exp) pair per (row, 256-column tile, epilogue part) instead of the [rows, N] fp32 logits (312 MB at batch 32);
 * a finishing pass combines them, recomputes logits[label] from the bf16 operands, and reduces in a fixed order:
 * loss_sum[0] = sum of per-row losses over rows with label != ignore_index, loss_sum[1] = their number.  A label outside [0, N)
 * makes the loss NaN.  t, w bf16 (row strides ldt, ldw), bias fp32 or NULL, labels int64 [rows]; workspace (caller-owned,
 * 16-byte aligned) of mvlt_mlm_ce_workspace_bytes(rows, N) bytes. */
long long mvlt_mlm_ce_workspace_bytes(long long rows, int N);
int mvlt_mlm_ce_fused(const void* t, long long ldt, const void* w, long long ldw, const float* bias, const long long* labels,
                      float* loss_sum, void* workspace, long long workspace_bytes, long long rows, int N, int K,
                      long long ignore_index, mvlt_stream_t stream);

/* out[r, n] = x[r,:] . w[n,:] + bias[n], N <= 16 (fp32 weights/outputs).  model.py:435, :363. */
int mvlt_linear_small(const void* x, int x_dtype, long long ldx, const float* w, const float* bias, float* out,
                      long long rows, int N, int K, mvlt_stream_t stream);

/* Row softmax, fp32.  model.py:348, :468. */
int mvlt_softmax_rows(const float* in, float* out, long long rows, int N, mvlt_stream_t stream);

/* loss_sum[0] = sum over rows with label != ignore_index of (logsumexp(row) - row[label]); loss_sum[1] = their count.
 * F.cross_entropy(..., ignore_index=-100) of model.py:410 and :418 (mean = [0]/[1]). */
int mvlt_masked_ce_rows(const float* logits, long long ld, const long long* labels, float* loss_sum, long long rows,
                        int N, long long ignore_index, mvlt_stream_t stream);

/* run_retrieval.py:220-249 `compute_ranks` on the device.  scores fp32 [R, C] (row stride lds) = the image x caption
 * matching probabilities, labels uint8 [R, C] (1 = matching pair).  row_ranks int32 [R] (or NULL): position of the first
 * matching caption of each image in descending-score order; col_ranks int32 [C] (or NULL): the same per caption over the
 * images.  Equal scores are ordered by decreasing index (np.argsort(sim)[::-1] with a stable sort); a line without a
 * positive ranks C (`num_captions_per_img`).  Two O(n) passes per line, no sort. */
int mvlt_rank_first_positive(const float* scores, long long lds, const unsigned char* labels, long long ldl,
                             int* row_ranks, int* col_ranks, int R, int C, mvlt_stream_t stream);

/* ---- first slice of the training step: the non-GEMM kernels of the backward of ONE BertLayer
 * (HF modeling_bert.py:359-421 under run_pretrain.py:177-184 `loss.backward()`).  dgrad / wgrad themselves are mvlt_gemm_bf16_tc
 * calls on transposed bf16 operands (host: medical_vision_langauge_transformer_b200/training.py). ---- */

/* out[c, r] = bf16(in[r, c]) for r < rows, 0 for rows <= r < ld_out: the M-contiguous operands of the wgrad GEMMs (ld_out a
 * multiple of 8 makes the rows TMA-aligned).  in: fp32 or bf16 [rows, cols], row stride ld_in. */
int mvlt_transpose_to_bf16(const void* in, int in_dtype, long long ld_in, void* out, long long ld_out, long long rows, int cols,
                           mvlt_stream_t stream);

/* torch.nn.functional.layer_norm backward over dense fp32 rows: x is the PRE-normalisation input, dx fp32 (+ optional bf16
 * copy), dgamma = sum_rows dy * xhat, dbeta = sum_rows dy, reduced in a fixed order.  C % 128 == 0, C <= 1024. */
long long mvlt_layernorm_bwd_workspace_bytes(long long rows, int C);
int mvlt_layernorm_bwd_rows(const float* dy, const float* x, const float* gamma, float eps, float* dx, void* dx_bf16,
                            float* dgamma, float* dbeta, void* workspace, long long rows, int C, mvlt_stream_t stream);

/* out[c] = sum_r x[r, c] (bias gradients); x fp32 or bf16, fixed summation order. */
long long mvlt_colsum_workspace_bytes(long long rows, int cols);
int mvlt_colsum(const void* x, int dtype, long long ld, float* out, void* workspace, long long rows, int cols, mvlt_stream_t stream);

/* du = df * d/du erf-GELU(u) (HF:330-342 BertIntermediate), bf16 in / out, n % 4 == 0. */
int mvlt_gelu_bwd(const void* u, const void* df, void* du, long long n, mvlt_stream_t stream);

/* dqkv (bf16 [B*S, 3C] = dq | dk | dv) of mvlt_joint_attention given dctx (bf16 [B*S, C]); probabilities are recomputed from qkv
 * and the masks (same mask arguments as the forward).  head_dim 64, S <= 160. */
int mvlt_joint_attention_bwd(const void* qkv, const float* kmask, const void* dctx, void* dqkv, int B, int S, int heads,
                             int head_dim, int seq2seq, int obj_end, float scale, mvlt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVLT_B200_H_ */
