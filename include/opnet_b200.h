/* libopnet_b200 -- C ABI of the B200-native OPNet temporal-reasoning hot path.
 *
 * The reference (ofrikleinfeld/ObjectPermanence) has no FFI layer: its hot path is the
 * forward()/backward of the nn.Modules in baselines/learned_models.py, whose arithmetic
 * lives in PyTorch library calls (nn.LSTM, nn.Linear, F.softmax, torch.einsum,
 * nn.TransformerEncoder).  Each entry point below names the reference call site(s) it
 * replaces.  The Python host side (objectpermanence_b200/) binds these with ctypes; the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous float32 unless stated otherwise
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it:
 *     no allocation, no host synchronisation (except opn_lstm_status, which says so)
 *   - return value: OPN_OK (0) or a negative OPN_ERR_*; opn_last_error() gives the text
 *   - batch-first layouts everywhere: a "row" r = b*T + t
 *   - sm_100a only; there is no CPU path
 */
#ifndef OPNET_B200_H
#define OPNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OPN_API __attribute__((visibility("default")))
#else
#define OPN_API
#endif

#define OPN_OK 0
#define OPN_ERR_BAD_ARG (-1)
#define OPN_ERR_UNSUPPORTED (-2)
#define OPN_ERR_CUDA (-3)
#define OPN_ERR_TIMEOUT (-4)

#define OPN_MAX_OBJECTS 15 /* baselines/learned_models.py:13 */
#define OPN_BOX_FEATURES 6 /* baselines/learned_models.py:21 (OPNet family) */

/* ---- library -------------------------------------------------------------------- */
OPN_API int opn_version(void);
OPN_API const char* opn_last_error(void);
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
OPN_API unsigned long long opn_launch_count(void);
/* SM count and compute capability of the current device */
OPN_API int opn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- dense fp32 contraction ----------------------------------------------------------
 * Two implementations behind one entry point: a CUDA-core FFMA kernel (any shape) and, for
 * large shapes when the caller provides scratch, a tcgen05 tensor-core kernel that splits
 * every fp32 operand into bf16 hi+lo and accumulates hi*hi + hi*lo + lo*hi in fp32 TMEM
 * (relative operand error 2^-17).  opn_sgemm_workspace_bytes() returns the scratch size the
 * tensor-core path wants for a shape, or 0 when the FFMA kernel will be used.
 * Short-K input projections (x * W_ih^T: trans_a == 0, trans_b != 0, 16 <= K <= 128, M >= 1024, alpha 1, beta 0, no
 * bias / ReLU) take a third kernel that splits the fp32 operands on the way into shared memory and needs no scratch
 * (opn_gemm_proj.cu; same bf16 hi+lo arithmetic).  Skinny shapes (one operand at most 16 wide) have their own kernels too.
 * C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C + bias[N], optional ReLU.
 *   trans_a == 0: A is [M,K] row-major (lda);  trans_a != 0: A is [K,M] row-major
 *   trans_b == 0: B is [K,N] row-major (ldb);  trans_b != 0: B is [N,K] row-major
 *   beta must be 0 or 1; bias may be NULL.  lda/ldb/ldc are arbitrary row strides (>= the row length),
 *   which is how strided row sets are contracted (e.g. slot 0 of every frame, or the last frame of
 *   every video).
 * Replaces: nn.Linear / nn.LSTM input projections and every autograd matmul behind them
 *   (baselines/learned_models.py:30,33,39,46,47,67,69,70,83,84,100,102,113,116,130,133,
 *    139,146,149,167,172,178,184,192,195).
 */
OPN_API int64_t opn_sgemm_workspace_bytes(int64_t M, int64_t N, int64_t K);
OPN_API int opn_sgemm(int trans_a, int trans_b, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
              const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int relu,
              void* workspace, int64_t workspace_bytes, void* stream);

/* ---- persistent LSTM recurrence ---------------------------------------------------
 * One bias-free, unidirectional LSTM layer with zero initial state, PyTorch gate order
 * i,f,g,o.  The input projection xproj = x * W_ih^T is computed beforehand (opn_sgemm).
 * H must be one of 32, 64, 128, 256, 512.
 * Replaces: the nn.LSTM calls at baselines/learned_models.py:39,46,76,113,146,192 (forward)
 *   and their autograd backward (baselines/training_main.py:216).
 *
 * workspace: opn_lstm_workspace_bytes(B,T,H) bytes, owned by the caller, must stay alive
 *   until the stream has drained; contents are scratch (status word, the two exchange rings,
 *   transposed W_hh).
 */
OPN_API int64_t opn_lstm_workspace_bytes(int64_t B, int64_t T, int64_t H);

/* forward: xproj [B,T,4H], w_hh [4H,H]  ->  hs [B,T,H]; if gates/cells are non-NULL the
 * post-activation gates [B,T,4H] (i,f,g,o blocks) and cell states [B,T,H] are stashed for
 * the backward pass (pass NULL for inference). */
OPN_API int opn_lstm_fwd(int64_t B, int64_t T, int64_t H, const float* xproj, const float* w_hh, float* hs, float* gates,
                 float* cells, void* workspace, int64_t workspace_bytes, void* stream);

/* backward recurrence: dh_out [B,T,H] is dLoss/dh from the consumers of hs.  Writes
 * dgates [B,T,4H] = dLoss/d(pre-activation gates).  Parameter and input gradients are
 * time-parallel contractions of dgates (opn_sgemm):
 *   dW_ih = dgates^T x ; dW_hh = sum_t dgates[:,t]^T hs[:,t-1] ; dx = dgates W_ih. */
OPN_API int opn_lstm_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates, const float* cells,
                 const float* dh_out, float* dgates, void* workspace, int64_t workspace_bytes, void* stream);

/* 1 when opn_lstm_fwd / opn_lstm_bwd run the batch-wide tcgen05 kernels for this (B, H) (groups of 128 videos, weights in
 * shared memory, accumulators in TMEM; opn_lstm_tc.cu), 0 for the 8-video mma.sync / FP32-FMA kernels.  The gates / cells
 * stash of the batch-wide forward has an internal layout: hand it to opn_lstm_bwd of the same (B, H) only. */
OPN_API int opn_lstm_batchwide(int64_t B, int64_t H);

/* SYNCHRONISES the device.  Reads back the status word the persistent kernels leave in
 * `workspace`: OPN_OK, or OPN_ERR_TIMEOUT if an inter-CTA wait expired (detail words in
 * info[0..2] if info != NULL).  For tests and debugging. */
OPN_API int opn_lstm_status(const void* workspace, uint32_t* info);

/* ---- arithmetic mode ---------------------------------------------------------------
 * The tensor-core kernels (recurrences, fused OPNet forward / backward, opn_sgemm's tcgen05 path) multiply 16-bit
 * operands with fp32 accumulation.  OPN_PRECISION_FP32 (default) carries every fp32 operand as a hi + lo pair and runs
 * three products per slice (hi.hi + hi.lo + lo.hi, 22 significand bits): predicted boxes within 1e-4 of the reference.
 * OPN_PRECISION_16BIT runs the hi.hi product alone (one pass, a third of the tensor work; cell state, accumulation and
 * the stash stay fp32): the north star's "1e-2 bf16" mode -- boxes within 1e-2.  The setting is per calling thread and
 * applies to the launches enqueued after it. */
#define OPN_PRECISION_FP32 0
#define OPN_PRECISION_16BIT 1
OPN_API int opn_set_precision(int mode);
OPN_API int opn_get_precision(void);

/* Sticky status page.  `device_page`: >= 4096 zeroed bytes of device memory owned by the caller, registered for the
 * CURRENT device (NULL unregisters).  While one is registered the persistent kernels (opn_lstm_*, opn_opnet_*) report
 * time-outs into it instead of into the head of their per-launch workspace: word 0 = code (0 = fine), words 1..3 =
 * step / CTA / thread.  The word is never cleared by the library, so the caller can read it wherever it synchronises
 * anyway (e.g. together with the loss of a training step, as objectpermanence_b200.training does) and must zero it after
 * reporting; while it is non-zero later launches give up at their first slow wait.  opn_lstm_status accepts the page. */
OPN_API int opn_set_status_page(void* device_page);

/* ---- OPNet "who to track" stage ---------------------------------------------------
 * forward (baselines/learned_models.py:40-43,50):
 *   logits = hs1 * w_pred^T            [B,T,15]
 *   probs  = softmax(logits, -1)
 *   frames_boxes[b,t,:] = sum_o probs[b,t,o] * boxes[b,t,o,:]        [B,T,6]
 *   logits_bpt = logits.permute(0,2,1).contiguous()                  [B,15,T]
 */
OPN_API int opn_wtt_fwd(int64_t B, int64_t T, int64_t H1, const float* boxes, const float* hs1, const float* w_pred,
                float* logits_bpt, float* probs, float* frames_boxes, void* stream);
/* backward: d_frames_boxes [B,T,6], optional d_logits_bpt [B,15,T] (NULL = no gradient
 * arrives on the logits output, as in baselines/training_main.py:186-216) ->
 *   d_logits [B,T,15] (row layout, for dW_pred = d_logits^T hs1 via opn_sgemm),
 *   d_hs1 [B,T,H1]. */
OPN_API int opn_wtt_bwd(int64_t B, int64_t T, int64_t H1, const float* boxes, const float* probs, const float* w_pred,
                const float* d_frames_boxes, const float* d_logits_bpt, float* d_logits, float* d_hs1,
                void* stream);

/* ---- OPNet fused forward -------------------------------------------------------------
 * LSTM1 (x-projection given), the who-to-track stage and LSTM2 (its K = 6 input projection included) as ONE
 * persistent kernel that advances all three frame by frame: the function of opn_lstm_fwd(H1) -> opn_wtt_fwd ->
 * opn_sgemm(W_ih2) -> opn_lstm_fwd(H2), with the same outputs (so the backward entry points above apply
 * unchanged), for the shipped OPNet config H1 = 256, H2 = 512 (OPN_ERR_UNSUPPORTED otherwise).
 * Replaces: OPNet.forward, baselines/learned_models.py:36-46 (configs/opnet_model_config.json).
 *   boxes [B,T,15,6], xproj1 [B,T,4*H1] = boxes W_ih1^T, w_hh1 [4*H1,H1], w_pred [15,H1], w_ih2 [4*H2,6],
 *   w_hh2 [4*H2,H2]  ->  hs1 [B,T,H1], logits_bpt [B,15,T], probs [B,T,15], frames_boxes [B,T,6], hs2 [B,T,H2];
 *   gates1/cells1/gates2/cells2: stash for the backward pass, all given or all NULL (inference).
 * workspace: opn_opnet_fwd_workspace_bytes(B,T) bytes of scratch (status word as for opn_lstm_status, two rings). */
OPN_API int64_t opn_opnet_fwd_workspace_bytes(int64_t B, int64_t T);
OPN_API int opn_opnet_fwd(int64_t B, int64_t T, int64_t H1, int64_t H2, const float* boxes, const float* xproj1,
                  const float* w_hh1, const float* w_pred, const float* w_ih2, const float* w_hh2, float* hs1,
                  float* gates1, float* cells1, float* logits_bpt, float* probs, float* frames_boxes, float* hs2,
                  float* gates2, float* cells2, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- OPNet fused backward ------------------------------------------------------------
 * The reverse recurrences of LSTM2 and LSTM1 and the who-to-track backward between them as ONE persistent kernel (the
 * mirror image of opn_opnet_fwd): the function of opn_lstm_bwd(H2) -> opn_sgemm(d frames_boxes) -> opn_wtt_bwd ->
 * opn_lstm_bwd(H1) when no gradient arrives on the who-to-track logits (as in baselines/training_main.py:186-216).
 * Shipped OPNet config only (H1 = 256, H2 = 512; OPN_ERR_UNSUPPORTED otherwise).
 *   stash of opn_opnet_fwd (probs, gates1, cells1, gates2, cells2), d_hs2 [B,T,H2] = dLoss / d h2  ->
 *   d_gates1 [B,T,4*H1], d_gates2 [B,T,4*H2], d_logits [B,T,15] (row layout): the inputs of the time-parallel
 *   weight-gradient contractions (opn_sgemm).
 * workspace: opn_opnet_bwd_workspace_bytes(B,T) bytes of scratch (status word as for opn_lstm_status, three rings). */
OPN_API int64_t opn_opnet_bwd_workspace_bytes(int64_t B, int64_t T);
OPN_API int opn_opnet_bwd(int64_t B, int64_t T, int64_t H1, int64_t H2, const float* boxes, const float* probs,
                  const float* w_hh1, const float* w_pred, const float* w_ih2, const float* w_hh2, const float* gates1,
                  const float* cells1, const float* gates2, const float* cells2, const float* d_hs2, float* d_gates1,
                  float* d_gates2, float* d_logits, void* workspace, int64_t workspace_bytes, void* stream);
/* The same in two calls.  In its default form on a B200 the backward runs as two concurrent kernels: the LSTM2 reverse loop on
 * `stream` and the head backward + LSTM1 reverse recurrence on a library-owned side stream, which finishes ~0.15 ms later.
 * opn_opnet_bwd_begin returns with d_gates2 ordered on `stream` but d_gates1 / d_logits still outstanding, so the caller can
 * enqueue work that needs only d_gates2 (the weight gradients of LSTM2: opn_wgrad) on the 128 SMs that are already free;
 * opn_opnet_bwd_join(stream) orders d_gates1 / d_logits on `stream` (a no-op when the single-kernel form ran). */
OPN_API int opn_opnet_bwd_begin(int64_t B, int64_t T, int64_t H1, int64_t H2, const float* boxes, const float* probs,
                  const float* w_hh1, const float* w_pred, const float* w_ih2, const float* w_hh2, const float* gates1,
                  const float* cells1, const float* gates2, const float* cells2, const float* d_hs2, float* d_gates1,
                  float* d_gates2, float* d_logits, void* workspace, int64_t workspace_bytes, void* stream);
OPN_API int opn_opnet_bwd_join(void* stream);

/* ---- LSTM weight gradients -------------------------------------------------------------
 * The time-parallel contractions autograd forms for the weights of nn.LSTM (baselines/learned_models.py:39,46,76,113,
 * 146,192), several per launch, straight from the fp32 tensors on tcgen05 (opn_wgrad_tc.cu): no operand pre-pass, no zero
 * fill, no atomics (the partial sums of the row ranges are added in a fixed order).  One job:
 *   out[M, N] (ldc) = sum over rows r of a[r, 0..M)^T b[r - shift, 0..N)
 *   a [rows, M] (lda) = d(gates), M = 4H a multiple of 128;  b [rows, N] (ldb) = the layer input (shift 0: dW_ih) or its
 *   hidden states (shift 1: dW_hh; rows with r % T == 0, the first frame of every video, take no part).
 * Arithmetic as opn_sgemm's tensor-core path: bf16 hi + lo operands, three products (one in OPN_PRECISION_16BIT).
 * workspace: opn_wgrad_workspace_bytes(n_jobs, jobs) bytes (status words as for opn_lstm_status, partial sums);
 * the job array is a HOST array read during the call. */
typedef struct {
    const float* a;
    const float* b;
    float* out;
    int64_t lda, ldb, ldc;
    int64_t rows, T, M, N;
    int32_t shift;
    int32_t trans_out; /* != 0: out is written transposed, [N, M] (ldc >= M) */
} opn_wgrad_job;
OPN_API int64_t opn_wgrad_workspace_bytes(int32_t n_jobs, const opn_wgrad_job* jobs);
OPN_API int opn_wgrad(int32_t n_jobs, const opn_wgrad_job* jobs, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- fused self-attention (transformer_lstm encoder) -----------------------------------
 * softmax(Q K^T / sqrt(d)) V per head over ONE sequence of S rows, as nn.MultiheadAttention computes it inside
 * nn.TransformerEncoderLayer (baselines/learned_models.py:166-168,184; the sequence axis of the reference is B*T).
 *   qkv [S, 3D] packed q|k|v (in_proj output), nhead heads of d = D / nhead = 128 (OPN_ERR_UNSUPPORTED otherwise)
 *   -> ctx [S, D].  No [S,S] tensor is formed: flash-style tcgen05 kernels (opn_attention_tc.cu); the backward pass
 *   recomputes the scores from the row statistics the forward call leaves in `workspace`, which must therefore be
 *   handed unchanged from opn_attention_fwd to opn_attention_bwd of the same call.
 *   p_drop > 0: attention-weight dropout with the mask of opn_dropout (element (q,k) of head h = element q*S + k of the
 *   stream (seed, offset + h * ceil(S*S/4))). */
OPN_API int64_t opn_attention_workspace_bytes(int64_t S, int64_t D, int64_t nhead);
OPN_API int opn_attention_fwd(int64_t S, int64_t D, int64_t nhead, const float* qkv, float* ctx, void* workspace,
                      int64_t workspace_bytes, float p_drop, uint64_t seed, uint64_t offset, void* stream);
OPN_API int opn_attention_bwd(int64_t S, int64_t D, int64_t nhead, const float* ctx, const float* d_ctx, float* d_qkv,
                      void* workspace, int64_t workspace_bytes, float p_drop, uint64_t seed, uint64_t offset,
                      void* stream);

/* ---- element-wise / row-wise helpers (transformer_lstm encoder, MLP variant) -------- */
/* dy[i] = (y[i] > 0) ? dy[i] : 0   (ReLU backward, in place on dy) */
OPN_API int opn_relu_bwd(int64_t n, const float* y, float* dy, void* stream);
/* out[c] (+)= sum_r x[r*ld + c]   (bias gradients) */
OPN_API int opn_colsum(int64_t rows, int64_t cols, const float* x, int64_t ld, float* out, int accumulate, void* stream);
/* row softmax in place over x [rows, cols] (ld); scale applied to the logits first */
OPN_API int opn_softmax_rows(int64_t rows, int64_t cols, float* x, int64_t ld, float scale, void* stream);
/* softmax backward in place: dp <- scale * p * (dp - sum_j p_j dp_j) */
OPN_API int opn_softmax_rows_bwd(int64_t rows, int64_t cols, const float* p, float* dp, int64_t ld, float scale, void* stream);
/* y = LayerNorm(x + res) * w + b over the last dim D (res may be NULL).  Stashes the
 * normalised activations xhat [rows,D] and rstd [rows] for the backward pass.
 * Replaces norm1/norm2 of nn.TransformerEncoderLayer (baselines/learned_models.py:166). */
OPN_API int opn_layernorm_fwd(int64_t rows, int64_t D, const float* x, const float* res, const float* w, const float* b,
                      float eps, float* y, float* xhat, float* rstd, void* stream);
/* backward: dx [rows,D] (gradient w.r.t. the sum x+res); dw, db [D] are accumulated (+=) */
OPN_API int opn_layernorm_bwd(int64_t rows, int64_t D, const float* xhat, const float* rstd, const float* w, const float* dy,
                      float* dx, float* dw, float* db, void* stream);
/* out[i] = keep(i) ? x[i] / (1 - p) : 0, in place allowed (out == x).  keep(i) is a pure function of (seed, offset, i):
 * word i%4 of Philox4x32-10 block offset + i/4 under key `seed` is >= p * 2^32.  Forward and backward of a dropout
 * site are the same call with the same (seed, offset); nothing is stashed.  One call consumes (n+3)/4 blocks of the
 * offset space.  Replaces the four nn.Dropout(p=0.1) sites of nn.TransformerEncoderLayer in train mode
 * (baselines/learned_models.py:166; statistically, not bit-wise: the mask generator is not PyTorch's). */
OPN_API int opn_dropout(int64_t n, const float* x, float* out, float p, uint64_t seed, uint64_t offset, void* stream);
/* out = a + b */
OPN_API int opn_add(int64_t n, const float* a, const float* b, float* out, void* stream);

/* ---- training loss (baselines/training_main.py:192-210) ---------------------------
 * loss = mean(|y - labels| [* mask]) [+ 0.5 * mean_t ||y[t+1]-y[t]||_2]  and its gradient
 * dy, in one launch.  mask (uint8 [B,T,4]) may be NULL; consistency != 0 adds the second
 * term (the *_no_labels models).  loss_out: 3 floats {total, prediction, consistency}. */
OPN_API int opn_loss_fwd_bwd(int64_t B, int64_t T, const float* y, const float* labels, const uint8_t* mask, int consistency,
                     float* loss_out, float* dy, void* stream);

/* ---- bbox head + training loss, forward and backward, in one pass --------------------
 * y [B,T,4] = h [B,T,H] w^T (the bias-free prediction layer, baselines/learned_models.py:33,47,117,150,196), the loss of
 * opn_loss_fwd_bwd on it, and the backward of both: d_h [B,T,H] = dLoss/dy w, d_w [4,H] = dLoss/dy^T h.  h is read once
 * and d_h written once (the separate head, loss, dy w and column-reduction launches read h twice at a fraction of the HBM
 * rate).  H a multiple of 128 up to 512 (OPN_ERR_UNSUPPORTED otherwise: use opn_sgemm + opn_loss_fwd_bwd).
 * workspace: opn_head_loss_workspace_bytes(B,T,H) bytes (reserved; d_w is accumulated with fp32 reductions). */
OPN_API int64_t opn_head_loss_workspace_bytes(int64_t B, int64_t T, int64_t H);
OPN_API int opn_head_loss(int64_t B, int64_t T, int64_t H, const float* h, const float* w, const float* labels, const uint8_t* mask,
                  int consistency, float* y, float* loss_out, float* d_h, float* d_w, void* workspace, int64_t workspace_bytes,
                  void* stream);

/* ---- optimiser step (baselines/training_main.py:150,217) ---------------------------
 * torch.optim.Adam (amsgrad off) over one flat fp32 buffer of n parameters, in place:
 *   g' = g + weight_decay * p;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
 *   p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)         step = 1, 2, ...
 * params / grads / exp_avg / exp_avg_sq: device pointers, 16-byte aligned, n floats each. */
OPN_API int opn_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream);

/* ---- IoU evaluation (baselines/training_main.py:97-112, baselines/tracking_utils.py:138-159) ----
 * y, labels [N,T,4] normalised xyxy; pixels = int32(trunc(double(x) * [320,240,320,240])); per-frame IoU with the
 * +1 pixel convention in double;  video_mean[n] = mean_t IoU  (ResultsAnalyzer "video_mean_iou");
 * with mask (uint8 [N,T,4], a frame counts when any of its 4 entries is set, training_main.py:88):
 * masked_mean[n] = mean over the masked frames (NaN when none; "containment_mean_iou"), masked_frames[n] their
 * number.  frame_iou [N,T] (optional) receives the per-frame values.  masked_* / frame_iou may be NULL. */
OPN_API int opn_iou_eval(int64_t N, int64_t T, const float* y, const float* labels, const uint8_t* mask, double* video_mean,
                 double* masked_mean, int32_t* masked_frames, double* frame_iou, void* stream);

/* ---- pixel boxes (baselines/inference_main.py:214-215, baselines/training_main.py:98-101) ----
 * pixels[r][c] = int32(trunc(double(boxes[r][c]) * {320,240,320,240}[c])) for rows normalised xyxy boxes: the
 * `(np.array(predictions) * frame_shapes).astype(np.int32)` of the reference, on the device. */
OPN_API int opn_to_pixels(int64_t rows, const float* boxes, int32_t* pixels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPNET_B200_H */
