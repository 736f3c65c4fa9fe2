/* libpcaa_sm100 -- C ABI of the B200-native PCAA train / open-set-inference hot path.
 *
 * The reference (rmazzier/OpenSetGaitRecognition_PCAA) is 100 % Python and has no FFI; the seam
 * this library plugs into is the Python module surface of its models.py / utils.py.
 * Every entry point below names the reference code (file:line) whose arithmetic it replaces.
 *
 * Conventions
 *  - all pointers are DEVICE pointers owned by the caller (PyTorch allocations); the library never
 *    allocates or frees device memory and keeps no pointers across calls;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*), no internal synchronisation;
 *  - return value: 0 = PCAA_OK, otherwise a pcaa_status; text via pcaa_last_error() (thread local);
 *  - "rows" matrices are channels-last: X[R, C] with C contiguous; point rows are ordered (b, t, n);
 *  - dtype arguments are pcaa_dtype values; statistics accumulators are double and are ADDED to
 *    (the caller zeroes them), everything else is overwritten unless stated.
 */
#ifndef PCAA_H_
#define PCAA_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PCAA_OK = 0,
    PCAA_ERR_SHAPE = 1,
    PCAA_ERR_ALIGN = 2,
    PCAA_ERR_UNSUPPORTED = 3,
    PCAA_ERR_LAUNCH = 4,
    PCAA_ERR_DRIVER = 5
} pcaa_status;

typedef enum { PCAA_F32 = 0, PCAA_BF16 = 1 } pcaa_dtype;
typedef enum { PCAA_ACT_NONE = 0, PCAA_ACT_ELU = 1 } pcaa_act;

typedef void* pcaa_stream; /* cudaStream_t */

const char* pcaa_version(void);
const char* pcaa_last_error(void);

/* ---- generic CUDA-core GEMM (small layers: TCN, heads, discriminator; bring-up comparator) -------------
 * C(m,n) = act( sum_k A(m,k) B(k,n) + bias[n] ) [+ C(m,n) if accumulate]; element (i,j) of X at X[i*s0 + j*s1].
 * Replaces torch.nn.Linear / Conv1d matmuls of models.py:59-67, 252-277, 344-371, 409-415. */
int pcaa_gemm_simt(const void* A, int a_dtype, int64_t sam, int64_t sak,
                   const void* B, int b_dtype, int64_t sbk, int64_t sbn,
                   void* C, int c_dtype, int64_t scm, int64_t scn,
                   int64_t M, int64_t N, int64_t K,
                   const float* bias, int act, int accumulate, pcaa_stream stream);

/* ---- tensor-core GEMMs (tcgen05 + TMA + TMEM), bf16 x bf16 -> fp32 ---------------------------------------
 * mode PCAA_TC_BIAS_STATS : Y = A W^T + bias (bf16 out) and per-column sum / sum-of-squares of the fp32 result
 *                           added to stats[2N] (BatchNorm batch statistics of models.py:29 fused in the epilogue)
 * mode PCAA_TC_BIAS_ELU   : Y = ELU(A W^T + bias)       (eval mode with BatchNorm folded into W, bias)
 * mode PCAA_TC_PLAIN      : Y = A W^T (+ bias when non-null)  (bf16 out)
 * mode PCAA_TC_DGRAD_ELUBN: dZ = (A W^T) * ELU'(scale*Yprev + shift) (bf16 out) and sum dZ, sum dZ*xhat added to
 *                           stats[2N]  (backward of models.py:33-34 fused in the data-gradient GEMM's epilogue)
 * A [M,K] (lda), W [N,K] (ldw), out [M,N] (ldo); all bf16, leading dims multiples of 8 elements. */
typedef enum { PCAA_TC_BIAS_STATS = 0, PCAA_TC_BIAS_ELU = 1, PCAA_TC_PLAIN = 2, PCAA_TC_DGRAD_ELUBN = 3,
               PCAA_TC_WGRAD_ACC = 4, PCAA_TC_DGRAD_ELUOUT = 5, PCAA_TC_WGRAD_STORE = 6,
               PCAA_TC_T_BIAS_STATS = 7, PCAA_TC_T_AFFINE_ELU = 8, PCAA_TC_T_DGRAD_ELUBN = 9,
               PCAA_TC_T_AFFINE_ELU_POOL = 10 } pcaa_tc_mode;
/* operand storage: PCAA_OP_K     A(m,k) at A[m*lda + k]  (B(n,k) at B[n*ldb + k]),  "K-major"
 *                  PCAA_OP_MN    A(m,k) at A[k*lda + m]  (B(n,k) at B[k*ldb + n]),  "MN-major"
 *                  PCAA_OP_T256_* the channel-major activation format of the PointNet path: element (channel c, point p)
 *                  at X[((p / 256) * C + c) * 256 + p % 256] -- 256-point tiles, each a contiguous [C, 256] block, the last
 *                  tile zero-padded.  _K: the points are the k index (weight gradient); _MN: the points are the n index. */
typedef enum { PCAA_OP_K = 0, PCAA_OP_MN = 1, PCAA_OP_T256_K = 2, PCAA_OP_T256_MN = 3 } pcaa_operand_layout;
/* General form.  out[m,n] = epilogue( sum_k A(m,k) B(n,k) ), bf16 operands, fp32 accumulation in TMEM.
 * a_mn / b_mn are pcaa_operand_layout values; leading dimensions in elements, multiples of 8 (ignored for T256).  out_dtype: PCAA_BF16 or PCAA_F32 (modes 1, 2).
 * Extra modes: PCAA_TC_WGRAD_ACC   out (fp32) += A B^T, split over k across the SMs (weight gradient, k = rows);
 *              PCAA_TC_WGRAD_STORE out (fp32)  = A B^T (no split, plain stores: decoder weight gradient, k = batch);
 *              PCAA_TC_DGRAD_ELUOUT out = (A B^T) * (a > 0 ? 1 : a + 1) with a = yprev [M, ldy] the saved bf16 OUTPUT
 *                                  of the previous layer's ELU (decoder data gradient, models.py:373-382).
 * Channel-major ("T") modes -- the PointNet path keeps activations channels-first in 256-point tiles (T256, above), so
 * the output ROW is the channel: per-channel parameters are per-row, BatchNorm statistics are per-row sums over the
 * points n < N, kept in registers across all tiles of a CTA and added to stats[2M] once.  B must be PCAA_OP_T256_MN,
 * out (and yprev) are T256 with C = M; points >= N of the last tile are written as zeros:
 *   PCAA_TC_T_BIAS_STATS  out[m,n] = acc + bias[m] (bf16); stats[m] += sum_n out, stats[M+m] += sum_n out^2
 *                         (forward of models.py:21-29: A = W [Cout,Cin], B = aT [Cin,P] read MN-major)
 *   PCAA_TC_T_AFFINE_ELU  out[m,n] = ELU(scale[m]*(acc + bias[m]) + shift[m])  (eval mode: BatchNorm with running
 *                         statistics applied in the epilogue -> the next layer's activation in one pass)
 *   PCAA_TC_T_AFFINE_ELU_POOL  the same activation, mean-pooled over groups of ldo consecutive points in the epilogue
 *                         (models.py:242-243, 282 fused into the last shared-MLP layer of the eval forward): out is
 *                         fp32 [N / ldo, M] (out_dtype PCAA_F32, ldo >= 32, ldo | N), zeroed by the call; the [M, N]
 *                         activation is never written
 *   PCAA_TC_T_DGRAD_ELUBN out[m,n] = acc * ELU'(scale[m]*yprev[m,n] + shift[m]); stats += [sum dz, sum dz*xhat]
 *                         (A = W read MN-major = W^T, B = dyT [Cout,P] MN-major; yprev = yT of the previous layer)
 * Instantiated layouts (a,b): (K,K) modes 0-4, 6; (K,MN) 2, 5; (MN,MN) 4, 6; (K,T256_MN) 7, 8, 10; (MN,T256_MN) 9;
 * (T256_K,T256_K) 4 (PointNet weight gradient, k = points). */
int pcaa_gemm_tc(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* out, int64_t ldo,
                 int out_dtype, int64_t M, int64_t N, int64_t K, int mode, const float* bias, double* stats,
                 const void* yprev, int64_t ldy, const float* scale, const float* shift, const float* mean,
                 const float* invstd, pcaa_stream stream);
int pcaa_gemm_tc_tn(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                    int64_t M, int64_t N, int64_t K, int mode,
                    const float* bias, double* stats,
                    const void* yprev, const float* scale, const float* shift, const float* mean, const float* invstd,
                    pcaa_stream stream);
/* weight gradient: dW[N1,N2] (fp32, ADDED to) = A^T B with A [K,N1] (lda), B [K,N2] (ldb) bf16, K = rows (points). */
int pcaa_gemm_tc_nt_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw,
                          int64_t N1, int64_t N2, int64_t K, pcaa_stream stream);
/* number of SMs the persistent kernels size their grids with (0 if no device) */
int pcaa_sm_count(void);

/* ---- PointNet layer 1 (Cin = 4: CUDA cores), models.py:86-88 -----------------------------------------------
 * x (B,4,T,N) fp32 NCHW -> y [B*TN, Cout] bf16 = W x + b, and column statistics of y added to stats[2*Cout]. */
int pcaa_pointnet_l1_fwd(const float* x, const float* w, const float* bias, void* y, double* stats,
                         int64_t B, int64_t TN, int Cout, pcaa_stream stream);
/* dW[Cout,4] = dy^T x (overwritten). */
int pcaa_pointnet_l1_wgrad(const float* x, const void* dy, float* dW, int64_t B, int64_t TN, int Cout,
                           pcaa_stream stream);

/* ---- channel-major PointNet kernels on the T256 activation format (see pcaa_operand_layout): bf16, element (c, p) at
 * X[((p / 256) * C + c) * 256 + p % 256], P = B*TN points ordered (b, t, n), pad points of the last tile stored as zeros.
 * layer 1, models.py:86-88: y(c,p) = sum_f w[c,f] x[b,f,tn] + bias[c]; stats (nullable) += [sum_p y, sum_p y^2];
 * with scale/shift (eval mode, BatchNorm folded) the stored value is ELU(scale[c]*y + shift[c]) instead. */
int pcaa_pointnet_l1_fwd_t(const float* x, const float* w, const float* bias, const float* scale, const float* shift,
                           void* yT, double* stats, int64_t B, int64_t TN, int Cout, pcaa_stream stream);
/* train mode with the BatchNorm coefficients known up front (pcaa_bn_from_input_moments): ONE pass writes both the
 * pre-BatchNorm y (yT, kept for the backward) and a = ELU(scale[c]*y + shift[c]) (aT) -- models.py:21-34 for the K = 4 layer. */
int pcaa_pointnet_l1_fwd_bn_t(const float* x, const float* w, const float* bias, const float* scale, const float* shift,
                              void* yT, void* aT, int64_t B, int64_t TN, int Cout, pcaa_stream stream);
/* mom[14] (double, overwritten) = { sum_p x_f (f = 0..3), sum_p x_f x_g (f <= g, row-major upper triangle) } over the B*TN
 * points of x (B,4,TN): all that BatchNorm 1's batch statistics depend on, because y1 = W1 x + b1 is linear in x. */
int pcaa_input_moments(const float* x, int64_t B, int64_t TN, double* mom, pcaa_stream stream);
/* BatchNorm2d batch statistics + running-statistics update (models.py:29; momentum, unbiased variance) of the layer
 * y = w x + bias, w [C,4], from the input moments of R points: mean_c = w_c.mu + bias_c, var_c = w_c^T Cov(x) w_c.
 * Outputs as pcaa_bn_finalize: scale = gamma*invstd, shift = beta - mean*scale, mean, invstd (each [C], mean/invstd nullable). */
int pcaa_bn_from_input_moments(const double* mom, int64_t R, int C, const float* w, const float* bias, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                               float* scale, float* shift, float* mean, float* invstd, pcaa_stream stream);
/* dW[Cout,4] (overwritten) = sum_p dy(c,p) x[f,p], dy = c1[c]*dzT + c2[c]*yT + c3[c] (BatchNorm backward fused; yT and
 * the coefficients null: dy = dzT) */
int pcaa_pointnet_l1_wgrad_t(const float* x, const void* dzT, const void* yT, const float* c1, const float* c2,
                             const float* c3, float* dW, int64_t B, int64_t TN, int Cout, pcaa_stream stream);
/* outT = ELU(scale[c]*yT + shift[c])  (models.py:29, 33-34) */
int pcaa_bn_elu_apply_t(const void* yT, const float* scale, const float* shift, void* outT, int64_t P, int C,
                        pcaa_stream stream);
/* dyT = c1[c]*dzT + c2[c]*yT + c3[c]  (BatchNorm backward; may run in place on dzT) */
int pcaa_bn_bwd_apply_t(const void* dzT, const void* yT, const float* c1, const float* c2, const float* c3, void* dyT,
                        int64_t P, int C, pcaa_stream stream);
/* pooled[g,c] = mean_{i<n} ELU(scale[c]*y(c, g*n+i) + shift[c])  (AvgPool2d((1,nmax)), models.py:242-243, 282; scale
 * null: plain mean of yT).  Training: e1[g,c] = sum_i ELU'(z), e2[g,c] = sum_i ELU'(z)*xhat (needs mean/invstd). */
int pcaa_bn_elu_meanpool_t(const void* yT, const float* scale, const float* shift, const float* mean,
                           const float* invstd, float* pooled, float* e1, float* e2, int64_t G, int n, int C,
                           pcaa_stream stream);
/* BatchNorm-backward statistics of the pooled layer from the group sums: stats2 += [sum_g dpool/n*e1, sum_g dpool/n*e2] */
int pcaa_pool_bwd_stats(const float* dpool, const float* e1, const float* e2, double* stats2, int64_t G, int n, int C,
                        pcaa_stream stream);
/* dy(c,p) = c1[c]*(dpool[g(p),c]/n)*ELU'(scale[c]*y+shift[c]) + c2[c]*y + c3[c]: mean-pool, ELU and BatchNorm
 * backward of the last PointNet layer in one pass (P = G*n) */
int pcaa_pool_bwd_apply_t(const float* dpool, const void* yT, const float* scale, const float* shift, const float* c1,
                          const float* c2, const float* c3, void* dyT, int64_t G, int n, int C, pcaa_stream stream);

/* ---- BatchNorm (train) + ELU on channels-last rows, models.py:29,33-34,72-78 ------------------------------ */
int pcaa_colstats(const void* y, int dtype, int64_t R, int C, double* stats, pcaa_stream stream);
/* scale = gamma*invstd, shift = beta - mean*scale; running stats updated in place when non-null
 * (momentum, unbiased variance) exactly as torch.nn.BatchNorm does in training mode. */
int pcaa_bn_finalize(const double* stats, int64_t R, int C, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps,
                     float* scale, float* shift, float* mean, float* invstd, pcaa_stream stream);
/* eval mode: scale/shift from running statistics */
int pcaa_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                        float eps, float* scale, float* shift, int C, pcaa_stream stream);
/* out = ELU(scale*y + shift) */
int pcaa_bn_elu_apply(const void* y, int y_dtype, const float* scale, const float* shift, void* out, int out_dtype,
                      int64_t R, int C, pcaa_stream stream);
/* pooled[g,c] = mean_{i<n} ELU(scale*y[g*n+i,c] + shift)   (AvgPool2d((1,nmax)), models.py:242-243,282) */
int pcaa_bn_elu_meanpool(const void* y, int y_dtype, const float* scale, const float* shift, float* pooled,
                         int64_t G, int n, int C, pcaa_stream stream);
/* dz = dout * ELU'(scale*y+shift); stats2 += [sum dz, sum dz*xhat].  pooled_n > 0: dout is fp32 [R/pooled_n, C]
 * (gradient of the mean pool, broadcast and divided by pooled_n), else dout is [R,C] of dout_dtype. */
int pcaa_elu_bwd_colstats(const void* dout, int dout_dtype, int pooled_n, const void* y, int y_dtype,
                          const float* scale, const float* shift, const float* mean, const float* invstd,
                          void* dz, int dz_dtype, double* stats2, int64_t R, int C, pcaa_stream stream);
/* dy = c1*dz + c2*y + c3 (BatchNorm backward); dgamma = sum dz*xhat, dbeta = sum dz */
int pcaa_bn_bwd_finalize(const double* stats2, int64_t R, int C, const float* scale, const float* mean,
                         const float* invstd, float* c1, float* c2, float* c3, float* dgamma, float* dbeta,
                         pcaa_stream stream);
int pcaa_bn_bwd_apply(const void* dz, int dz_dtype, const void* y, int y_dtype, const float* c1, const float* c2,
                      const float* c3, void* dy, int dy_dtype, int64_t R, int C, pcaa_stream stream);

/* ---- small helpers ----------------------------------------------------------------------------------------- */
/* dz = dout * (out > 0 ? 1 : out + 1): ELU backward from the saved OUTPUT (Linear+ELU heads, decoder) */
int pcaa_elu_bwd_from_out(const float* dout, const float* out, float* dz, int64_t n, pcaa_stream stream);
int pcaa_colsum(const float* x, int64_t R, int C, float* out, pcaa_stream stream);
/* same for a fp32 / bf16 matrix with leading dimension ld (bias gradients of the tensor-core decoder path) */
int pcaa_colsum_ld(const void* x, int dtype, int64_t R, int C, int64_t ld, float* out, pcaa_stream stream);
int pcaa_convert(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, pcaa_stream stream);
/* out[r, c] (ld_out, bf16) = in[r, c] (ld_in, fp32), or the transpose when transpose != 0; pad columns zeroed */
int pcaa_pack_bf16(const float* in, int64_t R, int64_t C, int64_t ld_in, void* out, int64_t ld_out, int transpose,
                   pcaa_stream stream);
/* element-wise family for the twice-differentiable critic path (create_graph=True, PCAA_ablation.py:955-962):
 * out = a*b | a+b | ELU(a) | ELU'(a) | ELU''(a) | a + b[col]  (b is a row vector of ncols for ADD_ROWVEC; a may be
 * null = zeros for ADD_ROWVEC, which then broadcasts b over the rows) */
typedef enum { PCAA_EW_MUL = 0, PCAA_EW_ADD = 1, PCAA_EW_ELU = 2, PCAA_EW_ELU_GRAD = 3, PCAA_EW_ELU_GRAD2 = 4,
               PCAA_EW_ADD_ROWVEC = 5 } pcaa_ew_op;
int pcaa_ew(int op, const float* a, const float* b, float* out, int64_t n, int ncols, pcaa_stream stream);
/* causal dilated Conv1d as GEMM: col[(b,t), ci*3+k] = x[b, t-(2-k)*dil, ci] (0 for negative time), models.py:59-76;
 * col is fp32 or bf16 (col_dtype: the tensor-core GEMM operand) */
int pcaa_tcn_im2col(const float* x, void* col, int col_dtype, int64_t B, int T, int Cin, int dil, pcaa_stream stream);
int pcaa_tcn_col2im(const float* dcol, float* dx, int64_t B, int T, int Cin, int dil, pcaa_stream stream);
/* ---- fused steps of one DilTempConv1d layer (models.py:37-79; the [B*T, C] tensors are tiny, the fusion removes launches)
 * forward, after the layer's GEMM y[B*T, C] (+ column statistics from its epilogue): BatchNorm1d + ELU (models.py:72-78).
 *   training: stats[2*C] = [sum_r y, sum_r y^2]; coef_out[4*C] receives scale | shift | mean | invstd, the running
 *   statistics are updated (momentum, unbiased variance);  eval: stats null, scale / shift given.
 *   Outputs (either may be null): col = bf16 im2col operand of the NEXT layer (dilation dil_next), col[(b,t), c*3+k] =
 *   a[b, t-(2-k)*dil_next, c] (0 for negative time); act = fp32 activation a[B*T, C]. */
int pcaa_tcn_bn_elu_next(const float* y, const double* stats, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float momentum, float eps, const float* scale,
                         const float* shift, float* coef_out, int64_t B, int T, int C, int dil_next, void* col, float* act,
                         pcaa_stream stream);
/* backward pass 1: dz[B*T, C] = d * ELU'(scale*y + shift), stats2[2*C] += [sum dz, sum dz*xhat]; d comes from
 *   src_mode 0: src[B*T, C];  1: the col2im of the layer above's d-im2col src[B*T, C*3] (its dilation dil_up);
 *   2: src[B, C] / T broadcast over the T frames (backward of AvgPool1d(NSTEPS), models.py:249, 284). */
int pcaa_tcn_elu_bwd_stats(const float* src, int src_mode, int dil_up, const float* y, const float* scale,
                           const float* shift, const float* mean, const float* invstd, float* dz, double* stats2,
                           int64_t B, int T, int C, pcaa_stream stream);
/* backward pass 2: BatchNorm1d backward from the completed sums, dy (bf16 [R, C]) = c1*dz + c2*y + c3; d gamma = sum dz*xhat,
 *   d beta = sum dz (nullable) */
int pcaa_tcn_bn_bwd_apply(const float* dz, const float* y, const double* stats2, const float* scale, const float* mean,
                          const float* invstd, float* dgamma, float* dbeta, void* dy, int64_t R, int C,
                          pcaa_stream stream);
/* out[g,c] = mean_i x[g,i,c] (AvgPool1d(NSTEPS), models.py:249,284) and its backward */
int pcaa_mean_rows(const float* x, float* out, int64_t G, int n, int C, pcaa_stream stream);
int pcaa_mean_rows_bwd(const float* g, float* dx, int64_t G, int n, int C, pcaa_stream stream);
/* mean cross-entropy of torch.nn.CrossEntropyLoss (PCAA_ablation.py:1009), its gradient * gscale, argmax class */
int pcaa_softmax_ce(const float* logits, const int64_t* gt, float* loss, float* dlogits, float gscale, int32_t* pred,
                    int64_t B, int C, pcaa_stream stream);

/* ---- SeqChamferLoss, utils.py:98-132 ------------------------------------------------------------------------
 * preds, gts (B,F,T,N) fp32.  frame_loss[b,t] = sum_j min_i P + sum_i min_j P with
 * P[i,j] = |gt_i|^2 + |pred_j|^2 - 2 gt_i.pred_j; argmins (lowest index on ties) are optional outputs. */
int pcaa_chamfer_fwd(const float* preds, const float* gts, int64_t B, int F, int T, int N, float* frame_loss,
                     int32_t* idx_gt_for_pred, int32_t* idx_pred_for_gt, pcaa_stream stream);
/* avg_out != 0: out[0] = mean over (b,t); else out[b] = mean over t (utils.py:104-107) */
int pcaa_chamfer_reduce(const float* frame_loss, int64_t B, int T, int avg_out, float* out, pcaa_stream stream);
/* grad_preds = gout * d loss / d preds through the saved argmins (gout: scalar if avg_out else [B]) */
int pcaa_chamfer_bwd(const float* preds, const float* gts, const int32_t* idx_gt_for_pred,
                     const int32_t* idx_pred_for_gt, const float* gout, int avg_out, int64_t B, int F, int T, int N,
                     float* grad_preds, pcaa_stream stream);

/* P[b,t,i,j] of SeqChamferLoss.batch_pairwise_dist(x, y), utils.py:109-132; x, y (B,F,T,N) -> P (B,T,N,N) */
int pcaa_pairwise_dist(const float* x, const float* y, int64_t B, int F, int T, int N, float* P, pcaa_stream stream);

/* ---- conditional critic (CGDiscriminator, models.py:405-421) and the WGAN-GP step ---------------------------
 * Weights: W1 [64, 32+C], W2 [32,64], W3 [1,32].  labels int64 [B]; one-hot is formed in-kernel.
 * pcaa_wgangp_dstep = PCAA_ablation.py:905-973: z = z0 + means[label]; d_loss = mean D(fv) - mean D(z)
 *   + gp_weight * mean((||grad_x D(z + alpha (fv - z))|| - 1)^2), gradients w.r.t. all critic weights in closed
 *   form (double backward through Linear/ELU done analytically).  losses[4] = {d_loss, gp, mean_fake, mean_real}.
 *   grads (ADDED to, caller zeroes): gW1,gb1,gW2,gb2,gW3,gb3. */
int pcaa_wgangp_dstep(const float* fv, const float* z0, const float* means, const int64_t* labels,
                      const float* alphas, const float* W1, const float* b1, const float* W2, const float* b2,
                      const float* W3, const float* b3, float gp_weight, float* losses,
                      float* gW1, float* gb1, float* gW2, float* gb2, float* gW3, float* gb3,
                      int64_t B, int C, pcaa_stream stream);
/* out[b] = D(x_b, onehot(label_b)); dx[b,:] = dx_scale * d out[b] / d x_b; out_sum[0] = out_scale * sum_b out[b]
 * (any of the three outputs may be null).  With dx_scale = out_scale = -ADV_WEIGHT/B this is the generator's
 * adversarial loss and its gradient w.r.t. the embeddings (PCAA_ablation.py:996-1000). */
int pcaa_disc_fwd(const float* x, const int64_t* labels, const float* W1, const float* b1, const float* W2,
                  const float* b2, const float* W3, const float* b3, float* out, float* dx, float dx_scale,
                  float* out_sum, float out_scale, int64_t B, int C, pcaa_stream stream);

/* ---- Adam, torch.optim.Adam defaults as used at PCAA_ablation.py:821-833 ------------------------------------
 * flat buffers of n floats; step is the 1-based step count; g is multiplied by grad_scale first (1/world for DP);
 * when shadow_bf16 is non-null the updated parameter is also written there as bf16 (tensor-core operand copy). */
int pcaa_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int step, float grad_scale, void* shadow_bf16, pcaa_stream stream);
/* ---- batch assembly: the step before the path (MSRadarDataset.__getitem__ + default collate, datasets.py:466-479) ----
 * dst[r, :] = src[idx[r], :] for rows of row_bytes bytes (a crop is one row of 4*30*nmax fp32): gathers a batch from a
 * packed crop store resident in HBM.  row_bytes % 16 == 0; an index outside [0, n_src) gives a zero row. */
int pcaa_gather_rows(const void* src, const int64_t* idx, void* dst, int64_t n_idx, int64_t row_bytes, int64_t n_src,
                     pcaa_stream stream);
/* dst[i] += sum_{k < nsrc} src[k * src_stride + i], i < n: local reduction step of the data-parallel gradient exchange
 * over NVLink peer memory (the only exchange of the path; the reference is single-GPU).  n, src_stride
 * multiples of 4; 16-byte aligned buffers. */
int pcaa_sum_into(float* dst, const float* src, int64_t n, int64_t src_stride, int nsrc, pcaa_stream stream);
/* dst[i] = sum_k src[k*src_stride + i], k = 0 .. nsrc-1 in that order (overwrite): the one-shot reduction of a small span
 * whose per-rank copies were gathered in RANK order, so that every rank computes bit-identical sums */
int pcaa_sum_rows(float* dst, const float* src, int64_t n, int64_t src_stride, int nsrc, pcaa_stream stream);
/* CUDA-graph form of the same update: the step counter lives on the device.  pcaa_adam_advance increments
 * step_dev[0] and writes coef_dev = { lr / (1 - beta1^step), 1 / sqrt(1 - beta2^step) } (the bias corrections
 * torch.optim.Adam computes on the host); pcaa_adam_flat_dev reads them, so a captured train step advances the
 * optimizer on every replay. */
int pcaa_adam_advance(int32_t* step_dev, float* coef_dev, float lr, float beta1, float beta2, pcaa_stream stream);
int pcaa_adam_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2, float eps,
                       const float* coef_dev, float grad_scale, void* shadow_bf16, pcaa_stream stream);

/* ---- open-set scoring, inference_PCAA.py:129-136, 255-271 ---------------------------------------------------
 * loglik[i] = log( (1/C) sum_c N(emb_i; mu_c, I_D) ) in float64 (log-domain restatement).
 * vote: windows of k consecutive samples; n_above = #(loglik > log_thr); n_above > k/2 -> lowest most-frequent
 * class among pred, else n_labels ("unknown"). */
int pcaa_openset_score(const float* emb, const float* means, int64_t M, int C, int D, double* loglik,
                       pcaa_stream stream);
int pcaa_openset_vote(const double* loglik, const int32_t* pred, int64_t n_windows, int k, double log_thr,
                      int n_labels, int32_t* out, pcaa_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* PCAA_H_ */
