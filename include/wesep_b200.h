/* wesep_b200 — C-ABI of the B200-native (sm_100a) target-speaker-extraction train-step kernels.
 *
 * This is the drop-in boundary: the reference (wenet-e2e/wesep) has no native
 * code on its training path — it calls stock torch.nn modules.  Each entry point below replaces the
 * ATen/cuDNN/cuBLAS work behind one reference nn.Module.forward (+ its autograd backward); the
 * reference file:line is cited per function.  Host side = Python nn.Modules with the reference's
 * class names / ctor kwargs / state_dict keys (wesep_b200/models, wesep_b200/modules) that bind
 * these symbols with ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, POD argument structs, no torch types.
 *   - every pointer is a DEVICE pointer unless named host_*; `stream` is a cudaStream_t.
 *   - asynchronous on `stream`; no allocation, no host sync; caller owns all memory.
 *   - activations are fp32 `[n][C][ld]`: time contiguous, row stride `ld` floats (ld % 4 == 0,
 *     ld >= T, base 16-byte aligned), batch stride = C*ld.  Columns t in [T, ld) are padding:
 *     never read as data; their contents are unspecified.
 *   - return 0 on success, -1 bad shape/alignment, -2 unsupported configuration, -3 CUDA error
 *     (text via wesep_b200_last_error()).
 *   - gLN statistics buffers are `double[n][2]` = (sum, sum of squares) over (C,T) of the
 *     normalised tensor; they must be zeroed by the caller before the producing call unless noted.
 *   - parameter-gradient outputs ACCUMULATE (+=) into caller-zeroed buffers.
 */
#ifndef WESEP_B200_H_
#define WESEP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WESEP_B200_VERSION 100

int wesep_b200_version(void);
/* Last error text of the calling thread ("" if none). */
const char* wesep_b200_last_error(void);
/* Number of kernels launched by this library in this process (diagnostic; bench.py gpu_launches). */
uint64_t wesep_b200_launch_count(void);
/* GEMM precision: 0 = fp32-grade split products (default), 1 = single-pass TF32. Process-wide DEFAULT: every GEMM entry point
 * also takes the choice per call (`mode_sel` / `backend_sel` of WesepGemmArgs / WesepGemmDwArgs). */
int wesep_b200_set_gemm_mode(int mode);
/* GEMM backend: 0 = legacy tensor path (mma.sync), 1 (default) = tcgen05/UMMA + TMA + TMEM where the shape is eligible
 * (M % 128 == 0, Kd % 16 == 0, Kd <= 512, a workspace is supplied); other shapes stay on backend 0. Process-wide. */
int wesep_b200_set_gemm_backend(int backend);
/* tcgen05 debug flags. bit 0: also store the explicitly truncated "hi" operand tile (default off: the tensor core
 * ignores the 13 low mantissa bits of tf32 inputs — measured identical results — so the raw tile serves as hi).
 * bit 1: disable the 2-CTA (cta_group::2) GEMM variants.  bits 2-3: transform-warp groups (0 = default 2, 1 = one group, 2 = two, 3 = four; clamped so it divides the ring depth).
 * bits 4-7: timing experiments on the 2-CTA conv GEMM (16 no epilogue loads, 32 no epilogue stores, 64 no operand
 * transform, 128 single store box) — results are WRONG with 16/32/64 set: those three are compiled out of release builds
 * (-DWESEP_TC_DEBUG enables them; otherwise the call returns -2).  bit 8 (256): force the balanced stream-K split in the
 * weight-gradient GEMMs.  bit 9 (512): weight-gradient GEMMs on the 1-CTA kernel.  bit 10 (1024): 2-CTA kernels in 3xTF32
 * instead of the mixed tf32 + bf16 split product (A/B timing, tests). */
int wesep_b200_set_tc_flags(int flags);
/* Workspace bytes the tcgen05 GEMM needs for an [M x Kd] weight (split hi/lo copies). */
int64_t wesep_b200_gemm_ws_bytes(int M, int Kd);

/* ------------------------------------------------------------------------------------------------
 * SI-SDR loss (replaces auraloss.time.SISDRLoss used at wesep/utils/losses.py:24-25, called at
 * wesep/utils/executor.py:115-122).  Up to 4 estimates share one target (Spex+ est1..3).
 * loss_i = -(1/n) sum_rows 10 log10( |a t~|^2 / (|x~ - a t~|^2 + eps) + eps ),  zero-mean, eps=1e-8.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n_est;             /* 1..4 */
  int n;                 /* rows */
  int L;                 /* samples per row */
  const float* est[4];   /* [n][ld_est] */
  int64_t ld_est[4];
  const float* tgt;      /* [n][ld_tgt] */
  int64_t ld_tgt;
  double* sums;          /* workspace [n_est][n][5] (Sx,St,Sxt,Sxx,Stt); zeroed by the call */
  float* sisdr_rows;     /* out [n_est][n] per-row SI-SDR in dB (positive = good) */
  float* loss;           /* out [n_est] = -mean_rows(sisdr) */
} WesepSisdrFwdArgs;
int wesep_b200_sisdr_fwd(const WesepSisdrFwdArgs* a, void* stream);

typedef struct {
  int n_est, n, L;
  const float* est[4];
  int64_t ld_est[4];
  const float* tgt;
  int64_t ld_tgt;
  const double* sums;    /* from the forward call */
  const float* gloss;    /* [n_est] upstream d(total)/d(loss_i) (device) */
  float* gest[4];        /* out [n][ld_gest] gradient wrt est_i (overwritten) */
  int64_t ld_gest[4];
} WesepSisdrBwdArgs;
int wesep_b200_sisdr_bwd(const WesepSisdrBwdArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-tensor gradient clip + Adam with coupled L2 decay over a flat fp32 arena (replaces
 * clip_gradients wesep/utils/funcs.py:79-88 + torch.optim.Adam(weight_decay) wesep/bin/train.py:237).
 * Tensors are segments [seg_off[i], seg_off[i+1]) of the arenas (offsets multiples of 4).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t total;            /* arena length in floats */
  int n_seg;
  const int64_t* seg_off;   /* device [n_seg+1] */
  const int32_t* chunk_seg; /* device [n_chunk]: segment of each 4096-float chunk (chunks never straddle) */
  const int64_t* chunk_off; /* device [n_chunk]: first element of the chunk */
  int n_chunk;
  float* param;             /* [total] */
  float* grad;              /* [total]  (scaled by grad_scale before clipping; modified in place when clipped) */
  float* exp_avg;           /* [total] */
  float* exp_avg_sq;        /* [total] */
  double* sumsq;            /* workspace [n_seg]; zeroed by the call */
  float* norms;             /* out [n_seg] per-tensor L2 norm (after grad_scale) */
  float grad_scale;         /* e.g. 1/world_size after an all-reduce(sum) */
  float clip;               /* <= 0 disables clipping */
  float lr, beta1, beta2, eps, weight_decay;
  int step;                 /* 1-based */
  const float* dyn;         /* optional device [3] = {lr, 1 - beta1^step, sqrt(1 - beta2^step)}: when given, these replace
                               `lr` / `step` so that a CUDA graph of the train step can be replayed with a moving schedule */
} WesepClipAdamArgs;
int wesep_b200_clip_adam(const WesepClipAdamArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 1x1-conv GEMM  Y[n][M][T] = W[M][Kd] * f(X[n][Kd][T]) (+ epilogue)  — the pointwise convs of
 * wesep/modules/tasnet/{convs,encoder,decoder,speaker}.py.  Exposed for tests and for the host
 * modules that are not fused TCN blocks.
 *   pro:  0 identity | 1 prelu(alpha) | 2 scale_c*prelu(x; alpha)+shift_c with per-(row,channel)
 *         scale/shift built from (ch_scale, ch_shift, row_stats): PReLU then gLN-apply (TCN blocks)
 *         | 3 prelu(scale_c*x+shift_c; alpha): BatchNorm-apply then PReLU (speaker ResBlock).
 *   epi:  0 Y=acc+bias | 1 Y=relu(acc+bias) | 2 Y=acc+bias+R (residual) |
 *         3 Y=aux*relu(acc+bias), Y2=relu(acc+bias) (decoder mask, decoder.py:96-102)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n, M, Kd, T;
  const float* W; int64_t ldw; int w_trans;   /* w_trans: W stored [Kd][M] (ldw >= M) */
  const float* X; int64_t ldx;                /* [n][Kd][ldx] */
  float* Y; int64_t ldy;                      /* [n][M][ldy] */
  const float* bias;                          /* [M] or NULL */
  const float* row_bias;                      /* [n][M] or NULL (added to bias) */
  int pro;
  const float* alpha;                         /* PReLU slope (device scalar) for pro 1/2; NULL = 1 */
  const float* ch_scale; const float* ch_shift; /* [Kd] or NULL */
  const double* row_stats; double stat_count; float stat_eps; /* gLN of X: [n][2], count=C*T */
  int epi;
  const float* R; int64_t ldr;                /* residual [n][M][ldr] (epi 2) / aux (epi 3) */
  float* Y2; int64_t ldy2;                    /* second output (epi 3) */
  double* out_stats;                          /* optional [n][2]: += sum / sumsq of prelu(Y; out_alpha) */
  const float* out_alpha;
  double* ch_stats;                           /* optional [M][2]: += per-channel sum / sumsq of Y (BatchNorm) */
  int64_t bsx, bsy, bsr, bsy2;                /* batch strides in floats; 0 = dense (C*ld): channel-slices of wider tensors */
  void* ws; int64_t ws_bytes;                 /* optional workspace (>= wesep_b200_gemm_ws_bytes(M, Kd)) enabling backend 1 */
  int mode_sel, backend_sel;                  /* per call: 0 = the process default (wesep_b200_set_gemm_mode / _backend), 1 / 2 = mode or
                                                 backend 0 / 1 for this launch only */
} WesepGemmArgs;
int wesep_b200_conv1x1(const WesepGemmArgs* a, void* stream);

/* Weight-gradient GEMM  C[M][N] += sum_n sum_t fa(A[n][M][t]) * fb(B[n][N][t]);  per_row: C is
 * [n][M][N] (no sum over n).  pro_* as above (row_stats give per-row mean/rstd: fb = (prelu(b)-mu)*r
 * when ch_scale==NULL). */
typedef struct {
  int n, M, N, T;
  const float* A; int64_t lda;  /* [n][M][lda] */
  const float* B; int64_t ldb;  /* [n][N][ldb] */
  float* C; int per_row;
  int pro_b;
  const float* alpha_b;
  const float* ch_scale_b; const float* ch_shift_b;
  const double* row_stats_b; double stat_count; float stat_eps;
  int64_t bsa, bsb;             /* batch strides in floats; 0 = dense */
  int64_t ldc;                  /* row stride of C; 0 = N */
  int mode_sel, backend_sel;    /* per call, as in WesepGemmArgs */
} WesepGemmDwArgs;
int wesep_b200_conv1x1_dw(const WesepGemmDwArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused TCN block (Conv1DBlock wesep/modules/tasnet/convs.py:43-104, skip_con False, non-causal,
 * gLN; and Conv1DBlock4Fuse :107-160 when aux != NULL: the speaker half of conv1x1.weight is
 * folded into a per-row bias).
 *   u   = W1[:, :B] x + b1 (+ W1[:, B:] aux_n)       y1 = prelu(u, a1)      z1 = gLN1(y1)
 *   d   = dwconv(z1; wd, bd, dilation)               y2 = prelu(d, a2)      z2 = gLN2(y2)
 *   out = x + W3 z2 + b3
 * Saved for backward: u, d (pre-activations), stats1, stats2.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n, B, H, T, dil, E;       /* E = speaker-embedding dim (0 for plain block) */
  int64_t ld;                   /* row stride of every activation below */
  const float* x;               /* [n][B][ld] */
  const float* aux;             /* [n][E] or NULL */
  const float* W1; int64_t ldw1;/* [H][B+E] */
  const float* b1;              /* [H] */
  const float* a1;              /* [1] */
  const float* g1; const float* be1; /* gLN1 weight/bias [H] */
  const float* wd;              /* [H][3] */
  const float* bd;              /* [H] */
  const float* a2;
  const float* g2; const float* be2;
  const float* W3; int64_t ldw3;/* [B][H] */
  const float* b3;              /* [B] */
  float* u; float* d;           /* out [n][H][ld] */
  float* out;                   /* out [n][B][ld] */
  double* stats1; double* stats2; /* out [n][2] each (zeroed by the call) */
  float* row_bias;              /* workspace [n][H] (fuse block only) */
  void* ws; int64_t ws_bytes;   /* optional GEMM workspace (>= wesep_b200_gemm_ws_bytes(H, max(B,H))) enabling backend 1 */
} WesepTcnFwdArgs;
int wesep_b200_tcn_block_fwd(const WesepTcnFwdArgs* a, void* stream);

typedef struct {
  WesepTcnFwdArgs f;            /* same tensors as forward (out unused) */
  const float* gout;            /* [n][B][ld] dL/d out */
  float* dx;                    /* out [n][B][ld] */
  /* parameter gradients (+=) */
  float* dW1; float* db1; float* da1; float* dg1; float* dbe1;
  float* dwd; float* dbd; float* da2; float* dg2; float* dbe2;
  float* dW3; float* db3;
  float* daux;                  /* [n][E] (overwritten) or NULL */
  /* workspace */
  float* dd; float* du;         /* [n][H][ld] each */
  float* Gn;                    /* [n][B][H] */
  float* sg;                    /* [n][B] */
  float* sdu;                   /* [n][H] */
  double* rowsc;                /* [n][8] */
} WesepTcnBwdArgs;
int wesep_b200_tcn_block_bwd(const WesepTcnBwdArgs* a, void* stream);
/* bytes of the zero-initialised scratch the backward call needs are all caller-provided above. */

/* ------------------------------------------------------------------------------------------------
 * Channel-wise LayerNorm over C per frame (ChannelWiseLayerNorm wesep/modules/common/norm.py:51-66).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n, C, T; int64_t ldx, ldy;
  const float* x; float* y;
  const float* gamma; const float* beta; float eps;
  float* mean; float* rstd;     /* out [n][T] (saved for backward) */
} WesepClnFwdArgs;
int wesep_b200_cln_fwd(const WesepClnFwdArgs* a, void* stream);
typedef struct {
  int n, C, T; int64_t ldx, ldg, lddx;
  const float* x; const float* gy; float* dx;
  const float* gamma; const float* mean; const float* rstd;
  float* dgamma; float* dbeta;  /* += [C] */
} WesepClnBwdArgs;
int wesep_b200_cln_bwd(const WesepClnBwdArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Framing / overlap-add for the strided encoder / transposed decoder convs
 * (MultiEncoder wesep/modules/tasnet/encoder.py:95-110, MultiDecoder decoder.py:104-108).
 *   frames:      F[n][j][k] = x[n][k*hop + j]  (0 beyond the signal), j < J, k < K
 *   overlap_add: y[n][s]    = bias + sum_{j,k: k*hop+j == s} F[n][j][k],  s < S
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n, J, K, hop; int64_t S;
  const float* x; int64_t ldx;  /* [n][ldx] signal of S valid samples */
  float* F; int64_t ldf;        /* [n][J][ldf] */
} WesepFrameArgs;
int wesep_b200_frames(const WesepFrameArgs* a, void* stream);
typedef struct {
  int n, J, K, hop; int64_t S;
  const float* F; int64_t ldf;
  const float* bias;            /* device scalar or NULL */
  float* y; int64_t ldy;
} WesepOlaArgs;
int wesep_b200_overlap_add(const WesepOlaArgs* a, void* stream);

/* Elementwise helpers used by the host modules' backward passes. */
typedef struct {
  int n, C, T; int64_t ld;
  const float* gS;              /* [n][C][ld] dL/dS, S = w * m, m = relu(mask conv) */
  const float* w; const float* m;
  float* gw;                    /* out dL/dw += ... (accumulate: w feeds three paths) or overwrite, see acc_w */
  float* gm;                    /* out dL/d(pre-relu mask) */
  int acc_w;
} WesepMaskBwdArgs;
int wesep_b200_mask_bwd(const WesepMaskBwdArgs* a, void* stream);

typedef struct {
  int n, C, T; int64_t ld;
  const float* y;               /* relu output */
  const float* gy; float* gx;   /* gx = gy * (y > 0) */
} WesepReluBwdArgs;
int wesep_b200_relu_bwd(const WesepReluBwdArgs* a, void* stream);

/* Row sums over time: out[n][c] = sum_t x[n][c][t]  (bias gradients). */
typedef struct {
  int n, C, T; int64_t ld;
  const float* x; float* out;   /* out [n][C] overwritten */
} WesepRowSumArgs;
int wesep_b200_rowsum(const WesepRowSumArgs* a, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Speaker-encoder ResBlock pieces (ResBlock wesep/modules/tasnet/speaker.py:7-45): BatchNorm1d with batch
 * statistics (the pointwise convs accumulate per-channel (sum, sumsq) in their GEMM epilogue: WesepGemmArgs.ch_stats),
 * BN-apply + residual + PReLU + MaxPool1d(3), and the two-pass backward.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int C; double count;              /* count = n*T elements per channel */
  const double* ch_stats;           /* [C][2] (sum, sumsq) — training only */
  const float* weight; const float* bias;
  float* running_mean; float* running_var; float momentum; float eps; int training;
  float* scale; float* shift;       /* out [C]: y = scale*x + shift */
  float* mean; float* rstd;         /* out [C] (saved for backward) */
} WesepBnFinalizeArgs;
int wesep_b200_bn_finalize(const WesepBnFinalizeArgs* a, void* stream);

typedef struct {
  int n, C, T, pool;                /* pool = 1 (none) or 3 (MaxPool1d(3)) */
  int64_t ldx, ldr, ldy;
  const float* x; const float* res; /* res may be NULL */
  const float* scale; const float* shift; const float* alpha;
  float* y;                         /* [n][C][ldy], T/pool frames */
} WesepBnActPoolFwdArgs;
int wesep_b200_bn_act_pool_fwd(const WesepBnActPoolFwdArgs* a, void* stream);

typedef struct {
  int n, C, T, pool;
  int64_t ldx, ldr, ldgy, ldgv;
  const float* x; const float* res;
  const float* scale; const float* shift; const float* alpha; const float* mean; const float* rstd;
  const float* gy;                  /* [n][C][ldgy] gradient of the pooled output */
  float* gv;                        /* out [n][C][ldgv]: gradient w.r.t. v = scale*x+shift(+res) */
  double* ch_sums;                  /* += [C][2]: sum gv, sum gv*xhat (zeroed by the caller) */
  float* dalpha;                    /* += [1] */
} WesepBnActPoolBwdArgs;
int wesep_b200_bn_act_pool_bwd(const WesepBnActPoolBwdArgs* a, void* stream);

typedef struct {
  int n, C, T, training; int64_t ldgv, ldx, lddx; double count;
  const float* gv; const float* x; float* dx;
  const float* scale; const float* mean; const float* rstd;
  const double* ch_sums;
} WesepBnBwdArgs;
int wesep_b200_bn_bwd(const WesepBnBwdArgs* a, void* stream);

/* pred_linear (nn.Linear, wesep/models/convtasnet.py:115,194) and nn.CrossEntropyLoss (wesep/utils/losses.py:11). */
typedef struct {
  int n, J, K;
  const float* x;                   /* [n][K] */
  const float* W; const float* b;   /* [J][K], [J] or NULL */
  float* y;                         /* fwd out [n][J] */
  const float* gy;                  /* bwd in  [n][J] */
  float* dW; float* db; float* dx;  /* bwd out (overwritten) */
} WesepLinearArgs;
int wesep_b200_linear_fwd(const WesepLinearArgs* a, void* stream);
int wesep_b200_linear_bwd(const WesepLinearArgs* a, void* stream);

typedef struct {
  int n, J;
  const float* logits;              /* [n][J] */
  const int64_t* labels;            /* [n] */
  float* loss;                      /* out [1] mean over rows (zeroed by the call) */
  float* dlogits;                   /* out [n][J] d loss / d logits */
} WesepCeArgs;
int wesep_b200_cross_entropy(const WesepCeArgs* a, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Alternative speaker fusion in front of each TCN repeat (FuseSeparation, spk_fuse_type in {concat, additive,
 * multiply, FiLM}: wesep/modules/tasnet/separation.py:116-135,172-181; SpeakerFuseLayer common/speaker.py:81-125;
 * FiLM common/norm.py:118-139) followed by the stand-alone nn.PReLU and gLN of the reference:
 *     z = gLN( PReLU( ra[n][c] * x + rb[n][c] ; alpha ) ; gamma, beta )
 * fwd fills `stats` (double [n][2], saved); bwd needs them plus a [n][2] scratch `rowsums`.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n, C, T; int64_t ldx, ldy, ldg, lddx;
  const float* x;                   /* [n][C][ldx] */
  const float* ra; const float* rb; /* [n][C] per-row channel scale / shift, NULL = 1 / 0 */
  const float* alpha; const float* gamma; const float* beta;
  double* stats;                    /* [n][2] */
  float* y;                         /* fwd out [n][C][ldy] */
  const float* gz;                  /* bwd in  [n][C][ldg] */
  double* rowsums;                  /* bwd scratch [n][2] */
  float* dx;                        /* bwd out [n][C][lddx] */
  float* dra; float* drb;           /* bwd out [n][C] (overwritten; may be NULL) */
  float* dalpha; float* dgamma; float* dbeta;  /* bwd += [1], [C], [C] */
  float eps;                        /* variance epsilon of the normalisation; 0 = 1e-5 (gLN).  With alpha = 1 and eps =
                                       FLT_EPSILON this is nn.GroupNorm(1, C) of pBSRNN (bsrnn.py:24,256,274) */
} WesepFuseArgs;
int wesep_b200_fuse_prelu_gln_fwd(const WesepFuseArgs* a, void* stream);
int wesep_b200_fuse_prelu_gln_bwd(const WesepFuseArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * pBSRNN building blocks (wesep/models/bsrnn.py).  ResRNN runs in a TIME-MAJOR layout
 * [S][C][ld(Q)] (rows = time steps, columns = sequences) so that one step of the recurrence is a conv1x1 GEMM
 * [4Hd x Hd] . [Hd x Q]; these entry points provide the layout change and the LSTM cell.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int nb, Q, C, S; int64_t ld_in, ld_out;
  const float* in;    /* [nb][Q][C][ld_in],  S valid columns */
  float* out;         /* [nb][S][C][ld_out], Q valid columns: out[b][s][c][q] = in[b][q][c][s] (+ res[b][s][c][q]) */
  const float* res;   /* optional, laid out like `out` (the ResRNN residual, bsrnn.py:46) */
} WesepTransposeArgs;
int wesep_b200_swap_outer_inner(const WesepTransposeArgs* a, void* stream);

/* One time step of one LSTM direction (nn.LSTM gate order i, f, g, o; bsrnn.py:25-31) for Q sequences.
 * fwd: G = gate pre-activations in, activations out (saved); c = f*c_prev + i*g; h = o*tanh(c).
 * bwd: G = saved activations in, d(pre-activations) out; dh / dc_in = gradients w.r.t. h_s / c_s; dc_prev out. */
typedef struct {
  int Hd, Q; int64_t ld;
  float* G;               /* [4*Hd][ld] */
  const float* c_prev;    /* [Hd][ld] or NULL (zero state) */
  float* c;               /* [Hd][ld]: fwd out, bwd in */
  float* h;               /* [Hd][ld]: fwd out */
  const float* dh;        /* bwd in  [Hd][ld] */
  const float* dc_in;     /* bwd in  [Hd][ld] or NULL */
  float* dc_prev;         /* bwd out [Hd][ld] */
} WesepLstmCellArgs;
int wesep_b200_lstm_cell_fwd(const WesepLstmCellArgs* a, void* stream);
int wesep_b200_lstm_cell_bwd(const WesepLstmCellArgs* a, void* stream);

/* nn.GroupNorm(1, C, eps) over (C, T) of every row of [n][C][ld] (the norm in front of every ResRNN, band split and mask
 * head: wesep/models/bsrnn.py:23,39,204,274), one CTA per row.  fwd writes stats [n][2] = (sum, sum of squares) for the
 * backward.  bwd: dx, and dgamma / dbeta [C] ACCUMULATED (+=) into the caller's (zero-initialised or running) buffers.
 * x / y / gy / dx may be channel slices of larger tensors (batch strides bs*). */
typedef struct {
  int n, C, T;
  int64_t ldx, bsx, ldy, bsy, ldg, bsg, lddx, bsdx;   /* row and batch strides (floats), multiples of 4 */
  const float* x; const float* gamma; const float* beta; float eps;
  float* y; double* stats;
  const float* gy; float* dx; float* dgamma; float* dbeta;
  double* bsum;          /* bwd workspace [n][2] for rows of >= 2^19 elements (a (chunk, channel, row) grid with fp64 atomics instead
                            of one CTA per row: TF-GridNet's utterance-wide GroupNorm); may be NULL for shorter rows */
} WesepGroupNorm1Args;
int wesep_b200_groupnorm1_fwd(const WesepGroupNorm1Args* a, void* stream);
int wesep_b200_groupnorm1_bwd(const WesepGroupNorm1Args* a, void* stream);

/* ---- wespeaker ResNet speaker encoder (joint training of pBSRNN: wesep/models/bsrnn.py:217,352-356; bsrnn.yaml:56-64).
 * Feature maps are [n][C][H*W] with W (time) contiguous.  conv3x3 (pad 1, stride 1 / 2, no bias) = im2col3x3 + conv1x1 over the
 * 9 C gathered channels (weight [Cout][Cin*9] = the Conv2d weight viewed 2-D); the adjoint of im2col is a gather. */
typedef struct {
  int n, C, H, W, stride, Ho, Wo;       /* Ho = (H - 1) / stride + 1, likewise Wo */
  int64_t ldx, ldc, bsc;                /* row strides of x / gx ([n][C][ldx]) and col / gcol (rows of ldc, batch stride bsc >= 9 C ldc;
                                           subsample: [n][C][ldc], bsc = C ldc) */
  const float* x; float* col;           /* fwd */
  const float* gcol; float* gx;         /* bwd */
  int stride_w;                         /* im2col only: stride along W when it differs from `stride` (DPCCN: (1, 2)); 0 = same */
} WesepIm2colArgs;
int wesep_b200_im2col3x3_fwd(const WesepIm2colArgs* a, void* stream);
int wesep_b200_im2col3x3_bwd(const WesepIm2colArgs* a, void* stream);
int wesep_b200_subsample2d_fwd(const WesepIm2colArgs* a, void* stream);   /* x[.., ::s, ::s] for the 1x1 stride-s shortcut conv */
int wesep_b200_subsample2d_bwd(const WesepIm2colArgs* a, void* stream);

/* BatchNorm2d with batch statistics over (n, H, W) on [n][C][ld] (T = H*W valid columns), fused with the residual add and
 * the ReLU of a BasicBlock.  stats / bsum: fp64 [C][2] accumulators the CALLER zeroes; the host side turns stats into
 * scale / shift / mean / rstd ([C] vectors) and updates the running buffers.
 *   bn2_stats : stats[c] += (sum x, sum x^2)
 *   bn2_apply : y = act(scale[c] x + shift[c] (+ res)), act = ReLU if relu
 *   bn2_bwd   : g' = gy * (y > 0) if relu; bsum[c] += (sum g', sum g' xhat); gx = gamma rstd (g' - bsum0/count - xhat bsum1/count);
 *               gres = g' if not NULL (gradient of the residual branch); dgamma = bsum[c][1], dbeta = bsum[c][0] */
typedef struct {
  int n, C, T, relu; int64_t ld;        /* every tensor [n][C][ld] */
  double count;                         /* n * T (bwd) */
  const float* x; const float* res; float* y;
  const float* scale; const float* shift; double* stats;
  const float* gy; const float* mean; const float* rstd; const float* gamma; double* bsum; float* gx; float* gres;
} WesepBn2Args;
int wesep_b200_bn2_stats(const WesepBn2Args* a, void* stream);
int wesep_b200_bn2_apply(const WesepBn2Args* a, void* stream);
int wesep_b200_bn2_bwd(const WesepBn2Args* a, void* stream);

/* TSTP pooling (wespeaker pooling_layers.TSTP): x [n][R][ld] -> out [n][2R] = (mean over time | sqrt(unbiased var + 1e-7)). */
typedef struct {
  int n, R, T; int64_t ld;
  const float* x; float* out;
  const float* gout; float* gx;
  float eps;             /* added to the variance under the square root; 0 = 1e-7 (TSTP); ASTP's global context uses 1e-10 */
} WesepTstpArgs;
int wesep_b200_tstp_fwd(const WesepTstpArgs* a, void* stream);
int wesep_b200_tstp_bwd(const WesepTstpArgs* a, void* stream);

/* The LSTM recurrence of a bidirectional layer as ONE persistent cluster kernel per pass (time-major tensors, see above):
 * replaces the recurrent half of nn.LSTM inside ResRNN (wesep/models/bsrnn.py:25-31,41-44); the input projection
 * W_ih x + b_ih + b_hh of both directions is a GEMM done beforehand into G.  A cluster of Hd / 32 CTAs keeps W_hh (split
 * fp16 hi / lo, per-row scaled) resident in distributed shared memory for all S steps, runs the per-step product on
 * tcgen05 and exchanges h_t through DSMEM.  fwd: G holds the gate pre-activations on entry and the gate activations on
 * return, H / C receive h_t / c_t.  bwd: G (activations) is overwritten by d(pre-activations); dH holds dL/dh_t from the
 * layers above on entry (read only; the recurrent contribution stays inside the kernel).
 * Supported: Hd in {32, 64, 128, 192, 256}  (wesep_b200_lstm_rec_supported). */
typedef struct {
  int S, Q, Hd;             /* steps, sequences (columns), hidden units per direction */
  int64_t ld;               /* row stride (floats) of every tensor below, multiple of 4, >= Q */
  int64_t bsG, bsH;         /* step strides: G blocks [8*Hd][ld] (forward dir rows [0,4Hd), reverse [4Hd,8Hd)); H / C / dH blocks [2*Hd][ld] */
  float* G; float* H; float* C;
  const float* Whh_f; const float* Whh_r;   /* [4*Hd][Hd] each, nn.LSTM weight_hh_l0 / weight_hh_l0_reverse */
  const float* dH;          /* bwd only */
  int seqs_per_cluster;     /* 0 = automatic; 64 / 128 force the grouping (tests, tuning) */
  void* prof;               /* optional (NULL = off): int64 [16][16] SM-clock stamps of cluster 0 for 16 steps (tools/time_lstm_rec.py) */
} WesepLstmRecArgs;
int wesep_b200_lstm_rec_supported(int Hd);
int wesep_b200_lstm_rec_max_clusters(int Hd, int bwd);   /* co-resident clusters on the current device (-1: query failed) */
int wesep_b200_lstm_rec_fwd(const WesepLstmRecArgs* a, void* stream);
int wesep_b200_lstm_rec_bwd(const WesepLstmRecArgs* a, void* stream);

/* y = ra[n][c] * x + rb[n][c] (NULL = 1 / 0): SpeakerFuseLayer multiply / additive with the Linear hoisted out of the
 * (band, frame) loop (wesep/modules/common/speaker.py:103-121).  bwd: dx = ra * gy, dra = sum_t gy * x, drb = sum_t gy. */
typedef struct {
  int n, C, T; int64_t ld;          /* x, y, gy, dx: [n][C][ld] */
  const float* x; const float* ra; const float* rb;
  float* y;
  const float* gy; float* dx; float* dra; float* drb;   /* dra / drb [n][C], overwritten, may be NULL */
} WesepRowAffineArgs;
int wesep_b200_rowaffine_fwd(const WesepRowAffineArgs* a, void* stream);
int wesep_b200_rowaffine_bwd(const WesepRowAffineArgs* a, void* stream);

/* tanh over `rows` rows of T valid columns (mask MLP, bsrnn.py:276-280).  bwd: dx = gy * (1 - y^2). */
typedef struct {
  int64_t rows; int T; int64_t ld;
  const float* x; float* y;
  const float* gy; float* dx;
} WesepTanhArgs;
int wesep_b200_tanh_fwd(const WesepTanhArgs* a, void* stream);
int wesep_b200_tanh_bwd(const WesepTanhArgs* a, void* stream);

/* Mask head tail (bsrnn.py:365-379): o [n][4*bw][ldo] = (value | gate) x (re | im) x bw; m = value * sigmoid(gate);
 * e = s * m as a complex product, s / e = band slices (bw re rows, then bw im rows) of the mixture / estimate spectra. */
typedef struct {
  int n, bw, T; int64_t ldo, lds, lde, bso, bss, bse;   /* row and batch strides (floats) of o, s, e */
  const float* o; const float* s; float* e;
  const float* ge; float* go;                           /* bwd: ge laid out like e, go like o */
} WesepMaskApplyArgs;
int wesep_b200_mask_apply_fwd(const WesepMaskApplyArgs* a, void* stream);
int wesep_b200_mask_apply_bwd(const WesepMaskApplyArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation scoring: replaces the per-utterance numpy path of wesep/bin/infer.py:124-129
 * (peak rule: if every row of the batch has a positive sample, est = est / max|est| * 0.9 per row) and
 * wesep/utils/score.py:7-36 (cal_SISNR / cal_SISNRi) for a whole batch in three launches.
 *   sisnr[r]  = 20 log10(eps + |t| / (|x~ - t| + eps)),  t = <x~,r~> r~ / (|r~|^2 + eps), zero-mean, eps = 1e-8
 *   sisnri[r] = sisnr(est_r, ref_r) - sisnr(mix_r, ref_r)
 * Moments are taken over the first len[r] samples (infer.py:147-152 trims the three waves to the shortest);
 * the peak search and the scaling run over all L samples of est, as the reference does before trimming.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n;                 /* rows (<= 65535) */
  int L;                 /* samples per row */
  const int* len;        /* [n] valid samples per row for the scores, or NULL = L */
  float* est;            /* [n][ld_est] separated waves; scaled IN PLACE when the peak rule fires */
  int64_t ld_est;
  const float* ref;      /* [n][ld_ref] clean targets */
  int64_t ld_ref;
  const float* mix;      /* [n][ld_mix] mixtures */
  int64_t ld_mix;
  int peak_norm;         /* 0: leave est untouched */
  double* ws;            /* workspace, wesep_b200_score_ws_bytes(n) bytes; zeroed by the call */
  float* sisnr;          /* out [n] dB */
  float* sisnri;         /* out [n] dB */
  int* normed;           /* out [1] 1 if the peak rule fired (may be NULL) */
} WesepScoreArgs;
int64_t wesep_b200_score_ws_bytes(int n);
int wesep_b200_score(const WesepScoreArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side data front end.
 *
 * wesep_b200_mix: random_chunk + snr_mixer for M mixtures of S speakers in three streaming launches
 * (wesep/dataset/processor.py:536-573 get_random_chunk with host-drawn offsets; processor.py:276-320 snr_mixer):
 *   chunk_s[j] = utt_s[chunk0 + j]            (utterance >= T samples)   |   utt_s[j mod ulen]   (shorter: tiled)
 *   v_0 = chunk_0;  v_s = chunk_s * sqrt(E_0 / E_s) * 10^(snr_s / 20);  mix = v_0 + v_1 + ...
 *   everything * 1 / max(max|mix|, max_s max|v_s|)          (left unscaled if that maximum is 0)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int M;                 /* mixtures (<= 65535) */
  int S;                 /* speakers per mixture, 1..4; speaker 0 is the target of the energy match */
  int T;                 /* chunk length in samples */
  const float* pool;     /* source utterances, back to back */
  const int64_t* start;  /* [M][S] index in pool of sample 0 of each utterance */
  const int* ulen;       /* [M][S] utterance lengths (> 0) */
  const int* chunk0;     /* [M][S] chunk start inside the utterance (ignored when ulen < T) */
  const float* snr_db;   /* [M][S] SNR of speaker s against speaker 0 in dB (entry 0 unused), or NULL = 0 dB */
  float* mix;            /* out [M][ld] */
  float* spk;            /* out [S][M][ld] the scaled sources (wav_spk1.. of the reference sample) */
  int64_t ld;
  double* ws;            /* workspace, wesep_b200_mix_ws_bytes(M) bytes; zeroed by the call */
} WesepMixArgs;
int64_t wesep_b200_mix_ws_bytes(int M);
int wesep_b200_mix(const WesepMixArgs* a, void* stream);

/* wesep_b200_fbank: compute_fbank + apply_cmvn of the enrollment waves (processor.py:480-535), i.e.
 * torchaudio.compliance.kaldi.fbank(wave * 2^15, num_mel_bins, frame_length 25 ms, frame_shift 10 ms, dither,
 * window_type "hamming", use_energy False) with its defaults (snip_edges, remove_dc_offset, pre-emphasis 0.97,
 * round_to_power_of_two, power spectrum, log) followed by the subtraction of the per-utterance mean over frames.
 * Frames of row r: 1 + (len[r] - frame_len) / frame_shift (0 if shorter than one frame); rows of `out` past that
 * are zero (the "max" collate mode pads with zeros, wesep/dataset/dataset.py:229-245).  The window and the mel
 * filterbank are small host-built tables (fp32, as torchaudio builds them). */
typedef struct {
  int n;                 /* utterances (<= 65535) */
  int T;                 /* samples per row of wav */
  const float* wav;      /* [n][ld_wav] */
  int64_t ld_wav;
  const int* len;        /* [n] valid samples per row, or NULL = T */
  int frame_len;         /* 400 */
  int frame_shift;       /* 160 */
  int n_fft;             /* 512 (the only size built) */
  int num_mel;           /* 80 */
  float scale;           /* 32768 */
  float dither;          /* std of the white noise added to every frame sample; 0 = none */
  uint64_t seed;         /* dither stream */
  float preemph;         /* 0.97 */
  int remove_dc;         /* 1 */
  const float* window;   /* [frame_len] */
  const float* mel;      /* [num_mel][n_fft/2+1] triangular weights */
  const int* mel_lo;     /* [num_mel] first non-zero bin */
  const int* mel_hi;     /* [num_mel] one past the last non-zero bin */
  float log_floor;       /* fp32 epsilon */
  float* out;            /* [n][bs_out] = [n][max_frames][num_mel] */
  int64_t bs_out;
  int max_frames;
  int cmn;               /* 1: subtract the per-utterance mean over frames */
} WesepFbankArgs;
int wesep_b200_fbank(const WesepFbankArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * pDPCCN building blocks (wesep/models/dpccn.py:206-290, wesep/modules/dpccn/convs.py:28-152).
 * Feature maps are act tensors [n][C][T*F] (F contiguous = the reference's NCHW maps); the 3x3 convolutions and
 * transposed convolutions are wesep_b200_im2col3x3_{fwd,bwd} (stride (1, 2)) around wesep_b200_conv1x1.
 * ---------------------------------------------------------------------------------------------- */
/* mode 0: y = InstanceNorm(ELU(x)) (Conv2dBlock / ConvTrans2dBlock, convs.py:44-47,66-69);
 * mode 1: y = ELU(InstanceNorm(x)) (TCNBlock, convs.py:144,148).  Per row (one (n, c) plane), biased variance, eps 1e-5, no affine. */
typedef struct {
  int64_t rows; int L; int64_t ld;     /* rows = n * C planes of L valid samples, ld floats apart */
  int mode; float eps;
  const float* x; float* y;
  double* stats;                       /* workspace [rows][2], zeroed by the call */
  float* mr;                           /* [rows][2] (mean, rstd): written by fwd, read by bwd */
  const float* gy; float* gx;          /* bwd */
} WesepEluInArgs;
int wesep_b200_elu_in_fwd(const WesepEluInArgs* a, void* stream);
int wesep_b200_elu_in_bwd(const WesepEluInArgs* a, void* stream);

/* depthwise Conv1d(C, C, 3, padding = dil, dilation = dil, groups = C) of TCNBlock (convs.py:122-131) */
typedef struct {
  int n, C, L; int64_t ld; int dil;
  const float* x; const float* w; const float* b;   /* w [C][3], b [C] or NULL */
  float* y;
  const float* gy; float* gx; float* gw; float* gb; /* bwd: gw [C][3], gb [C] (or NULL) zeroed by the call */
} WesepDwConv1dArgs;
int wesep_b200_dwconv1d_fwd(const WesepDwConv1dArgs* a, void* stream);
int wesep_b200_dwconv1d_bwd(const WesepDwConv1dArgs* a, void* stream);

/* nn.AvgPool2d(k) on [rows][H*W] planes: Ho = H / k, Wo = W / k (floor; the remainder is dropped) — dpccn.py:196-204 */
typedef struct {
  int64_t rows; int H, W, k, Ho, Wo; int64_t ldx, ldy;
  const float* x; float* y;
  const float* gy; float* gx;
} WesepPool2dArgs;
int wesep_b200_avgpool2d_fwd(const WesepPool2dArgs* a, void* stream);
int wesep_b200_avgpool2d_bwd(const WesepPool2dArgs* a, void* stream);

/* nn.Upsample(size=(Ho, Wo), mode="bilinear") (align_corners False) — dpccn.py:262-265 */
typedef struct {
  int64_t rows; int Hi, Wi, Ho, Wo; int64_t ldi, ldo;
  const float* x; float* y;
  const float* gy; float* gx;          /* bwd: gx zeroed by the call */
} WesepUpsample2dArgs;
int wesep_b200_upsample2d_fwd(const WesepUpsample2dArgs* a, void* stream);
int wesep_b200_upsample2d_bwd(const WesepUpsample2dArgs* a, void* stream);

/* 4-D "multiply" speaker fusion (wesep/modules/common/speaker.py:117-121 as called at dpccn.py:240):
 * y[n][c][t*F + f] = x[n][c][t*F + f] * s[n][f];  bwd: gx = gy * s, gs[n][f] = sum over (c, t) of gy * x. */
typedef struct {
  int n, C, T, F; int64_t ld;
  const float* x; const float* s; float* y;
  const float* gy; float* gx; float* gs;            /* gs [n][F] zeroed by the call */
} WesepColScaleArgs;
int wesep_b200_colscale_fwd(const WesepColScaleArgs* a, void* stream);
int wesep_b200_colscale_bwd(const WesepColScaleArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * TF-GridNet building blocks (wesep/modules/tfgridnet/gridnet_block.py:118-284,
 * wesep/models/tfgridnet.py:197-302).  Maps are act tensors [B][H*E][T*F] (F contiguous).
 * ---------------------------------------------------------------------------------------------- */
/* PReLU (slope per head, or one slope) then LayerNorm over (E, F) of every (b, h, t) with affine gamma / beta [H][E][F]:
 * AllHeadPReLULayerNormalization4DCF (gridnet_block.py:255-284); with H = 1, E = C: nn.PReLU + LayerNormalization4DCF of
 * attn_concat_proj (gridnet_block.py:103-110,229-252).  Biased variance, eps inside the square root. */
typedef struct {
  int B, H, E, T, F; int64_t ld;
  int alpha_per_head; float eps;
  const float* x; const float* alpha; const float* gamma; const float* beta;
  float* y;
  float* mr;                           /* [B][H][T][2] (mean, rstd): written by fwd, read by bwd */
  const float* gy; float* gx;          /* bwd */
  float* dgamma; float* dbeta;         /* [H][E][F], zeroed by the call */
  float* dalpha;                       /* [H] or [1], zeroed by the call */
} WesepHeadLnArgs;
int wesep_b200_head_ln_fwd(const WesepHeadLnArgs* a, void* stream);
int wesep_b200_head_ln_bwd(const WesepHeadLnArgs* a, void* stream);

/* y[r][c] = softmax over c < C of scale * x[r][c] (attention matrix, gridnet_block.py:213-214); columns [C, ld) of y and gx are
 * zeroed so the matrix can be used as a GEMM operand with its padded row length.  bwd: gx = scale * y * (gy - sum_c gy y). */
typedef struct {
  int64_t rows; int C; int64_t ld; float scale;
  const float* x; float* y;
  const float* gy; float* gx;
} WesepSoftmaxArgs;
int wesep_b200_softmax_fwd(const WesepSoftmaxArgs* a, void* stream);
int wesep_b200_softmax_bwd(const WesepSoftmaxArgs* a, void* stream);

/* torch.std(x, dim=1) (unbiased) of every row and its reciprocal: the RMS normalisation of tfgridnet.py:217-218,297 */
typedef struct {
  int n; int L; int64_t ld;
  const float* x; float* std; float* inv_std;
} WesepRowStdArgs;
int wesep_b200_rowstd(const WesepRowStdArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * NCCL wrappers for the one collective of the path: the in-place all-reduce (SUM) of the flat gradient arena (replaces
 * torch DDP's bucketed all-reduce behind wesep/bin/train.py:227-228).  libnccl.so.2 is bound with dlopen at the first call.
 * The caller moves the 128-byte unique id from rank 0 to the other ranks over its own side channel.
 * ---------------------------------------------------------------------------------------------- */
int wesep_b200_nccl_available(void);                                      /* 1 if libnccl could be loaded */
int wesep_b200_nccl_unique_id(void* id128);                               /* rank 0: ncclGetUniqueId into 128 bytes */
int wesep_b200_nccl_comm_init_rank(void** comm, int nranks, const void* id128, int rank);   /* collective over all ranks */
int wesep_b200_nccl_allreduce_flat(float* buf, int64_t count, void* comm, void* stream);    /* in place, SUM, async on `stream` */
int wesep_b200_nccl_comm_destroy(void* comm);

/* "Consistent" in-model enrollment features (`spk_feat: False`, `feat_type: consistent`; wesep/models/bsrnn.py:231-241,343-351 and
 * the same block in convtasnet.py / dpccn.py / tfgridnet.py): PreEmphasis (wesep/modules/common/speaker.py:10-23), then
 * torchaudio.transforms.MelSpectrogram(n_fft = win, hop = stride, f_min 20, hamming window, power 2) + 1e-8, log, minus the mean
 * over frames, permuted to [B, frames, n_mels].  The STFT and the mel projection run on wesep_b200_frames / wesep_b200_conv1x1. */
typedef struct { int n, L; const float* x; int64_t ldx; float coef; float* y; int64_t ldy; } WesepPreEmphArgs;
int wesep_b200_preemphasis(const WesepPreEmphArgs* a, void* stream);     /* y[t] = x[t] - coef x[t-1], x[-1] := x[1] (reflect) */
typedef struct { int n, F, T; const float* spec; int64_t ld, bs; float* pw; int64_t ldp, bsp; } WesepPowerSpecArgs;
int wesep_b200_power_spec(const WesepPowerSpecArgs* a, void* stream);    /* pw[n][f][t] = re^2 + im^2 (spec rows [0,F) re, [F,2F) im) */
typedef struct { int n, M, T; const float* mel; int64_t ld; float eps; float* out; } WesepLogCmnArgs;
int wesep_b200_log_cmn(const WesepLogCmnArgs* a, void* stream);          /* out[n][t][m] = log(mel[n][m][t] + eps) - its mean over t */

/* ------------------------------------------------------------------------------------------------
 * wespeaker ECAPA-TDNN building blocks (speaker encoder `ECAPA_TDNN_GLOB_c512` of dpccn.yaml:59-64 /
 * bsrnn_feats.yaml; wespeaker/models/ecapa_tdnn.py + pooling_layers.ASTP, an external package: parity unpinned).
 * ---------------------------------------------------------------------------------------------- */
/* patches of nn.Conv1d(C, Co, K, dilation = dil, padding = dil (K - 1) / 2): col[n][c*K + k][t] = x[n][c][t + (k - (K-1)/2) dil] (0 outside);
 * the convolution is wesep_b200_conv1x1 over the K C gathered rows; bwd = the adjoint gather. */
typedef struct {
  int n, C, T, K, dil; int64_t ldx, ldc, bsc;
  const float* x; float* col;
  const float* gcol; float* gx;
  int unfold, stride;    /* unfold = 1: F.unfold(x[..., None], (K, 1), stride = (stride, 1)) — no padding, dilation 1, (T - K) / stride + 1
                            columns (TF-GridNet with emb_ks > 1, gridnet_block.py:147-160; its adjoint is ConvTranspose1d's overlap-add) */
} WesepIm2col1dArgs;
int wesep_b200_im2col1d_fwd(const WesepIm2col1dArgs* a, void* stream);
int wesep_b200_im2col1d_bwd(const WesepIm2col1dArgs* a, void* stream);

/* elementwise y = f(x) over `count` contiguous floats: mode 0 ReLU, 1 sigmoid (the SE block, ecapa_tdnn.py SE_Connect); bwd from y */
typedef struct { int64_t count; int mode; const float* x; float* y; const float* gy; float* gx; } WesepUnaryArgs;
int wesep_b200_unary_fwd(const WesepUnaryArgs* a, void* stream);
int wesep_b200_unary_bwd(const WesepUnaryArgs* a, void* stream);     /* gx = gy * f'(.) expressed through y */

/* attentive statistics (pooling_layers.ASTP.forward tail): alpha [n][C][ld] = softmax over time (given), x [n][C][ld] ->
 * out [n][2C] = (sum_t alpha x | sqrt(clamp(sum_t alpha x^2 - mean^2, 1e-10))); bwd: gx and galpha (both [n][C][ld]). */
typedef struct {
  int n, C, T; int64_t ld;
  const float* x; const float* alpha; float* out;
  const float* gout; float* gx; float* galpha;
} WesepAstpArgs;
int wesep_b200_astp_fwd(const WesepAstpArgs* a, void* stream);
int wesep_b200_astp_bwd(const WesepAstpArgs* a, void* stream);

/* Plain elementwise helpers so that residual sums and the iSTFT envelope normalisation stay inside the library:
 * out = a + b over `count` floats (16-byte aligned buffers; may alias), and y[r][t] = x[r][t] * v[t] (the 1 / sum w^2 envelope of
 * torch.istft, wesep/models/bsrnn.py:382-389; its adjoint is the same call on the gradient). */
typedef struct { int64_t count; const float* a; const float* b; float* out; } WesepAddArgs;
int wesep_b200_add(const WesepAddArgs* a, void* stream);
typedef struct { int64_t rows; int L; const float* x; int64_t ldx; const float* v; float* y; int64_t ldy; } WesepColVecArgs;
int wesep_b200_colvec_mul(const WesepColVecArgs* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WESEP_B200_H_ */
