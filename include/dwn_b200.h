/* dwn_b200.h — C ABI of libdwn_b200.so: hand-written sm_100a kernels for the DwiseNeuro hot path.
 *
 * The reference (lRomul/sensorium) has no FFI: every op below replaces a torch-eager library call made
 * from /root/reference/src/models/dwiseneuro.py, src/losses.py, src/ema.py, src/argus_models.py or
 * src/predictors.py (file:line cited per entry point).  The Python side (sensorium_b200/) binds these
 * with ctypes and passes raw device pointers + the current CUDA stream.
 *
 * Contract: every function returns 0 on success, negative on error (text via dwn_last_error()); nothing
 * throws, allocates or frees; all launches are asynchronous on `stream`; all pointers are device
 * pointers owned by the caller unless stated otherwise.  dtype codes: 0 = fp32, 1 = bf16.
 * Activations are channels-last: [B][T][H][W][C] == row-major [M][C].
 * BN coefficient tables: coef[4][C] = {scale, shift, mean, rstd}; bcoef[2][C] = {sum(dy)/N, sum(dy*xhat)/N}.
 */
#ifndef DWN_B200_H
#define DWN_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* dwn_last_error(void);
int dwn_abi_version(void);
int dwn_sm_count(void);
int dwn_set_sm_budget(int n);   /* SMs the persistent kernels size their grids for (0 = all): leaves room for NCCL */

/* ---- GEMM: D[z] (MxN) = A[z] (MxK) * B[z]^T (NxK) -------------------------------------------------
 * replaces nn.Conv3d 1x1x1 / nn.Conv1d k=1 (grouped) forward, dgrad and wgrad
 * (dwiseneuro.py:90-93 conv_pw, :117-120 conv_pwl, :206 ShuffleLayer.conv, :277 Readout conv). */
typedef struct {
  int dtype;                 /* operand type: 0 fp32 (SIMT FFMA), 1 bf16 (tcgen05 + TMA, fp32 accumulate) */
  const void* A;
  const void* B;
  int a_mn, b_mn;            /* 0: K-major (row-major [rows][K]); 1: MN-major (row-major [K][rows]) */
  long lda, ldb;             /* row pitch in elements */
  long a_zstride, b_zstride; /* element offset between z slices */
  int a_zmode, b_zmode;      /* 0: operand shared by all z; 1: slice z; 2 (B only): slice (m_blk*128)/b_batch_rows */
  int b_batch_rows;
  int M, N, K, Z;            /* per-slice problem */
  int epi;                   /* 0: row-major store; 1: readout epilogue (bias + softplus -> out[b][n][t]) */
  void* D;
  int d_dtype;               /* 0 fp32, 1 bf16 (epi 0) */
  long ldd, d_zstride;
  int m_limit, n_limit;      /* valid rows / cols (0 = M / N) */
  const float* bias;         /* epi 1: [n_out_total rounded up to groups] */
  float beta;                /* epi 1: softplus beta */
  int Tn;                    /* epi 1: frames per sample (columns are (b,t)) */
  int n_out_total;           /* epi 1: neurons of this readout */
  int row_offset_per_z;      /* epi 1: rows per group (ceil(n/groups)) */
  int block_n;               /* 0 = auto */
  const void* A2;            /* optional second operand pair (bf16 only, same majors / z-modes): D += A2 (MxK2) * B2^T */
  const void* B2;
  long lda2, ldb2;
  int K2;
  unsigned dbg_lbo_a, dbg_sbo_a, dbg_lbo_b, dbg_sbo_b; /* debug override of MN-major descriptor strides */
  int split;                 /* 3: A and B each hold three bf16 planes (hi, mid, lo of an fp32 operand, dwn_split3) with
                                the layout described above, a_pstride / b_pstride elements apart; the six significant
                                plane products are accumulated in fp32 -> fp32-accurate product on the tensor cores */
  long a_pstride, b_pstride;
} dwn_gemm_desc;
int dwn_gemm(const dwn_gemm_desc* d, void* stream);
/* fp32 -> three bf16 planes: dst[p][i], p = 0..2, x = dst[0] + dst[1] + dst[2] up to 2^-25 |x| (n % 4 == 0) */
int dwn_split3(const float* src, void* dst, long n, void* stream);

/* ---- stem (dwiseneuro.py:306-309) + positional encoding (:147-192) --------------------------------- */
int dwn_input_moments(const float* x, int B, int cin, long plane, double* partial, int P, double* mom, void* stream);
int dwn_stem_coef(const double* mom, int cin, double count, const float* w, const float* gamma, const float* beta,
                  float* rmean, float* rvar, long long* nbt, float momentum, float eps, float* coef, int C, void* stream);
int dwn_stem_fwd(const float* x, const float* w, const float* coef, const float* pe_t, const float* pe_h,
                 const float* pe_w, float* out, void* out_bf, float* partial, int P, int next_stride, int B, int cin,
                 int Tn, int H, int W, int C0, void* stream);

/* ---- BatchNorm statistics (dwiseneuro.py:9-22) -------------------------------------------------------- */
int dwn_bn_finalize(const float* partial, int P, double count, const float* gamma, const float* beta, float* rmean,
                    float* rvar, long long* nbt, float momentum, float eps, int training, float* coef, int C, int Cp,
                    int NQ, void* stream);   /* partial[P][NQ][Cp], quantities 0/1 = sum / sumsq */
int dwn_colstats(const void* x, long M, int ld, int C, float* partial, int J, int dtype, void* stream);

/* ---- depth-wise convolutions fused with BN+SiLU-on-load (dwiseneuro.py:96-111) ------------------------ */
int dwn_sdw_fwd(const void* in, const float* coef, const float* wgt, void* out, float* partial, int P, int NP, int H,
                int W, int C, int stride, int dtype, void* stream);
int dwn_tdw_fwd(const void* in, const float* coef, const float* wgt, void* out, float* partial, int P, int B, int Tn,
                int HW, int C, int dtype, void* stream);

/* ---- squeeze-excite (dwiseneuro.py:25-43) ---------------------------------------------------------------- */
int dwn_se_pool(const void* in, const float* coef, void* act, float* partial, int J, int B, int Nsp, int C, int dtype,
                void* stream);   /* dtype 2: fp32 in, act = three bf16 planes [3][ceil8(B*Nsp*C)] (dwn_split3 layout) */
int dwn_se_mlp(const float* partial, int J, int Nsp, const float* w1, const float* b1, const float* w2, const float* b2,
               float* mean_out, float* hpre_out, float* gate_out, int B, int C, int RD, void* stream);
int dwn_fold_gate(const float* w, const float* gate, void* out, int B, int N, int K, int dtype, void* stream);

/* ---- residual epilogue: drop-path + nearest/cyclic shortcut + BN_sc + next PE (dwiseneuro.py:125-144) --- */
int dwn_block_out(const void* y_raw, const float* coef4, const float* dp, const float* xin, const float* coef_sc,
                  const float* pe_t, const float* pe_h, const float* pe_w, float* out, void* out_bf, float* partial,
                  int P, int next_stride, int B, int Tn, int Ho, int Wo, int Ci, int Co, int stride, int Hi, int Wi,
                  int dtype, void* stream);   /* Ho = ceil(Hi / stride): the nearest map floor(dst * Hi / Ho) of
                                                  F.interpolate (dwiseneuro.py:127-129) also for sizes the stride does not divide */
int dwn_pool_hw(const float* in, float* out, void* out_bf, int BT, int HW, int C, void* stream);

/* ---- cortex (dwiseneuro.py:195-234) and readout input (:276) ------------------------------------------ */
int dwn_cortex_out(const void* y, const float* coef, const float* dp, const float* xin, const float* coef_sc, float* out,
                   void* out_bf, int M, int Tn, int I, int O, int G, int dtype, void* stream);
int dwn_readout_prep(const float* x, const float* mask, void* xm, void* xt, int M, int K, int Tn, int dtype,
                     void* stream);
int dwn_cast_bf16(const float* in, void* out, long n, void* stream);

/* ==== backward ============================================================================================
 * The reference gets every gradient from torch autograd through the library kernels of dwiseneuro.py; the entry
 * points below are the hand-written backward of the same ops (formulas: SURVEY.md 7.4). */
int dwn_bn_bwd_finalize(const float* partial, int P, int NQ, int q0, double count, float* dgamma, float* dbeta,
                        float* bcoef, int C, void* stream);
/* two BatchNorms whose sums share one partial table (quantities q0a.. and q0b..): one launch */
int dwn_bn_bwd_finalize2(const float* partial, int P, int NQ, int q0a, int q0b, double count, float* dgammaA,
                         float* dbetaA, float* bcoefA, float* dgammaB, float* dbetaB, float* bcoefB, int C, void* stream);                                /* dwiseneuro.py:9-22 */
int dwn_block_bwd_reduce(const float* dO, const void* y_raw, const float* coef4, const float* dp, const float* xin,
                         const float* coef_sc, float* partial, int P, int B, int Tn, int Ho, int Wo, int Ci, int Co,
                         int stride, int Hi, int Wi, int dtype, void* stream);                              /* dwiseneuro.py:136-144 */
int dwn_block_bwd_dy(const float* dO, const void* y_raw, const float* coef4, const float* bcoef4, const float* dp,
                     void* dY, long Mo, long rows_per_b, int Co, int dtype, void* stream);
int dwn_block_in_bwd(const float* dXpw, const float* dO, const float* xin, const float* coef_sc, const float* bcoef_sc,
                     const float* colbias, float* dXin, int B, int Tn, int Hi, int Wi, int Ci, int Co, int stride,
                     void* stream);                                                        /* dwiseneuro.py:125-134 */
int dwn_block_in_bwd_stem(const float* dXpw, const float* dO, const float* xin, const float* coef_sc,
                          const float* bcoef_sc, const float* colbias, const float* x_in, float* stem_partial, int P, int B, int Tn,
                          int Hi, int Wi, int Ci, int Co, int stride, void* stream);
int dwn_block_in_bwd_stem_rows(int Ci);   /* P that fills exactly one resident wave (stem_partial has P rows) */  /* block 0 + stem reductions fused */
int dwn_stem_bwd_finalize(const float* partial, int P, const double* mom, const float* w, const float* coef, float* dw,
                          float* dgamma, float* dbeta, int B, int cin, long plane, int C0, void* stream);
int dwn_pool_bwd(const float* dP, float* dX, long BT, int HW, int C, void* stream);       /* dwiseneuro.py:374,400 */
int dwn_se_bwd(const float* Pp, const float* wt, const float* gate, const float* hpre, const float* mean,
               const float* w1, const float* w2, float* dpre2, float* dhpre, float* dmean, float* dwpwl, float* dw2,
               float* db2, float* dw1, float* db1, int B, int C, int Co, int RD, void* stream); /* dwiseneuro.py:25-43 */
int dwn_tdw_bwd_reduce(const void* da, const void* tm, const float* coef3, const float* dmean, int Nsp, float* partial,
                       int J, int B, int C, int dtype, void* stream); /* dwiseneuro.py:105-111; statistics only; partial: B*J rows */
int dwn_tdw_bwd(void* dth, const void* tm, const void* s_raw, const float* coef3, const float* bcoef3,
                const float* coef2, const float* wgt, const float* dmean, float* partial, int P, int B, int Tn, int HW,
                int C, int dtype, void* stream); /* dth holds da on entry (the SE/act backward is recomputed), d s_hat on exit */
int dwn_sdw_bwd(const void* dsh, const void* s_raw, const void* e_raw, const float* coef2, const float* bcoef2,
                const float* coef1, const float* wgt, void* dE, float* partial, int P, int NP, int H, int W, int C,
                int stride, int dtype, void* stream);                                      /* dwiseneuro.py:96-102 */
int dwn_bn_bwd_apply(void* g, const void* x, const float* coef, const float* bcoef, long M, int C, int dtype,
                     void* stream);
int dwn_reduce_rows(const float* partial, int Z, long n, float* out, void* stream);
int dwn_dw_wgrad_finalize(const float* partial, int P, int NQ, int q0, int KK, float* dw, int C, void* stream);
int dwn_stem_bwd(const float* dy, const float* x, float* partial, int P, const double* mom, const float* w,
                 const float* coef, float* dw, float* dgamma, float* dbeta, int B, int cin, long plane, int C0,
                 void* stream);                                                            /* dwiseneuro.py:306-309 */
int dwn_cortex_bwd_reduce(const float* dOut, const void* y, const float* coef, const float* dp, const float* xin,
                          const float* coef_sc, float* partial, int J, int M, int Tn, int I, int O, int G, int dtype,
                          void* stream);                                                   /* dwiseneuro.py:195-234 */
int dwn_cortex_bwd_dy(const float* dOut, const void* y, const float* coef, const float* bcoef, const float* dp, void* dY,
                      int M, int Tn, int O, int G, int dtype, void* stream);
int dwn_cortex_in_bwd(const float* dXc, const float* dOut, const float* xin, const float* coef_sc,
                      const float* bcoef_sc, float* dX, int M, int I, int O, void* stream);

/* ==== conv_pw algebra (dwiseneuro.py:90-93): BatchNorm statistics of E = X W^T from the Gram matrix of X, and the
 * BatchNorm backward folded into the dgrad / wgrad GEMMs (no pass over E) ================================== */
int dwn_partial_colsum(const float* partial, int P, int NQ, int q, int C, float* out, void* stream);
/* split-K partials part[Z][ci][ci] -> gram (raw, fp64-accumulated) and cgram = gram/count - mu mu^T (centred in fp64) */
int dwn_gram_finalize(const float* part, int Z, int ci, const float* sx, double count, float* gram, float* cgram,
                      void* stream);
int dwn_pw_stats(const float* cgram, const float* sx, const void* w_bf16, double count, const float* gamma,
                 const float* beta, float* rmean, float* rvar, long long* nbt, float momentum, float eps, float* coef,
                 int mid, int ci, void* stream);
int dwn_pw_bwd_prep(const float* coef, const float* bcoef, const void* w_bf16, void* wprime, void* negq, float* r,
                    float* scratch, int mid, int ci, void* stream);   /* scratch: dwn_pw_bwd_prep_scratch() floats */
int dwn_pw_bwd_prep_scratch(int mid, int ci);
int dwn_pw_wgrad_finalize(const float* Psum, const float* coef, const float* bcoef, const void* w_bf16, const float* gram,
                          const float* sx, float* dw, int mid, int ci, void* stream);

/* ==== loss (src/losses.py:5-21), readout backward prep (dwiseneuro.py:266-287) ============================== */
int dwn_poisson_fwd(const float* pred, const float* tgt, const float* wn, int wstride, int B, long per_b, float eps,
                    double* partial, int J, void* stream);
int dwn_poisson_bwd(const float* pred, const float* tgt, const float* wn, int wstride, const float* gout, int B,
                    long per_b, float eps, float* dpred, void* stream);
int dwn_readout_bwd_prep(const float* pred, const float* dpred, float beta, void* dz_nm, void* dz_mn, float* dbias,
                         int B, int Tn, int n_out, int half, int half_pad, int G, int dtype, void* stream);
int dwn_readout_dx_combine(const float* dxm, const float* masks, int nlive, float* dX, int M, int K, int Tn,
                           void* stream);

/* ==== distillation target fill (src/argus_models.py:31-41) ================================================== */
int dwn_distill_prepare(const float* w, int n, float ratio, void* mask, float* dweight, void* stream);
int dwn_distill_fill(float* tgt, const float* teacher, const void* mask, int nmice, int mouse, int B, long per_b,
                     void* stream);
int dwn_distill_weights(float* w, const void* mask, const float* dweight, int n, void* stream);

/* ==== sliding-window predictor (src/predictors.py:36-55, src/indexes.py:23-30) ============================== */
int dwn_window_gather(const float* inp, float* clips, int Cn, int L, long HW, int size, int step, int first_index,
                      int nwin, void* stream);
/* StackInputsProcessor (inputs.py:22-36) fused with the window gather of predict_trial
 * (predictors.py:42-51).  video: (Hv, Wv, L) row-major fp32 (video_dtype 0) or uint8 (2); behavior, pupil: (2, L) fp32;
 * clips: (nw, 5, size, H, W) fp32, window i ends at frame last0 + i. */
int dwn_assemble_clips(const void* video, int video_dtype, const float* behavior, const float* pupil, float* clips,
                       int L, int Hv, int Wv, int H, int W, float fill, int size, int step, int last0, int nw,
                       void* stream);
/* batch form: video (B, T, Hv, Wv) fp32 / uint8, scalars (B, 4, T) fp32 -> clips (B, 5, T, H, W) fp32 (inputs.py:22-36) */
int dwn_assemble_batch(const void* video, int video_dtype, const float* scalars, float* clips, int B, int T, int Hv,
                       int Wv, int H, int W, float fill, void* stream);
/* streaming CorrelationMetric (metrics.py:11-31, 49-74).  acc: (n, 5) doubles
 * {sum x, sum y, sum xy, sum x^2, sum y^2}, cnt: 1 double; samples with weights[b*wstride] == 0 are skipped. */
int dwn_corr_update(const float* pred, const float* target, const float* weights, int wstride, int B, int n, int T,
                    double* acc, double* cnt, void* stream);
int dwn_corr_finalize(const double* acc, const double* cnt, int n, double eps, float* out, float* out_mean,
                      void* stream);
/* CutMix (mixers.py:52-67) and batch collation (datasets.py:172-187) on the device.
 * x1, x2, out: (B, planes, H, W) fp32; boxes: (B, 4) int32 {bbx1, bby1, bbx2, bby2} exactly as rand_bbox returns them
 * (bbx runs over H, bby over W, like the reference's slicing); lam: (B) fp32. */
int dwn_cutmix(const float* x1, const float* x2, const int* boxes, float* out, int B, long planes, int H, int W,
               void* stream);
int dwn_lerp_rows(const float* t1, const float* t2, const float* lam, float* out, int B, long row, void* stream);
int dwn_scatter_mouse_targets(const float* compact, const int* ids, int m, float* out, int B, int n_m, int n_max, int T,
                              void* stream);
int dwn_window_blend(const float* pred, const float* blend, float* out, int n_out, int L, int size, int step, int win0,
                     int nwin, long pred_wstride, void* stream);

/* ==== optimizer (torch.optim.AdamW as configured in configs/true_batch_001.py:45-48) and EMA (src/ema.py:47-55)
 * tab: device array of 8 x int64 rows {p, g, m, v, shadow_bf16, ema, n, flags}; chunk tables map CTAs to
 * (tensor, offset) in units of dwn_opt_chunk() elements. */
int dwn_opt_chunk(void);
int dwn_adamw(const void* tab, const int* chunk_tensor, const long* chunk_off, int nchunks, int* steps,
              const int* active, int nt, float lr, float wd, float b1, float b2, float eps, float ema_decay,
              const float* lr_dev, float* bc_scratch, void* stream);
/* lr_dev != NULL: learning rate read from device memory (CUDA graphs); bc_scratch: 2*nt floats (bias corrections) */
int dwn_ema(const void* tab, const int* chunk_tensor, const long* chunk_off, int nchunks, float decay, void* stream);
int dwn_scale(float* x, long n, float s, void* stream);

/* ==== data-parallel gradient exchange behind the C ABI (the reference has no distributed path at all:
 * scripts/train.py:173-189 trains one fold on one GPU).  NCCL is bound at run time (dlopen of the libnccl.so.2 the
 * process already holds, e.g. torch's), so the library loads on machines without NCCL and these calls then fail loudly.
 * One communicator per process, one process per GPU; the Python host of this repo keeps using torch.distributed's
 * communicator (sensorium_b200/parallel.py) - these entry points are for hosts without torch.distributed.
 *   dwn_comm_unique_id: 128-byte rendezvous token created on rank 0 and handed to the other ranks by the host;
 *   dwn_allreduce_bucket: in place, asynchronous on comm_stream; dtype 0 = fp32, 1 = bf16, 2 = int32; avg != 0 -> mean
 *   over ranks (DDP semantics), op_max != 0 -> maximum (per-mouse has-grad flags); several buckets between
 *   dwn_comm_group_begin / dwn_comm_group_end are issued as one NCCL group. */
int dwn_comm_unique_id(void* out128);
int dwn_comm_init(int rank, int nranks, const void* unique_id128);
int dwn_allreduce_bucket(void* ptr, long count, int dtype, int avg, int op_max, void* comm_stream);
int dwn_comm_group_begin(void);
int dwn_comm_group_end(void);
int dwn_comm_destroy(void);

#ifdef __cplusplus
}
#endif
#endif
