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
  unsigned dbg_lbo_a, dbg_sbo_a, dbg_lbo_b, dbg_sbo_b; /* debug override of MN-major descriptor strides */
} dwn_gemm_desc;
int dwn_gemm(const dwn_gemm_desc* d, void* stream);

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
                    void* stream);
int dwn_colstats(const void* x, long M, int ld, int C, float* partial, int J, int dtype, void* stream);

/* ---- depth-wise convolutions fused with BN+SiLU-on-load (dwiseneuro.py:96-111) ------------------------ */
int dwn_sdw_fwd(const void* in, const float* coef, const float* wgt, void* out, float* partial, int P, int NP, int H,
                int W, int C, int stride, int dtype, void* stream);
int dwn_tdw_fwd(const void* in, const float* coef, const float* wgt, void* out, float* partial, int P, int B, int Tn,
                int HW, int C, int dtype, void* stream);

/* ---- squeeze-excite (dwiseneuro.py:25-43) ---------------------------------------------------------------- */
int dwn_se_pool(const void* in, const float* coef, void* act, float* partial, int J, int B, int Nsp, int C, int dtype,
                void* stream);
int dwn_se_mlp(const float* partial, int J, int Nsp, const float* w1, const float* b1, const float* w2, const float* b2,
               float* mean_out, float* hpre_out, float* gate_out, int B, int C, int RD, void* stream);
int dwn_fold_gate(const float* w, const float* gate, void* out, int B, int N, int K, int dtype, void* stream);

/* ---- residual epilogue: drop-path + nearest/cyclic shortcut + BN_sc + next PE (dwiseneuro.py:125-144) --- */
int dwn_block_out(const void* y_raw, const float* coef4, const float* dp, const float* xin, const float* coef_sc,
                  const float* pe_t, const float* pe_h, const float* pe_w, float* out, void* out_bf, float* partial,
                  int P, int next_stride, int B, int Tn, int Ho, int Wo, int Ci, int Co, int stride, int dtype,
                  void* stream);
int dwn_pool_hw(const float* in, float* out, void* out_bf, int BT, int HW, int C, void* stream);

/* ---- cortex (dwiseneuro.py:195-234) and readout input (:276) ------------------------------------------ */
int dwn_cortex_out(const void* y, const float* coef, const float* dp, const float* xin, const float* coef_sc, float* out,
                   void* out_bf, int M, int Tn, int I, int O, int G, int dtype, void* stream);
int dwn_readout_prep(const float* x, const float* mask, void* xm, void* xt, int M, int K, int Tn, int dtype,
                     void* stream);
int dwn_cast_bf16(const float* in, void* out, long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif
