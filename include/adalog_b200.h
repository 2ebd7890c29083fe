/*
 * adalog_b200 -- C ABI of the B200 (sm_100a) kernels behind AdaLog's FPCS calibration sweep and
 * fake-quant forward.
 *
 * The reference (GoatWu/AdaLog) has no FFI: its hot path is eager PyTorch inside
 * quantizers/*.py and quant_layers/*.py.  Each entry point below replaces one eager-op chain of the
 * reference; the citation after "replaces:" is the reference file:line whose arithmetic it computes.
 * The host side that binds these (ctypes) is adalog_b200/_lib.py; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless named h_*; no allocation, no ownership transfer, no
 *    global mutable state; workspaces are passed in by the caller.
 *  - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream); all work is enqueued
 *    on it and the call returns without synchronising.
 *  - return 0 on success, negative on error; adalog_last_error() returns a thread-local message.
 *  - P (candidates per sweep) is at most 128 = one TMEM lane per candidate; callers pad.
 *  - zero points are passed as float (the reference promotes int64 zp to float32 in
 *    `(x/s).round_() + zp`, linear.py:304) and, for the quantizer forward, already passed through
 *    round_ste (quantizers/_ste.py:5-6).
 */
#ifndef ADALOG_B200_H_
#define ADALOG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADALOG_P 128          /* candidate rows of one unit tile */
#define ADALOG_BK 64          /* bf16 elements per 128-byte swizzled K block (128 for int8 operands) */
#define ADALOG_BF16 0         /* operand dtype: bf16 (any quantizer; exact for |int| <= 256 and m*2^-e) */
#define ADALOG_I8 1           /* operand dtype: int8 (uniform quantizers up to 7 bits; tcgen05 kind::i8, S32 accumulate) */
#define ADALOG_ORDER_UNIT_FAST 0   /* CTA order: neighbours differ in unit list (share fixed-operand tiles through L2) */
#define ADALOG_ORDER_SPLIT_FAST 1  /* CTA order: neighbours differ in N split (share one unit's candidate rows through L2) */

int adalog_version(void);
const char* adalog_last_error(void);

/* ---------------------------------------------------------------- quantizer forwards (K1/K2) */

/* replaces: quantizers/uniform.py:25-36 (UniformQuantizer.forward, inference branch).
 * group(i) = (i / inner) % ngroups selects scale[g], zp[g].  y and codes are optional (NULL).
 * asym: c = clamp(rint(x/s) + zp, 0, 2n-1), y = (c - zp) * s ; sym: c = clamp(rint(x/s), -n, n-1), y = c*s */
int adalog_uniform_fakequant_f32(const float* x, float* y, int16_t* codes, int64_t n, const float* scale,
                                 const float* zp, int64_t inner, int64_t ngroups, int n_levels, int symmetric,
                                 void* stream);

/* replaces: quantizers/logarithm.py:25-35 (kind 0, Log2), :45-62 (kind 1, LogSqrt2), :83-99 (kind 2, AdaLog
 * with table1/table2 LUTs and base q) and the Shift* wrappers :105-135.
 * v = clamp((x + shift)/s, 1e-15, 1); c = rint(-log2(v) [*2 | *37/q]); y = dequant(c) * s * (c < 2n) [- shift].
 * scale: [1]; q: int64 [1] (kind 2); shift: [1] or NULL; sub_shift: subtract the shift from the result. */
int adalog_log_fakequant_f32(const float* x, float* y, int16_t* codes, int64_t n, const float* scale, int kind,
                             int n_levels, const long long* q, const float* table1, const float* table2,
                             const float* shift, int sub_shift, void* stream);

/* replaces: quantizers/uniform.py:57-68 (TwinUniformQuantizer.forward); scale2 = {positive, negative}. */
int adalog_twin_fakequant_f32(const float* x, float* y, int64_t n, const float* scale2, int n_levels, void* stream);

/* ---------------------------------------------------------------- self-error sweeps (K3/K4) */

/* replaces: quant_layers/linear.py:296-309 (_search_best_w_scale_self, error part).
 * W [R,K]; candidates cs/cz laid out [P,R]; err_sum[p*R + r] = sum_k (w - dequant_p(w))^2 in FP64. */
int adalog_sweep_err_w_self(const float* W, int R, int K, const float* cs, const float* cz, int P, int n_levels,
                            double* err_sum, void* stream);

/* replaces: quant_layers/linear.py:320-345 (_search_best_a_scale_self, error part).
 * x is n_total floats viewed as rows of C columns.  per_channel: candidates [C,P] (column c uses row c);
 * else candidates [1,P] shared.  partial is [nsplit, Cw, P] FP64 sums of (x - dequant_p(x))^2 with
 * Cw = C (per_channel) or 32 (per tensor: the flat array is swept 32 lanes wide); caller sums splits. */
int adalog_sweep_err_a_self(const float* x, int64_t n_total, int C, int per_channel, const float* cs,
                            const float* cz, int P, int n_levels, double* partial, int nsplit, void* stream);

/* ---------------------------------------------------------------- operand generators for the candidate GEMM
 * All generators read FP32 rows with unit K stride and write operand rows of pitch `kpad` ELEMENTS (a multiple of
 * 64 for bf16, of 128 for int8, i.e. whole 128-byte swizzle rows; zero filled beyond K).  Values are the INTEGER
 * part of the fake-quantised tensor (code - zp, or m*2^-e for the log family) which the operand type holds
 * exactly; scales are applied in the GEMM epilogue.  `dtype` is ADALOG_BF16 or ADALOG_I8 (uniform only). */

/* fixed operand, uniform quantizer (the non-searched side: linear.py:373 quant_input / :406 quant_weight_bias,
 * matmul.py:141,179).  group(r) = (r / g_div) % g_mod.  rowsum (optional) = sum_k value. */
int adalog_gen_uniform_fixed(const float* x, int64_t R, int K, int64_t ldx, const float* scale, const float* zp,
                             int64_t g_div, int64_t g_mod, int n_levels, void* out, int kpad, float* rowsum,
                             int dtype, void* stream);

/* candidate operand, uniform quantizer (linear.py:369-370, :409-410, matmul.py:150-151, :188-189, conv.py:237-238).
 * out row (u*128 + p); candidate p of unit u uses cs/cz[p*pstride + ((u_base+u)/g_div % g_mod)*gstride].
 * krep in {1,3}: the K block is repeated krep times along the row (pitch krep*kpad) to pair with a split-3 operand. */
int adalog_gen_uniform_cand(const float* x, int64_t U, int K, int64_t ldx, const float* cs, const float* cz, int P,
                            int64_t pstride, int64_t gstride, int64_t g_div, int64_t g_mod, int64_t u_base,
                            int n_levels, void* out, int kpad, int krep, float* rowsum, int dtype, void* stream);

/* candidate operand, AdaLog search form (linear.py:872-878, :913-919; matmul.py:337-342).
 * per candidate p: scale cs[p] (NULL: unscaled & unclamped, the post-softmax form) and base cq[p] (int64);
 * value = mtab[(c*q) % 37] * 2^-floor(c*q/37), 0 where c >= 2n.  mtab: 37 floats holding integers. */
int adalog_gen_log_cand(const float* x, int64_t U, int K, int64_t ldx, const float* cs, const long long* cq, int P,
                        const float* shift, const float* mtab, int n_levels, uint16_t* out, int kpad, void* stream);

/* fixed operand, AdaLog inference form (logarithm.py:87-99 via matmul.py:179 / linear.py:373):
 * value = m2[c] * 2^-table1[c] with m2 = table2*(4n-2) (integers), 0 where masked. */
int adalog_gen_log_fixed(const float* x, int64_t R, int K, int64_t ldx, const float* scale, const long long* q,
                         const float* shift, const float* table1, const float* m2, int n_levels, uint16_t* out,
                         int kpad, void* stream);

/* fixed operand, unquantised FP32 (conv.py:55-58 with a_bit >= 8): x = h + m + l, three bf16 pieces laid out
 * [R, 3*kpad] as [l | m | h] (smallest first: the tensor core's FP32 accumulator truncates late small addends) so
 * that a krep=3 candidate operand reproduces the FP32 product to 2^-24. */
int adalog_gen_split3(const float* x, int64_t R, int K, int64_t ldx, uint16_t* out, int kpad, void* stream);

/* ---------------------------------------------------------------- exact order statistics by radix selection (K11)
 * replaces: the full sorts behind torch.quantile / Tensor.sort of the candidate seeding -- linear.py:432-481
 * (calculate_percentile_*_candidates), :763-814 (positive_percentile), matmul.py:211-240.
 *
 * x: `rows` rows of n contiguous FP32 elements (row r at x + r*row_stride); ranks[T] (T <= 8, 0-based, the same for
 * every row); out[rows*T] = the ranks[t]-th smallest element of each row, bit for bit what sort(row)[ranks[t]] holds
 * (NaN last; a selected zero comes back as +0.0).  positive_only: elements <= 0 count as +inf (linear.py:763-798).
 * Four passes of 8 bits: adalog_select_hist(pass) accumulates the histogram of the next digit, adalog_select_scan(pass)
 * descends.  A data-parallel caller whose ranks each hold local_rows of the `rows` (row0 = its offset; or the same rows,
 * sharded along n, with row0 = 0) all-reduces (SUM) the uint32 histograms at adalog_select_hist_ptr between the two
 * calls: every rank then selects from the union of the shards.  The workspace (adalog_select_workspace_bytes) is
 * caller-owned; nothing is allocated inside. */
int64_t adalog_select_workspace_bytes(int64_t rows, int T);
void* adalog_select_hist_ptr(void* workspace, int64_t rows, int T);
int adalog_select_init(void* workspace, int64_t rows, int T, const long long* ranks, void* stream);
int adalog_select_hist(const float* x, int64_t local_rows, int64_t n, int64_t row_stride, void* workspace, int64_t rows,
                       int64_t row0, int T, int pass, int positive_only, void* stream);
int adalog_select_scan(void* workspace, int64_t rows, int T, int pass, void* stream);
int adalog_select_finish(const void* workspace, int64_t rows, int T, float* out, void* stream);

/* ---------------------------------------------------------------- candidate-batched GEMM + fused error (K5-K10)
 * replaces: the F.linear / @ / F.conv2d + _get_similarity + mean/sum chains of linear.py:355-384, :394-423,
 * :856-890, :898-931, matmul.py:135-163, :173-201, :321-351, conv.py:226-256.
 *
 * A  [U*128, ka] : candidate operand, one 128-row tile per unit (row = u*128 + p); ka = KB*64 (bf16) / KB*128 (int8).
 * Bm [G*brpg, ka]: fixed operand, group g = (g_base + u/UG) owns rows [g*brpg, g*brpg + N).
 * D[p, n] = sum_k A[u*128+p, k] * Bm[g*brpg + n, k]   (tcgen05.mma kind::f16 -> FP32, or kind::i8 -> S32, in TMEM)
 * yhat    = rs[ri+p] * (cs ? cs[n]*D : D) + (rb ? rb[ri+p] : 0),  ri = ((u_base+u)/rs_div % rs_mod)*128
 * e[p]   += (y[u*ldy + n] - (cb ? cb[n] : 0) - yhat)^2      summed over the CTA's units and N tiles
 * partial[(y*gridX + x)*128 + p] = e[p] (FP64).  Logical grid = (gridX = G_chunk * cpg, S), launched 1-D in `order`:
 * CTA x handles units [g*UG + ci*UG/cpg, g*UG + (ci+1)*UG/cpg) of group g = x / cpg, ci = x % cpg, cpg = ceil(UG/upc);
 * CTA y handles N tiles
 * [y*NT/S, (y+1)*NT/S).  The partition is static, so equal candidates produce bit-equal sums. */
typedef struct {
  const void* A;      const void* Bm;
  int64_t a_rows;     int64_t b_rows;      /* allocated rows of A / Bm (TMA bounds) */
  int32_t KB;         /* K blocks of 64 */
  int32_t N;          /* valid columns per group */
  int32_t BN;         /* tile width, multiple of 16, <= 256 */
  int32_t U;          /* units in this launch */
  int32_t UG;         /* units per group */
  int32_t upc;        /* units per CTA (upper bound: a group is dealt evenly to ceil(UG/upc) CTAs) */
  int32_t S;          /* N-tile splits */
  int32_t dtype;      /* ADALOG_BF16 | ADALOG_I8 */
  int32_t order;      /* ADALOG_ORDER_*: which CTAs are co-resident; does not change any result bit */
  int32_t reserved;   /* 0 */
  int64_t brpg;       /* Bm rows per group */
  int64_t g_base;     int64_t u_base;
  const float* y;     int64_t ldy;
  const float* rs;    const float* rb;  int64_t rs_div; int64_t rs_mod;
  const float* cs;    const float* cb;
  double* partial;    /* [S * gridX * 128] */
} adalog_gemm_err_args;

/* returns gridDim.x for the given args (so the caller can size `partial`), or negative on error */
int adalog_cand_gemm_err_grid(const adalog_gemm_err_args* a);
int adalog_cand_gemm_err(const adalog_gemm_err_args* a, void* stream);

/* ---------------------------------------------------------------- fused generator + GEMM + error (attention matmuls)
 * replaces: quant_layers/matmul.py:135-163 (_search_best_A_scale), :173-201 (_search_best_B_scale), :321-351
 * (post-softmax AdaLog base search).  Same arithmetic and same result layout as adalog_gen_*_cand followed by
 * adalog_cand_gemm_err, for the shapes of the attention products (one N tile, K <= 256 bf16 / 512 int8), but the
 * 128x-expanded candidate operand is generated inside the GEMM kernel, straight into shared memory, and never
 * touches HBM; the group's fixed operand is loaded once per CTA.
 *
 * x  [U, K] FP32 (row pitch ldx): source rows of the candidate side, unit u = row u.
 * Bm [G*brpg, KB*64|128]: fixed operand as written by adalog_gen_uniform_fixed / adalog_gen_log_fixed.
 * gen = ADALOG_GEN_UNIFORM: candidate p of unit u uses cs/cz[p*pstride + ((u_base+u)/g_div % g_mod)*gstride];
 * gen = ADALOG_GEN_LOG: unscaled AdaLog code rint(-log2(x)*37/cq[p]) (the post-softmax form), value from mtab.
 * yhat = rs[ri*128+p] * D, ri = (u_base+u)/rs_div % rs_mod;  partial[x*128 + p] (FP64), x < grid.
 * g_div, rs_div and u_base must be multiples of UG (a CTA works inside one group). */
#define ADALOG_GEN_UNIFORM 0
#define ADALOG_GEN_LOG 1
typedef struct {
  const float* x;     int64_t ldx;
  const void* Bm;     int64_t b_rows;
  int32_t K;          /* true reduction length */
  int32_t KB;         /* ceil(K / 64) (bf16) or ceil(K / 128) (int8), <= 4 */
  int32_t N;          /* valid columns per group, <= BN */
  int32_t BN;         /* tile width, multiple of 16, <= 256 */
  int32_t U;          int32_t UG;          int32_t upc;
  int32_t P;          int32_t n_levels;    /* of the candidate-side quantizer */
  int32_t gen;        int32_t dtype;
  int32_t epi_warps;  /* 4 or 8 of the kernel's 14 worker warps reduce the error, the others generate candidates */
  int64_t brpg;       int64_t g_base;      int64_t u_base;
  const float* cs;    const float* cz;     int64_t pstride; int64_t gstride; int64_t g_div; int64_t g_mod;
  const long long* cq; const float* mtab;
  const float* y;     int64_t ldy;
  const float* rs;    int64_t rs_div;      int64_t rs_mod;
  double* partial;    /* [grid * 128] */
} adalog_fused_args;

/* returns the grid size (so the caller can size `partial`), or negative (e.g. -3: does not fit in shared memory) */
int adalog_fused_cand_gemm_err_grid(const adalog_fused_args* a);
int adalog_fused_cand_gemm_err(const adalog_fused_args* a, void* stream);

/* ---------------------------------------------------------------- fused generator + GEMM + error (linear activation sweeps)
 * replaces: quant_layers/linear.py:394-423 (_search_best_a_scale), :816-854, :856-890, :898-931 (post-GELU AdaLog
 * scale / base / scale x base searches).  Same arithmetic as adalog_gen_uniform_cand / adalog_gen_log_cand followed by
 * adalog_cand_gemm_err with column scales, but ONE launch scores all units: persistent CTAs generate each unit's
 * 128-candidate K-block tiles in shared memory (never in HBM) and multiply them against the whole fixed operand.
 *
 * x  [U, K] FP32 (row pitch ldx): the layer input, unit u = token u; candidates are per-tensor:
 *   gen = ADALOG_GEN_UNIFORM: cs[p], cz[p] (scale, zero point);
 *   gen = ADALOG_GEN_LOG: cs[p] (scale), cq[p] (base), shift[0] (post-GELU shift or NULL), mtab[37] search-LUT numerators.
 * Bm [b_rows >= N, KB*64|128]: integer part of the quantised weight (adalog_gen_uniform_fixed); its K padding MUST be
 *   zero: the candidate side is only kept finite there, not zeroed.
 * yhat[p, n] = rs[p] * ccs[n] * D[p, n];  e[p] += (y[u*ldy + n] - ccb[n] - yhat)^2;  partial[x*128 + p] (FP64), x < grid.
 * N must be a multiple of 4, ccs 16-byte aligned.  Returns -3 when no schedule fits in shared memory (the caller then
 * takes the generator -> workspace -> adalog_cand_gemm_err path). */
typedef struct {
  const float* x;     int64_t ldx;
  const void* Bm;     int64_t b_rows;
  int32_t K;          int32_t N;           int32_t U;
  int32_t P;          int32_t n_levels;    /* of the candidate-side (activation) quantizer */
  int32_t gen;        int32_t dtype;       int32_t reserved;
  const float* cs;    const float* cz;
  const long long* cq; const float* shift; const float* mtab;
  const float* y;     int64_t ldy;
  const float* rs;    const float* ccs;    const float* ccb;
  double* partial;    /* [grid * 128] */
} adalog_lin_fused_args;

int adalog_lin_fused_cand_gemm_err_grid(const adalog_lin_fused_args* a);
/* how often the schedule for this shape generates each unit's candidate operand: 1 when all K blocks of a unit stay in
 * shared memory or its N columns fit in one 512-column TMEM pass, ceil(N/512) otherwise (the caller may then prefer
 * the generator -> workspace -> adalog_cand_gemm_err path); negative on error */
int adalog_lin_fused_cand_gemm_err_passes(const adalog_lin_fused_args* a);
int adalog_lin_fused_cand_gemm_err(const adalog_lin_fused_args* a, void* stream);

/* ---------------------------------------------------------------- fake-quant INFERENCE forward of a linear layer
 * replaces: F.linear(Q_a(x), Q_w(W), b) of quant_layers/linear.py:46-51 / :90-92 (quant_forward).
 * Same kernel, pipeline and argument block as adalog_cand_gemm_err, but the 128 rows of a unit are 128 consecutive
 * TOKENS (A [m_rows, ka] = integer part of Q_a(x), a_rows = m_rows, U = UG = ceil(m_rows/128)), Bm [N, ka] the integer
 * part of Q_w(W), and the epilogue writes  out[m, n] = rs[p] * cs[n] * D[m, n] + cb[n]  (rs = activation scale,
 * cs = per-row weight scales, cb = bias with any shift term folded in) instead of accumulating an error.
 * y, rb, partial are ignored. */
int adalog_gemm_dequant(const adalog_gemm_err_args* a, float* out, int64_t ldo, int64_t m_rows, void* stream);

/* plain (non-candidate) debug GEMM through the same tcgen05 pipeline: D[m,n] FP32 for A [128,ka], Bm [N,ka];
 * used by the tests to validate descriptors / swizzle / TMEM addressing in isolation.  Scratch comes from the caller
 * (zeros [N] floats holding 0, ones [128] floats holding 1, partial [64 * 128] doubles): the library owns no device state. */
int adalog_debug_gemm_tile(const void* A, const void* Bm, int KB, int N, float* D, const float* zeros, const float* ones,
                           double* partial, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADALOG_B200_H_ */
