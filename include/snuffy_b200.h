/* libsnuffy_b200 — C ABI of the B200-native Snuffy / DSMIL MIL-aggregator hot path.
 *
 * The reference (jafarinia/snuffy @ 4b5b918) has no FFI: its hot path is a set of nn.Module classes whose
 * forward() methods call PyTorch ATen ops.  Each entry point below replaces the ATen op
 * sequence at the cited reference lines; the drop-in nn.Modules in snuffy_b200/{snuffy,snuffy_multiclass,
 * dsmil}.py bind them through ctypes (snuffy_b200/_lib.py) — see INTEGRATION.md for the binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types, no exceptions, no exit().
 *   - every function returns 0 on success; on failure a non-zero status and snuffy_last_error() (per host
 *     thread) describes it.  Arguments are validated before any launch.
 *   - all pointers are DEVICE pointers to contiguous row-major fp32 unless stated; indices are int64;
 *     every function takes an explicit cudaStream_t, is asynchronous, re-entrant and never synchronises
 *     (a whole forward can be captured into a CUDA graph).
 *   - "planes" = split-bf16 operand planes (hi, lo) pre-tiled for tcgen05 (layout: csrc/common.cuh).
 */
#ifndef SNUFFY_B200_H
#define SNUFFY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* snuffy_stream_t;   /* == cudaStream_t */

/* ---- library ------------------------------------------------------------------------------------- */
int         snuffy_version(void);
const char* snuffy_last_error(void);
int         snuffy_sm_count(void);
long long   snuffy_launch_count(void);          /* kernels launched since load (bench.py: gpu_launches) */

/* ---- a1: instance scores  c = x W^T + b            (snuffy.py:39-41 FCLayer.forward, nn.Linear)      */
int snuffy_scores_fwd(const float* x, const float* W, const float* bias, float* c,
                      int64_t rows, int64_t d, int64_t C, snuffy_stream_t stream);

/* ---- a6/a12: patch selection                                                                         */
/* per (bag, class) the K best rows, descending score, ties -> lower index. Replaces torch.sort + slice
 * (snuffy.py:128-129, snuffy_multiclass.py:136-138, dsmil.py:78-81 with K = 1).
 * scores [B,N,C]; idx_out [B,C,K] int64; flags [B,N] uint8 (optional, caller-zeroed) marks winners.   */
int snuffy_select_topk(const float* scores, int64_t B, int64_t N, int64_t C, int64_t K,
                       int64_t* idx_out, uint8_t* flags, snuffy_stream_t stream);
/* K distinct un-flagged rows per bag, uniform without replacement (Philox4x32-10 keyed by seed/offset).
 * Replaces .tolist() + python set difference + np.random.choice + .to(device)  (snuffy.py:136-143,
 * snuffy_multiclass.py:152-157) without the D2H/H2D round trip.  idx_out [B,K] int64.  The caller guarantees
 * that every bag has at least K un-flagged rows (the reference's Krand = min(int(K r), N - Ktop) does).          */
int snuffy_select_random(const uint8_t* flags, int64_t B, int64_t N, int64_t K, uint64_t seed,
                         uint64_t offset, int64_t* idx_out, snuffy_stream_t stream);
/* The whole binary-path selection in one launch (snuffy.py:128-147 selection, 131 / 145-147 gathers, 152-155 row
 * replacement): per bag the k_top highest-scoring rows (descending, ties to the lower index), k_rand distinct rows drawn
 * from the rest (the Philox stream of snuffy_select_random: identical indices), sel[B, k_top + k_rand] = T ++ R,
 * flags[B, N] and row_map[B * N] (row -> slot or -1) written from scratch, xs[B, k_top + k_rand, d] = x[b, sel[b, k], :]. */
int snuffy_select_gather(const float* scores, const float* x, int64_t B, int64_t N, int64_t d, int64_t k_top,
                         int64_t k_rand, uint64_t seed, uint64_t offset, int64_t* sel, uint8_t* flags,
                         int32_t* row_map, float* xs, snuffy_stream_t stream);
/* ascending distinct flagged rows per bag = torch.unique of the flattened per-class top-K
 * (snuffy_multiclass.py:139-141).  out [B,cap] int64, counts [B] int32.                                */
int snuffy_compact_flags(const uint8_t* flags, int64_t B, int64_t N, int64_t cap, int64_t* out,
                         int32_t* counts, snuffy_stream_t stream);
/* out[B,K,d] = x[b, idx[b,k], :]      (torch.index_select / gather: snuffy.py:131,145-147,103-106)     */
int snuffy_gather_rows(const float* x, const int64_t* idx, int64_t B, int64_t N, int64_t K, int64_t d,
                       float* out, snuffy_stream_t stream);
/* row_map[B*N] int32 = -1 | slot b*K+k.  Replaces y = x.clone(); y[:, S, :] = x_sel (snuffy.py:152-155):
 * later kernels read "x with the selected rows replaced" through this map, x is never copied.          */
int snuffy_build_row_map(const int64_t* idx, int64_t B, int64_t N, int64_t K, int32_t* row_map,
                         snuffy_stream_t stream);

/* ---- LayerNorm (eps 1e-5, affine)                                                                    */
/* rows of y = (row_map ? x with mapped rows taken from alt : x) -> LN -> fp32 and/or planes and/or
 * (mean, rstd).  apply_ln = 0: plain convert/split.  Replaces nn.LayerNorm at snuffy.py:107,110 and
 * produces the tcgen05 A operand in the same pass.  plane_rc: 128 (activations) or
 * snuffy_gemm_tc_block_n(rows) (weights).                                                              */
int snuffy_ln_rows_fwd(const float* x, const int32_t* row_map, const float* alt, const float* gamma,
                       const float* beta, int64_t rows, int64_t d, int apply_ln, float* out_f32,
                       void* planes, int64_t plane_stride, int plane_rc, float* stats,
                       snuffy_stream_t stream);
/* Layer-0 fusion (inference): one pass over the bag gives the instance scores (snuffy.py:39-41) AND the
 * normalised operand planes shared by LN1 / LN2 (snuffy.py:107,110); scores are bit-identical to snuffy_scores_fwd. */
int snuffy_scores_ln_planes_fwd(const float* x, const float* W, const float* bias, int64_t rows, int64_t d,
                                int64_t C, float* c, void* planes, int64_t plane_stride, float* stats,
                                snuffy_stream_t stream);
/* Overwrite plane rows b*N + idx[b,k] with split(LN(src[b*K+k])): LN1 and LN2 (snuffy.py:107,110) share one set
 * of normalised planes, only the Ksel rows changed by the attention sub-layer (snuffy.py:152-155) are redone.
 * apply_ln everywhere: 0 = convert (optional per-column gain gamma), 1 = LN with affine, 2 = normalise only.     */
int snuffy_ln_rows_scatter_planes(const float* src, const int64_t* idx, int64_t B, int64_t N, int64_t K,
                                  int64_t d, const float* gamma, const float* beta, int apply_ln, void* planes,
                                  int64_t plane_stride, snuffy_stream_t stream);
/* bag[B,C] = head(mean_n LN_f(x[b,n,:]))   (Encoder.norm + BClassifier: snuffy.py:86 + 71) in one pass.
 * partials: B*chunks*d floats (chunks = snuffy_ln_mean_head_chunks); tickets: B uint32 zeroed once.     */
int64_t snuffy_ln_mean_head_chunks(int64_t B, int64_t N);
int snuffy_ln_mean_head_fwd(const float* x, const float* gamma, const float* beta, const float* Wh,
                            const float* bh, int64_t B, int64_t N, int64_t d, int64_t C,
                            float* partials, uint32_t* tickets, float* stats, float* pooled,
                            float* bag_out, snuffy_stream_t stream);

/* ---- GEMMs                                                                                            */
/* fp32 SIMT: C[M,N] = dropout(act(alpha * A.B^T + bias)) + resid(row_map).  a_kc/b_kc = 1: operand stored
 * [rows,K]; 0: stored [K,rows].  Replaces nn.Linear on [Ksel,d] rows (snuffy.py:188 key proj, 205 out proj),
 * dsmil's q/v MLPs (dsmil.py:56-66) and all backward contractions.                                     */
int snuffy_gemm_f32(const float* A, int64_t lda, int a_kc, const float* B, int64_t ldb, int b_kc,
                    float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha,
                    const float* bias, int act, const float* resid, int64_t ldr, const int32_t* row_map,
                    const float* resid_alt, float* preact, float dropout_p, uint64_t seed,
                    uint64_t offset, snuffy_stream_t stream);
/* tcgen05 split-bf16 (3-pass) GEMM on planes: the Q|V projection (snuffy.py:188) and the FFN
 * (snuffy.py:225) over all N tokens.  Outputs: fp32 out (+residual through row_map), fp32 preact,
 * and/or the activated result as A planes for the next GEMM.                                           */
int     snuffy_gemm_tc_block_n(int64_t N);
int64_t snuffy_plane_elems(int64_t rows, int64_t K, int rc);
int snuffy_gemm_tc(const void* A_planes, int64_t a_plane_stride, const void* B_planes,
                   int64_t b_plane_stride, int64_t M, int64_t N, int64_t K, int passes,
                   const float* bias, int act, const float* resid, int64_t ldr, const int32_t* row_map,
                   const float* resid_alt, float* out, int64_t ldc, float* preact, void* out_planes,
                   int64_t out_plane_stride, float dropout_p, uint64_t seed, uint64_t offset,
                   snuffy_stream_t stream);

/* Split-K form for weight gradients dW[M,N] = A B^T contracting over all patches (few output tiles, long K):
 * ksplit <= 0 = automatic; workspace: snuffy_gemm_tc_splitk_workspace(M, N, ksplit) bytes (deterministic fold).     */
int64_t snuffy_gemm_tc_auto_ksplit(int64_t M, int64_t N, int64_t K);
int64_t snuffy_gemm_tc_splitk_workspace(int64_t M, int64_t N, int64_t ksplit);
int snuffy_gemm_tc_splitk(const void* A_planes, int64_t a_plane_stride, const void* B_planes,
                          int64_t b_plane_stride, int64_t M, int64_t N, int64_t K, int passes, int64_t ksplit,
                          float* out, void* workspace, int64_t workspace_bytes, snuffy_stream_t stream);
/* All derived weight operands of a layer (or several) in one launch: what a training step re-derives from the updated
 * parameters before each forward (the reference re-reads nn.Linear.weight directly, snuffy.py:188, 225).  kind 0: planes of
 * src [rows, cols] (plane row = src row); kind 1: planes of src^T (plane row = src column, k = src row); kind 2: fp32 copy
 * to (float*)dst + dst_row0.  Several jobs may fill disjoint row tiles of one plane set (Wq and Wv -> the fused Q|V
 * operand): dst_row0 is the first plane row, a multiple of plane_rc, dst_k0 the first k (a multiple of 32) and k_total the
 * contraction length of the whole plane set; padding rows / k of the job's tiles are zeroed.                          */
#define SNUFFY_MAX_PLANE_JOBS 16
typedef struct {
    const float* src; int64_t ld, rows, cols;
    int32_t kind, plane_rc;
    int64_t dst_row0, dst_k0, k_total;
    void* dst; int64_t plane_stride;
} snuffy_plane_job_t;
int snuffy_plane_job_bytes(void);          /* sizeof(snuffy_plane_job_t) as the library was built: bindings check their struct */
int snuffy_weight_planes_batch(const snuffy_plane_job_t* jobs, int64_t n_jobs, snuffy_stream_t stream);
/* dW[M, N] = dY^T X straight from the ROW planes of dY [R, M] and X [R, N] (128 rows per chunk, as the GEMM epilogues and
 * snuffy_ln_rows_fwd write them) through MN-major tcgen05 descriptors: no transposed copy of either operand (autograd of
 * nn.Linear.weight at snuffy.py:188, 225).  M % 128 == 0, N % snuffy_gemm_tc_block_n(N) == 0; plane rows R .. ceil16(R)
 * must be zero when R % 16 != 0 (snuffy_planes_zero_rows).  Workspace: snuffy_gemm_tc_splitk_workspace(M, N, ksplit).   */
int snuffy_gemm_tc_splitk_rows(const void* dY_planes, int64_t a_plane_stride, const void* X_planes,
                               int64_t b_plane_stride, int64_t M, int64_t N, int64_t R, int passes, int64_t ksplit,
                               float* out, void* workspace, int64_t workspace_bytes, snuffy_stream_t stream);
int snuffy_planes_zero_rows(void* planes, int64_t plane_stride, int64_t K, int plane_rc, int64_t row0, int64_t row1,
                            snuffy_stream_t stream);
/* Operand planes of X^T for fp32 X [R, C] (plane row = column of X, k = row of X) with an optional prologue:
 * mode 0 plain, 1 LayerNorm from saved (mean, rstd) through row_map, 2 dropout(act(x)).  Feeds the transposed
 * operands of dW = dY^T X (autograd of nn.Linear at snuffy.py:188, 225) to snuffy_gemm_tc_splitk.                   */
int snuffy_planes_t_fwd(const float* x, int64_t ldx, int64_t R, int64_t C, int plane_rc, int mode,
                        const float* stats, const float* gamma, const float* beta, const int32_t* row_map,
                        const float* alt, int act, float dropout_p, uint64_t seed, uint64_t offset, void* planes,
                        int64_t plane_stride, snuffy_stream_t stream);

/* A . B^T where A is a 32-aligned K window of a wider A-plane set (e.g. the Q or the V half of the Q|V planes).         */
/* dX product with the activation backward fused into its epilogue (autograd of snuffy.py:225):
 * result[m, n] = (A . B^T)[m, n] * act'(gate[m, n]) * dropout_mask(m * N + n), as fp32 `out` and / or A-operand planes
 * and / or its column sums `colsum` [N] (the bias gradient; `colsum_partials`: ceil(M / 128) * 4 * N floats).          */
int snuffy_gemm_tc_actgrad(const void* A_planes, int64_t a_plane_stride, const void* B_planes,
                           int64_t b_plane_stride, int64_t M, int64_t N, int64_t K, int passes, const float* gate,
                           int64_t ldg, int gate_act, float dropout_p, uint64_t seed, uint64_t offset, float* out,
                           int64_t ldc, void* out_planes, int64_t out_plane_stride, float* colsum,
                           float* colsum_partials, snuffy_stream_t stream);
/* The same for ReLU, gated by the forward's own activated planes (snuffy_gemm_tc's out_planes of FFN-up, hi plane):
 * result = (A . B^T) * (a[m, n] > 0 ? 1 / (1 - dropout_p) : 0): neither the fp32 pre-activation nor the dropout draw is
 * needed again (a clamped or dropped element is exactly 0 in the planes).  Autograd of snuffy.py:225 with --activation relu. */
int snuffy_gemm_tc_relugrad(const void* A_planes, int64_t a_plane_stride, const void* B_planes,
                            int64_t b_plane_stride, int64_t M, int64_t N, int64_t K, int passes,
                            const void* act_planes, float dropout_p, float* out, int64_t ldc, void* out_planes,
                            int64_t out_plane_stride, float* colsum, float* colsum_partials, snuffy_stream_t stream);
/* out = A_window . B^T against a block-diagonal B (B[n, k] != 0 only where n / group_n == k / group_k: the head-block
 * operands of the attention backward, autograd of snuffy.py:160-168 with the head split of 187-201): every column tile
 * contracts only over the k-blocks of the groups it touches.
 * b_rc = rows per chunk of the B planes (128 or 256).                                                             */
int snuffy_gemm_tc_blockdiag(const void* A_planes, int64_t a_plane_stride, int64_t a_cols_total, int64_t a_col0,
                             const void* B_planes, int64_t b_plane_stride, int b_rc, int64_t M, int64_t N,
                             int64_t K, int passes, int64_t group_n, int64_t group_k, float* out, int64_t ldc,
                             snuffy_stream_t stream);
/* Split-K product of which only the diagonal blocks (row / diag_m == col / diag_n) are wanted (dKp = diagonal blocks of
 * dS^T Q): tiles that meet no such block are skipped, their part of `out` is unspecified.  ksplit <= 0: automatic
 * (snuffy_gemm_tc_diag_ksplit); workspace: snuffy_gemm_tc_splitk_workspace(M, N, ksplit) bytes.                     */
int64_t snuffy_gemm_tc_diag_ksplit(int64_t M, int64_t N, int64_t K, int b_rc, int64_t diag_m, int64_t diag_n);
int snuffy_gemm_tc_splitk_blockdiag(const void* A_planes, int64_t a_plane_stride, const void* B_planes,
                                    int64_t b_plane_stride, int b_rc, int64_t M, int64_t N, int64_t K, int passes,
                                    int64_t ksplit, int64_t diag_m, int64_t diag_n, float* out, void* workspace,
                                    int64_t workspace_bytes, snuffy_stream_t stream);
int snuffy_gemm_tc_awindow(const void* A_planes, int64_t a_plane_stride, int64_t a_cols_total, int64_t a_col0,
                           const void* B_planes, int64_t b_plane_stride, int64_t M, int64_t N, int64_t K,
                           int passes, float* out, int64_t ldc, snuffy_stream_t stream);

/* ---- a9: sparse attention  O = concat_j softmax_keys(Q_j Kp_j^T / sqrt(dk))^T V_j                      */
/* Replaces matmul / div / softmax / dropout / matmul / transpose+contiguous (snuffy.py:160-168, 187-201).
 * Q,V [B*N,d] with row strides ldq/ldv; Kp [B*Ksel,d]; O [B*Ksel,d]; P_out [B,h,N,Ksel] optional
 * (pre-dropout, as the reference returns it); stats_out [B,h,N,2] = (row max, 1/row sum) optional.     */
int64_t snuffy_sparse_attn_workspace(int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d);
int snuffy_sparse_attn_fwd(const float* Q, int64_t ldq, const float* V, int64_t ldv, const float* Kp,
                           int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d, float dropout_p,
                           uint64_t seed, uint64_t offset, float* O, float* P_out, float* stats_out,
                           void* workspace, int64_t workspace_bytes, snuffy_stream_t stream);

/* Tensor-core version of the same contract (tcgen05, split-bf16 3-pass, P never leaves the SM): Q and V arrive as
 * the planes the Q|V projection wrote over [B*N, ldk] (Q at column q_col0, V at v_col0).  The workspace query
 * returns -1 for shapes it does not serve (needs dk % 32 == 0, dk <= 128, operands within 227 KB of smem).     */
int64_t snuffy_sparse_attn_tc_workspace(int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d);
int snuffy_sparse_attn_tc_fwd(const void* qv_planes, int64_t plane_stride, int64_t ldk, int64_t q_col0,
                              int64_t v_col0, const float* Kp, int64_t B, int64_t N, int64_t Ksel,
                              int64_t h, int64_t d, float dropout_p, uint64_t seed, uint64_t offset,
                              float* O, float* P_out, float* stats_out, uint8_t* drop_mask, void* workspace,
                              int64_t workspace_bytes, snuffy_stream_t stream);
/* drop_mask (optional, with dropout_p > 0): [B, h, N, ceil(Ksel / 8)] bytes, bit e of byte g = keep flag of key 8 g + e —
 * the draw as the forward made it, read back by snuffy_sparse_attn_bwd_tc instead of hashing every score again.        */

/* ---- a15: DSMIL critical-instance pooling   (dsmil.py:83-91: mm, softmax over dim 0, mm, Conv1d)       */
int64_t snuffy_dsmil_workspace(int64_t N, int64_t d, int64_t C);
int snuffy_dsmil_pool_fwd(const float* Q, const float* qmax, const float* V, const float* Wfcc,
                          const float* bfcc, int64_t N, int64_t d, int64_t dq, int64_t C, float* A,
                          float* Bm, float* logits, float* stats_out, void* workspace,
                          int64_t workspace_bytes, snuffy_stream_t stream);

/* ---- packed variable-length bags (BASELINE configs[3]; the reference loops over bags one by one, train.py:249-258)
 * x is the packed [T, d] concatenation, cu_seqlens [B+1] int64 on the device, max_n = longest bag.  Selection returns
 * GLOBAL row indices, so every row-wise entry point above runs on the packed tensor as one bag of T rows; only the
 * per-bag reductions need the offsets: selection, the attention's softmax^T V reduction, and the mean-pool head.      */
int snuffy_select_topk_varlen(const float* scores, const int64_t* cu_seqlens, int64_t B, int64_t max_n, int64_t C,
                              int64_t K, int64_t* idx_out, uint8_t* flags, snuffy_stream_t stream);
int snuffy_select_random_varlen(const uint8_t* flags, const int64_t* cu_seqlens, int64_t B, int64_t max_n, int64_t K,
                                uint64_t seed, uint64_t offset, int64_t* idx_out, snuffy_stream_t stream);
int snuffy_sparse_attn_tc_varlen_fwd(const void* qv_planes, int64_t plane_stride, int64_t ldk, int64_t q_col0,
                                     int64_t v_col0, const float* Kp, const int64_t* cu_seqlens, int64_t B,
                                     int64_t max_n, int64_t Ksel, int64_t h, int64_t d, float* O, void* workspace,
                                     int64_t workspace_bytes, snuffy_stream_t stream);
int snuffy_ln_mean_head_varlen_fwd(const float* x, const int64_t* cu_seqlens, const float* gamma, const float* beta,
                                   const float* Wh, const float* bh, int64_t B, int64_t max_n, int64_t d, int64_t C,
                                   float* partials, uint32_t* tickets, float* pooled, float* bag_out,
                                   snuffy_stream_t stream);

/* ---- backward (train.py:259 loss.backward() through the drop-in modules; autograd only sequences these)   */
/* Batched / split-K form of snuffy_gemm_f32 (no epilogue but alpha and bias): batch z = (zo, zi) offsets the
 * operands by zo*s?_o + zi*s?_i elements (outer = bag, inner = head slice); ksplit > 1 splits the contraction
 * over CTAs with a deterministic fold (dW = dY^T X and dKp = dS^T Q contract over all N patches).            */
int64_t snuffy_gemm_f32_batched_workspace(int64_t nbatch, int64_t M, int64_t N, int64_t ksplit);
int64_t snuffy_gemm_f32_auto_ksplit(int64_t nbatch, int64_t M, int64_t N, int64_t K);
int snuffy_gemm_f32_batched(const float* A, int64_t lda, int a_kc, const float* B, int64_t ldb, int b_kc,
                            float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha,
                            const float* bias, int64_t nb_outer, int64_t nb_inner, int64_t sa_o, int64_t sa_i,
                            int64_t sb_o, int64_t sb_i, int64_t sc_o, int64_t sc_i, int64_t ksplit,
                            void* workspace, int64_t workspace_bytes, snuffy_stream_t stream);
/* LayerNorm backward (autograd of nn.LayerNorm at snuffy.py:107,110,86) with saved (mean, rstd):
 * dx = add + LN'(dy), dgamma_dbeta [2, d].  dy == null: upstream is dy_bcast[row / rows_per_bag] * bscale
 * (the mean-pool of snuffy.py:71).  partials: snuffy_ln_rows_bwd_blocks(rows) * 2 * d floats.               */
int64_t snuffy_ln_rows_bwd_blocks(int64_t rows);
int snuffy_ln_rows_bwd(const float* dy, const float* dy_bcast, int64_t rows_per_bag, float bscale,
                       const float* x, const int32_t* row_map, const float* alt, const float* stats,
                       const float* gamma, const float* add, int64_t rows, int64_t d, float* dx,
                       float* dgamma_dbeta, float* partials, snuffy_stream_t stream);
/* dgamma_dbeta == NULL above leaves the per-CTA partial sums [snuffy_ln_rows_bwd_blocks(rows), 2 d] in `partials`; this folds
 * them (or any [splits, n] partial rows, n % 4 == 0) in a fixed order.  The parameter gradients are off the critical chain of
 * the backward pass, so the fold may be issued on another stream.                                                         */
int snuffy_fold_partials(const float* partials, int64_t splits, int64_t n, float* out, snuffy_stream_t stream);
/* dh = da * dropout_mask * act'(hpre), a_out = act(hpre) * dropout_mask   (snuffy.py:216-225 backward)      */
int snuffy_act_bwd(const float* hpre, const float* da, int act, float dropout_p, uint64_t seed,
                   uint64_t offset, int64_t total, float* dh, float* a_out, snuffy_stream_t stream);
/* out = x + y * dropout_mask: the residual of a stand-alone SublayerConnection.forward (snuffy.py:108,110)  */
int snuffy_residual_dropout(const float* x, const float* y, float dropout_p, uint64_t seed, uint64_t offset,
                            int64_t total, float* out, snuffy_stream_t stream);
/* out[c, :] = sum_rows w[row, c] * X[row, :]  (bias grads; FCLayer weight grad snuffy.py:37)                 */
int64_t snuffy_colsum_chunks(int64_t rows);
int snuffy_colsum(const float* X, int64_t ldx, const float* w, int64_t rows, int64_t d, int64_t C, float* out,
                  float* partials, snuffy_stream_t stream);
/* row-local pieces of the attention backward (snuffy.py:160-168) on materialised [B*h*N, Ksel] scores:
 * mode 0: Pd = softmax(S) * dropout_mask;  mode 1: G (= V dO^T on entry) <- dS                               */
int snuffy_attn_rows_bwd(const float* S, const float* stats, int64_t nrows, int64_t Ksel, int64_t N, int mode,
                         float scale, float dropout_p, uint64_t seed, uint64_t offset, float* Pd, float* G,
                         snuffy_stream_t stream);
/* dx[b, idx[b,k], :] += src[b*K + k, :]   (backward of the raw-row gather, snuffy.py:131,145-147)            */
int snuffy_scatter_add_rows(float* dx, const int64_t* idx, const float* src, int64_t B, int64_t N, int64_t K,
                            int64_t d, snuffy_stream_t stream);

/* Backward of snuffy_sparse_attn_tc_fwd in ONE kernel (autograd of snuffy.py:160-168): S and P are recomputed per 128-query
 * tile from the saved (row max, 1 / row sum); dV = P~ dO, G = V dO^T, dS = P o (D o G - V . dV) / sqrt(dk), dQ = dS Kp and
 * dKp^T = Q^T dS (accumulated in TMEM over the tiles) never leave the SM.  dQV [B*N, 2d] = dQ | dV, dKp [B*Ksel, d].
 * With dropout_p > 0 the keep bits come from `drop_mask` as snuffy_sparse_attn_tc_fwd wrote them.
 * snuffy_sparse_attn_bwd_tc_workspace returns -1 for shapes it does not serve (Ksel > 224, head size not a multiple of 32 or
 * > 128): use the block-diagonal GEMM formulation below.                                                          */
int64_t snuffy_sparse_attn_bwd_tc_workspace(int64_t B, int64_t N, int64_t Ksel, int64_t h, int64_t d);
int snuffy_sparse_attn_bwd_tc(const void* qv_planes, int64_t plane_stride, int64_t ldk, int64_t q_col0,
                              int64_t v_col0, const float* Kp, const float* dO, const float* stats, int64_t B,
                              int64_t N, int64_t Ksel, int64_t h, int64_t d, float dropout_p,
                              const uint8_t* drop_mask, float* dQV, float* dKp, void* workspace,
                              int64_t workspace_bytes, snuffy_stream_t stream);
/* Attention backward on tensor cores (autograd of snuffy.py:160-168): all heads of a bag are contracted by one dense
 * tcgen05 GEMM against head-block operands Kbd[(j,k), c] = Kp[k, c] on head j's columns, 0 elsewhere.
 * snuffy_attn_seg_bwd: the row-local pieces on S_all [N, h*Ksel] (mode 0: P~, mode 1: G <- dS).  planes != NULL
 * (Ksel % 8 == 0, Ksel <= 256, (h*Ksel) % 32 == 0): the result is also written as split-bf16 A-operand planes over
 * [ceil128(N), h*Ksel] (rc = 128) for the next product, and Pd may then be NULL in mode 0.                          */
int snuffy_block_diag_rows(const float* src, int64_t Ksel, int64_t h, int64_t d, float* out, snuffy_stream_t stream);
int snuffy_block_diag_extract(const float* bd, int64_t Ksel, int64_t h, int64_t d, float* out,
                              snuffy_stream_t stream);
int snuffy_attn_seg_bwd(const float* S, const float* stats, int64_t N, int64_t h, int64_t Ksel, int64_t bag, int mode,
                        float scale, float dropout_p, uint64_t seed, uint64_t offset, float* Pd, float* G,
                        void* planes, int64_t plane_stride, snuffy_stream_t stream);
/* DSMIL: backward of A = softmax over the N instances (dsmil.py:86): dS = A (dA - sum_n A dA) / scale          */
int snuffy_softmax_cols_bwd(const float* A, const float* dA, int64_t N, int64_t C, float scale, float* dS,
                            snuffy_stream_t stream);

/* ---- training-loop glue on the device (caller side of the path: train.py:828-846, 468-473)                  */
/* loss = w BCEwL(bag, y) + (1 - w) BCEwL(max_n classes, y), its gradients and the mixed prediction in one launch.
 * terms: 2*B*C floats, ticket: one zeroed uint32; loss[3] = (mixed, bag term, max term).  w_dev (optional): the mix
 * weight read from the device (train.py:804 `single_weight_parameter`, learnable with --soft_average); dw (optional):
 * gscale * d loss / d w.  The arg-max follows torch.max: a NaN score wins and propagates into the loss.          */
int snuffy_mil_loss(const float* classes, const float* bag, const float* label, const float* weight,
                    int64_t B, int64_t N, int64_t C, float w, const float* w_dev, float gscale, float* terms,
                    uint32_t* ticket, float* loss, float* pred, float* dclasses, float* dbag, float* dw,
                    snuffy_stream_t stream);
/* Random draws under CUDA-graph replay.  Every (seed, offset) pair of this header (dropout, random patches) may be given
 * INDIRECTLY: seed bit 63 set  =>  `offset` is the address of a device uint64 step counter and the draw used is
 * (seed & 0xFFFFFFFF, *counter + ((seed >> 32) & 0x7FFFFFFF)).  snuffy_rng_advance(counter, delta) is enqueued as the
 * last node of the captured step: replays then differ, while forward and backward of one replay agree.  Direct
 * (seed, offset) pairs must therefore keep seed bit 63 clear.                                                     */
int snuffy_rng_advance(uint64_t* counter, uint64_t delta, snuffy_stream_t stream);
int64_t snuffy_sumsq_blocks(int64_t n);
int snuffy_sumsq(const float* x, int64_t n, float* partials, float* out, snuffy_stream_t stream);
/* torch.optim.AdamW step over flat fp32 buffers (train.py:809-826), gradient pre-scale and optional global-norm
 * clip (train.py:469-470) folded in.                                                                            */
int snuffy_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int64_t step, float gscale,
                      const float* gnorm_sq, float max_norm, snuffy_stream_t stream);
/* The same step as a node of a captured CUDA graph: the 1-based step count, the number of contributing ranks (optional;
 * the update divides by it and is skipped at 0) and the learning rate (optional) are read from the device; clamp_lo <=
 * clamp_hi clamps the updated values (train.py:852-854: the learnable mix weight stays in [0, 1]).               */
int snuffy_adamw_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                          float beta2, float eps, float weight_decay, const int64_t* step_dev,
                          const float* contributors, const float* lr_dev, float gscale, const float* gnorm_sq,
                          float max_norm, float clamp_lo, float clamp_hi, snuffy_stream_t stream);

/* dst[offsets[i] .. +sizes[i]) = srcs[i][0 .. sizes[i])  for n tensors (host arrays of device pointers / element counts):
 * one launch per 32 tensors packs the per-parameter gradients into the flat all-reduce / optimizer buffer.         */
int snuffy_pack_f32(const void* const* srcs, const int64_t* sizes, const int64_t* offsets, int64_t n, float* dst,
                    snuffy_stream_t stream);

/* ---- patch-level outputs on the device (step after the path in valid/test)                                    */
/* probs[i] = sigmoid(scores[i]), i < n  (train.py:913-916): the per-bag `attentions` the reference copies to the host
 * one bag at a time (train.py:271, 345, 354); the caller points probs at a row offset of one epoch-wide buffer.   */
int snuffy_patch_probs(const float* scores, int64_t n, float* probs, snuffy_stream_t stream);
/* FROC detection tuples (train.py:342-345 + mp_thresholding train.py:138-141).  Slide b owns rows
 * [cu_seqlens[b], cu_seqlens[b+1]) (cu_seqlens NULL: one slide of `total` rows); probs[r * prob_stride] is row r's
 * probability, positions int32 [total, 2] its patch grid coordinates.  Rows with probability > threshold (strict) are
 * kept in patch order: det_prob[start + j], det_xy[start + j] = (x * tile + half, y * tile + half), count[b] = kept.  */
int snuffy_froc_detections(const float* probs, int64_t prob_stride, const int32_t* positions,
                           const int32_t* cu_seqlens, int64_t slides, int64_t total, float threshold,
                           int32_t tile, int32_t half, float* det_prob, int32_t* det_xy, int32_t* count,
                           snuffy_stream_t stream);

/* ---- (e) data-parallel exchange: gradient all-reduce over NVLink peer memory (csrc/comm.cu) ---------------------------
 * The reference trains one process on one GPU (train.py:249-264); sharding slides over ranks adds ONE exchange per optimizer
 * step: the sum of the flat gradient.  Every rank allocates one block with snuffy_comm_alloc (plain cudaMalloc, zero filled),
 * exports it (snuffy_comm_export -> snuffy_comm_handle_bytes() opaque bytes, exchanged by the host through any channel) and
 * maps its peers' blocks (snuffy_comm_import).  snuffy_peer_allreduce is then one kernel: barrier, every rank sums its slice
 * of all buffers in rank order reading peer memory directly, writes the sum into all buffers, barrier.  In place, identical
 * bits on every rank, capturable in a CUDA graph; all ranks must make the same sequence of calls.                        */
int snuffy_comm_alloc(int64_t bytes, void** ptr);
int snuffy_comm_free(void* ptr);
int snuffy_comm_handle_bytes(void);
int snuffy_comm_export(void* ptr, void* handle_out);
int snuffy_comm_import(const void* handle, void** ptr);
int snuffy_comm_close(void* ptr);
int snuffy_comm_counter_bytes(void);
int snuffy_peer_allreduce(void* const* bufs, void* const* counters, void* state, int rank, int world, int64_t n,
                          snuffy_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SNUFFY_B200_H */
