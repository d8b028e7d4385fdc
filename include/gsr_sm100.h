/*
 * gsr_sm100.h -- C ABI of libgsr_sm100.so, the B200 (sm_100a) kernels behind the
 * GSR-Net hot path of denisfouchard/fMRI-connectivity-super-resolution.
 *
 * The reference has no FFI of its own (it is pure Python on PyTorch eager ops);
 * the boundary a maintainer binds is therefore "one entry point per eager op (or
 * fused group of eager ops) the reference calls on this path".  Every
 * declaration below cites the reference lines it replaces (paths relative to
 * the reference repository root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     its name ends in _host; matrices are fp32, row-major, contiguous, with a
 *     leading batch ("subjects") dimension; indices are int64 (torch's dtype);
 *   - `stream` is a cudaStream_t passed as void*; calls never synchronise,
 *     never allocate, never throw; the caller owns all memory and keeps it
 *     alive until the stream has consumed it;
 *   - return value: 0 = ok, < 0 = bad argument (GSR_E_*), > 0 = cudaError_t of
 *     the launch.  gsr_last_error() gives a static, thread-local description.
 */
#ifndef GSR_SM100_H
#define GSR_SM100_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define GSR_API __attribute__((visibility("default")))
#else
#define GSR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_ABI_VERSION 1

#define GSR_E_ARG        (-1)   /* null pointer / negative size                     */
#define GSR_E_UNSUPPORTED (-2)  /* size outside what the kernels are built for      */
#define GSR_E_ALIGN      (-3)   /* pointer / leading dimension not 16-byte aligned  */

GSR_API int         gsr_abi_version(void);
GSR_API const char* gsr_last_error(void);
/* Number of kernels launched by this library since load (bench.py "gpu_launches"). */
GSR_API uint64_t    gsr_launch_count(void);
/* Kernel-variant selection for benchmarking and tests: "gemm_variant" 0, 1, 2 (csrc/bgemm.cu; identical
 * results), "tcgemm_pairs" 1 / 0 (CTA-pair tensor kernel or one CTA per tile; identical results),
 * "tcgemm_ragged_rows" 1 / 0 (FP32 kernels or the tensor kernel for the <= 16 rows after the pairs;
 * equal to rounding), "eigh_ql_max_iter", "eigh_oddeven", "eigh_max_polish".        */
GSR_API int         gsr_set_tuning(const char* key, int value);

/* ------------------------------------------------------------------ a1
 * utils/preprocessing.py:16-25  normalize_adj_torch
 *   r = rowsum(M)^-1/2 (inf -> 0);  out[i,j] = (M[j,i] * r_i) * r_j   (no +I, transposed)
 * pre = 0: M = mx.   pre = 1: M = |mx| with unit diagonal, i.e. the
 *   `f_d = abs(f_d); f_d.fill_diagonal_(1)` of models/layers.py:55-56 fused in
 *   front (mx is then the raw W.U^T.f product).
 * r_out [batch,n] is kept for the backward; m_out (nullable) receives M.        */
GSR_API int gsr_normalize_adj_fwd(const float* mx, float* out, float* r_out, float* m_out,
                          int batch, int n, int pre, void* stream);
/* autograd of the above: grad_mx = d/d mx  sum(out * grad_out).  With pre = 1
 * the gradient passes through abs (sgn(0) = 0) and is zero on the diagonal
 * (fill_diagonal_ overwrites it).                                     */
GSR_API int gsr_normalize_adj_bwd(const float* mx, const float* r, const float* grad_out,
                          float* grad_mx, int batch, int n, int pre, void* stream);
/* Same gradient, tile-pair kernels (two passes over 64 x 64 tile pairs with a deterministic partial-sum
 * workspace of gsr_normalize_adj_bwd_workspace(batch, n) FLOATS): the fast path for n % 4 == 0 and 16-byte
 * aligned operands; anything else (or workspace == NULL) runs gsr_normalize_adj_bwd.                    */
GSR_API size_t gsr_normalize_adj_bwd_workspace(int batch, int n);
GSR_API int gsr_normalize_adj_bwd_tiles(const float* mx, const float* r, const float* grad_out, float* grad_mx,
                                float* workspace, int batch, int n, int pre, void* stream);
/* The same with the incoming gradient given as the SUM of two tensors (grad_out2 may be NULL), added as the passes read
 * them: d adj = 2 dX0 adj + (dH2 S2^T + dH1 S1^T) (models/layers.py:57-59 and 85-86) without an addend epilogue on the
 * contraction or a separate addition pass.  Tile path only (GSR_E_UNSUPPORTED otherwise when grad_out2 != NULL).        */
GSR_API int gsr_normalize_adj_bwd_tiles2(const float* mx, const float* r, const float* grad_out, const float* grad_out2,
                                 float* grad_mx, float* workspace, int batch, int n, int pre, void* stream);


/* ------------------------------------------------------------------ a2 / a12
 * models/layers.py:27 and utils/gsr_train.py:34  torch.linalg.eigh(A, UPLO='U')
 * Batched one-sided (Hestenes) Jacobi on A + sigma*I, one CTA per graph, the working matrix
 * resident in shared memory (n <= GSR_EIGH_MAX_N_CTA); for n <= ~170 a second, polishing
 * phase (implicit two-sided Jacobi on a freshly computed A.U) brings the eigenvectors to
 * LAPACK's accuracy class.  Larger n (the 268-node HR target, up to GSR_EIGH_MAX_N) keep the
 * working matrix in the caller's `workspace` (gsr_eigh_workspace(batch, n) floats, 0 when none
 * is needed; L2 resident) -- the once-per-dataset, hoisted path.  Only the upper triangle of
 * A is read.  Eigenvalues ascending in w[batch,n]; eigenvector k in COLUMN k of
 * U[batch,n,n].  canonical_sign != 0 applies the documented sign rule: the
 * component of largest magnitude (lowest row on ties) of every eigenvector is
 * positive.  info[batch]: > 0 = sweeps used (bits 0-7: one-sided sweeps, bits 8-15: polish
 * sweeps); -1 = non-finite input (outputs are
 * NaN); -2 = not converged after max_sweeps (torch raises LinAlgError in both
 * cases; the Python wrapper does the same).  tol <= 0 / max_sweeps <= 0 select
 * the defaults (1e-6, 30).                                                       */
#define GSR_EIGH_MAX_N_CTA 224
#define GSR_EIGH_MAX_N     288
GSR_API size_t gsr_eigh_workspace(int batch, int n);
GSR_API int gsr_eigh_jacobi_f32(const float* A, float* w, float* U, int32_t* info, float* workspace,
                        int batch, int n, int canonical_sign, float tol, int max_sweeps,
                        void* stream);

/* Same contract (torch.linalg.eigh(A, UPLO='U'), models/layers.py:27 and utils/gsr_train.py:34), tridiagonal
 * route for n <= GSR_EIGH_TRIDIAG_MAX_N: Householder reduction, implicit-shift QL on (d, e) with one warp per
 * graph recording its plane rotations, and a replay of the rotation stream on the rows of Q.  Up to 160 nodes
 * (the in-step LR decomposition) the matrices live in REGISTERS (one thread per row); 161..288 nodes (the
 * 268-node HR target) keep the packed triangle / the rows of Z in shared memory.  ~6 n^3 flops instead of
 * the ~18 n^3 of the Jacobi kernel, LAPACK-class accuracy.  `workspace` holds
 * gsr_eigh_tridiag_workspace(batch, n) BYTES (16-byte aligned).
 * info[batch]: > 0 = QL iterations used; -1 = non-finite input (outputs are NaN); -2 = QL did
 * not converge in 30 iterations per eigenvalue or needed more than 2 n^2 rotations (outputs
 * are NaN; callers fall back to gsr_eigh_jacobi_f32 for those graphs).                       */
#define GSR_EIGH_TRIDIAG_MAX_N 288
GSR_API size_t gsr_eigh_tridiag_workspace(int batch, int n);
GSR_API int gsr_eigh_tridiag_f32(const float* A, float* w, float* U, int32_t* info, void* workspace,
                         size_t workspace_bytes, int batch, int n, int canonical_sign, void* stream);

/* ------------------------------------------------------------------ a3-a5, a7
 * Every torch.mm / torch.matmul / nn.Linear on the path:
 *   models/layers.py:37,45,54 (W.U^T.f, without the zero padding), :59 (adj.adj^T),
 *   :85-86 (adj @ (x @ W)), models/ops.py:60 (Linear) and their autograd.
 * C[b] = alpha * op(A[b]) . op(B[b]) (+ epilogue), b < batch, fp32 SIMT FMA.
 *   op = transpose when trans* != 0; A is M x K (or K x M when transposed) with
 *   leading dimension lda; stride* are ELEMENT strides between subjects, 0 = the
 *   operand is shared by all subjects (a weight).
 * epilogue: GSR_EPI_NONE; GSR_EPI_BIAS adds extra[n] (Linear bias);
 *   GSR_EPI_ADD adds extra[b][m][n] (ld_extra, stride_extra) -- the "+=" of a
 *   gradient with several producers and the U-Net skip connection.
 * Leading dimensions and strides must be multiples of 4 and pointers 16-byte
 * aligned for the fast path; anything else takes a scalar-load variant of the
 * same kernel (same arithmetic order).                                           */
#define GSR_EPI_NONE 0
#define GSR_EPI_BIAS 1
#define GSR_EPI_ADD  2
GSR_API int gsr_bgemm_f32(const float* A, const float* B, float* C,
                  int M, int N, int K, int lda, int ldb, int ldc,
                  long long strideA, long long strideB, long long strideC,
                  int transA, int transB, int batch, float alpha,
                  int epilogue, const float* extra, int ld_extra, long long stride_extra,
                  void* stream);
/* Weight gradients: C = alpha * sum_b op(A[b]) . op(B[b]) (+ extra[m][n] when
 * epilogue == GSR_EPI_ADD).  Subjects are dealt to `splits` CTAs per tile which
 * each accumulate in registers; partial tiles go to workspace
 * [splits, M, N] and are summed in a fixed order (deterministic).
 * gsr_bgemm_reduce_workspace() returns the floats needed and the split count
 * the library would pick for (M, N, batch).                                      */
GSR_API size_t gsr_bgemm_reduce_workspace(int M, int N, int batch, int* splits_out);
GSR_API int gsr_bgemm_reduce_f32(const float* A, const float* B, float* C,
                         int M, int N, int K, int lda, int ldb, int ldc,
                         long long strideA, long long strideB,
                         int transA, int transB, int batch, float alpha,
                         int epilogue, const float* extra, int ld_extra,
                         float* workspace, int splits, void* stream);
/* The same two entry points on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
 * accumulators in tensor memory) at fp32 accuracy: every operand is split hi + lo into two
 * TF32 values by the producer warps and three MMA passes (lo.hi + hi.lo + hi.hi) feed one fp32
 * accumulator.  Same arguments and semantics as gsr_bgemm_f32 / gsr_bgemm_reduce_f32; operands
 * must be 16-byte aligned with leading dimensions % 4 == 0 (GSR_E_ALIGN otherwise -- callers
 * fall back to the SIMT entry, e.g. for the ragged Graph U-Net shapes).              */
GSR_API int gsr_tcgemm_supported(int M, int N, int K, int lda, int ldb);
GSR_API int gsr_tcgemm_f32(const float* A, const float* B, float* C,
                   int M, int N, int K, int lda, int ldb, int ldc,
                   long long strideA, long long strideB, long long strideC,
                   int transA, int transB, int batch, float alpha,
                   int epilogue, const float* extra, int ld_extra, long long stride_extra,
                   void* stream);
/* C = alpha (op(A1) op(B1) + op(A2) op(B2)): two products of one shape (M x N, inner sizes K1 and K2) and the same operand
 * majors summed into one result -- d adj = dH2 S2^T + dH1 S1^T of the two graph convolutions (models/layers.py:85-86 applied
 * twice, models/gsr_model.py:30-32).  Both products accumulate in tensor memory inside one launch of the TMA kernel, so the
 * first product is never written out and read back as an addend.                                                          */
GSR_API int gsr_tcgemm_dual_f32(const float* A1, const float* B1, const float* A2, const float* B2, float* C,
                        int M, int N, int K1, int K2, int lda1, int ldb1, int lda2, int ldb2, int ldc,
                        long long strideA1, long long strideB1, long long strideA2, long long strideB2, long long strideC,
                        int transA, int transB, int batch, float alpha, void* stream);
GSR_API size_t gsr_tcgemm_reduce_workspace(int M, int N, int batch, int* splits_out);
GSR_API int gsr_tcgemm_reduce_f32(const float* A, const float* B, float* C,
                          int M, int N, int K, int lda, int ldb, int ldc,
                          long long strideA, long long strideB,
                          int transA, int transB, int batch, float alpha,
                          int epilogue, const float* extra, int ld_extra,
                          float* workspace, int splits, void* stream);
/* K-split contractions (one skinny product with a long inner dimension: the notebook MLP's 512 x 1024 x 12 720 /
 * 35 778 layers, training_template.ipynb:123-145): the caller runs the K slices as the items of ONE batched
 * gsr_tcgemm_f32 / gsr_bgemm_f32 call (strideA / strideB = the slice, strideC = M * N into workspace
 * [splits, M, N]); this sums the partial products in a fixed order and applies the product's epilogue:
 *   C[m][n] = alpha * sum_z workspace[z][m][n] (+ extra[n] | + extra[m][n]);  extra may alias C.               */
GSR_API int gsr_splitk_reduce_f32(const float* workspace, float* C, int M, int N, int ldc, int splits, float alpha,
                          int epilogue, const float* extra, int ld_extra, void* stream);
/* Column sums of a [rows, cols] matrix (Linear bias gradient, models/ops.py:60). */
GSR_API int gsr_colsum_f32(const float* X, float* out, int rows, int cols, int ld, void* stream);

/* ------------------------------------------------------------------ a4 / a6
 * models/layers.py:60-63 and models/gsr_model.py:32-37:
 *   X = (X + X^T)/2; X[diag] = 1; return |X|                                     */
GSR_API int gsr_sym_diag_abs_fwd(const float* x, float* out, int batch, int n, void* stream);
/* autograd of the above: with z = (x + x^T)/2, dz = grad_out * sgn(z) (0 on the
 * diagonal), grad_x = (dz + dz^T)/2.                                             */
GSR_API int gsr_sym_diag_abs_bwd(const float* x, const float* grad_out, float* grad_x,
                         int batch, int n, void* stream);
/* models/gsr_model.py:32-37 followed by nn.L1Loss against the HR target (utils/gsr_train.py:9,39-43) in ONE pass:
 *   loss_out[b] = mean | |(x_b + x_b^T)/2| (unit diagonal) - y_b |
 * x, y [batch, n, n]; the prediction itself is written to pred_out only when it is not NULL.  partials: double
 * [batch * gsr_sym_diag_abs_l1_partials(n)] (deterministic two-stage sum).  Needs n % 4 == 0 and 16-byte aligned
 * tensors (GSR_E_UNSUPPORTED otherwise: callers then use gsr_sym_diag_abs_fwd + gsr_l1_loss_fwd).               */
GSR_API int gsr_sym_diag_abs_l1_partials(int n);
GSR_API int gsr_sym_diag_abs_l1_fwd(const float* x, const float* y, float* pred_out, float* loss_out, double* partials,
                            int batch, int n, void* stream);
/* autograd of the above with respect to x, given grad_loss [batch]: bit identical to gsr_l1_loss_bwd followed by
 * gsr_sym_diag_abs_bwd, without the [batch, n, n] gradient tensor between them.                                */
GSR_API int gsr_sym_diag_abs_l1_bwd(const float* x, const float* y, const float* grad_loss, float* grad_x,
                            int batch, int n, void* stream);
/* models/layers.py:55-56 alone (the stored `self.f_d`): out = |x|, unit diagonal. */
GSR_API int gsr_abs_diag1_fwd(const float* x, float* out, int batch, int n, void* stream);

/* ------------------------------------------------------------------ a8-a10
 * models/ops.py:35-38  scores = sigmoid((X . p + c) / 100)                       */
GSR_API int gsr_pool_scores(const float* X, const float* p, const float* c, float* scores,
                    int n, int d, void* stream);
/* models/ops.py:42  torch.topk(scores, k) under the stated tie rule (score
 * descending, index ascending; NaN sorts first as in torch).  One CTA, n <= 4096. */
GSR_API int gsr_pool_topk(const float* scores, int n, int k, int64_t* idx, float* values,
                  void* stream);
/* models/ops.py:43-45  new_X = X[idx, :] * values[:, None]                       */
GSR_API int gsr_pool_gather_scale_fwd(const float* X, const int64_t* idx, const float* values,
                              float* new_X, int k, int d, void* stream);
/* autograd of pool_scores + topk values + gather_scale w.r.t. X, p, c.           */
GSR_API int gsr_pool_bwd(const float* X, const float* p, const float* scores, const int64_t* idx,
                 const float* grad_new_X, float* grad_X, float* grad_p, float* grad_c,
                 float* ds_workspace /* n floats */, int n, int k, int d, void* stream);
/* models/ops.py:46-47  A[idx, :][:, idx]  (bit exact gather)                     */
GSR_API int gsr_gather_rows_cols(const float* A, const int64_t* idx, float* out,
                         int batch, int n, int k, void* stream);
/* models/ops.py:22-23  new_X = zeros(n, d); new_X[idx] = X   and its autograd
 * (row gather).                                                                  */
GSR_API int gsr_unpool_fwd(const float* X, const int64_t* idx, float* new_X, int n, int k, int d,
                   void* stream);
GSR_API int gsr_unpool_bwd(const float* grad_new_X, const int64_t* idx, float* grad_X, int k, int d,
                   void* stream);

/* ------------------------------------------------------------------ a13
 * utils/gsr_train.py:9,39-43  nn.L1Loss(): per-subject mean |x - y|.
 *   x, y: [batch, count]; y_stride = 0 shares y across the batch.
 *   loss_out[batch] (fp32).  Deterministic two-stage reduction; partials needs
 *   gsr_l1_partials(count) * batch doubles.                                      */
GSR_API int gsr_l1_partials(long long count);
GSR_API int gsr_l1_loss_fwd(const float* x, const float* y, long long y_stride, float* loss_out,
                    double* partials, int batch, long long count, void* stream);
/* grad_x[b,i] = sgn(x - y) * grad_loss[b] / count   (sgn(0) = 0)                 */
GSR_API int gsr_l1_loss_bwd(const float* x, const float* y, long long y_stride, const float* grad_loss,
                    float* grad_x, int batch, long long count, void* stream);
/* y shared by the batch (the |W - U_hr| term, utils/gsr_train.py:41, batched): the gradient of y alone,
 * grad_y[i] = - sum_b sgn(x[b,i] - y[i]) * grad_loss[b] / count, in one pass over x (fixed order). */
GSR_API int gsr_l1_loss_bwd_shared_y(const float* x, const float* y, const float* grad_loss, float* grad_y,
                             int batch, long long count, void* stream);

/* ------------------------------------------------------------------ a16 / a17
 * utils/matrix_vectorizer.py:24-58  vectorize: strict upper triangle walked
 * COLUMN-major; include_diagonal appends the SUB-diagonal element (c+1, c)
 * after each column.  elem_size in {1,2,4,8,16} bytes (dtype is preserved, so
 * this is a byte gather).  out: [batch, gsr_vectorize_len(n, include_diagonal)]. */
GSR_API long long gsr_vectorize_len(int n, int include_diagonal);
GSR_API int gsr_vectorize(const void* matrix, void* out, int batch, int n, int elem_size,
                  int include_diagonal, void* stream);
/* utils/matrix_vectorizer.py:61-95  anti_vectorize: ROW-major strict upper fill
 * (torch.triu_indices order) mirrored; include_diagonal puts
 * vec[n_el : n_el + n] on the MAIN diagonal; fp32 out.  vec_len >= n(n-1)/2 is
 * the reference's assert (checked on the host side as well).                     */
GSR_API int gsr_anti_vectorize(const float* vec, float* out, int batch, int n, long long vec_len,
                       int include_diagonal, void* stream);
/* training_gsr.ipynb:112-121 (submission file)  matrix[np.triu_indices(n, k=1)]: the strict upper
 * triangle in ROW-major order (the inverse of anti_vectorize's fill; utils/data.py:214-231 writes
 * the COLUMN-major gsr_vectorize order instead).  out: [batch, n(n-1)/2] fp32.                  */
GSR_API int gsr_triu_flatten(const float* matrix, float* out, int batch, int n, void* stream);


/* ------------------------------------------------------------------ a18
 * utils/evaluation.py:96-112  MAE, Pearson r, Jensen-Shannon distance over the
 * concatenated vectorised upper triangles of pred / gt.  The three numbers do
 * not depend on the order of the concatenation, so the kernels read the strict
 * upper triangle of [batch, n, n] directly (upper != 0) or flat vectors
 * (upper == 0, n = row length).  fp64 accumulators.
 *   pass 1 -> sums[0..5] = { sum|p-g|, sum p, sum g, sum p^2, sum g^2, sum p*g }
 *   pass 2 -> sums[6..7] = { sum a*log(a/m), sum b*log(b/m) },  a = g/sum g,
 *             b = p/sum p, m = (a+b)/2   (rel_entr conventions of scipy)
 * results_out[0..2] = { mae, pcc, js_distance } (device, fp64).
 * workspace: gsr_eval_workspace() doubles.                                       */
GSR_API size_t gsr_eval_workspace(void);
GSR_API int gsr_eval_reduce(const float* pred, const float* gt, long long batch, int n, int upper,
                    double* workspace, double* results_out, void* stream);

/* ------------------------------------------------------------------ (e)
 * Pack / unpack the 31 gradient tensors into one flat fp32 bucket for the
 * single NCCL all-reduce of the subject-data-parallel step.  ptrs[n_tensors] is a HOST
 * array of device pointers, offsets[n_tensors + 1] the HOST array of element offsets of
 * each tensor inside the bucket (both are read during the call and travel in the kernel
 * parameters, so the launches can be captured in a CUDA graph); values are multiplied by
 * scale. */
GSR_API int gsr_flat_pack(const float* const* ptrs, const long long* offsets, int n_tensors,
                  float* flat, long long total, float scale, void* stream);
GSR_API int gsr_flat_unpack(float* const* ptrs, const long long* offsets, int n_tensors,
                    const float* flat, long long total, float scale, void* stream);

/* ------------------------------------------------------------------ N2 (SURVEY 8 f)
 * MLP baseline, models/mlp.py:7-27: the BatchNorm1d -> Dropout(p) -> ReLU run between two
 * Linear layers (the Linear layers themselves are gsr_tcgemm_f32 / gsr_bgemm_f32 with the
 * BIAS epilogue).  x, y, dy, dx are [batch][features] row-major fp32.
 * gsr_bn_stats_f32: training statistics of nn.BatchNorm1d -- mean[features], biased
 *   var[features]; when running_mean / running_var are given they are updated in place as
 *   torch does ((1 - momentum) * running + momentum * batch value, unbiased variance).
 * gsr_bn_act_fwd_f32: y = act( drop( gamma * (x - mean) / sqrt(var + eps) + beta ) ), act(z) = z > 0 ? z :
 *   negative_slope * z -- 0 is nn.ReLU (models/mlp.py), 0.01 the notebook model's nn.LeakyReLU
 *   (training_template.ipynb:117), 1 no activation.
 *   (mean, var) are the batch statistics when training, the running estimates in eval mode.
 *   Dropout keeps element e iff u(seed, e) >= p_drop and scales by 1 / (1 - p_drop), u being
 *   the top 24 bits of splitmix64(seed + e * 0x9E3779B97F4A7C15) / 2^24; p_drop = 0 disables
 *   it.  Nothing is stored for the backward pass.
 * gsr_bn_act_bwd_f32: dx, dgamma[features], dbeta[features] from dy with the same seed;
 *   batch_stats = 1 when (mean, var) were computed from x (training), 0 for constants.  */
GSR_API int gsr_bn_stats_f32(const float* x, int batch, int features, float* mean, float* var,
                     float* running_mean, float* running_var, float momentum, void* stream);
GSR_API int gsr_bn_act_fwd_f32(const float* x, const float* mean, const float* var, const float* gamma,
                       const float* beta, float eps, float p_drop, unsigned long long seed, float negative_slope,
                       float* y, int batch, int features, void* stream);
GSR_API int gsr_bn_act_bwd_f32(const float* x, const float* mean, const float* var, const float* gamma,
                       const float* beta, const float* dy, float eps, float p_drop,
                       unsigned long long seed, float negative_slope, int batch_stats, float* dx, float* dgamma,
                       float* dbeta, int batch, int features, void* stream);
/* The notebook MLP of BASELINE configs[3] (training_template.ipynb:100-216: 12 720 -> hidden -> 35 778 on vectorised
 * triangles) adds, around the same Linear / BatchNorm passes:
 * gsr_add_act_fwd_f32: out = act(a + b)  -- the residual connection and F.leaky_relu of :176-180;
 * gsr_act_bwd_f32:     grad_in = grad_out * (out > 0 ? 1 : negative_slope)  (the same for both addends);
 * gsr_sigmoid_fwd_f32 / gsr_sigmoid_bwd_f32: nn.Sigmoid of the output layer (:132), grad_x = grad_y * y * (1 - y).   */
GSR_API int gsr_add_act_fwd_f32(const float* a, const float* b, float* out, long long count, float negative_slope, void* stream);
GSR_API int gsr_act_bwd_f32(const float* out, const float* grad_out, float* grad_in, long long count, float negative_slope, void* stream);
GSR_API int gsr_sigmoid_fwd_f32(const float* x, float* y, long long count, void* stream);
GSR_API int gsr_sigmoid_bwd_f32(const float* y, const float* grad_y, float* grad_x, long long count, void* stream);
/* spectral_norm(nn.Linear) of the same model (:125,134,143; torch.nn.utils.spectral_norm.compute_weight): the weight the
 * layer uses is W / sigma with sigma = u . (W v) (u, v: the power-iteration buffers, constants for autograd).
 * gsr_scale_div_f32:    y = x / *sigma  (sigma is a DEVICE scalar: nothing synchronises);
 * gsr_spectral_bwd_f32: grad_w = grad_w_eff / sigma + dsigma u v^T with dsigma = -sum(grad_w_eff o w_eff) / sigma (the
 *   sum as fp64 partials folded in a fixed order; workspace: gsr_spectral_bwd_workspace() doubles).                    */
GSR_API int gsr_scale_div_f32(const float* x, const float* sigma, float* y, long long count, void* stream);
GSR_API size_t gsr_spectral_bwd_workspace(void);
GSR_API int gsr_spectral_bwd_f32(const float* grad_w_eff, const float* w_eff, const float* u, const float* v, const float* sigma,
                         float* grad_w, int rows, int cols, double* workspace, void* stream);

/* ------------------------------------------------------------------ N3 (SURVEY 8 f)
 * torch.optim.Adam(model.parameters(), lr) of training_gsr.ipynb:72 over all parameter tensors in
 * one launch per 40 tensors: m = lerp(m, g, 1 - beta1); v = beta2 v + (1 - beta2) g^2;
 * p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps); weight_decay adds wd * p to g (L2).
 * params / grads / exp_avg / exp_avg_sq / sizes are HOST arrays of n_tensors entries (device pointers and
 * element counts; read during the call).  step is a DEVICE fp32 scalar holding t - 1; it is incremented on
 * the stream before the update, so the call needs no host synchronisation and can be captured in a graph. */
GSR_API int gsr_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                  const long long* sizes, int n_tensors, float* step, float lr, float beta1, float beta2,
                  float eps, float weight_decay, void* stream);

/* ------------------------------------------------------------------ measurement aid
 * Launches an FFMA-only kernel (8 independent chains per thread, 148*8 CTAs of 256 threads)
 * and returns the flops it executes; timed by the caller with CUDA events to obtain the
 * FP32 SIMT peak that bounds the contraction kernels (bench.py roofline).           */
GSR_API double gsr_ffma_probe(int iters, float* scratch, void* stream);
/* The same for the tensor pipe: every CTA pair (74 per B200) issues iters x 4 tcgen05.mma.cta_group::2 kind::tf32 of
 * M 256 x N 256 x K 8 on resident operands; returns the dense TF32 flops executed.  The fp32-accurate (3xTF32)
 * contraction kernels issue three such MMAs per algorithmic product: their ceiling is a third of this rate.        */
GSR_API double gsr_tf32_probe(int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSR_SM100_H */
