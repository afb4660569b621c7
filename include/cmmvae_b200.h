/* cmmvae_b200.h -- C ABI of libcmmvae_b200.so: hand-written sm_100a kernels for the CMMVAE
 * training step.  Plain pointers + sizes; no torch types.  Every pointer is a DEVICE pointer
 * unless it says "host".  Every launch is asynchronous on `stream` (a cudaStream_t passed as
 * void*), performs no hidden synchronisation and allocates nothing: the caller owns all buffers.
 * Return value: 0 on success, negative on error (message via cmmvae_last_error(), thread local).
 *
 * The reference (zdebruine/MMVAE) has no FFI layer: its boundary is the Python module API of
 * src/cmmvae.  Each entry point below cites the reference code whose arithmetic
 * it replaces; the Python host side in mmvae_b200/ mirrors the reference's nn.Module /
 * LightningModule interface and is the only caller (through ctypes).
 *
 * dtype enums: CMMVAE_F32 = 0, CMMVAE_BF16 = 1.  Matrices are row-major, leading dimension in
 * elements.
 */
#ifndef CMMVAE_B200_H
#define CMMVAE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMMVAE_ABI_VERSION 2
#define CMMVAE_F32 0
#define CMMVAE_BF16 1

int cmmvae_abi_version(void);
const char* cmmvae_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py gpu_launches) */
long long cmmvae_launch_count(void);
/* SMs the persistent tensor kernels plan their grids / split-K factors for (default 148).  Lower it while
 * communication kernels (NCCL) share the GPU so that all planned CTAs are co-resident. */
int cmmvae_set_sm_budget(int sms);
/* programmatic dependent launch between consecutive kernels of a stream (default on; CMMVAE_PDL=0 disables it for
 * the process).  The host turns it off while it CAPTURES a CUDA graph: measured on B200, graph nodes with
 * programmatic edges replay slower (1.65 ms per step) than plain nodes (1.55 ms), while stream launches gain from
 * it (1.58 -> 1.50 ms). */
int cmmvae_set_pdl(int on);

/* ---- K1: expert-encoder first layer on a CSR batch ---------------------------------------
 * replaces nn.Linear applied to torch.sparse_csr input: components.py:276,306 (FCBlock layer 0
 * of Expert.encoder, components.py:853); batch layout from cellxgene_datapipe.py:178-183.
 * Y[B,H] (f32) = X_csr[B,G] * Wt[G,H] + bias.  Wt is the TRANSPOSED weight (physical [G,H],
 * f32 or bf16).  crow int32[B+1], col int32[nnz] sorted & duplicate-free per row, val f32. */
int cmmvae_csr_linear_fwd(const int32_t* crow, const int32_t* col, const float* val,
                          int B, int G, int H, const void* Wt, int w_dtype, const float* bias,
                          float* Y, void* stream);

/* CSR -> CSC of the same batch (needed by the weight gradient).  cptr int32[G+1], ridx int32[nnz],
 * cval f32[nnz]; `cursor` is an int32[G+1] scratch.  Order inside a column is unspecified. */
int cmmvae_csr_transpose(const int32_t* crow, const int32_t* col, const float* val,
                         int B, int G, long long nnz, int32_t* cptr, int32_t* ridx, float* cval,
                         int32_t* cursor, void* stream);

/* K1b: autograd of the sparse addmm: dWt[G,H] = X^T dY, via the CSC arrays.
 * Genes absent from the batch get exactly 0.  dY f32 [B,H].  beta=0 overwrites. */
int cmmvae_csr_linear_bwd_w(const int32_t* cptr, const int32_t* ridx, const float* cval,
                            int B, int G, int H, const float* dY, float* dWt, void* stream);

/* Tensor-pipe form of K1/K1b (bf16 policy): the CSR batch is densified tile by tile inside shared memory,
 * directly in the tcgen05 operand layout, and multiplied on the tensor cores; HBM sees only the CSR
 * arrays, the bf16 weight (or dY) and the output.  `tile_ptr` (device, cmmvae_csr_tile_ptr_bytes) is
 * the per-(cell, 64-gene window) CSR pointer table built by cmmvae_csr_tile_ptr, bit-exact with
 * crow/col.  Wt_bf16 [G,H], dY_bf16 [B,H], H % 8 == 0.  Faster than the gather kernels above
 * ~1.5 % density; identical contract otherwise (fwd: Y = X Wt + bias; bwd: dWt = X^T dY, genes
 * absent from the batch get exactly 0). */
size_t cmmvae_csr_tile_ptr_bytes(int B, int G);
size_t cmmvae_csr_packed_bytes(long long nnz);
/* builds tile_ptr ([G/64+1][B], window-major) and `packed` (one 4-byte record per non-zero: 16-bit gene id |
 * bf16 value, CSR order; needs G <= 65536) */
int cmmvae_csr_tile_ptr(const int32_t* crow, const int32_t* col, const float* val, int B, int G, long long nnz,
                        int32_t* tile_ptr, void* packed, void* stream);
int cmmvae_csr_linear_fwd_tc(const void* packed, const int32_t* tile_ptr, int B, int G, int H,
                             const void* Wt_bf16, const float* bias, float* Y, void* stream);
/* [g_begin, g_end): gene rows of dWt computed by this launch (128-aligned; g_end <= 0 means G), so a caller
 * can hand finished row ranges to a collective while the rest is still being computed */
int cmmvae_csr_linear_bwd_w_tc(const void* packed, const int32_t* tile_ptr, int B, int G, int H,
                               const void* dY_bf16, float* dWt, int g_begin, int g_end, double* sumsq_out,
                               void* stream);
/* Same product for ONE gene shard [g_begin, g_end) of a data-parallel job (DDP gradient mean of
 * cmmvae_model.py:191-213, taken on the inputs instead of the outputs): `packed` / `dY_bf16` hold the cells of
 * ALL ranks (B = global cells), `tile_ptr_shard` holds only the windows of the shard (first row = window
 * g_begin/64, plus the closing row) and `dWt_shard` only the shard's rows (first row = gene g_begin).  The
 * result is the shard of the SUMMED gradient; no reduce-scatter of dWt is needed. */
int cmmvae_csr_linear_bwd_w_tc_shard(const void* packed, const int32_t* tile_ptr_shard, int B, int G, int H,
                                     const void* dY_bf16, float* dWt_shard, int g_begin, int g_end,
                                     double* sumsq_out, void* stream);

/* ---- K2/K3: BatchNorm1d(momentum, eps) + ReLU + Dropout, components.py:279-288 ------------- */
/* column statistics of Y[B,H]: mean[H], rstd[H] = 1/sqrt(biased var + eps); updates running
 * stats (unbiased var, momentum) when running_mean != NULL.  One launch: the last row-chunk block
 * of each column group finishes the statistics.  `scratch` = cmmvae_bn_stats_scratch_bytes(H) bytes that are
 * ZERO on entry; the call leaves them zero again (allocate + zero once, reuse for every call). */
size_t cmmvae_bn_stats_scratch_bytes(int H);
int cmmvae_bn_stats(const float* Y, int B, int H, float eps, float momentum, float* mean, float* rstd,
                    float* running_mean, float* running_var, double* scratch, void* stream);
/* out = drop(act(gamma*(Y-mean)*rstd+beta)); gamma==NULL -> no normalisation (plain act/dropout).
 * relu: 0/1.  dropout: keep-prob scaling 1/(1-p), Bernoulli mask from a counter-based hash of
 * (seed, element index) (recomputed in backward), or from `mask` (uint8 [B,H], injected) if
 * non-NULL.  Writes out_f32 and/or out_bf16 (either may be NULL). */
int cmmvae_bn_act_drop_fwd(const float* Y, int B, int H, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, int relu, float p_drop,
                           unsigned long long seed, const uint8_t* mask,
                           float* out_f32, void* out_bf16, void* stream);
/* backward of the above.  dOut f32 [B,H]; `out` = forward output (sign gives the ReLU mask).
 * Produces dY f32 (+bf16 copy), dgamma, dbeta (if gamma != NULL) and dbias = colsum(dY).
 * accumulate != 0: the three vector gradients are added to (the caller zeroed them: one fill per step
 * instead of three memsets per layer). */
int cmmvae_bn_act_drop_bwd(const float* dOut, const float* Y, const float* out, int B, int H,
                           const float* mean, const float* rstd, const float* gamma,
                           int relu, float p_drop, unsigned long long seed, const uint8_t* mask,
                           float* dY, void* dY_bf16, float* dgamma, float* dbeta, float* dbias,
                           int accumulate, void* stream);
/* eval-mode BN uses running stats: call bn_act_drop_fwd with mean=running_mean and
 * rstd computed by this helper. */
int cmmvae_rstd_from_var(const float* var, int H, float eps, float* rstd, void* stream);

/* ---- dense GEMMs (K4/K12), nn.Linear: components.py:276 ------------------------------------
 * C[M,N] = act( opA(A) * opB(B) + bias ) (+ C if accumulate).  transA: A stored [K,M];
 * transB=0: B stored [N,K] (nn.Linear weight layout), transB=1: B stored [K,N].
 * bias (f32 [N]) may be NULL.  relu: 0/1.  Outputs: C_f32 and/or C_bf16 (either may be NULL). */
/* fp32 CUDA-core path: any shape; inputs f32. */
int cmmvae_gemm_f32(const float* A, int lda, int transA, const float* Bm, int ldb, int transB,
                    int M, int N, int K, const float* bias, int relu, int accumulate,
                    float* C_f32, void* C_bf16, int ldc, void* stream);
/* tcgen05/TMEM/TMA path: inputs bf16, fp32 accumulate.  Requires 16-byte aligned bases and
 * leading dimensions that are multiples of 8 elements; any M, N, K (TMA zero-fills edges).
 * sumsq_out (double[1], optional): += sum of squares of the stored C, so a weight-gradient GEMM also
 * yields its contribution to the clip norm (log_gradient_norms, base_model.py:111-123). */
int cmmvae_gemm_bf16_tc(const void* A, int lda, int transA, const void* Bm, int ldb, int transB,
                        int M, int N, int K, const float* bias, int relu, int accumulate,
                        float* C_f32, void* C_bf16, int ldc, double* sumsq_out, void* stream);
/* same pipeline with f32 operands read as TF32 (tcgen05 kind::tf32, 10-bit mantissa): used for the small GEMMs
 * between the two gene-sized layers, where operand rounding -- not throughput -- decides gradient parity.
 * lda/ldb multiples of 4 elements. */
int cmmvae_gemm_tf32_tc(const float* A, int lda, int transA, const float* Bm, int ldb, int transB,
                        int M, int N, int K, const float* bias, int relu, int accumulate,
                        float* C_f32, void* C_bf16, int ldc, double* sumsq_out, void* stream);
/* column sums: out[N] (=|+=) sum_m X[m,n]  (bias gradients) */
int cmmvae_colsum(const void* X, int x_dtype, int M, int N, int ldx, float* out, int accumulate, void* stream);

/* ---- K8+K9(+Mean/Variance): reparameterisation + KL, components.py:790-801, vae.py:136-138 --
 * ML f32 [B,2Z] = [mu | logvar];  z = mu + eps*sqrt(exp(lv)+var_eps).
 * sums (double[3], zeroed by the call): [sum_b KL_b, sum mu, sum var].  z_f32/z_bf16 optional. */
int cmmvae_reparam_kl_fwd(const float* ML, const float* eps, int B, int Z, float var_eps,
                          float* z_f32, void* z_bf16, double* sums, void* stream);
/* dML[B,2Z] from dz (f32 [B,Z]) and the KL term: kl_scale = kl_weight / B. */
int cmmvae_reparam_kl_bwd(const float* ML, const float* eps, const float* dz, int B, int Z, float var_eps,
                          float kl_scale, float* dML, void* dML_bf16, void* stream);

/* ---- K5-K7: expert-decoder output + ReLU + sum-MSE against the CSR batch --------------------
 * replaces Linear(H->G)+ReLU (components.py:840,857), x.to_dense() (cmmvae_model.py:162-163) and
 * F.mse_loss(reduction='sum') (vae.py:143).
 * unfused form (f32 logits already materialised, used by the fp32 path and by validation):
 * xhat = relu(logits) in place (optional), loss_sum (double[1], zeroed by the call) =
 * sum (xhat - x)^2, dlogits = 2 (xhat - x) 1[logit>0] (f32 and/or bf16 with leading dim ldd). */
int cmmvae_mse_relu_csr(float* logits, int ldl, int B, int G, const int32_t* crow, const int32_t* col,
                        const float* val, int write_xhat, float* dlogits_f32, void* dlogits_bf16, int ldd,
                        double* loss_sum, void* stream);
/* fused form: logits tile stays in TMEM; epilogue adds bias, applies ReLU, reduces the loss against
 * the CSR entries of the tile and emits dlogits bf16 [B,ldd]; xhat never reaches HBM.
 * h bf16 [B,H] (ldh), Wout bf16 [G,H] (ldw), bout f32 [G].  `workspace` (device, size from
 * cmmvae_decoder_mse_fused_workspace_bytes) receives the per-(cell, 64-gene window) CSR pointer table when
 * `tile_ptr` is NULL; pass the table built by cmmvae_csr_tile_ptr as `tile_ptr` to share it.
 * loss_sum (double[1]) is zeroed by the call. */
size_t cmmvae_decoder_mse_fused_workspace_bytes(int B, int G);
int cmmvae_decoder_mse_fused(const void* h, int ldh, const void* Wout, int ldw, const float* bout,
                             int B, int G, int H, const int32_t* crow, const int32_t* col, const float* val,
                             const int32_t* tile_ptr, void* dlogits_bf16, int ldd, double* loss_sum,
                             void* workspace, void* stream);

/* ---- K11: adversary heads, CrossEntropyLoss(reduction='sum') (cmmvae_model.py:54,85) --------
 * logits f32 [B,C] (ldl), labels int64[B]; loss_sum double[1] += ; dlogits = scale*(softmax-onehot). */
int cmmvae_softmax_ce_sum(const float* logits, int ldl, int B, int C, const long long* labels,
                          float scale, float* dlogits, int ldd, double* loss_sum, void* stream);

/* ---- K13-K15: grad-norm, clip-by-norm, Adam (base_model.py:111-123, cmmvae_model.py:203-213,
 * 306-319; torch.optim.Adam with L2 weight decay) --------------------------------------------
 * norm_sq (double[1]) += sum g^2 over n elements (caller zeroes it). */
int cmmvae_sumsq(const float* g, long long n, double* norm_sq, void* stream);
/* coef = max_norm>0 ? min(1, max_norm/(sqrt(*norm_sq)+1e-6)) : 1, read ON DEVICE (no host sync);
 * g' = coef*grad_scale*g + wd*p; m,v,p updated; bf16 shadow written if non-NULL.
 * bc1 = 1-beta1^t, bc2 = 1-beta2^t computed by the host. */
int cmmvae_clip_adam(float* p, const float* g, float* m, float* v, void* p_bf16, long long n,
                     const double* norm_sq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                     float eps, float wd, float bc1, float bc2, void* stream);
/* Same update launched as many short-lived 128-thread CTAs: meant for a low-priority stream that runs the
 * output-layer update underneath the next step's forward pass (engine.pipeline_optimizer). */
int cmmvae_clip_adam_bg(float* p, const float* g, float* m, float* v, void* p_bf16, long long n,
                        const double* norm_sq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                        float eps, float wd, float bc1, float bc2, void* stream);

/* ---- small utilities ------------------------------------------------------------------------ */
int cmmvae_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);
/* dst[c,r] = src[r,c] for a row-major [R,C] f32/bf16 matrix */
int cmmvae_transpose(const void* src, void* dst, int dtype, int R, int C, int lds, int ldd, void* stream);
/* a[i] += alpha * b[i] */
int cmmvae_axpy(float* a, const float* b, float alpha, long long n, void* stream);


/* ---- output discriminator on the reconstruction (BASELINE config 4; MLP of runners/meta_discriminators.py:33-49:
 * Linear(G,128) Sigmoid Linear(128,64) Sigmoid Linear(64,1) Sigmoid, binary_cross_entropy(mean), 112-148) ------
 * xhat is never in HBM; it follows from the decoder's dlogits (bf16) and the CSR batch:
 *   xhat = dlogits/2 + x where dlogits != 0, else 0   =>   xhat W^T = 1/2 dlogits W^T + Xm W^T,
 * Xm = the batch restricted to entries with non-zero dlogits.  val_masked[i] = dlogits[row(i), col[i]] != 0 ?
 * val[i] : 0; the two products then run on the GEMM / SpMM entry points above. */
int cmmvae_mask_vals_by_dl(const int32_t* crow, const int32_t* col, const float* val, int B,
                           const void* dlogits_bf16, int ldd, float* val_masked, void* stream);
/* the same over rows given as [row_begin[b], row_end[b]) (the data-parallel route's received slabs) */
int cmmvae_mask_vals_by_dl_rows(const int32_t* row_begin, const int32_t* row_end, const int32_t* col,
                                const float* val, int B, const void* dlogits_bf16, int ldd, float* val_masked,
                                void* stream);
int cmmvae_sigmoid_fwd(const float* x, long long n, float* out_f32, void* out_bf16, void* stream);
/* dx = dout * out * (1 - out) */
int cmmvae_sigmoid_bwd(const float* dout, const float* out, long long n, float* dx, void* dx_bf16, void* stream);
/* p = sigmoid(a[B]); loss (double[1], zeroed by the call) = mean BCE(p, label) with torch's log clamp at -100;
 * da = (p - label) / B (gradient of the mean BCE through the sigmoid); p_out optional */
int cmmvae_bce_sigmoid(const float* a, int B, float label, float* p_out, float* da, double* loss, void* stream);

/* ---- launches that can be REPLAYED from a captured CUDA graph ---------------------------------------------------
 * The per-step scalars a replayed launch cannot carry by value live in device memory instead: the batch's nnz is
 * read from crow[B]; the dropout seed is *seed_base + seed; the KL weight is *kl_weight (kl_scale then = 1/B); Adam's
 * bias corrections are bc[0] = 1 - beta1^t, bc[1] = 1 - beta2^t.  Otherwise identical to the functions above. */
int cmmvae_csr_tile_ptr_dyn(const int32_t* crow, const int32_t* col, const float* val, int B, int G, long long cap,
                            int32_t* tile_ptr, void* packed, void* stream);
int cmmvae_bn_act_drop_fwd_dyn(const float* Y, int B, int H, const float* mean, const float* rstd,
                               const float* gamma, const float* beta, int relu, float p_drop,
                               unsigned long long seed, const unsigned long long* seed_base, const uint8_t* mask,
                               float* out_f32, void* out_bf16, void* stream);
int cmmvae_bn_act_drop_bwd_dyn(const float* dOut, const float* Y, const float* out, int B, int H,
                               const float* mean, const float* rstd, const float* gamma,
                               int relu, float p_drop, unsigned long long seed, const unsigned long long* seed_base,
                               const uint8_t* mask, float* dY, void* dY_bf16, float* dgamma, float* dbeta,
                               float* dbias, int accumulate, void* stream);
int cmmvae_reparam_kl_bwd_dyn(const float* ML, const float* eps, const float* dz, int B, int Z, float var_eps,
                              float kl_scale, const float* kl_weight, float* dML, void* dML_bf16, void* stream);
int cmmvae_clip_adam_dyn(float* p, const float* g, float* m, float* v, void* p_bf16, long long n,
                         const double* norm_sq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                         float eps, float wd, const float* bc, int background, void* stream);

/* ---- data parallel over NVLink peer memory (DDP gradient mean of cmmvae_model.py:191-213, taken on a
 * gene-sharded first / last layer) ---------------------------------------------------------
 * No reference counterpart (the reference relies on Lightning DDP = NCCL all-reduce of 500 MB per step).  Every
 * rank owns symmetric buffers mapped into all peers (CUDA IPC, set up by the host side); producers store into the
 * consumers' buffers and raise flags there, consumers spin on local flags.  `route` / `dst_slots` / `peer_flags`
 * are HOST arrays of n device pointers (one per rank, own rank included). */
/* same product as cmmvae_csr_linear_fwd_tc, no bias; output row r is stored at row r % route_rows of
 * route[r / route_rows] (partial sums of a gene shard for the cells of ALL ranks, delivered to their owners) */
/* n_split > 1: the gene range is cut into n_split pieces (more CTAs when few cell tiles exist); piece z is stored
 * split_stride elements behind route[..] -- one more slab for the owner to sum, no atomics */
int cmmvae_csr_linear_fwd_tc_routed(const void* packed, const int32_t* tile_ptr, int B, int G, int H,
                                    const void* Wt_bf16, float* const* route, int n_route, int route_rows,
                                    int n_split, long long split_stride, void* stream);
/* C[M,N] (f32) = opA(A) opB(B) with the same row routing (dh partials of a gene shard -> owners of the cells) */
int cmmvae_gemm_bf16_tc_routed(const void* A, int lda, int transA, const void* Bm, int ldb, int transB,
                               int M, int N, int K, int ldc, float* const* route, int n_route, int route_rows,
                               int n_split, long long split_stride, void* stream);
/* fused decoder with one loss sum per block of `loss_rows` cells (loss_sums double[ceil(B/loss_rows)], zeroed by
 * the call): a gene-sharded rank sees the cells of all ranks and reports each rank's share separately */
int cmmvae_decoder_mse_fused_blocks(const void* h, int ldh, const void* Wout, int ldw, const float* bout,
                                    int B, int G, int H, const int32_t* crow, const int32_t* col, const float* val,
                                    const int32_t* tile_ptr, void* dlogits_bf16, int ldd, double* loss_sums,
                                    int loss_rows, void* workspace, void* stream);
/* copy src[nbytes] (16-byte multiple) into dst_slots[i] on every rank i, then store `step` to peer_flags[i]
 * (system-scope release; peer_flags == NULL: no flags, for all but the last part of a multi-part push).
 * `ticket`: zeroed device uint32 owned by the caller (last-block detection).  step_dev (device uint32, optional, in
 * all three flag functions): the step number is read from device memory instead, so that the launch can be replayed
 * from a captured CUDA graph. */
int cmmvae_peer_push(const void* src, long long nbytes, void* const* dst_slots, void* const* peer_flags,
                     int n_peers, unsigned int step, const unsigned int* step_dev, unsigned int* ticket, void* stream);
/* flags only: after a kernel whose epilogue already stored to the peers (the routed kernels above) */
int cmmvae_peer_signal(void* const* peer_flags, int n_peers, unsigned int step, const unsigned int* step_dev,
                       void* stream);
/* block the stream until local_flags[0..n_peers) (uint32, this rank's memory) have all reached `step` */
int cmmvae_peer_wait(const void* local_flags, int n_peers, unsigned int step, const unsigned int* step_dev,
                     void* stream);
/* out[i] = sum_s slabs[s * slab_stride + i] (+ bias[i % H]), i < n; f32 and/or bf16 output */
int cmmvae_slab_sum(const float* slabs, int n_slabs, long long slab_stride, long long n, const float* bias, int H,
                    float* out_f32, void* out_bf16, void* stream);
/* All-to-all of this rank's batch by gene shard: rows are cut at the shard boundaries q * per (rows are sorted, so
 * piece q is a contiguous range) and piece q is stored into rank q's buffer: dst_crow[q] int32[B+1] (offsets of
 * the pieces, starting at 0), dst_col[q] / dst_val[q] [cap] with columns REBASED to the shard (col - q * per).
 * info[0] = largest piece, info[1] = 1 if a piece exceeds cap (its tail rows then arrive empty -- the caller must
 * treat that as an error).  cnt/start/offs: int32[B * n_dst] scratch.  dst_* are HOST arrays of n_dst device
 * (peer) pointers.  Bit-exact with the rows' sorted, duplicate-free column lists. */
int cmmvae_csr_scatter_shards(const int32_t* crow, const int32_t* col, const float* val, int B, int n_dst, int per,
                              int cap, void* const* dst_crow, void* const* dst_col, void* const* dst_val,
                              int32_t* cnt, int32_t* start, int32_t* offs, int32_t* info, void* stream);
/* n_src received slabs, `slab_bytes` apart, each starting with its crow int32[B+1]: row_begin / row_end
 * [B * n_src] = positions of every row in ONE int32/f32 array spanning all slabs (position = slab * slab_bytes/4
 * + offset inside the slab's col / val section) */
int cmmvae_slab_rows(const void* slabs, long long slab_bytes, int B, int n_src, int32_t* row_begin,
                     int32_t* row_end, void* stream);
/* cmmvae_csr_tile_ptr for rows given as (begin, end) position arrays instead of one crow array; n_records =
 * positions covered by col / val (records of all positions are packed, gaps included) */
int cmmvae_csr_tile_ptr_rows(const int32_t* row_begin, const int32_t* row_end, const int32_t* col, const float* val,
                             int B, int G, long long n_records, int32_t* tile_ptr, void* packed, void* stream);
/* slabs: n_src rows of `stride` doubles = [loss share of rank 0..n_src-1 | sumsq of the source's shard | ...];
 * out_recon = sum_s slab_s[rank]; out_norm += sum_s slab_s[n_src] */
int cmmvae_dp_scalars(const double* slabs, int n_src, int stride, int rank, double* out_recon, double* out_norm,
                      void* stream);

/* ---- conditional layers on the latent ------------------------------------------------------
 * ConditionalLayer.forward (components.py:369-413) sends every cell through the FCBlock of its metadata value;
 * ConditionalLayers.forward (components.py:581-631) chains the batch keys or concatenates their outputs.  Blocks are
 * the shipped one-layer kind (configs/model/human_only.yaml:61-68): Linear(Zin, Zout) [+ LayerNorm without affine,
 * eps 1e-5] [+ ReLU].  All blocks live in one parameter bank: slot s at params + s * slot_stride = [W (Zout x Zin,
 * torch layout) | b (Zout)]; grads / m / v alike.  The host sorts the rows of the batch by value:
 *   tiles  int32[n_tiles][4] = (slot, start into rows, count <= 32, index c of the batch key)
 *   rows   int32: cell indices, grouped as the tiles say
 * fwd:  out[row, out_col[c] + j] = act(LN(x[row, :] W_s^T + b_s)) (fp32, optional bf16 copy), pre = LN output,
 *       rstd[c * B + row]
 * bwd:  grads (pre-zeroed for the slots present) += dW_s, db_s;  dx[row, dx_col[c] + k] = (dy W_s)[row, k]
 * zero_grads / sumsq / adam run over the slots PRESENT in the batch only: torch.optim.Adam skips parameters whose
 * grad is None (unused modules after zero_grad(set_to_none=True)), so every slot carries its own step count
 * (steps[slot], advanced by cond_adam); sumsq ADDS the present slots' share to the group's clip norm. */
int cmmvae_cond_fwd(const float* params, long long slot_stride, int Zin, int Zout, const int32_t* tiles, int n_tiles,
                    const int32_t* rows, const float* x, int ldx, float* out, void* out_bf16, float* pre, int ldo,
                    const int32_t* out_col, float* rstd, int B, int layer_norm, int relu, void* stream);
int cmmvae_cond_bwd(const float* params, float* grads, long long slot_stride, int Zin, int Zout, const int32_t* tiles,
                    int n_tiles, const int32_t* rows, const float* x, int ldx, const float* dout, const float* pre,
                    int ldo, const int32_t* out_col, const float* rstd, int B, float* dx, int lddx,
                    const int32_t* dx_col, int layer_norm, int relu, void* stream);
int cmmvae_cond_zero_grads(float* grads, long long slot_stride, const int32_t* present, int n_present, void* stream);
int cmmvae_cond_sumsq(const float* grads, long long slot_stride, const int32_t* present, int n_present, double* out,
                      void* stream);
int cmmvae_cond_adam(float* params, const float* grads, float* m, float* v, long long slot_stride,
                     const int32_t* present, int n_present, int32_t* steps, const double* norm_sq, float max_norm,
                     float grad_scale, float lr, double beta1, double beta2, float eps, float weight_decay,
                     void* stream);
/* dz[b, j] = sum_k dcat[b, k*Z + j]: the gradients of the n parallel branches, which all read z, add up */
int cmmvae_fold_cols(const float* dcat, int B, int Z, int n, float* dz, void* stream);

/* ---- host -> HBM feed of CSR batches (batch format of cellxgene_datapipe.py:169-193) ---------
 * HOST function (all pointers are host pointers): rows [lo, hi) of a CSR chunk -- what scipy's chunk[lo:hi]
 * yields in SparseCSRMatrixBatcherDataPipe -- written into a (pinned) staging block: crow int32 rebased to 0,
 * col as uint16 (col_u16 != 0, needs n_genes <= 65536) or int32, val fp32 bit for bit.  indptr / indices are
 * int32 or int64 (width in bytes).  Returns the number of non-zeros, or < 0 (bad arguments / gene id out of
 * range).  Thread safe; releases no locks, allocates nothing. */
long long cmmvae_host_slice_rows(const void* indptr, int indptr_width, const void* indices, int indices_width,
                                 const float* data, long long lo, long long hi, long long n_genes,
                                 int32_t* crow_out, void* col_out, int col_u16, float* val_out);
/* device: dst int32[n] = src uint16[n] (gene ids shipped narrow over PCIe, widened once they are in HBM) */
int cmmvae_widen_u16_i32(const void* src_u16, int32_t* dst, long long n, void* stream);
/* Zero-copy route for chunks that serve many batches (the reference cuts every ~100k-cell .npz chunk into
 * consecutive row ranges, cellxgene_datapipe.py:169-193): page-lock the chunk's indices / data arrays IN PLACE
 * once (cudaHostRegister), then every batch is two DMA transfers straight out of them -- no host packing.
 * register / unregister: 0 on success, < 0 with cmmvae_last_error() set (e.g. the lock limit of the process). */
int cmmvae_host_register(const void* host_ptr, long long nbytes);
int cmmvae_host_unregister(const void* host_ptr);
/* device -> device copy by a kernel (16-byte aligned buffers): does not queue behind H2D transfers on a copy engine */
int cmmvae_copy_bytes(void* dst, const void* src, long long nbytes, void* stream);
/* cudaMemcpyAsync host -> device on `stream` (asynchronous when the host range is page-locked) */
int cmmvae_h2d_async(void* dst_dev, const void* src_host, long long nbytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
