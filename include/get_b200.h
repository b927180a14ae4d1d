/*
 * get_b200 — C-ABI of the B200-native GET hot path (libget_b200.so).
 *
 * Every entry point is `extern "C"`, takes raw DEVICE pointers + sizes + a `cudaStream_t` passed as
 * `void*`, allocates nothing, never synchronises, and returns 0 on success or a non-zero status
 * (negative = invalid argument, positive = cudaError_t of the launch). `get_b200_last_error()` returns a
 * thread-local message for the last non-zero status. The Python host (`get_b200/_lib.py`) raises
 * RuntimeError on non-zero status.
 *
 * The reference (CRIPAC-DIG/GET) has no native code; each function below names the reference Python call
 * site(s) it replaces (paths relative to the reference root):
 *   wrapper.py  = Models/BiDAF/wrapper.py
 *   tba.py      = thirdparty/two_branches_attention.py
 *   sa.py       = thirdparty/self_attention.py
 *   bfm.py      = Models/FCWithEvidences/basic_fc_model.py
 *   gbss.py     = Models/FCWithEvidences/graph_based_semantic_structure.py
 *
 * All matrices are fp32 row-major unless stated. "Graph" = one claim or evidence word graph: N node slots,
 * dense normalised adjacency (N x N), node features (N x H).
 */
#ifndef GET_B200_H
#define GET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GET_B200_ABI_VERSION 2

/* ------------------------------------------------------------------------------------------------
 * Dense contraction with fused epilogues, exact fp32 on the CUDA cores: the small / odd `Linear`s (output MLP, per-claim
 * projections, head-mix weight gradients, op-level surfaces with sizes the tensor-core path does not take). The large
 * contractions of GGNN / attention run through get_gemm_bp below.
 * Replaces: wrapper.py:191,194-204 (seven nn.Linear per GGNN), tba.py:140-141, sa.py:89-90,
 *           gbss.py:121 (output MLP), gbss.py:100,150 (embedding gather fused into the A operand),
 *           wrapper.py:189-190 (nn.Dropout fused into the A operand), and their autograd backward.
 *
 *   acc[m,n] = sum over segments s < nseg of  sum_{k < K[s]}  opA_s(m,k) * opB_s(n,k)
 *
 * operand element access (row-major storage with leading dimension ld):
 *   trans == 0 :  op(i,k) = ptr[row(i) * ld + k]        (contiguous along the contraction index)
 *   trans == 1 :  op(i,k) = ptr[row(k) * ld + i]        (contiguous along the output index)
 *   row(r) = rowidx ? rowidx[r] : r                      (row gather; A operand only)
 * A-operand dropout (segment-wise, when drop_p > 0): element at logical (row r, col c) of the stored
 * matrix with `drop_cols` columns is multiplied by keep(seed, r*drop_cols + c) / (1 - p); see
 * get_dropout_mask_f32 for the exact keep() function.
 * ---------------------------------------------------------------------------------------------- */
#define GET_GEMM_MAX_SEG 3

typedef struct get_gemm_operand {
  const float* ptr;
  int64_t ld;
  int32_t trans;
  int32_t _pad;
  const int64_t* rowidx; /* optional gather (A only), NULL otherwise */
} get_gemm_operand;

enum get_gemm_epilogue {
  GET_EPI_STORE = 0,        /* C = v                                    v = alpha*acc + bias0[n] + bias1[n] (+ C if accumulate) */
  GET_EPI_SIGMOID = 1,      /* C = s = sigmoid(v); if out1: out1 = s * aux0                     (z / r gates, wrapper.py:194-200) */
  GET_EPI_TANH_BLEND = 2,   /* h = tanh(v); if out1: out1 = h;  C = h*aux0 + aux1*(1-aux0)      (aux0 = z, aux1 = x; wrapper.py:202-206) */
  GET_EPI_TANH_ROWGROUP = 3,/* C = tanh(v + aux0[(m / group_rows) * ld_aux0 + n])               (tba.py:137-140 with the left half hoisted) */
  GET_EPI_DGATE_R = 4,      /* g = v; C = g*aux0*aux1*(1-aux1) (aux0 = x, aux1 = r); out1 += g*aux1   (GGNN backward) */
  GET_EPI_DROPOUT_OUT = 5,  /* C = v * keep(seed_out, m*N + n)/(1-p_out)  (+ C if accumulate)    (dX through nn.Dropout) */
  GET_EPI_TANH = 6          /* C = tanh(v) */
};

typedef struct get_gemm_desc {
  get_gemm_operand A[GET_GEMM_MAX_SEG];
  get_gemm_operand B[GET_GEMM_MAX_SEG];
  int32_t K[GET_GEMM_MAX_SEG];
  int32_t nseg;
  int32_t M, N;
  float* C;
  int64_t ldc;
  float alpha;
  int32_t accumulate;      /* C += ... (STORE / DROPOUT_OUT only) */
  int32_t epilogue;        /* enum get_gemm_epilogue */
  const float* bias0;      /* [N] or NULL */
  const float* bias1;      /* [N] or NULL */
  const float* aux0; int64_t ld_aux0;
  const float* aux1; int64_t ld_aux1;
  float* out1; int64_t ld_out1;
  int32_t group_rows;      /* TANH_ROWGROUP */
  /* A-operand dropout, segment 0 only */
  float drop_p; uint32_t drop_seed; int32_t drop_cols; int32_t _pad0;
  /* epilogue dropout (DROPOUT_OUT) */
  float drop_out_p; uint32_t drop_out_seed;
  /* split-K: if split_k > 1, `workspace` must hold split_k*M*N floats; partial sums are reduced in a fixed
   * order by a second kernel (deterministic), then the epilogue is applied. */
  int32_t split_k; int32_t _pad1;
  float* workspace;
} get_gemm_desc;

int get_gemm_f32(const get_gemm_desc* desc, void* stream);

/* Number of kernels get_gemm_f32 will launch for this descriptor (1, or 2 with split-K). */
int get_gemm_f32_launches(const get_gemm_desc* desc);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core contraction on bf16 PLANES (tcgen05.mma.kind::f16, TMA-fed, persistent) -- the main path of every large
 * `Linear` of GGNN / attention and of their backward passes (same reference call sites as get_gemm_f32).
 *
 * A fp32 value v is carried as up to three bf16 planes p0 = bf16(v), p1 = bf16(v - p0), p2 = bf16(v - p0 - p1).
 * A plane tensor is bf16 [planes][rows][ld] (ld % 8 == 0, plane_stride % 8 == 0, 16-byte aligned). Contractions never
 * read beyond the logical width of an operand (the TMA boxes are clipped there), so padding columns carry no contract
 * except one: a producer asked for `pad_one` writes 1.0 in column `width` (and zeros up to the next multiple of 8) -- the
 * "ones column" through which a weight-gradient contraction also yields the bias gradient.
 * Planes are produced by the kernels that produce the activation (GEMM epilogues, graph kernels, element-wise
 * kernels) or by get_to_planes_bf16; weights are packed once per optimizer step by get_pack_planes_multi.
 *
 *   acc[m,n] = sum_s sum_{k<K[s]} A_s(m,k) * B_s(n,k)      with the plane products selected by `mode`:
 *     mode 1:  a0.b0                                        (plain bf16: the 1e-2 parity class, BASELINE configs[2])
 *     mode 2:  a0.b0 + a0.b1 + a1.b0                        (16-bit operands: relative error ~1e-5, fp32 parity class)
 *     mode 3:  a0.b0  |  a0.b1 + a1.b0 + a1.b1 + a0.b2 + a2.b0   (fp32-exact class; the small terms accumulate in
 *              their own TMEM accumulator and are added in the epilogue: the layer that feeds the GSL top-k)
 *   operand (trans == 0, "K-major"):  op(i,k) = plane[i*ld + k]      (activations as A, packed weights as B)
 *   operand (trans == 1, "MN-major"): op(i,k) = plane[k*ld + i]      (both operands of a weight gradient dW = dG^T X)
 * All segments share `trans` per side. split_k > 1: raw partial tiles go to `workspace` ([split][M][ws_ld] fp32,
 * ws_ld = get_gemm_bp_ws_ld(desc)) and the caller finishes with get_bp_splitk_reduce (fixed order: deterministic).
 * ---------------------------------------------------------------------------------------------- */
typedef struct get_bp_tensor {
  const void* ptr;        /* bf16 */
  int64_t ld;             /* elements per row */
  int64_t plane_stride;   /* elements between planes */
  int32_t planes;         /* planes available (>= what `mode` needs) */
  int32_t trans;
} get_bp_tensor;

enum get_bp_epilogue {
  GET_BPE_STORE = 0,        /* v = acc + bias[n] (+ C if accumulate), optional dropout-out mask; C = v; planes_out = v          */
  GET_BPE_ZR = 1,           /* fused z|r gates (wrapper.py:194-200): s = sigmoid(acc + bias[n]); g = n / zr_group_stride,
                               c = n % zr_group_stride (< zr_cols): g == 0: C[m,c] = s (z); g == 1: out1[m,c] = s (r),
                               planes_out[m,c] = s * aux0[m,c] (r*x)                                                          */
  GET_BPE_TANH_BLEND = 2,   /* h = tanh(acc + bias); out1 = h; C = planes_out = h*aux0 + aux1*(1-aux0) (aux0 = z, aux1 = x)   */
  GET_BPE_TANH_ROWGROUP = 3,/* C = tanh(acc + aux0[(m / group_rows) * ld_aux0 + n])                                           */
  GET_BPE_DGATE_R = 4,      /* g = acc; planes_out = g*aux0*aux1*(1-aux1) (aux0 = x, aux1 = r); out1 += g*aux1               */
  GET_BPE_TANH = 5          /* C = tanh(acc)                                                                                  */
};

typedef struct get_gemm_bp_desc {
  get_bp_tensor A[GET_GEMM_MAX_SEG];
  get_bp_tensor B[GET_GEMM_MAX_SEG];
  int32_t K[GET_GEMM_MAX_SEG];
  int32_t nseg;
  int32_t M, N;
  int32_t mode;             /* 1, 2 or 3 */
  int32_t tile_n;           /* CTA tile along N (multiple of 16, <= 256); 0 = library choice (get_gemm_bp_tile_n) */
  int32_t epilogue;         /* enum get_bp_epilogue */
  int32_t accumulate;       /* STORE only: C += v */
  float* C; int64_t ldc;    /* fp32 output or NULL */
  float* out1; int64_t ld_out1;
  const float* bias;        /* [N] or NULL */
  const float* aux0; int64_t ld_aux0;
  const float* aux1; int64_t ld_aux1;
  void* planes_out;         /* bf16 plane tensor written by the epilogue, or NULL */
  int64_t ld_planes_out, planes_out_stride;
  int32_t planes_out_n;     /* 1..3 */
  int32_t planes_out_pad_one;
  int32_t group_rows;       /* TANH_ROWGROUP */
  int32_t zr_group_stride, zr_cols;   /* ZR */
  float drop_out_p; uint32_t drop_out_seed;   /* STORE: C = v * keep(seed, m*N + n)/(1-p) when drop_out_p > 0 */
  int32_t split_k;
  int32_t kblock;           /* k elements per pipeline stage: 32 or 64; 0 = library choice */
  float* workspace; int64_t workspace_floats;
  /* TANH_BLEND only: row-dot by-product of the blended output (the `proj` of the GSL scorer GGNN(H -> 1),
   * wrapper.py:158,167,191, with the scorer's own nn.Dropout draw): partial sums per N tile and tile half,
   * rowdot_out[(n_tile*2 + half)*M + m] = sum_n dropout(out[m,n]; rowdot_p, rowdot_seed, index m*N + n) * rowdot_w[n];
   * get_gemm_bp_rowdot_parts(desc) partial vectors in total, summed (fixed order) by the fused GSL kernel. */
  const float* rowdot_w; float* rowdot_out;
  float rowdot_p; uint32_t rowdot_seed;
} get_gemm_bp_desc;

/* 0 = launched; negative = invalid / unsupported descriptor (the Python host then raises: there is no silent fallback) */
int get_gemm_bp(const get_gemm_bp_desc* desc, void* stream);
/* N tile the library picks for (M, N, mode) (what zr_group_stride must be a multiple of); row pitch of the split-K
 * workspace; number of k splits the launch will really use. */
int get_gemm_bp_tile_n(int M, int N, int mode);
int64_t get_gemm_bp_ws_ld(const get_gemm_bp_desc* desc);
int get_gemm_bp_splits(const get_gemm_bp_desc* desc);
int get_gemm_bp_rowdot_parts(const get_gemm_bp_desc* desc);   /* 2 * number of N tiles */

/* Fixed-order reduction of split-K partial tiles into up to GET_BP_MAX_DST destination blocks (the weight / bias
 * gradients, written straight into the flat gradient bucket): for every block b and (r, c) inside it
 *   dst_b[(r - row0) * ld + (c - col0)] = (accumulate ? dst_b[..] : 0) + sum_z ws[(z*M + r)*ws_ld + c]. */
#define GET_BP_MAX_DST 12
typedef struct get_bp_dst {
  float* dst; int64_t ld;
  int32_t row0, nrows, col0, ncols;
} get_bp_dst;
int get_bp_splitk_reduce(const float* workspace, int splits, int M, int64_t ws_ld, const get_bp_dst* dsts, int ndst,
                         int accumulate, void* stream);

/* fp32 (rows, cols) matrix with row stride ld_src -> bf16 planes [nplanes][rows][ld_out]; the pad columns from `cols`
 * up to the next multiple of 8 of cols (+1 with pad_one), at most ld_out, are written too (zeros; 1.0 at column `cols`
 * when pad_one). */
int get_to_planes_bf16(const float* src, int64_t ld_src, int rows, int cols, void* planes, int64_t ld_out,
                       int64_t plane_stride, int nplanes, int pad_one, void* stream);

/* Weight packing, MANY jobs in one launch (all weights of the model once per optimizer step). `jobs` lives in DEVICE
 * memory. kind 0: 3 bf16 planes of the logical (rows, cols) matrix src[r*ld_r + c*ld_c] into dst (bf16, row pitch
 * ld_out, plane pitch plane_stride; dst already points at the block's first element, so several matrices can be packed
 * side by side / stacked into one plane tensor whose padding was zeroed once). kind 1: fp32 vector
 * dst[i] = src[i] + (src2 ? src2[i] : 0), i < rows (fused bias vectors). Job j covers blocks
 * [first_block, first_block + nblocks) of the launch, nblocks = ceil(rows/32)*ceil(cols/32) (kind 0: one block per
 * 32 x 32 tile) or ceil(rows/1024) (kind 1). */
typedef struct get_pack_job {
  const float* src; const float* src2;
  int64_t ld_r, ld_c;
  int32_t rows, cols;
  void* dst; int64_t ld_out, plane_stride;
  int64_t first_block;
  int32_t kind, _pad;
} get_pack_job;
int get_pack_planes_multi(const get_pack_job* jobs, int n_jobs, int64_t total_blocks, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Neighbour aggregation  out[g] (+)= op(adj'[g]) @ x[g]   with adj'[i,j] = adj[i,j] * (keep[i] | keep[j]).
 * Replaces: wrapper.py:192 `adj.matmul(x)`; with keep != NULL also wrapper.py:221-225 (the dense mask is
 * never materialised); transpose=1 gives adj'^T @ x for the backward pass (SURVEY A.1 `A^T da`).
 * adj (G,N,N), x (G,N,H), out (G,N,H), keep (G,N) uint8 or NULL.
 * ---------------------------------------------------------------------------------------------- */
int get_graph_aggregate_f32(const float* adj, const float* x, const uint8_t* keep, float* out,
                            int G, int N, int H, int transpose, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused graph-structure-learning kernel (the metric's "GSL-GAT kernel"):
 *   s_p = drop_s(F) . wp ; s_a = adj @ s_p ; scalar GRU gates -> score          (wrapper.py:167, GGNN with out_features=1)
 *   keep = top-k(score), k = int(rate*N)                                        (wrapper.py:215-219)
 *   out  = (adj * (keep_i | keep_j)) @ drop_2(F)                                (wrapper.py:221-225 + :192 of feat_prop2)
 * F (G,N,H), adj (G,N,N), wp (H), gate (12) = {wz0,bz0,wz1,bz1,wr0,br0,wr1,br1,wh0,bh0,wh1,bh1},
 * score (G,N) or NULL, keep (G,N) uint8, out (G,N,H).
 * drop_s / drop_2 are the two independent nn.Dropout draws of word_scorer1 / feat_prop2 (p = 0 in eval).
 * Ties are broken towards the lower node index.
 * ---------------------------------------------------------------------------------------------- */
int get_gsl_fused_f32(const float* adj, const float* F, const float* wp, const float* gate,
                      int G, int N, int H, int k,
                      float drop_p, uint32_t seed_scorer, uint32_t seed_layer2,
                      float* score, uint8_t* keep, float* out, void* stream);

/* The two graph kernels with the output rows written as bf16 planes [nplanes][G*N][ld_p] for the next tensor-core
 * contraction (see get_gemm_bp; pad columns H..round_up(H,8)-1 written as zeros, 1.0 in column H when pad_one) and / or
 * as fp32 (`out` may be NULL when planes are given; accumulate needs `out`). */
int get_graph_aggregate_bp(const float* adj, const float* x, const uint8_t* keep, float* out, void* planes,
                           int64_t ld_p, int64_t plane_stride, int nplanes, int pad_one, int G, int N, int H,
                           int transpose, int accumulate, void* stream);
int get_gsl_fused_bp(const float* adj, const float* F, const float* wp, const float* gate, int G, int N, int H, int k,
                     float drop_p, uint32_t seed_scorer, uint32_t seed_layer2, float* score, uint8_t* keep,
                     float* out, void* planes, int64_t ld_p, int64_t plane_stride, int nplanes, void* stream);

/* out[m] = dropout(F[m,:]; drop_p, seed, index m*H + c) . w   (the `proj` of GGNN(H -> 1), wrapper.py:191). */
int get_rowdot_f32(const float* F, const float* w, int64_t M, int H, float drop_p, uint32_t seed, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Packed neighbour lists: the graph path of the model. The dense (G,N,N) adjacency (4-13 % dense) is read ONCE per
 * step; every aggregation of the step (wrapper.py:192 in both GGNN layers, the scorer's adj @ s_p, and the two adj^T
 * products of the backward pass) then walks per-graph CSR records, for adj (orientation 0) and adj^T (orientation 1):
 *   ent    : (2, G, ecap) entries of 8 bytes {int32 neighbour index, float weight}, ecap = get_neighbor_lists_entry_capacity(N)
 *            (N*N rounded up to even: no overflow case); a graph's entries are compact from the start of its block,
 *            rows in order, neighbours in increasing index.
 *   rowptr : (2, G, pitch) int32, pitch = get_neighbor_lists_rowptr_pitch(N); rowptr[i] .. rowptr[i+1] = entries of row i.
 *   used   : (2, G) int32: 1 + the highest neighbour index any list of graph g refers to -- feature rows at or beyond it
 *            (the pad nodes at the end of a text) are never gathered, so the graph kernels neither load nor mask them.
 * N <= 232. The gather entry points take the pointers of ONE orientation (ent + o*G*ecap, rowptr + o*G*pitch, used + o*G).
 * ---------------------------------------------------------------------------------------------- */
int get_neighbor_lists_rowptr_pitch(int N);
int get_neighbor_lists_entry_capacity(int N);
int get_build_neighbor_lists(const float* adj, int G, int N, void* ent, int32_t* rowptr, int32_t* used, void* stream);
/* out[g,i,:] (+)= sum_e w_e * x[g, j_e, :], edges between two dropped nodes skipped when keep != NULL (same semantics as
 * get_graph_aggregate_bp; pass the records of orientation 1 for the transposed product; used may be NULL). out and / or
 * bf16 planes. */
int get_graph_gather(const void* ent, const int32_t* rowptr, const int32_t* used, const float* x, const uint8_t* keep,
                     float* out, void* planes, int64_t ld_p, int64_t plane_stride, int nplanes, int pad_one, int G, int N,
                     int H, int accumulate, void* stream);
/* The fused GSL kernel on lists: same outputs as get_gsl_fused_bp, with the scorer projection s_p = dropout_s(F) . w_p
 * PRECOMPUTED (n_sp partial vectors of G*N floats, summed in order): a by-product of the epilogue of the contraction that
 * wrote F (get_gemm_bp, rowdot_out), or of get_rowdot_f32. drop_p / seed_layer2 = the nn.Dropout draw of feat_prop2 (the
 * scorer's draw went into s_p). */
int get_gsl_gather(const void* ent, const int32_t* rowptr, const int32_t* used, const float* F, const float* sp_parts, int n_sp,
                   const float* gate, int G, int N, int H, int k, float drop_p, uint32_t seed_layer2, float* score,
                   uint8_t* keep, float* out, void* planes, int64_t ld_p, int64_t plane_stride, int nplanes, void* stream);

/* GSL.forward as a stand-alone op (wrapper.py:215-227): adj_out = adj * mask(top-k(score)). score (G,N). */
int get_gsl_mask_adj_f32(const float* adj, const float* score, int G, int N, int k,
                         float* adj_out, uint8_t* keep, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-head additive attention pooling (tail of ConcatNotEqualSelfAtt / MultiHeadSelfAttentionICLR2017Extend).
 * Replaces tba.py:141-147, sa.py:90-96: e = t @ W2^T; e[mask==0] = -inf; att = softmax over positions;
 * pooled[d,c] = sum_p right[p,d]*att[p,c].
 * t (G,P,H) = tanh(linear1(...)), right (G,P,Dr) with row stride ld_right, W2 (C,H), mask (G,P) uint8,
 * att (G,P,C), pooled (G, Dr*C) with row stride ld_pooled, element index d*C + c (gbss.py:191,220).
 * ---------------------------------------------------------------------------------------------- */
int get_att_pool_fwd_f32(const float* t, const float* right, int64_t ld_right, const float* W2,
                         const uint8_t* mask, int G, int P, int H, int Dr, int C,
                         float* att, float* pooled, int64_t ld_pooled, void* stream);

/* Backward of the above. d_pooled (G, Dr*C) row stride ld_dpooled; d_att (G,P,C) or NULL.
 * Writes de (G,P,C), du (G,P,H) = (de @ W2) * (1 - t^2), dright (G,P,Dr) row stride ld_dright
 * (= att @ d_pooled^T; overwritten unless accumulate), du_sum (G,H) = sum_p du. */
int get_att_pool_bwd_f32(const float* t, const float* right, int64_t ld_right, const float* W2,
                         const float* att, const float* d_pooled, int64_t ld_dpooled, const float* d_att,
                         int G, int P, int H, int Dr, int C,
                         float* de, float* du, float* du_sum, float* dright, int64_t ld_dright,
                         int accumulate, void* stream);

/* The same backward with du written as bf16 planes [nplanes][G*P][ld_dup] (operand of the tensor-core contractions
 * dright += du @ W1_right and dW1_right = du^T @ right, see get_gemm_bp) instead of fp32. */
int get_att_pool_bwd_bp(const float* t, const float* right, int64_t ld_right, const float* W2,
                        const float* att, const float* d_pooled, int64_t ld_dpooled, const float* d_att,
                        int G, int P, int H, int Dr, int C, float* de, void* du_planes, int64_t ld_dup,
                        int64_t plane_stride, int nplanes, float* du_sum, float* dright, int64_t ld_dright,
                        int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GGNN backward, element-wise stage: from dout, z, h, x  (all (M,H) contiguous)
 *   dhp = dout*z*(1-h^2) ; dzp = dout*(h-x)*z*(1-z) ; dx = dout*(1-z).
 * dhp / dzp are (M,H) views with row stride ld_g: columns of the (M,3H) gate-gradient buffer [dz'|dr'|dh'] whose
 * weight gradients are then single contractions; dx is (M,H) contiguous.
 * ---------------------------------------------------------------------------------------------- */
int get_ggnn_gate_bwd_f32(const float* dout, const float* z, const float* h, const float* x,
                          int M, int H, int64_t ld_g, float* dhp, float* dzp, float* dx, void* stream);

/* Column sums out[n] = sum_m a[m*ld + n] (bias gradients), deterministic two-stage reduction.
 * workspace: at least get_colsum_workspace_floats(M, N) floats. */
int get_colsum_f32(const float* a, int64_t ld, int M, int N, float* out, float* workspace, void* stream);
int64_t get_colsum_workspace_floats(int M, int N);

/* ------------------------------------------------------------------------------------------------
 * Segment helpers (bfm.py:80-121) driven by seg_of_row (B1,) int32 = claim index of every flattened
 * evidence and slot_of_row (B1,) int32 = claim*n + j:
 *   expand:       out[r,:]  = src[seg_of_row[r],:]                 (_pad_left_tensor)
 *   expand_bwd:   dsrc[c,:] = sum_{r: seg(r)=c} dout[r,:]          (rows of a claim are contiguous; offsets (B+1,))
 *   scatter_rows: out[slot_of_row[r],:] = src[r,:]                 (_pad_right_tensor, out pre-zeroed)
 *   gather_rows:  out[r,:] = src[slot_of_row[r],:]                 (its backward)
 * ---------------------------------------------------------------------------------------------- */
int get_rows_gather_f32(const float* src, int64_t ld_src, const int32_t* idx, int R, int W,
                        float* out, int64_t ld_out, void* stream);
int get_rows_scatter_f32(const float* src, int64_t ld_src, const int32_t* idx, int R, int W,
                         float* out, int64_t ld_out, void* stream);
int get_segment_sum_f32(const float* src, int64_t ld_src, const int32_t* offsets, int S, int W,
                        float* out, int64_t ld_out, void* stream);

/* The index vectors above from the per-claim evidence counts, in one launch (replaces the per-claim Python loops and
 * host syncs of bfm.py:86-90,113-118): offsets (B+1,) = exclusive prefix sums of evd_cnt (int64 or int32, B <= 8192),
 * seg_of_row / slot_of_row (B1,) with B1 = sum(evd_cnt) known on the host (shape of the flattened evidence tensor). */
int get_segments_i32(const void* evd_cnt, int cnt_is_int64, int B, int B1, int n, int32_t* seg_of_row,
                     int32_t* slot_of_row, int32_t* offsets, void* stream);

/* Attention masks from token ids (int64 or int32), one warp per row: mask[r] = (sum_j ids[r*W + j] >= 1).
 * W = 1 gives the word-level mask `doc >= 1` (gbss.py:98; ids are never negative), W = R the evidence-level mask
 * `sum(document, -1) >= 1` (gbss.py:215). */
int get_ids_mask_u8(const void* ids, int ids_is_int64, int64_t rows, int W, uint8_t* mask, void* stream);

/* Trainable source embeddings (base_model.py:184-188): out[r,:] = table[max(idx[r], 0), :] (the -1 padding id is looked
 * up as row 0, gbss.py:166-168); backward = the dense (V,E) table gradient, summed in index order by one block per
 * table row (deterministic; no atomics). */
int get_embedding_rows_fwd_f32(const float* table, int V, int E, const int64_t* idx, int R, float* out, int64_t ld_out,
                               void* stream);
int get_embedding_rows_bwd_f32(const float* g, int64_t ld_g, const int64_t* idx, int R, int V, int E, float* dtable,
                               int accumulate, void* stream);

/* Masked mean over the nodes of each claim graph (gbss.py:145-153):
 * out[g,:] = sum_i [ids[g,i] > 0] * h[g,i,:] / lens[g];  and its backward dh[g,i,:] = [ids>0]*dout[g,:]/lens[g]. */
int get_masked_mean_fwd_f32(const float* h, const int64_t* ids, const int64_t* lens, int G, int N, int H,
                            float* out, void* stream);
int get_masked_mean_bwd_f32(const float* dout, const int64_t* ids, const int64_t* lens, int G, int N, int H,
                            float* dh, void* stream);

/* Input of a GGNN projection materialised once: out[r,:] = dropout(src[idx ? idx[r] : r, :]) -- the embedding gather
 * (gbss.py:100,150) fused with nn.Dropout (wrapper.py:189-190); mask index r*W + c, keep() as below. idx int64 or NULL.
 * Forward projection and weight gradient then read the same (R, W) matrix through plain TMA tiles. */
int get_rows_gather_dropout_f32(const float* src, int64_t ld_src, const int64_t* idx, int R, int W, float p,
                                uint32_t seed, float* out, int64_t ld_out, void* stream);

/* Plane-emitting variants for the bf16-plane contraction (get_gemm_bp): the same computations with the result written
 * as bf16 planes [nplanes][rows][ld] (pad columns up to the next multiple of 8 written as zeros).
 * get_ggnn_gate_bwd_bp: dz' -> column block `col_z`, dh' -> column block `col_h` of the gate-gradient plane buffer
 * [dz' | dr' | dh'] with row pitch ld_g; dx stays fp32 (M,H).
 * get_rows_gather_dropout_bp: src_rows > 0 bounds-checks the gathered row ids (an id outside [0, src_rows) traps the
 * kernel, like the device assert of the reference's nn.Embedding). */
int get_ggnn_gate_bwd_bp(const float* dout, const float* z, const float* h, const float* x, int M, int H, void* dg,
                         int64_t ld_g, int64_t plane_stride, int nplanes, int col_z, int col_h, float* dx, void* stream);
int get_rows_gather_dropout_bp(const float* src, int64_t ld_src, int src_rows, const int64_t* idx, int R, int W, float p,
                               uint32_t seed, void* planes, int64_t ld_out, int64_t plane_stride, int nplanes,
                               void* stream);

/* Dropout salt: one device word per process, added to EVERY dropout seed inside the kernels (effective seed =
 * seed + salt mod 2^32). It is 0 unless set. A captured CUDA graph replays fixed kernel arguments; recording
 * get_dropout_salt_advance at the head of the graph (salt <- salt*1664525 + 1013904223) gives every replay fresh masks.
 * get_dropout_salt_get copies the word to the host (synchronises; tests). */
int get_dropout_salt_set(uint32_t value, void* stream);
int get_dropout_salt_advance(void* stream);
int get_dropout_salt_get(uint32_t* host_value);

/* The dropout keep-mask used by every fused dropout site, materialised (tests / externally supplied masks):
 * out[i] = keep(seed + salt, i) ? 1/(1-p) : 0. One 32-bit hash serves an aligned pair of elements, 16 bits each:
 * keep(s, i) = ((mix32(lo32(i>>1)*0x9E3779B1 + hi32(i>>1)*0x632BE5AB + s*0x85EBCA6B + 0x6A09E667) >> 16*(i&1)) & 0xFFFF)
 *              >= floor(p * 65536)   (csrc/common.cuh; host mirror get_b200/dropout.py). */
int get_dropout_mask_f32(float* out, int64_t numel, float p, uint32_t seed, void* stream);

/* Mean cross-entropy over claims + gradient w.r.t. logits (losses.py:29-32). logits (B,C), labels (B,) int64.
 * loss (1,), dlogits (B,C) = (softmax - onehot)/B. */
int get_cross_entropy_f32(const float* logits, const int64_t* labels, int B, int C,
                          float* loss, float* dlogits, void* stream);

/* Optimizer step of the reference fitter, torch.optim.Adam(lr, weight_decay) (declare_fitter.py:57-61), over FLAT
 * fp32 buffers of n elements (parameters, gradients, first / second moments): g' = g + wd*p; m = b1 m + (1-b1) g';
 * v = b2 v + (1-b2) g'^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps). `step` is a device counter (float),
 * advanced by one before use, so a captured CUDA graph replays a correct bias correction. */
int get_adam_flat_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, float* step, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Word-graph construction on the device (the host code interactions.py:334-351 `convert_text` + :11-18
 * `_laplacian_normalize`): tokens (G,T) int64 raw token ids, lengths (G,) int32 valid tokens per text.
 * nodes (G,N) int64 = distinct tokens of the first min(length, N) positions in first-occurrence order (0 elsewhere),
 * n_nodes (G,) int32, adj (G,N,N) fp32 = D^-1/2 A D^-1/2 of the sliding-window co-occurrence graph
 * (|i-j| <= window-1, self loops), evaluated in double like the reference and rounded once to fp32.
 * ---------------------------------------------------------------------------------------------- */
int get_build_word_graphs(const int64_t* tokens, const int32_t* lengths, int G, int T, int N, int window,
                          int64_t* nodes, float* adj, int32_t* n_nodes, void* stream);

/* Library info */
int get_b200_abi_version(void);
const char* get_b200_last_error(void);
/* Number of kernel launches issued through this library by the calling process so far (bench `gpu_launches`). */
int64_t get_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GET_B200_H */
