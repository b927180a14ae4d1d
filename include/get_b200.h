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

#define GET_B200_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------------
 * Dense contraction with fused epilogues (the `Linear`s of GGNN / attention / output MLP).
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
  /* Tensor-core path (tcgen05, 3xTF32 error-compensated: fp32-level accuracy), offered when tc_mode == 1; the library
   * takes it when the descriptor is eligible and otherwise falls back to the exact SIMT path:
   *  (a) persistent TMA-fed kernel: every A segment 16-byte aligned with ld % 4 == 0, no row gather, no A dropout, all
   *      segments with the same `trans`; B either given pre-split (below) or raw with the same `trans` in all segments.
   *      Operands with trans == 1 (both activations: weight gradients) are consumed MN-major, no transposition pass.
   *      split_k > 1 is honoured (k blocks of 32) through `workspace`.
   *  (b) register-staged kernel for gathered / dropped-out A operands: A contiguous along k (trans == 0, ld % 4 == 0,
   *      K % 4 == 0, K >= 32), no split-K, B pre-split.
   * Pre-split B: two k-contiguous (N, K[s]) row-major matrices B_hi[s] (tf32-rounded) and B_lo[s] (= B - B_hi) with
   * leading dimension ld_split[s] (see get_split_tf32_f32; a transposed B is simply split into a k-contiguous copy).
   * B[s] stays the original operand for the fallback. tc_n_tiles: CTA tiles along N (0 = auto).
   * tc_mode == 2: reduced-precision mode of kernel (a): ONE tf32 pass on the raw operands (no hi/lo split, B_lo unused),
   * relative error ~1e-3 -- for the 1e-2 parity class (BASELINE.json configs[2]) and never for the layer that feeds the
   * GSL top-k. */
  const float* B_hi[GET_GEMM_MAX_SEG];
  const float* B_lo[GET_GEMM_MAX_SEG];
  int64_t ld_split[GET_GEMM_MAX_SEG];
  int32_t tc_mode; int32_t tc_n_tiles;
} get_gemm_desc;

int get_gemm_f32(const get_gemm_desc* desc, void* stream);

/* 2 / 1 if the descriptor would run on the tcgen05 path (persistent TMA-fed / register-staged kernel), 0 if on the
 * SIMT path, <0 on invalid descriptors. */
int get_gemm_f32_uses_tc(const get_gemm_desc* desc);

/* Error-compensated TF32 split of a weight matrix for the tensor-core path:
 * hi = tf32_round_nearest(src), lo = tf32_round_nearest(src - hi). src is a logical (rows, cols) matrix addressed as
 * src[r*ld_r + c*ld_c] (so a transposed view can be split into a k-contiguous copy); hi/lo are (rows, cols)
 * row-major with leading dimension ld_out. */
int get_split_tf32_f32(const float* src, int64_t ld_r, int64_t ld_c, int rows, int cols,
                       float* hi, float* lo, int64_t ld_out, void* stream);

/* The same split for MANY matrices in one launch (all weights of the model once per optimizer step). `jobs` is an array
 * in DEVICE memory; job j covers blocks [first_block, first_block + ceil(rows*cols/256)) of the launch, first_block
 * ascending from 0; total_blocks = sum over jobs. */
typedef struct get_split_job {
  const float* src; int64_t ld_r, ld_c;
  int32_t rows, cols;
  float* hi; float* lo; int64_t ld_out;
  int64_t first_block;
} get_split_job;
int get_split_tf32_multi_f32(const get_split_job* jobs, int n_jobs, int64_t total_blocks, void* stream);

/* Number of kernels get_gemm_f32 will launch for this descriptor (1, or 2 with split-K). */
int get_gemm_f32_launches(const get_gemm_desc* desc);

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
