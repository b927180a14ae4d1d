// FP32 SIMT contraction with fused epilogues: the exact-fp32 path for every `Linear` of the GET hot path
// (reference Models/BiDAF/wrapper.py:191,194-204; thirdparty/two_branches_attention.py:140-141;
// Models/FCWithEvidences/graph_based_semantic_structure.py:121) and for their backward passes.
//
// acc[m,n] = sum_s sum_k A_s(m,k) * B_s(n,k); operands are row-major with either index contiguous
// (see include/get_b200.h). 128x64x16 CTA tile, 256 threads, 8x4 register tile per thread, double-buffered
// shared memory with register prefetch, 128-bit global loads whenever the operand allows it. Row gather
// (embedding lookup, gbss.py:100) and dropout (wrapper.py:189-190) are applied while the A tile is loaded.
// Split-K writes per-split partial tiles and a second kernel reduces them in a fixed order (deterministic).
#include "gemm_common.cuh"
#include <string.h>

namespace getb {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
constexpr int NTHREADS = (BM / TM) * (BN / TN);  // 256
constexpr int PAD = 4;
static_assert(NTHREADS == 256, "tile config");

// ---- tile loaders -------------------------------------------------------------------------------
// A tile of ROWS x BK elements is fetched as NV "quads" per thread. Quad f = tid + j*NTHREADS:
//   trans == 0 : row = f / (BK/4), kq = f % (BK/4); elements (row, 4*kq + e)    (contiguous along k)
//   trans == 1 : k   = f / (ROWS/4), iq = f % (ROWS/4); elements (4*iq + e, k)   (contiguous along rows)
template <int ROWS, bool IS_A>
__device__ __forceinline__ void load_tile(const GemmParams& p, const GemmOp& op, int row0, int nrows, int k0, int kseg,
                                          float (&reg)[ROWS * BK / NTHREADS], int tid) {
  constexpr int NV = ROWS * BK / NTHREADS / 4;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int f = tid + j * NTHREADS;
    float* r = &reg[j * 4];
    if (op.trans == 0) {
      const int row = row0 + f / (BK / 4);
      const int k = k0 + (f % (BK / 4)) * 4;
      if (row < nrows && k < kseg) {
        const int64_t srow = (IS_A && op.rowidx) ? op.rowidx[row] : (int64_t)row;
        const float* src = op.ptr + srow * op.ld + k;
        if (op.vec) {
          float4 t = __ldg(reinterpret_cast<const float4*>(src));
          r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) r[e] = (k + e < kseg) ? __ldg(src + e) : 0.f;
        }
        if (IS_A && p.drop_thr) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            bool keep = drop_keep(p.drop_seed + __ldg(p.salt), (uint64_t)row * (uint64_t)p.drop_cols + (uint64_t)(k + e), p.drop_thr);
            r[e] = keep ? r[e] * p.drop_scale : 0.f;
          }
        }
      } else {
        r[0] = r[1] = r[2] = r[3] = 0.f;
      }
    } else {
      const int k = k0 + f / (ROWS / 4);
      const int i = row0 + (f % (ROWS / 4)) * 4;
      if (k < kseg && i < nrows) {
        const int64_t srow = (IS_A && op.rowidx) ? op.rowidx[k] : (int64_t)k;
        const float* src = op.ptr + srow * op.ld + i;
        if (op.vec) {
          float4 t = __ldg(reinterpret_cast<const float4*>(src));
          r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) r[e] = (i + e < nrows) ? __ldg(src + e) : 0.f;
        }
        if (IS_A && p.drop_thr) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            bool keep = drop_keep(p.drop_seed + __ldg(p.salt), (uint64_t)k * (uint64_t)p.drop_cols + (uint64_t)(i + e), p.drop_thr);
            r[e] = keep ? r[e] * p.drop_scale : 0.f;
          }
        }
      } else {
        r[0] = r[1] = r[2] = r[3] = 0.f;
      }
    }
  }
}

template <int ROWS>
__device__ __forceinline__ void store_tile(int trans, const float (&reg)[ROWS * BK / NTHREADS],
                                           float (*sm)[ROWS + PAD], int tid) {
  constexpr int NV = ROWS * BK / NTHREADS / 4;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int f = tid + j * NTHREADS;
    const float* r = &reg[j * 4];
    if (trans == 0) {
      const int row = f / (BK / 4);
      const int kq = (f % (BK / 4)) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) sm[kq + e][row] = r[e];
    } else {
      const int k = f / (ROWS / 4);
      const int i = (f % (ROWS / 4)) * 4;
      *reinterpret_cast<float4*>(&sm[k][i]) = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 2) gemm_f32_kernel(const __grid_constant__ GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tile_n = blockIdx.x % p.ntn;
  const int tile_m = blockIdx.x / p.ntn;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);

  // k-tile range of this split
  int t_begin = blockIdx.z * p.tiles_per_split;
  int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);

  // locate (segment, k0) of t_begin
  int s = 0, k0 = 0;
  {
    int t = t_begin;
    while (s < p.nseg) {
      int nt = (p.K[s] + BK - 1) / BK;
      if (t < nt) { k0 = t * BK; break; }
      t -= nt;
      ++s;
    }
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[BM * BK / NTHREADS], rb[BN * BK / NTHREADS];
  int buf = 0;
  int cur_transA = 0, cur_transB = 0;

  if (t_begin < t_end) {
    load_tile<BM, true>(p, p.A[s], m0, p.M, k0, p.K[s], ra, tid);
    load_tile<BN, false>(p, p.B[s], n0, p.N, k0, p.K[s], rb, tid);
    cur_transA = p.A[s].trans; cur_transB = p.B[s].trans;
    store_tile<BM>(cur_transA, ra, As[0], tid);
    store_tile<BN>(cur_transB, rb, Bs[0], tid);
    k0 += BK;
    if (k0 >= p.K[s]) { ++s; k0 = 0; }
  }
  __syncthreads();

  for (int t = t_begin; t < t_end; ++t) {
    const bool has_next = (t + 1 < t_end);
    if (has_next) {
      load_tile<BM, true>(p, p.A[s], m0, p.M, k0, p.K[s], ra, tid);
      load_tile<BN, false>(p, p.B[s], n0, p.N, k0, p.K[s], rb, tid);
      cur_transA = p.A[s].trans; cur_transB = p.B[s].trans;
      k0 += BK;
      if (k0 >= p.K[s]) { ++s; k0 = 0; }
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      store_tile<BM>(cur_transA, ra, As[buf ^ 1], tid);
      store_tile<BN>(cur_transB, rb, Bs[buf ^ 1], tid);
      __syncthreads();
      buf ^= 1;
    }
  }

  const int n = n0 + tx * TN;
  if (n >= p.N) return;
  if (p.split_k > 1) {
    float* ws = p.workspace + (int64_t)blockIdx.z * p.M * p.N;
    const bool vec = (p.N % 4) == 0;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m < p.M) store4(ws + (int64_t)m * p.N + n, vec, min(4, p.N - n), acc[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m < p.M) epilogue4(p, m, n, acc[i]);
    }
  }
}

__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(const __grid_constant__ GemmParams p) {
  const int nq = (p.N + 3) / 4;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (int64_t)p.M * nq) return;
  const int m = (int)(q / nq);
  const int n = (int)(q % nq) * 4;
  const bool vec = (p.N % 4) == 0;
  const int nvalid = min(4, p.N - n);
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, v[4];
  for (int z = 0; z < p.split_k; ++z) {
    load4(p.workspace + ((int64_t)z * p.M + m) * p.N + n, vec, nvalid, v);
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += v[e];
  }
  epilogue4(p, m, n, acc);
}

static bool op_vec_ok(const get_gemm_operand& o, int rows, int k) {
  if (!aligned16(o.ptr) || (o.ld % 4) != 0) return false;
  return o.trans == 0 ? (k % 4) == 0 : (rows % 4) == 0;
}

int gemm_build_params(const get_gemm_desc* d, GemmParams& p) {
  GETB_REQUIRE(d != nullptr, "get_gemm_f32: null descriptor");
  GETB_REQUIRE(d->nseg >= 1 && d->nseg <= GET_GEMM_MAX_SEG, "get_gemm_f32: nseg=%d out of range", d->nseg);
  GETB_REQUIRE(d->M >= 0 && d->N >= 0, "get_gemm_f32: negative M/N");
  GETB_REQUIRE(d->C != nullptr || d->M == 0 || d->N == 0, "get_gemm_f32: null C");
  GETB_REQUIRE(d->epilogue >= GET_EPI_STORE && d->epilogue <= GET_EPI_TANH, "get_gemm_f32: bad epilogue %d", d->epilogue);
  memset(&p, 0, sizeof(p));
  int tiles = 0;
  for (int s = 0; s < d->nseg; ++s) {
    GETB_REQUIRE(d->K[s] > 0, "get_gemm_f32: K[%d]=%d must be positive", s, d->K[s]);
    GETB_REQUIRE(d->A[s].ptr && d->B[s].ptr, "get_gemm_f32: null operand in segment %d", s);
    GETB_REQUIRE(d->B[s].rowidx == nullptr, "get_gemm_f32: row gather is supported on A only");
    p.A[s] = GemmOp{d->A[s].ptr, d->A[s].ld, d->A[s].rowidx, d->A[s].trans, op_vec_ok(d->A[s], d->M, d->K[s])};
    p.B[s] = GemmOp{d->B[s].ptr, d->B[s].ld, nullptr, d->B[s].trans, op_vec_ok(d->B[s], d->N, d->K[s])};
    p.K[s] = d->K[s];
    tiles += (d->K[s] + BK - 1) / BK;
  }
  p.nseg = d->nseg; p.M = d->M; p.N = d->N;
  p.C = d->C; p.ldc = d->ldc; p.alpha = d->alpha;
  p.accumulate = d->accumulate; p.epilogue = d->epilogue;
  p.bias0 = d->bias0; p.bias1 = d->bias1;
  p.aux0 = d->aux0; p.ld_aux0 = d->ld_aux0; p.aux1 = d->aux1; p.ld_aux1 = d->ld_aux1;
  p.out1 = d->out1; p.ld_out1 = d->ld_out1;
  p.group_rows = d->group_rows;
  switch (d->epilogue) {
    case GET_EPI_SIGMOID: GETB_REQUIRE(!d->out1 || d->aux0, "get_gemm_f32: SIGMOID with out1 needs aux0"); break;
    case GET_EPI_TANH_BLEND: GETB_REQUIRE(d->aux0 && d->aux1, "get_gemm_f32: TANH_BLEND needs aux0 (z) and aux1 (x)"); break;
    case GET_EPI_TANH_ROWGROUP: GETB_REQUIRE(d->aux0 && d->group_rows > 0, "get_gemm_f32: TANH_ROWGROUP needs aux0 and group_rows"); break;
    case GET_EPI_DGATE_R: GETB_REQUIRE(d->aux0 && d->aux1 && d->out1, "get_gemm_f32: DGATE_R needs aux0 (x), aux1 (r), out1 (dx)"); break;
    default: break;
  }
  GETB_REQUIRE(d->drop_p >= 0.f && d->drop_p < 1.f && d->drop_out_p >= 0.f && d->drop_out_p < 1.f,
               "get_gemm_f32: dropout probability must be in [0,1)");
  if (d->drop_p > 0.f) {
    GETB_REQUIRE(d->drop_cols > 0, "get_gemm_f32: drop_cols must be set with drop_p");
    p.drop_thr = drop_threshold(d->drop_p);
    p.drop_seed = d->drop_seed; p.drop_cols = d->drop_cols; p.drop_scale = 1.0f / (1.0f - d->drop_p);
  }
  if (d->epilogue == GET_EPI_DROPOUT_OUT) {
    p.drop_out_thr = drop_threshold(d->drop_out_p);
    p.drop_out_seed = d->drop_out_seed; p.drop_out_scale = 1.0f / (1.0f - d->drop_out_p);
  }
  p.salt = dropout_salt_ptr();
  p.split_k = d->split_k > 1 ? d->split_k : 1;
  if (p.split_k > tiles) p.split_k = tiles > 0 ? tiles : 1;
  p.tiles_total = tiles;
  p.tiles_per_split = (tiles + p.split_k - 1) / p.split_k;
  p.split_k = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.workspace = d->workspace;
  GETB_REQUIRE(p.split_k == 1 || d->workspace, "get_gemm_f32: split_k > 1 needs a workspace");
  bool ve = (d->N % 4) == 0 && (d->ldc % 4) == 0 && aligned16(d->C);
  if (d->bias0) ve = ve && aligned16(d->bias0);
  if (d->bias1) ve = ve && aligned16(d->bias1);
  if (d->aux0) ve = ve && aligned16(d->aux0) && (d->ld_aux0 % 4) == 0;
  if (d->aux1) ve = ve && aligned16(d->aux1) && (d->ld_aux1 % 4) == 0;
  if (d->out1) ve = ve && aligned16(d->out1) && (d->ld_out1 % 4) == 0;
  p.vec_epi = ve;
  p.ntn = (d->N + BN - 1) / BN;
  return 0;
}

}  // namespace getb

using namespace getb;

extern "C" int get_gemm_f32_launches(const get_gemm_desc* d) {
  GemmParams p;
  if (gemm_build_params(d, p) != 0) return -1;
  if (p.M == 0 || p.N == 0) return 0;
  return p.split_k > 1 ? 2 : 1;
}

extern "C" int get_gemm_f32(const get_gemm_desc* d, void* stream) {
  GemmParams p;
  int rc = gemm_build_params(d, p);
  if (rc != 0) return rc;
  if (p.M == 0 || p.N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ntm = (p.M + BM - 1) / BM;
  const int64_t nblk = ntm * p.ntn;
  GETB_REQUIRE(nblk < (int64_t)2147483647, "get_gemm_f32: too many tiles");
  dim3 grid((unsigned)nblk, 1, (unsigned)p.split_k);
  gemm_f32_kernel<<<grid, NTHREADS, 0, st>>>(p);
  GETB_CHECK_LAUNCH("gemm_f32_kernel");
  if (p.split_k > 1) {
    const int64_t nq = (int64_t)p.M * ((p.N + 3) / 4);
    gemm_splitk_reduce_kernel<<<ceil_div(nq, 256), 256, 0, st>>>(p);
    GETB_CHECK_LAUNCH("gemm_splitk_reduce_kernel");
  }
  return 0;
}
