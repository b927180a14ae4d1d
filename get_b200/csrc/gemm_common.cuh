// Shared between the SIMT (exact fp32) and tcgen05 (3xTF32) GEMM kernels: launch parameters and the fused epilogues.
#pragma once
#include "common.cuh"

namespace getb {

struct GemmOp {
  const float* ptr;
  int64_t ld;
  const int64_t* rowidx;
  int trans;
  int vec;
};

struct GemmParams {
  GemmOp A[GET_GEMM_MAX_SEG];
  GemmOp B[GET_GEMM_MAX_SEG];
  int K[GET_GEMM_MAX_SEG];
  int nseg, M, N;
  float* C;
  int64_t ldc;
  float alpha;
  int accumulate, epilogue;
  const float *bias0, *bias1, *aux0, *aux1;
  int64_t ld_aux0, ld_aux1;
  float* out1;
  int64_t ld_out1;
  int group_rows;
  uint32_t drop_thr, drop_seed;
  int drop_cols;
  float drop_scale;
  uint32_t drop_out_thr, drop_out_seed;
  float drop_out_scale;
  int split_k, tiles_per_split, tiles_total;
  float* workspace;
  int vec_epi;
  int ntn;  // number of tiles along N
};

// ---- epilogue on up to 4 consecutive columns (n .. n+3) of row m ------------------------------
__device__ __forceinline__ void load4(const float* base, bool vec, int nvalid, float v[4]) {
  if (vec) {
    float4 t = *reinterpret_cast<const float4*>(base);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = e < nvalid ? base[e] : 0.f;
  }
}
__device__ __forceinline__ void store4(float* base, bool vec, int nvalid, const float v[4]) {
  if (vec) {
    *reinterpret_cast<float4*>(base) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < nvalid) base[e] = v[e];
  }
}

__device__ __forceinline__ void epilogue4(const GemmParams& p, int m, int n, float acc[4]) {
  const int nvalid = min(4, p.N - n);
  const bool vec = p.vec_epi != 0;  // host guarantees N % 4 == 0 and 16-byte alignment of every pointer used
  float v[4], b[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = p.alpha * acc[e];
  if (p.bias0) {
    load4(p.bias0 + n, vec, nvalid, b);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] += b[e];
  }
  if (p.bias1) {
    load4(p.bias1 + n, vec, nvalid, b);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] += b[e];
  }
  float* crow = p.C + (int64_t)m * p.ldc + n;
  float a0[4], a1[4], o[4];
  switch (p.epilogue) {
    case GET_EPI_STORE: {
      if (p.accumulate) {
        load4(crow, vec, nvalid, o);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] += o[e];
      }
      store4(crow, vec, nvalid, v);
    } break;
    case GET_EPI_SIGMOID: {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = sigmoidf_(v[e]);
      store4(crow, vec, nvalid, v);
      if (p.out1) {
        load4(p.aux0 + (int64_t)m * p.ld_aux0 + n, vec, nvalid, a0);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = v[e] * a0[e];
        store4(p.out1 + (int64_t)m * p.ld_out1 + n, vec, nvalid, o);
      }
    } break;
    case GET_EPI_TANH_BLEND: {
      load4(p.aux0 + (int64_t)m * p.ld_aux0 + n, vec, nvalid, a0);  // z
      load4(p.aux1 + (int64_t)m * p.ld_aux1 + n, vec, nvalid, a1);  // x
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = tanhf(v[e]);
        o[e] = v[e] * a0[e] + a1[e] * (1.0f - a0[e]);
      }
      if (p.out1) store4(p.out1 + (int64_t)m * p.ld_out1 + n, vec, nvalid, v);
      store4(crow, vec, nvalid, o);
    } break;
    case GET_EPI_TANH_ROWGROUP: {
      load4(p.aux0 + (int64_t)(m / p.group_rows) * p.ld_aux0 + n, vec, nvalid, a0);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = tanhf(v[e] + a0[e]);
      store4(crow, vec, nvalid, v);
    } break;
    case GET_EPI_DGATE_R: {
      load4(p.aux0 + (int64_t)m * p.ld_aux0 + n, vec, nvalid, a0);  // x
      load4(p.aux1 + (int64_t)m * p.ld_aux1 + n, vec, nvalid, a1);  // r
      float* o1 = p.out1 + (int64_t)m * p.ld_out1 + n;
      load4(o1, vec, nvalid, o);
      float c[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        c[e] = v[e] * a0[e] * a1[e] * (1.0f - a1[e]);
        o[e] += v[e] * a1[e];
      }
      store4(crow, vec, nvalid, c);
      store4(o1, vec, nvalid, o);
    } break;
    case GET_EPI_DROPOUT_OUT: {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        bool keep = drop_keep(p.drop_out_seed, (uint64_t)m * (uint64_t)p.N + (uint64_t)(n + e), p.drop_out_thr);
        v[e] = keep ? v[e] * p.drop_out_scale : 0.f;
      }
      if (p.accumulate) {
        load4(crow, vec, nvalid, o);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] += o[e];
      }
      store4(crow, vec, nvalid, v);
    } break;
    case GET_EPI_TANH: {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = tanhf(v[e]);
      store4(crow, vec, nvalid, v);
    } break;
    default: break;
  }
}


// host side: validate a descriptor and fill the kernel parameters (gemm_simt.cu)
int gemm_build_params(const get_gemm_desc* d, GemmParams& p);
// tcgen05 path (gemm_tc.cu): returns 0 when launched, 1 when the descriptor is not eligible (caller falls back to SIMT)
int gemm_tc_launch(const get_gemm_desc* d, const GemmParams& p, cudaStream_t st);
// persistent TMA-fed tcgen05 path (gemm_tc2.cu): 0 launched, 2 launched with split-K partials in the workspace (p.split_k
// updated; the caller runs the reduction), 1 not eligible, <0 error
int gemm_tc2_launch(const get_gemm_desc* d, GemmParams& p, cudaStream_t st);
int gemm_tc2_plan_splits(const get_gemm_desc* d, const GemmParams& p);   // -1 not eligible, else number of k splits
int gemm_tc_eligible(const get_gemm_desc* d, const GemmParams& p);       // gemm_tc.cu: 1 eligible

}  // namespace getb
