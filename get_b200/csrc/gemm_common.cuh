// Exact-fp32 SIMT GEMM (small / odd contractions): launch parameters and the fused epilogues.
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
  const uint32_t* salt;   // device word added to both dropout seeds
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

// The epilogue is split in two so that callers with several (m, n) quads in flight can issue ALL their global loads
// before the first dependent use (the loads of one quad must not wait behind the stores of the previous one).
struct EpiIn {
  float a0[4], a1[4], o[4];   // aux0 / aux1 / previous C (accumulate) or out1 (DGATE_R)
};

// EPI >= 0: epilogue kind known at compile time (specialised kernels: no switch, small code); EPI < 0: p.epilogue
template <int EPI, bool VEC = false>
__device__ __forceinline__ void epilogue4_load_t(const GemmParams& p, int m, int n, EpiIn& in) {
  const int nvalid = VEC ? 4 : min(4, p.N - n);
  const bool vec = VEC || p.vec_epi != 0;
  switch (EPI >= 0 ? EPI : p.epilogue) {
    case GET_EPI_STORE:
    case GET_EPI_DROPOUT_OUT:
      if (p.accumulate) load4(p.C + (int64_t)m * p.ldc + n, vec, nvalid, in.o);
      break;
    case GET_EPI_SIGMOID:
      if (p.out1) load4(p.aux0 + (int64_t)m * p.ld_aux0 + n, vec, nvalid, in.a0);
      break;
    case GET_EPI_TANH_BLEND:
      load4(p.aux0 + (int64_t)m * p.ld_aux0 + n, vec, nvalid, in.a0);  // z
      load4(p.aux1 + (int64_t)m * p.ld_aux1 + n, vec, nvalid, in.a1);  // x
      break;
    case GET_EPI_TANH_ROWGROUP:
      load4(p.aux0 + (int64_t)(m / p.group_rows) * p.ld_aux0 + n, vec, nvalid, in.a0);
      break;
    case GET_EPI_DGATE_R:
      load4(p.aux0 + (int64_t)m * p.ld_aux0 + n, vec, nvalid, in.a0);  // x
      load4(p.aux1 + (int64_t)m * p.ld_aux1 + n, vec, nvalid, in.a1);  // r
      load4(p.out1 + (int64_t)m * p.ld_out1 + n, vec, nvalid, in.o);
      break;
    default: break;
  }
}

template <int EPI, bool VEC = false>
__device__ __forceinline__ void epilogue4_apply_t(const GemmParams& p, int m, int n, const float acc[4], const EpiIn& in) {
  const int nvalid = VEC ? 4 : min(4, p.N - n);
  const bool vec = VEC || p.vec_epi != 0;  // host guarantees N % 4 == 0 and 16-byte alignment of every pointer used
  float v[4], b[4], o[4];
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
  switch (EPI >= 0 ? EPI : p.epilogue) {
    case GET_EPI_STORE: {
      if (p.accumulate) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] += in.o[e];
      }
      store4(crow, vec, nvalid, v);
    } break;
    case GET_EPI_SIGMOID: {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = sigmoid_fast(v[e]);
      store4(crow, vec, nvalid, v);
      if (p.out1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = v[e] * in.a0[e];
        store4(p.out1 + (int64_t)m * p.ld_out1 + n, vec, nvalid, o);
      }
    } break;
    case GET_EPI_TANH_BLEND: {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = tanh_fast(v[e]);
        o[e] = v[e] * in.a0[e] + in.a1[e] * (1.0f - in.a0[e]);
      }
      if (p.out1) store4(p.out1 + (int64_t)m * p.ld_out1 + n, vec, nvalid, v);
      store4(crow, vec, nvalid, o);
    } break;
    case GET_EPI_TANH_ROWGROUP: {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = tanh_fast(v[e] + in.a0[e]);
      store4(crow, vec, nvalid, v);
    } break;
    case GET_EPI_DGATE_R: {
      float c[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        c[e] = v[e] * in.a0[e] * in.a1[e] * (1.0f - in.a1[e]);
        o[e] = in.o[e] + v[e] * in.a1[e];
      }
      store4(crow, vec, nvalid, c);
      store4(p.out1 + (int64_t)m * p.ld_out1 + n, vec, nvalid, o);
    } break;
    case GET_EPI_DROPOUT_OUT: {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        bool keep = drop_keep(p.drop_out_seed + __ldg(p.salt), (uint64_t)m * (uint64_t)p.N + (uint64_t)(n + e), p.drop_out_thr);
        v[e] = keep ? v[e] * p.drop_out_scale : 0.f;
      }
      if (p.accumulate) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] += in.o[e];
      }
      store4(crow, vec, nvalid, v);
    } break;
    case GET_EPI_TANH: {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = tanh_fast(v[e]);
      store4(crow, vec, nvalid, v);
    } break;
    default: break;
  }
}

__device__ __forceinline__ void epilogue4(const GemmParams& p, int m, int n, float acc[4]) {
  EpiIn in;
  epilogue4_load_t<-1>(p, m, n, in);
  epilogue4_apply_t<-1>(p, m, n, acc, in);
}


// host side: validate a descriptor and fill the kernel parameters (gemm_simt.cu)
int gemm_build_params(const get_gemm_desc* d, GemmParams& p);
}  // namespace getb
