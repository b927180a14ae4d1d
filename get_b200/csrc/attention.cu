// Multi-head additive attention pooling: the tail of ConcatNotEqualSelfAtt
// (reference thirdparty/two_branches_attention.py:141-147) and MultiHeadSelfAttentionICLR2017Extend
// (thirdparty/self_attention.py:90-96), forward and backward.
// HBM-bound streaming kernels: per group (an evidence at word level, a claim at evidence level) the (P, H) tanh
// activations and the (P, Dr) pooled operand are read exactly once with 128-bit loads; one 256-thread CTA per group, so
// a Snopes batch (~216 groups) keeps every SM busy. All reductions over positions are warp-shuffle or fixed-order
// serial sums (deterministic). The backward kernel can emit du as bf16 planes (operand of the tensor-core GEMMs that
// follow, see get_gemm_bp).
#include <mutex>
#include <set>

#include "common.cuh"
#include "tcgen05.cuh"

namespace getb {

constexpr int ATT_THREADS = 512;   // 220 groups on 148 SMs: the kernels are latency chains per warp, so more warps per group
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int ATT_MAX_HEADS = 8;

struct AttParams {
  const float* t;       // (G,P,H)
  const float* right;   // (G,P,Dr), row stride ld_right
  int64_t ld_right;
  const float* W2;      // (C,H)
  const uint8_t* mask;  // (G,P)
  const float* att_in;  // bwd
  const float* d_pooled; int64_t ld_dpooled;
  const float* d_att;
  int G, P, H, Dr, C;
  float* att;           // fwd out (G,P,C)
  float* pooled; int64_t ld_pooled;
  float* de; float* du; float* du_sum; float* dright; int64_t ld_dright;
  __nv_bfloat16* du_p; int64_t ld_dup, ps_dup; int np_dup;   // du as bf16 planes (optional)
  int accumulate;
  int vec;              // H % 4 == 0, Dr % 4 == 0 and 16-byte aligned rows: 128-bit path
  int dchunk;           // fwd: columns of `right` pooled per pass (bounds the partial-sum buffer)
};

// fwd smem: W2 (C*H) | att (P*C) | pooled partials (ATT_WARPS * Dr * C)   [the partials are reduced in a fixed order]
template <int C>
__global__ void __launch_bounds__(ATT_THREADS) att_pool_fwd_kernel(const __grid_constant__ AttParams p) {
  extern __shared__ __align__(16) float smem[];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = p.P, H = p.H, Dr = p.Dr;
  float* sW2 = smem;
  float* sE = smem + C * H;
  float* sPart = sE + P * C;
  for (int q = tid; q < C * H; q += ATT_THREADS) sW2[q] = __ldg(p.W2 + q);
  __syncthreads();
  const float* tg = p.t + (int64_t)g * P * H;
  // e[p,c] = t[p,:] . W2[c,:]   (one warp per position)
  for (int pp = warp; pp < P; pp += ATT_WARPS) {
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    const float* row = tg + (int64_t)pp * H;
    if (p.vec) {
      for (int q = lane; q < (H >> 2); q += 32) {
        const float4 tv = __ldg(reinterpret_cast<const float4*>(row) + q);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float4 w = *reinterpret_cast<const float4*>(sW2 + c * H + q * 4);
          acc[c] = fmaf(tv.x, w.x, acc[c]); acc[c] = fmaf(tv.y, w.y, acc[c]);
          acc[c] = fmaf(tv.z, w.z, acc[c]); acc[c] = fmaf(tv.w, w.w, acc[c]);
        }
      }
    } else {
      for (int h = lane; h < H; h += 32) {
        const float tv = __ldg(row + h);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(tv, sW2[c * H + h], acc[c]);
      }
    }
    const bool valid = p.mask[(int64_t)g * P + pp] != 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v = warp_sum(acc[c]);
      if (lane == 0) sE[pp * C + c] = valid ? v : -INFINITY;
    }
  }
  __syncthreads();
  // softmax over positions, one warp per head
  for (int c = warp; c < C; c += ATT_WARPS) {
    float mx = -INFINITY;
    for (int pp = lane; pp < P; pp += 32) mx = fmaxf(mx, sE[pp * C + c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int pp = lane; pp < P; pp += 32) {
      const float ev = expf(sE[pp * C + c] - mx);  // all-masked group: (-inf) - (-inf) = NaN, as in the reference
      sE[pp * C + c] = ev;
      sum += ev;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int pp = lane; pp < P; pp += 32) {
      const float a = sE[pp * C + c] * inv;
      sE[pp * C + c] = a;
      p.att[((int64_t)g * P + pp) * C + c] = a;
    }
  }
  __syncthreads();
  // pooled[d,c] = sum_p right[p,d] * att[p,c]: warp w takes positions w, w+8, ...; lanes take column quads; the ATT_WARPS
  // partial sums are reduced in a fixed order. Wide operands (evidence level: Dr = heads*H + E) go through the partial
  // buffer in column chunks of p.dchunk.
  const float* rg = p.right + (int64_t)g * P * p.ld_right;
  float* og = p.pooled + (int64_t)g * p.ld_pooled;
  const int DC = p.dchunk;
  for (int d0 = 0; d0 < Dr; d0 += DC) {
    const int dn = min(DC, Dr - d0);
    if (p.vec) {
      const int DQ = dn >> 2;
      for (int q0 = 0; q0 < DQ; q0 += 32) {
        const int q = q0 + lane;
        float acc[C][4];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
        if (q < DQ) {
          for (int pp = warp; pp < P; pp += ATT_WARPS) {
            const float4 rv = __ldg(reinterpret_cast<const float4*>(rg + (int64_t)pp * p.ld_right + d0) + q);
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const float a = sE[pp * C + c];
              acc[c][0] = fmaf(rv.x, a, acc[c][0]); acc[c][1] = fmaf(rv.y, a, acc[c][1]);
              acc[c][2] = fmaf(rv.z, a, acc[c][2]); acc[c][3] = fmaf(rv.w, a, acc[c][3]);
            }
          }
#pragma unroll
          for (int c = 0; c < C; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) sPart[((size_t)warp * DC + q * 4 + e) * C + c] = acc[c][e];
        }
      }
    } else {
      for (int d = lane; d < dn; d += 32) {
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        for (int pp = warp; pp < P; pp += ATT_WARPS) {
          const float rv = __ldg(rg + (int64_t)pp * p.ld_right + d0 + d);
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(rv, sE[pp * C + c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) sPart[((size_t)warp * DC + d) * C + c] = acc[c];
      }
    }
    __syncthreads();
    for (int q = tid; q < dn * C; q += ATT_THREADS) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < ATT_WARPS; ++w) v += sPart[(size_t)w * DC * C + q];
      og[(size_t)d0 * C + q] = v;
    }
    __syncthreads();
  }
}

// bwd smem: W2 (C*H) | dO (Dr*C) | att (P*C) | de (P*C) | dot (C, padded to 8) | du_sum partials (ATT_WARPS * H)
template <int C>
__global__ void __launch_bounds__(ATT_THREADS) att_pool_bwd_kernel(const __grid_constant__ AttParams p) {
  extern __shared__ __align__(16) float smem[];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = p.P, H = p.H, Dr = p.Dr;
  float* sW2 = smem;
  float* sdO = sW2 + C * H;
  float* sAtt = sdO + Dr * C;
  float* sDe = sAtt + P * C;
  float* sDot = sDe + P * C;
  float* sSum = sDot + 8;
  for (int q = tid; q < C * H; q += ATT_THREADS) sW2[q] = __ldg(p.W2 + q);
  const float* dOg = p.d_pooled + (int64_t)g * p.ld_dpooled;
  for (int q = tid; q < Dr * C; q += ATT_THREADS) sdO[q] = __ldg(dOg + q);
  for (int q = tid; q < P * C; q += ATT_THREADS) sAtt[q] = __ldg(p.att_in + (int64_t)g * P * C + q);
  __syncthreads();
  const float* rg = p.right + (int64_t)g * P * p.ld_right;
  // dalpha[p,c] = right[p,:] . dO[:,c] (+ d_att)     (one warp per position)
  for (int pp = warp; pp < P; pp += ATT_WARPS) {
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    const float* row = rg + (int64_t)pp * p.ld_right;
    if (p.vec) {
      for (int q = lane; q < (Dr >> 2); q += 32) {
        const float4 rv = __ldg(reinterpret_cast<const float4*>(row) + q);
        const float* o = sdO + q * 4 * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          acc[c] = fmaf(rv.x, o[c], acc[c]); acc[c] = fmaf(rv.y, o[C + c], acc[c]);
          acc[c] = fmaf(rv.z, o[2 * C + c], acc[c]); acc[c] = fmaf(rv.w, o[3 * C + c], acc[c]);
        }
      }
    } else {
      for (int d = lane; d < Dr; d += 32) {
        const float rv = __ldg(row + d);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(rv, sdO[d * C + c], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float v = warp_sum(acc[c]);
      if (p.d_att) v += __ldg(p.d_att + ((int64_t)g * P + pp) * C + c);
      if (lane == 0) sDe[pp * C + c] = v;  // holds dalpha for now
    }
  }
  __syncthreads();
  for (int c = warp; c < C; c += ATT_WARPS) {
    float s = 0.f;
    for (int pp = lane; pp < P; pp += 32) s = fmaf(sAtt[pp * C + c], sDe[pp * C + c], s);
    s = warp_sum(s);
    if (lane == 0) sDot[c] = s;
  }
  __syncthreads();
  for (int q = tid; q < P * C; q += ATT_THREADS) {
    const int c = q % C;
    const float a = sAtt[q];
    const float v = a == 0.f ? 0.f : a * (sDe[q] - sDot[c]);   // masked positions: att = 0 -> de = 0
    sDe[q] = v;
    p.de[(int64_t)g * P * C + q] = v;
  }
  __syncthreads();
  // du[p,h] = (de[p,:] @ W2[:,h]) * (1 - t^2);  du_sum[h] = sum_p du[p,h]: warp w takes positions w, w+8, ...
  const float* tg = p.t + (int64_t)g * P * H;
  if (p.vec) {
    const int HQ = H >> 2;
    const int HPQ = ((H + 7) & ~7) >> 2;            // quads of a plane row including the padding quad
    for (int q0 = 0; q0 < HPQ; q0 += 32) {
      const int q = q0 + lane;
      float4 w[C];
#pragma unroll
      for (int c = 0; c < C; ++c) w[c] = q < HQ ? *reinterpret_cast<const float4*>(sW2 + c * H + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
      if (q < HPQ) {
        for (int pp = warp; pp < P; pp += ATT_WARPS) {
          float d[4] = {0.f, 0.f, 0.f, 0.f};
          if (q < HQ) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const float e = sDe[pp * C + c];
              d[0] = fmaf(e, w[c].x, d[0]); d[1] = fmaf(e, w[c].y, d[1]); d[2] = fmaf(e, w[c].z, d[2]); d[3] = fmaf(e, w[c].w, d[3]);
            }
            const float4 tv = __ldg(reinterpret_cast<const float4*>(tg + (int64_t)pp * H) + q);
            d[0] *= 1.0f - tv.x * tv.x; d[1] *= 1.0f - tv.y * tv.y; d[2] *= 1.0f - tv.z * tv.z; d[3] *= 1.0f - tv.w * tv.w;
            if (p.du) *(reinterpret_cast<float4*>(p.du + ((int64_t)g * P + pp) * H) + q) = make_float4(d[0], d[1], d[2], d[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) sum[e] += d[e];
          }
          if (p.du_p) planes_store4(p.du_p + ((int64_t)g * P + pp) * p.ld_dup + q * 4, p.ps_dup, p.np_dup, d);
        }
        if (q < HQ) {
#pragma unroll
          for (int e = 0; e < 4; ++e) sSum[warp * H + q * 4 + e] = sum[e];
        }
      }
    }
  } else {
    for (int h = lane; h < H; h += 32) {
      float w[C];
#pragma unroll
      for (int c = 0; c < C; ++c) w[c] = sW2[c * H + h];
      float sum = 0.f;
      for (int pp = warp; pp < P; pp += ATT_WARPS) {
        float dt = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) dt = fmaf(sDe[pp * C + c], w[c], dt);
        const float tv = __ldg(tg + (int64_t)pp * H + h);
        const float duv = dt * (1.0f - tv * tv);
        p.du[((int64_t)g * P + pp) * H + h] = duv;
        sum += duv;
      }
      sSum[warp * H + h] = sum;
    }
  }
  __syncthreads();
  if (p.du_sum) {
    for (int h = tid; h < H; h += ATT_THREADS) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < ATT_WARPS; ++w) v += sSum[w * H + h];
      p.du_sum[(int64_t)g * H + h] = v;
    }
  }
  // dright[p,d] (+)= sum_c att[p,c] * dO[d,c]
  float* drg = p.dright + (int64_t)g * P * p.ld_dright;
  if (p.vec && (p.ld_dright & 3) == 0) {
    const int DQ = Dr >> 2;
    for (int e = tid; e < P * DQ; e += ATT_THREADS) {
      const int pp = e / DQ, q = e - pp * DQ;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const float* o = sdO + q * 4 * C;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float a = sAtt[pp * C + c];
        v[0] = fmaf(a, o[c], v[0]); v[1] = fmaf(a, o[C + c], v[1]); v[2] = fmaf(a, o[2 * C + c], v[2]); v[3] = fmaf(a, o[3 * C + c], v[3]);
      }
      float4* dst = reinterpret_cast<float4*>(drg + (int64_t)pp * p.ld_dright) + q;
      if (p.accumulate) {
        const float4 old = *dst;
        v[0] += old.x; v[1] += old.y; v[2] += old.z; v[3] += old.w;
      }
      *dst = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    for (int q = tid; q < P * Dr; q += ATT_THREADS) {
      const int pp = q / Dr, d = q % Dr;
      float v = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) v = fmaf(sAtt[pp * C + c], sdO[d * C + c], v);
      float* dst = drg + (int64_t)pp * p.ld_dright + d;
      *dst = p.accumulate ? *dst + v : v;
    }
  }
}

static int check_att(const char* name, int G, int P, int H, int Dr, int C) {
  GETB_REQUIRE(G >= 0 && P > 0 && H > 0 && Dr > 0, "%s: bad sizes", name);
  GETB_REQUIRE(C >= 1 && C <= ATT_MAX_HEADS, "%s: heads=%d not in [1,%d]", name, C, ATT_MAX_HEADS);
  return 0;
}

typedef void (*AttFn)(const AttParams);
static AttFn att_fwd_fn(int C) {
  switch (C) {
    case 1: return att_pool_fwd_kernel<1>; case 2: return att_pool_fwd_kernel<2>; case 3: return att_pool_fwd_kernel<3>;
    case 4: return att_pool_fwd_kernel<4>; case 5: return att_pool_fwd_kernel<5>; case 6: return att_pool_fwd_kernel<6>;
    case 7: return att_pool_fwd_kernel<7>; default: return att_pool_fwd_kernel<8>;
  }
}
static AttFn att_bwd_fn(int C) {
  switch (C) {
    case 1: return att_pool_bwd_kernel<1>; case 2: return att_pool_bwd_kernel<2>; case 3: return att_pool_bwd_kernel<3>;
    case 4: return att_pool_bwd_kernel<4>; case 5: return att_pool_bwd_kernel<5>; case 6: return att_pool_bwd_kernel<6>;
    case 7: return att_pool_bwd_kernel<7>; default: return att_pool_bwd_kernel<8>;
  }
}

static int att_launch(AttFn fn, const AttParams& p, size_t smem, cudaStream_t st, const char* name) {
  GETB_REQUIRE(smem <= 200 * 1024, "%s: shared memory %zu too large", name, smem);
  if (smem > 48 * 1024) {
    static std::mutex mu;
    static std::set<AttFn> done;
    std::lock_guard<std::mutex> lock(mu);
    if (!done.count(fn)) {
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("%s: cannot opt in to %zu bytes of shared memory", name, smem);
        return -2;
      }
      done.insert(fn);
    }
  }
  fn<<<p.G, ATT_THREADS, smem, st>>>(p);
  GETB_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace getb

using namespace getb;

extern "C" int get_att_pool_fwd_f32(const float* t, const float* right, int64_t ld_right, const float* W2,
                                    const uint8_t* mask, int G, int P, int H, int Dr, int C, float* att, float* pooled,
                                    int64_t ld_pooled, void* stream) {
  GETB_REQUIRE(t && right && W2 && mask && att && pooled, "get_att_pool_fwd_f32: null pointer");
  if (check_att("get_att_pool_fwd_f32", G, P, H, Dr, C)) return -1;
  if (G == 0) return 0;
  AttParams p;
  memset(&p, 0, sizeof(p));
  p.t = t; p.right = right; p.ld_right = ld_right; p.W2 = W2; p.mask = mask;
  p.G = G; p.P = P; p.H = H; p.Dr = Dr; p.C = C; p.att = att; p.pooled = pooled; p.ld_pooled = ld_pooled;
  p.vec = (H % 4) == 0 && (Dr % 4) == 0 && (ld_right % 4) == 0 && aligned16(t) && aligned16(right);
  int dchunk = (24576 / ATT_WARPS / C) / 128 * 128;      // <= 96 KB of partial sums
  if (dchunk < 128) dchunk = 128;
  if (dchunk > Dr) dchunk = Dr;
  p.dchunk = dchunk;
  const size_t smem = ((size_t)C * H + (size_t)P * C + (size_t)ATT_WARPS * dchunk * C) * sizeof(float);
  return att_launch(att_fwd_fn(C), p, smem, (cudaStream_t)stream, "get_att_pool_fwd_f32");
}

static int att_bwd_common(const float* t, const float* right, int64_t ld_right, const float* W2, const float* att,
                          const float* d_pooled, int64_t ld_dpooled, const float* d_att, int G, int P, int H, int Dr, int C,
                          float* de, float* du, void* du_planes, int64_t ld_dup, int64_t ps_dup, int np_dup, float* du_sum,
                          float* dright, int64_t ld_dright, int accumulate, void* stream, const char* name) {
  GETB_REQUIRE(t && right && W2 && att && d_pooled && de && (du || du_planes) && dright, "%s: null pointer", name);
  if (check_att(name, G, P, H, Dr, C)) return -1;
  if (G == 0) return 0;
  AttParams p;
  memset(&p, 0, sizeof(p));
  p.t = t; p.right = right; p.ld_right = ld_right; p.W2 = W2; p.att_in = att;
  p.d_pooled = d_pooled; p.ld_dpooled = ld_dpooled; p.d_att = d_att;
  p.G = G; p.P = P; p.H = H; p.Dr = Dr; p.C = C;
  p.de = de; p.du = du; p.du_sum = du_sum; p.dright = dright; p.ld_dright = ld_dright; p.accumulate = accumulate;
  p.du_p = reinterpret_cast<__nv_bfloat16*>(du_planes); p.ld_dup = ld_dup; p.ps_dup = ps_dup; p.np_dup = np_dup;
  p.vec = (H % 4) == 0 && (Dr % 4) == 0 && (ld_right % 4) == 0 && aligned16(t) && aligned16(right) && (!du || aligned16(du));
  if (du_planes)
    GETB_REQUIRE(p.vec && (((uintptr_t)du_planes) & 7u) == 0 && (ld_dup % 4) == 0 && (ps_dup % 4) == 0 && np_dup >= 1 && np_dup <= 3 &&
                     ld_dup >= ((H + 7) & ~7),
                 "%s: plane output needs H %% 4 == 0, Dr %% 4 == 0 and aligned tensors", name);
  const size_t smem = ((size_t)C * H + (size_t)Dr * C + (size_t)2 * P * C + 8 + (size_t)ATT_WARPS * H) * sizeof(float);
  return att_launch(att_bwd_fn(C), p, smem, (cudaStream_t)stream, name);
}

extern "C" int get_att_pool_bwd_f32(const float* t, const float* right, int64_t ld_right, const float* W2,
                                    const float* att, const float* d_pooled, int64_t ld_dpooled, const float* d_att,
                                    int G, int P, int H, int Dr, int C, float* de, float* du, float* du_sum,
                                    float* dright, int64_t ld_dright, int accumulate, void* stream) {
  return att_bwd_common(t, right, ld_right, W2, att, d_pooled, ld_dpooled, d_att, G, P, H, Dr, C, de, du, nullptr, 0, 0, 0, du_sum,
                        dright, ld_dright, accumulate, stream, "get_att_pool_bwd_f32");
}

extern "C" int get_att_pool_bwd_bp(const float* t, const float* right, int64_t ld_right, const float* W2,
                                   const float* att, const float* d_pooled, int64_t ld_dpooled, const float* d_att,
                                   int G, int P, int H, int Dr, int C, float* de, void* du_planes, int64_t ld_dup,
                                   int64_t plane_stride, int nplanes, float* du_sum, float* dright, int64_t ld_dright,
                                   int accumulate, void* stream) {
  return att_bwd_common(t, right, ld_right, W2, att, d_pooled, ld_dpooled, d_att, G, P, H, Dr, C, de, nullptr, du_planes, ld_dup,
                        plane_stride, nplanes, du_sum, dright, ld_dright, accumulate, stream, "get_att_pool_bwd_bp");
}
