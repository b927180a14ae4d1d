// Multi-head additive attention pooling: the tail of ConcatNotEqualSelfAtt
// (reference thirdparty/two_branches_attention.py:141-147) and MultiHeadSelfAttentionICLR2017Extend
// (thirdparty/self_attention.py:90-96), forward and backward.
// One CTA per group (an evidence at word level, a claim at evidence level); all reductions over positions
// are warp-shuffle or fixed-order serial sums (deterministic).
#include "common.cuh"

namespace getb {

constexpr int ATT_THREADS = 1024;     // the kernels are latency-bound chains per group: wide CTAs shorten every serial loop
constexpr int ATT_MAX_PARTS = 8;
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
  int accumulate;
};

// fwd smem: W2 (C*H) | e/att (P*C) | pooled partials (parts*Dr*C)
__global__ void __launch_bounds__(ATT_THREADS) att_pool_fwd_kernel(const __grid_constant__ AttParams p) {
  extern __shared__ __align__(16) float smem[];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = p.P, H = p.H, Dr = p.Dr, C = p.C;
  float* sW2 = smem;
  float* sE = smem + C * H;
  for (int q = tid; q < C * H; q += ATT_THREADS) sW2[q] = __ldg(p.W2 + q);
  __syncthreads();
  const float* tg = p.t + (int64_t)g * P * H;
  // e[p,c] = t[p,:] . W2[c,:]
  for (int pp = warp; pp < P; pp += ATT_WARPS) {
    float acc[ATT_MAX_HEADS];
#pragma unroll
    for (int c = 0; c < ATT_MAX_HEADS; ++c) acc[c] = 0.f;
    const float* row = tg + (int64_t)pp * H;
    for (int h = lane; h < H; h += 32) {
      const float tv = __ldg(row + h);
#pragma unroll
      for (int c = 0; c < ATT_MAX_HEADS; ++c)
        if (c < C) acc[c] = fmaf(tv, sW2[c * H + h], acc[c]);
    }
    const bool valid = p.mask[(int64_t)g * P + pp] != 0;
#pragma unroll
    for (int c = 0; c < ATT_MAX_HEADS; ++c) {
      if (c < C) {
        const float v = warp_sum(acc[c]);
        if (lane == 0) sE[pp * C + c] = valid ? v : -INFINITY;
      }
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
  // pooled[d,c] = sum_p right[p,d] * att[p,c]; positions are split over `parts` thread groups, partials reduced in smem
  const float* rg = p.right + (int64_t)g * P * p.ld_right;
  float* og = p.pooled + (int64_t)g * p.ld_pooled;
  const int dpad = (Dr + 31) & ~31;
  const int parts = max(1, min(ATT_MAX_PARTS, min(ATT_THREADS / dpad, P)));
  if (parts == 1) {
    for (int d = tid; d < Dr; d += ATT_THREADS) {
      float acc[ATT_MAX_HEADS];
#pragma unroll
      for (int c = 0; c < ATT_MAX_HEADS; ++c) acc[c] = 0.f;
      for (int pp = 0; pp < P; ++pp) {
        const float rv = __ldg(rg + (int64_t)pp * p.ld_right + d);
#pragma unroll
        for (int c = 0; c < ATT_MAX_HEADS; ++c)
          if (c < C) acc[c] = fmaf(rv, sE[pp * C + c], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < ATT_MAX_HEADS; ++c)
        if (c < C) og[(int64_t)d * C + c] = acc[c];
    }
  } else {
    float* sPart = sE + P * C;                       // [parts][Dr][C]
    const int part = tid / dpad, d = tid % dpad;
    if (part < parts && d < Dr) {
      const int per = (P + parts - 1) / parts;
      const int p0 = part * per, p1 = min(P, p0 + per);
      float acc[ATT_MAX_HEADS];
#pragma unroll
      for (int c = 0; c < ATT_MAX_HEADS; ++c) acc[c] = 0.f;
      for (int pp = p0; pp < p1; ++pp) {
        const float rv = __ldg(rg + (int64_t)pp * p.ld_right + d);
#pragma unroll
        for (int c = 0; c < ATT_MAX_HEADS; ++c)
          if (c < C) acc[c] = fmaf(rv, sE[pp * C + c], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < ATT_MAX_HEADS; ++c)
        if (c < C) sPart[((size_t)part * Dr + d) * C + c] = acc[c];
    }
    __syncthreads();
    for (int q = tid; q < Dr * C; q += ATT_THREADS) {
      float v = 0.f;
      for (int part2 = 0; part2 < parts; ++part2) v += sPart[(size_t)part2 * Dr * C + q];   // fixed order: deterministic
      og[q] = v;
    }
  }
}

// bwd smem: W2 (C*H) | dO (Dr*C) | att (P*C) | de (P*C) | dot (C) | du_sum partials (parts*H)
__global__ void __launch_bounds__(ATT_THREADS) att_pool_bwd_kernel(const __grid_constant__ AttParams p) {
  extern __shared__ __align__(16) float smem[];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = p.P, H = p.H, Dr = p.Dr, C = p.C;
  float* sW2 = smem;
  float* sdO = sW2 + C * H;
  float* sAtt = sdO + Dr * C;
  float* sDe = sAtt + P * C;
  float* sDot = sDe + P * C;
  for (int q = tid; q < C * H; q += ATT_THREADS) sW2[q] = __ldg(p.W2 + q);
  const float* dOg = p.d_pooled + (int64_t)g * p.ld_dpooled;
  for (int q = tid; q < Dr * C; q += ATT_THREADS) sdO[q] = __ldg(dOg + q);
  for (int q = tid; q < P * C; q += ATT_THREADS) sAtt[q] = __ldg(p.att_in + (int64_t)g * P * C + q);
  __syncthreads();
  const float* rg = p.right + (int64_t)g * P * p.ld_right;
  // dalpha[p,c] = right[p,:] . dO[:,c] (+ d_att)
  for (int pp = warp; pp < P; pp += ATT_WARPS) {
    float acc[ATT_MAX_HEADS];
#pragma unroll
    for (int c = 0; c < ATT_MAX_HEADS; ++c) acc[c] = 0.f;
    const float* row = rg + (int64_t)pp * p.ld_right;
    for (int d = lane; d < Dr; d += 32) {
      const float rv = __ldg(row + d);
#pragma unroll
      for (int c = 0; c < ATT_MAX_HEADS; ++c)
        if (c < C) acc[c] = fmaf(rv, sdO[d * C + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < ATT_MAX_HEADS; ++c) {
      if (c < C) {
        float v = warp_sum(acc[c]);
        if (p.d_att) v += __ldg(p.d_att + ((int64_t)g * P + pp) * C + c);
        if (lane == 0) sDe[pp * C + c] = v;  // holds dalpha for now
      }
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
  // du[p,h] = (de[p,:] @ W2[:,h]) * (1 - t^2);  du_sum[h] = sum_p du[p,h]
  const float* tg = p.t + (int64_t)g * P * H;
  float* dug = p.du + (int64_t)g * P * H;
  {
    const int hpad = (H + 31) & ~31;
    const int parts = max(1, min(ATT_MAX_PARTS, min(ATT_THREADS / hpad, P)));
    float* sSum = sDot + C;                          // [parts][H] partial du_sum
    const int per = (P + parts - 1) / parts;
    for (int idx = tid; idx < parts * hpad; idx += ATT_THREADS) {
      const int part = idx / hpad, h = idx % hpad;
      if (h >= H) continue;
      float w[ATT_MAX_HEADS];
#pragma unroll
      for (int c = 0; c < ATT_MAX_HEADS; ++c) w[c] = c < C ? sW2[c * H + h] : 0.f;
      float sum = 0.f;
      const int p0 = part * per, p1 = min(P, p0 + per);
      for (int pp = p0; pp < p1; ++pp) {
        float dt = 0.f;
#pragma unroll
        for (int c = 0; c < ATT_MAX_HEADS; ++c)
          if (c < C) dt = fmaf(sDe[pp * C + c], w[c], dt);
        const float tv = __ldg(tg + (int64_t)pp * H + h);
        const float duv = dt * (1.0f - tv * tv);
        dug[(int64_t)pp * H + h] = duv;
        sum += duv;
      }
      sSum[part * H + h] = sum;
    }
    __syncthreads();
    if (p.du_sum) {
      for (int h = tid; h < H; h += ATT_THREADS) {
        float v = 0.f;
        for (int part = 0; part < parts; ++part) v += sSum[part * H + h];
        p.du_sum[(int64_t)g * H + h] = v;
      }
    }
  }
  // dright[p,d] (+)= sum_c att[p,c] * dO[d,c]
  float* drg = p.dright + (int64_t)g * P * p.ld_dright;
  for (int q = tid; q < P * Dr; q += ATT_THREADS) {
    const int pp = q / Dr, d = q % Dr;
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < ATT_MAX_HEADS; ++c)
      if (c < C) v = fmaf(sAtt[pp * C + c], sdO[d * C + c], v);
    float* dst = drg + (int64_t)pp * p.ld_dright + d;
    *dst = p.accumulate ? *dst + v : v;
  }
}

static int check_att(const char* name, int G, int P, int H, int Dr, int C) {
  GETB_REQUIRE(G >= 0 && P > 0 && H > 0 && Dr > 0, "%s: bad sizes", name);
  GETB_REQUIRE(C >= 1 && C <= ATT_MAX_HEADS, "%s: heads=%d not in [1,%d]", name, C, ATT_MAX_HEADS);
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
  const size_t smem = ((size_t)C * H + (size_t)P * C + (size_t)ATT_MAX_PARTS * Dr * C) * sizeof(float);
  GETB_REQUIRE(smem <= 200 * 1024, "get_att_pool_fwd_f32: shared memory %zu too large", smem);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(att_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  att_pool_fwd_kernel<<<G, ATT_THREADS, smem, (cudaStream_t)stream>>>(p);
  GETB_CHECK_LAUNCH("get_att_pool_fwd_f32");
  return 0;
}

extern "C" int get_att_pool_bwd_f32(const float* t, const float* right, int64_t ld_right, const float* W2,
                                    const float* att, const float* d_pooled, int64_t ld_dpooled, const float* d_att,
                                    int G, int P, int H, int Dr, int C, float* de, float* du, float* du_sum,
                                    float* dright, int64_t ld_dright, int accumulate, void* stream) {
  GETB_REQUIRE(t && right && W2 && att && d_pooled && de && du && dright, "get_att_pool_bwd_f32: null pointer");
  if (check_att("get_att_pool_bwd_f32", G, P, H, Dr, C)) return -1;
  if (G == 0) return 0;
  AttParams p;
  memset(&p, 0, sizeof(p));
  p.t = t; p.right = right; p.ld_right = ld_right; p.W2 = W2; p.att_in = att;
  p.d_pooled = d_pooled; p.ld_dpooled = ld_dpooled; p.d_att = d_att;
  p.G = G; p.P = P; p.H = H; p.Dr = Dr; p.C = C;
  p.de = de; p.du = du; p.du_sum = du_sum; p.dright = dright; p.ld_dright = ld_dright; p.accumulate = accumulate;
  const size_t smem = ((size_t)C * H + (size_t)Dr * C + (size_t)2 * P * C + C + (size_t)ATT_MAX_PARTS * H) * sizeof(float);
  GETB_REQUIRE(smem <= 200 * 1024, "get_att_pool_bwd_f32: shared memory %zu too large", smem);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(att_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  att_pool_bwd_kernel<<<G, ATT_THREADS, smem, (cudaStream_t)stream>>>(p);
  GETB_CHECK_LAUNCH("get_att_pool_bwd_f32");
  return 0;
}
