// Per-graph kernels of the GET hot path (HBM-bound part).
//
//  * graph_kernel<FUSED=false>: out[g] (+)= op(adj'[g]) @ x[g]                       reference Models/BiDAF/wrapper.py:192
//  * graph_kernel<FUSED=true> : node scorer (GGNN with out_features=1, wrapper.py:167) -> top-k keep set
//                               (wrapper.py:215-219) -> refined aggregation (wrapper.py:221-225 + :192)
//                               in ONE pass; the dense (N,N) mask of the reference is never materialised.
//  * gsl_mask_adj_kernel      : stand-alone GSL.forward (wrapper.py:215-227) for the op-level surface.
//
// One CTA per graph. The adjacency tile (N x N fp32) is staged in shared memory once (row stride padded to an
// odd number of words so row- and column-wise scans are bank-conflict free); node-feature rows are streamed
// with 128-bit loads, one warp per output row, skipping zero adjacency entries by warp ballot.
#include "common.cuh"

namespace getb {

constexpr int GRAPH_THREADS = 256;
constexpr int GRAPH_WARPS = GRAPH_THREADS / 32;
constexpr int MAX_QUADS_PER_LANE = 8;  // H <= 32*4*8 = 1024

struct GraphParams {
  const float* adj;    // (G,N,N)
  const float* x;      // (G,N,H)
  const uint8_t* keep_in;
  float* out;          // (G,N,H)
  int G, N, H, NP;     // NP = padded smem row stride
  int transpose, accumulate;
  int adj_in_smem;
  // fused part
  const float* wp;     // (H)
  const float* gate;   // (12)
  int k;
  uint32_t thr;        // dropout threshold (0 = eval)
  float scale;
  uint32_t seed_s, seed_2;
  float* score;        // (G,N) or null
  uint8_t* keep_out;   // (G,N)
};

// smem layout: [adj N*NP floats (if adj_in_smem)] [sp N] [score N] [keep N bytes (padded)]
template <bool FUSED>
__global__ void __launch_bounds__(GRAPH_THREADS) graph_kernel(const __grid_constant__ GraphParams p) {
  extern __shared__ __align__(16) float smem[];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, H = p.H, NP = p.NP;
  const float* __restrict__ gadj = p.adj + (int64_t)g * N * N;
  const float* __restrict__ gx = p.x + (int64_t)g * N * H;
  float* __restrict__ gout = p.out + (int64_t)g * N * H;

  float* sadj = smem;
  float* s_sp = smem + (p.adj_in_smem ? (size_t)N * NP : 0);
  float* s_score = s_sp + N;
  uint8_t* s_keep = reinterpret_cast<uint8_t*>(s_score + N);

  // ---- stage adjacency -----------------------------------------------------------------------
  if (p.adj_in_smem) {
    if ((N & 3) == 0) {
      const int nq = N / 4;
      for (int q = tid; q < N * nq; q += GRAPH_THREADS) {
        const int i = q / nq, j = (q % nq) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(gadj + (int64_t)i * N + j));
        float* d = sadj + i * NP + j;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      for (int q = tid; q < N * N; q += GRAPH_THREADS) sadj[(q / N) * NP + (q % N)] = __ldg(gadj + q);
    }
  }
  const float* A = p.adj_in_smem ? sadj : gadj;
  const int lda = p.adj_in_smem ? NP : N;

  const bool vecH = ((H & 3) == 0) && aligned16(p.x) && aligned16(p.out);
  const int HQ = vecH ? H / 4 : 0;

  if (FUSED) {
    // ---- s_p[i] = drop_s(F[i,:]) . wp ----------------------------------------------------------
    for (int i = warp; i < N; i += GRAPH_WARPS) {
      float acc = 0.f;
      const float* row = gx + (int64_t)i * H;
      if (vecH) {
        for (int q = lane; q < HQ; q += 32) {
          float4 f = __ldg(reinterpret_cast<const float4*>(row) + q);
          const float4 w = __ldg(reinterpret_cast<const float4*>(p.wp) + q);
          if (p.thr) {
            const uint64_t base = ((uint64_t)g * N + i) * (uint64_t)H + (uint64_t)q * 4;
            f.x = drop_keep(p.seed_s, base + 0, p.thr) ? f.x * p.scale : 0.f;
            f.y = drop_keep(p.seed_s, base + 1, p.thr) ? f.y * p.scale : 0.f;
            f.z = drop_keep(p.seed_s, base + 2, p.thr) ? f.z * p.scale : 0.f;
            f.w = drop_keep(p.seed_s, base + 3, p.thr) ? f.w * p.scale : 0.f;
          }
          acc = fmaf(f.x, w.x, acc); acc = fmaf(f.y, w.y, acc);
          acc = fmaf(f.z, w.z, acc); acc = fmaf(f.w, w.w, acc);
        }
      } else {
        for (int c = lane; c < H; c += 32) {
          float f = __ldg(row + c);
          if (p.thr) f = drop_keep(p.seed_s, ((uint64_t)g * N + i) * (uint64_t)H + c, p.thr) ? f * p.scale : 0.f;
          acc = fmaf(f, __ldg(p.wp + c), acc);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) s_sp[i] = acc;
    }
    __syncthreads();
    // ---- s_a = adj @ s_p ; scalar GRU gates (GGNN with out_features = 1) ------------------------
    const float wz0 = __ldg(p.gate + 0), bz0 = __ldg(p.gate + 1), wz1 = __ldg(p.gate + 2), bz1 = __ldg(p.gate + 3);
    const float wr0 = __ldg(p.gate + 4), br0 = __ldg(p.gate + 5), wr1 = __ldg(p.gate + 6), br1 = __ldg(p.gate + 7);
    const float wh0 = __ldg(p.gate + 8), bh0 = __ldg(p.gate + 9), wh1 = __ldg(p.gate + 10), bh1 = __ldg(p.gate + 11);
    for (int i = tid; i < N; i += GRAPH_THREADS) {
      float sa = 0.f;
      const float* ar = A + (size_t)i * lda;
      for (int j = 0; j < N; ++j) sa = fmaf(ar[j], s_sp[j], sa);
      const float sp = s_sp[i];
      const float z = sigmoidf_((wz0 * sa + bz0) + (wz1 * sp + bz1));
      const float r = sigmoidf_((wr0 * sa + br0) + (wr1 * sp + br1));
      const float h = tanhf((wh0 * sa + bh0) + (wh1 * (r * sp) + bh1));
      const float sc = h * z + sp * (1.0f - z);
      s_score[i] = sc;
      if (p.score) p.score[(int64_t)g * N + i] = sc;
    }
    __syncthreads();
    // ---- top-k by rank counting; ties -> lower index first ---------------------------------------
    for (int i = tid; i < N; i += GRAPH_THREADS) {
      const float si = s_score[i];
      int rank = 0;
      for (int j = 0; j < N; ++j) {
        const float sj = s_score[j];
        rank += (sj > si) || (sj == si && j < i);
      }
      const uint8_t kp = rank < p.k;
      s_keep[i] = kp;
      p.keep_out[(int64_t)g * N + i] = kp;
    }
    __syncthreads();
  } else {
    if (p.keep_in)
      for (int i = tid; i < N; i += GRAPH_THREADS) s_keep[i] = p.keep_in[(int64_t)g * N + i];
    __syncthreads();
  }
  const bool masked = FUSED || (p.keep_in != nullptr);
  const bool drop2 = FUSED && p.thr != 0;

  // ---- out[i,:] = sum_j adj'[i,j] * x[j,:]  (one warp per output row) ----------------------------
  for (int i = warp; i < N; i += GRAPH_WARPS) {
    float4 acc[MAX_QUADS_PER_LANE];
#pragma unroll
    for (int u = 0; u < MAX_QUADS_PER_LANE; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool keep_i = masked ? (s_keep[i] != 0) : true;
    for (int c0 = 0; c0 < N; c0 += 32) {
      const int j = c0 + lane;
      float w = 0.f;
      if (j < N) {
        w = p.transpose ? A[(size_t)j * lda + i] : A[(size_t)i * lda + j];
        if (masked && !keep_i && !s_keep[j]) w = 0.f;
      }
      unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
      while (nz) {
        const int b = __ffs(nz) - 1;
        nz &= nz - 1;
        const float wj = __shfl_sync(0xffffffffu, w, b);
        const int jj = c0 + b;
        const float* row = gx + (int64_t)jj * H;
        if (vecH) {
#pragma unroll
          for (int u = 0; u < MAX_QUADS_PER_LANE; ++u) {
            const int q = lane + u * 32;
            if (q < HQ) {
              float4 f = __ldg(reinterpret_cast<const float4*>(row) + q);
              if (drop2) {
                const uint64_t base = ((uint64_t)g * N + jj) * (uint64_t)H + (uint64_t)q * 4;
                f.x = drop_keep(p.seed_2, base + 0, p.thr) ? f.x * p.scale : 0.f;
                f.y = drop_keep(p.seed_2, base + 1, p.thr) ? f.y * p.scale : 0.f;
                f.z = drop_keep(p.seed_2, base + 2, p.thr) ? f.z * p.scale : 0.f;
                f.w = drop_keep(p.seed_2, base + 3, p.thr) ? f.w * p.scale : 0.f;
              }
              acc[u].x = fmaf(wj, f.x, acc[u].x); acc[u].y = fmaf(wj, f.y, acc[u].y);
              acc[u].z = fmaf(wj, f.z, acc[u].z); acc[u].w = fmaf(wj, f.w, acc[u].w);
            }
          }
        } else {
          // scalar fallback: lane owns columns lane, lane+32, ... (up to 4*MAX_QUADS_PER_LANE of them)
#pragma unroll
          for (int u = 0; u < MAX_QUADS_PER_LANE; ++u) {
            float* a4 = reinterpret_cast<float*>(&acc[u]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = lane + (u * 4 + e) * 32;
              if (c < H) {
                float f = __ldg(row + c);
                if (drop2) f = drop_keep(p.seed_2, ((uint64_t)g * N + jj) * (uint64_t)H + c, p.thr) ? f * p.scale : 0.f;
                a4[e] = fmaf(wj, f, a4[e]);
              }
            }
          }
        }
      }
    }
    float* orow = gout + (int64_t)i * H;
    if (vecH) {
#pragma unroll
      for (int u = 0; u < MAX_QUADS_PER_LANE; ++u) {
        const int q = lane + u * 32;
        if (q < HQ) {
          float4 v = acc[u];
          if (p.accumulate) {
            const float4 o = *(reinterpret_cast<const float4*>(orow) + q);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          *(reinterpret_cast<float4*>(orow) + q) = v;
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < MAX_QUADS_PER_LANE; ++u) {
        const float* a4 = reinterpret_cast<const float*>(&acc[u]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = lane + (u * 4 + e) * 32;
          if (c < H) orow[c] = p.accumulate ? orow[c] + a4[e] : a4[e];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(GRAPH_THREADS) gsl_mask_adj_kernel(const float* __restrict__ adj,
                                                                    const float* __restrict__ score, int N, int k,
                                                                    float* __restrict__ adj_out,
                                                                    uint8_t* __restrict__ keep) {
  extern __shared__ __align__(16) float smem[];
  float* s_score = smem;
  uint8_t* s_keep = reinterpret_cast<uint8_t*>(smem + N);
  const int g = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < N; i += GRAPH_THREADS) s_score[i] = score[(int64_t)g * N + i];
  __syncthreads();
  for (int i = tid; i < N; i += GRAPH_THREADS) {
    const float si = s_score[i];
    int rank = 0;
    for (int j = 0; j < N; ++j) {
      const float sj = s_score[j];
      rank += (sj > si) || (sj == si && j < i);
    }
    s_keep[i] = rank < k;
    if (keep) keep[(int64_t)g * N + i] = rank < k;
  }
  __syncthreads();
  const float* a = adj + (int64_t)g * N * N;
  float* o = adj_out + (int64_t)g * N * N;
  for (int q = tid; q < N * N; q += GRAPH_THREADS) {
    const int i = q / N, j = q % N;
    o[q] = (s_keep[i] | s_keep[j]) ? a[q] : 0.f;
  }
}

static int launch_graph(GraphParams& p, bool fused, cudaStream_t st, const char* name) {
  GETB_REQUIRE(p.G >= 0 && p.N > 0 && p.H > 0, "%s: bad sizes G=%d N=%d H=%d", name, p.G, p.N, p.H);
  GETB_REQUIRE(p.H <= 32 * 4 * MAX_QUADS_PER_LANE, "%s: H=%d exceeds %d", name, p.H, 32 * 4 * MAX_QUADS_PER_LANE);
  if (p.G == 0) return 0;
  p.NP = p.N | 1;
  const size_t tail = (size_t)2 * p.N * sizeof(float) + ((p.N + 15) / 16) * 16;
  size_t smem = (size_t)p.N * p.NP * sizeof(float) + tail;
  p.adj_in_smem = 1;
  if (smem > 200 * 1024) {  // adjacency does not fit: read it through L2 instead
    p.adj_in_smem = 0;
    smem = tail;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(graph_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 4096);
    cudaFuncSetAttribute(graph_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 4096);
    attr_set = true;
  }
  if (fused)
    graph_kernel<true><<<p.G, GRAPH_THREADS, smem, st>>>(p);
  else
    graph_kernel<false><<<p.G, GRAPH_THREADS, smem, st>>>(p);
  GETB_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace getb

using namespace getb;

extern "C" int get_graph_aggregate_f32(const float* adj, const float* x, const uint8_t* keep, float* out, int G, int N,
                                       int H, int transpose, int accumulate, void* stream) {
  GETB_REQUIRE(adj && x && out, "get_graph_aggregate_f32: null pointer");
  GraphParams p;
  memset(&p, 0, sizeof(p));
  p.adj = adj; p.x = x; p.keep_in = keep; p.out = out;
  p.G = G; p.N = N; p.H = H; p.transpose = transpose; p.accumulate = accumulate;
  return launch_graph(p, false, (cudaStream_t)stream, "get_graph_aggregate_f32");
}

extern "C" int get_gsl_fused_f32(const float* adj, const float* F, const float* wp, const float* gate, int G, int N,
                                 int H, int k, float drop_p, uint32_t seed_scorer, uint32_t seed_layer2, float* score,
                                 uint8_t* keep, float* out, void* stream) {
  GETB_REQUIRE(adj && F && wp && gate && keep && out, "get_gsl_fused_f32: null pointer");
  GETB_REQUIRE(k >= 0 && k <= N, "get_gsl_fused_f32: k=%d out of [0,%d]", k, N);
  GETB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "get_gsl_fused_f32: dropout probability must be in [0,1)");
  GETB_REQUIRE(aligned16(wp) || (H & 3), "get_gsl_fused_f32: wp must be 16-byte aligned");
  GraphParams p;
  memset(&p, 0, sizeof(p));
  p.adj = adj; p.x = F; p.out = out; p.G = G; p.N = N; p.H = H;
  p.wp = wp; p.gate = gate; p.k = k;
  p.thr = drop_p > 0.f ? drop_threshold(drop_p) : 0;
  p.scale = 1.0f / (1.0f - drop_p);
  p.seed_s = seed_scorer; p.seed_2 = seed_layer2;
  p.score = score; p.keep_out = keep;
  return launch_graph(p, true, (cudaStream_t)stream, "get_gsl_fused_f32");
}

extern "C" int get_gsl_mask_adj_f32(const float* adj, const float* score, int G, int N, int k, float* adj_out,
                                    uint8_t* keep, void* stream) {
  GETB_REQUIRE(adj && score && adj_out, "get_gsl_mask_adj_f32: null pointer");
  GETB_REQUIRE(N > 0 && k >= 0 && k <= N, "get_gsl_mask_adj_f32: bad N/k");
  if (G == 0) return 0;
  const size_t smem = (size_t)N * sizeof(float) + ((N + 15) / 16) * 16;
  gsl_mask_adj_kernel<<<G, GRAPH_THREADS, smem, (cudaStream_t)stream>>>(adj, score, N, k, adj_out, keep);
  GETB_CHECK_LAUNCH("get_gsl_mask_adj_f32");
  return 0;
}
