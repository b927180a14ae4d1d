// Per-graph kernels of the GET hot path on the DENSE adjacency: the op-level entry points (get_graph_aggregate_f32 / _bp,
// get_gsl_fused_f32 / _bp, get_gsl_mask_adj_f32) and the fallback of the model for shapes the neighbour-list kernels do not
// cover (N > 232, feature widths that are not multiples of 4). The model's default graph path is graph_lists.cu.
//
//  * <FUSED=false>: out[g] (+)= op(adj'[g]) @ x[g]                                   reference Models/BiDAF/wrapper.py:192
//  * <FUSED=true> : node scorer (GGNN with out_features=1, wrapper.py:167) -> top-k keep set
//                   (wrapper.py:215-219) -> refined aggregation (wrapper.py:221-225 + :192)
//                   in ONE pass; the dense (N,N) mask of the reference is never materialised.
//  * gsl_mask_adj_kernel: stand-alone GSL.forward (wrapper.py:215-227) for the op-level surface.
//
// Two implementations behind the same entry points, one CTA per graph:
//  * graph_smem_kernel (the fast path, taken whenever one graph fits a CTA's shared memory -- Snopes / PolitiFact
//    shapes: 100 x 300 features = 120 KB + 100 x 100 adjacency = 40 KB): the adjacency and the node features are
//    staged ONCE by TMA bulk copies (cp.async.bulk + mbarrier complete_tx; features in row chunks so the scorer
//    starts on the first chunk while the rest is in flight); scorer dot products, SpMV, top-k, dropout and the
//    neighbour gather-and-weighted-sum then run entirely out of shared memory with 128-bit accesses and warp
//    shuffles; every HBM byte of the graph is read exactly once and every output row written once, coalesced.
//  * graph_kernel (generic fallback, any N/H/alignment): adjacency in shared memory when it fits, feature rows
//    streamed through L2 with 128-bit loads.
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

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
  const uint32_t* salt; // device word added to both seeds inside the kernels
  float* score;        // (G,N) or null
  uint8_t* keep_out;   // (G,N)
  // optional bf16-plane copy of the output rows (operand of the next tensor-core contraction); `out` may then be null
  __nv_bfloat16* out_p;
  int64_t ld_p, ps_p;
  int np_p, pad_one;
};

// smem layout: [adj N*NP floats (if adj_in_smem)] [sp N] [score N] [keep N bytes (padded)]
template <bool FUSED>
__global__ void __launch_bounds__(GRAPH_THREADS) graph_kernel(const __grid_constant__ GraphParams p) {
  extern __shared__ __align__(16) float smem[];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, H = p.H, NP = p.NP;
  const uint32_t salt = (FUSED && p.thr) ? __ldg(p.salt) : 0u;
  const uint32_t seed_s = p.seed_s + salt, seed_2 = p.seed_2 + salt;
  const float* __restrict__ gadj = p.adj + (int64_t)g * N * N;
  const float* __restrict__ gx = p.x + (int64_t)g * N * H;
  float* __restrict__ gout = p.out ? p.out + (int64_t)g * N * H : nullptr;

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
            f.x = drop_keep(seed_s, base + 0, p.thr) ? f.x * p.scale : 0.f;
            f.y = drop_keep(seed_s, base + 1, p.thr) ? f.y * p.scale : 0.f;
            f.z = drop_keep(seed_s, base + 2, p.thr) ? f.z * p.scale : 0.f;
            f.w = drop_keep(seed_s, base + 3, p.thr) ? f.w * p.scale : 0.f;
          }
          acc = fmaf(f.x, w.x, acc); acc = fmaf(f.y, w.y, acc);
          acc = fmaf(f.z, w.z, acc); acc = fmaf(f.w, w.w, acc);
        }
      } else {
        for (int c = lane; c < H; c += 32) {
          float f = __ldg(row + c);
          if (p.thr) f = drop_keep(seed_s, ((uint64_t)g * N + i) * (uint64_t)H + c, p.thr) ? f * p.scale : 0.f;
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
                f.x = drop_keep(seed_2, base + 0, p.thr) ? f.x * p.scale : 0.f;
                f.y = drop_keep(seed_2, base + 1, p.thr) ? f.y * p.scale : 0.f;
                f.z = drop_keep(seed_2, base + 2, p.thr) ? f.z * p.scale : 0.f;
                f.w = drop_keep(seed_2, base + 3, p.thr) ? f.w * p.scale : 0.f;
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
                if (drop2) f = drop_keep(seed_2, ((uint64_t)g * N + jj) * (uint64_t)H + c, p.thr) ? f * p.scale : 0.f;
                a4[e] = fmaf(wj, f, a4[e]);
              }
            }
          }
        }
      }
    }
    float* orow = gout + (int64_t)i * H;
    if (vecH) {
      __nv_bfloat16* prow = p.out_p ? p.out_p + ((int64_t)g * N + i) * p.ld_p : nullptr;
#pragma unroll
      for (int u = 0; u < MAX_QUADS_PER_LANE; ++u) {
        const int q = lane + u * 32;
        if (q < HQ) {
          float4 v = acc[u];
          if (p.accumulate) {
            const float4 o = *(reinterpret_cast<const float4*>(orow) + q);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          if (gout) *(reinterpret_cast<float4*>(orow) + q) = v;
          if (prow) {
            const float vv[4] = {v.x, v.y, v.z, v.w};
            planes_store4(prow + q * 4, p.ps_p, p.np_p, vv);
          }
        }
      }
      if (prow) {   // padding quads up to a multiple of 8 columns (room for the ones column when pad_one)
        const int npq = ((((H + (p.pad_one ? 1 : 0)) + 7) & ~7) - H) >> 2;
        if (lane < npq) {
          const float vv[4] = {(p.pad_one && lane == 0) ? 1.0f : 0.0f, 0.f, 0.f, 0.f};
          planes_store4(prow + H + lane * 4, p.ps_p, p.np_p, vv);
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

// =====================================================================================================
// Fast path: whole graph resident in shared memory, staged by TMA bulk copies.
// =====================================================================================================
constexpr int GS_THREADS = 1024;
constexpr int GS_WARPS = GS_THREADS / 32;
constexpr int GS_MAX_N = 128;           // <= 4 adjacency chunks of 32 columns, <= 4 rows per warp
constexpr int GS_CHUNKS = GS_MAX_N / 32;  // feature rows arrive in bulk copies of 32 rows, one mbarrier each
constexpr int GS_MAX_QUADS = 4;         // H <= 32*4*4 = 512 on the fast path
constexpr size_t GS_SMEM_LIMIT = 226 * 1024;

__device__ __forceinline__ uint32_t gs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gs_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void gs_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gs_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gs_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = gs_smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();   // a lost copy must fail loudly, never hang the GPU
  }
}
// global -> shared bulk copy (TMA engine, no tensor map needed for a contiguous block); bytes % 16 == 0
__device__ __forceinline__ void gs_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(gs_smem_u32(dst)), "l"(src), "r"(bytes), "r"(gs_smem_u32(bar))
               : "memory");
}

// smem layout: [F N*H f32] [L: N*N x 8 B -- the dense adjacency lands in its first half, then per-row neighbour lists
//              {byte offset of row j in F, weight} overwrite it] [sp N] [score N] [cnt N i32] [rank N i32] [keep N u8] [mbarriers]
// NQ = ceil(H/4/32): float4 "quads" of a feature row owned by each lane (compile-time so the row loops unroll exactly)
template <bool FUSED, int NQ>
__global__ void __launch_bounds__(GS_THREADS, 1) graph_smem_kernel(const __grid_constant__ GraphParams p) {
  extern __shared__ __align__(16) float smem[];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, H = p.H, HQ = H >> 2;
  const uint32_t salt = (FUSED && p.thr) ? __ldg(p.salt) : 0u;
  const uint32_t seed_s = p.seed_s + salt, seed_2 = p.seed_2 + salt;
  float* sF = smem;
  float2* sL = reinterpret_cast<float2*>(sF + (size_t)N * H);      // N*N entries {offset bits, weight}
  float* sA = reinterpret_cast<float*>(sL);                        // dense adjacency (first half of the list region)
  float* s_sp = reinterpret_cast<float*>(sL + (size_t)N * N);
  float* s_score = s_sp + N;
  int* s_cnt = reinterpret_cast<int*>(s_score + N);
  int* s_rank = s_cnt + N;
  uint8_t* s_keep = reinterpret_cast<uint8_t*>(s_rank + N);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_keep + ((N + 15) & ~15));   // [0] adjacency, [1..GS_CHUNKS] features

  const float* __restrict__ gadj = p.adj + (int64_t)g * N * N;
  const float* __restrict__ gx = p.x + (int64_t)g * N * H;
  float* __restrict__ gout = p.out ? p.out + (int64_t)g * N * H : nullptr;
  const bool last_ok = (lane + (NQ - 1) * 32) < HQ;   // does this lane own a quad in the last (partial) group?
  const int nchunks = (N + 31) >> 5;

  if (tid == 0) {
    for (int b = 0; b <= GS_CHUNKS; ++b) gs_mbar_init(&bars[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t abytes = (uint32_t)N * (uint32_t)N * 4u;
    gs_mbar_expect_tx(&bars[0], abytes);
    gs_bulk_g2s(sA, gadj, abytes, &bars[0]);
    for (int c = 0; c < nchunks; ++c) {
      const int r0 = c * 32;
      const uint32_t bytes = (uint32_t)(min(N, r0 + 32) - r0) * (uint32_t)H * 4u;
      gs_mbar_expect_tx(&bars[1 + c], bytes);
      gs_bulk_g2s(sF + (size_t)r0 * H, gx + (size_t)r0 * H, bytes, &bars[1 + c]);
    }
  }
  if (tid < N) {
    s_rank[tid] = 0;
    if (!FUSED) s_keep[tid] = p.keep_in ? p.keep_in[(int64_t)g * N + tid] : (uint8_t)1;
  }

  // ---- neighbour lists: row i of op(adj) -> {offset of F row j, w_ij}, e < cnt[i] -----------------------------
  gs_mbar_wait(&bars[0], 0);
  {
    float wv[4][4];   // [row slot][column chunk]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = warp + r * GS_WARPS;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = c * 32 + lane;
        wv[r][c] = (i < N && j < N) ? (p.transpose ? sA[(size_t)j * N + i] : sA[(size_t)i * N + j]) : 0.f;
      }
    }
    __syncthreads();   // every row is in registers before the lists overwrite the dense tile
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = warp + r * GS_WARPS;
      if (i < N) {
        int pos = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const unsigned nz = __ballot_sync(0xffffffffu, wv[r][c] != 0.f);
          if (wv[r][c] != 0.f) {
            const int e = pos + __popc(nz & ((1u << lane) - 1u));
            sL[(size_t)i * N + e] = make_float2(__int_as_float((c * 32 + lane) * H * 4), wv[r][c]);
          }
          pos += __popc(nz);
        }
        if (lane == 0) s_cnt[i] = pos;
      }
    }
  }

  if (FUSED) {
    // ---- s_p[i] = drop_s(F[i,:]) . wp ; the layer-2 dropout draw is applied in place on the way ----------------
    float4 wq[NQ];
#pragma unroll
    for (int u = 0; u < NQ; ++u) {
      const int q = lane + u * 32;
      wq[u] = q < HQ ? __ldg(reinterpret_cast<const float4*>(p.wp) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int c = 0; c < nchunks; ++c) {
      gs_mbar_wait(&bars[1 + c], 0);
      const int i = c * 32 + warp;
      if (i < N) {
        float4* row = reinterpret_cast<float4*>(sF + (size_t)i * H) + lane;
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
          if (u < NQ - 1 || last_ok) {
            float4 f = row[u * 32];
            if (p.thr) {
              const uint64_t base = ((uint64_t)g * N + i) * (uint64_t)H + (uint64_t)(lane + u * 32) * 4;
              float4 f2 = f;
              drop_apply4(seed_2, base, p.thr, p.scale, f2);
              row[u * 32] = f2;
              drop_apply4(seed_s, base, p.thr, p.scale, f);
            }
            acc = fmaf(f.x, wq[u].x, acc); acc = fmaf(f.y, wq[u].y, acc);
            acc = fmaf(f.z, wq[u].z, acc); acc = fmaf(f.w, wq[u].w, acc);
          }
        }
        acc = warp_sum(acc);
        if (lane == 0) s_sp[i] = acc;
      }
    }
    __syncthreads();
    // ---- s_a = adj @ s_p over the neighbour lists + scalar GRU gates (GGNN with out_features = 1) --------------
    if (tid < N) {
      const int i = tid;
      const float2* lr = sL + (size_t)i * N;
      const int cnt = s_cnt[i];
      const float inv_rowbytes = 1.0f / (float)(H * 4);   // offsets are exact multiples of H*4 < 2^24: rounding recovers j
      float sa = 0.f;
      for (int e = 0; e < cnt; ++e) {
        const float2 en = lr[e];
        sa = fmaf(en.y, s_sp[__float2int_rn(__int2float_rn(__float_as_int(en.x)) * inv_rowbytes)], sa);
      }
      const float wz0 = __ldg(p.gate + 0), bz0 = __ldg(p.gate + 1), wz1 = __ldg(p.gate + 2), bz1 = __ldg(p.gate + 3);
      const float wr0 = __ldg(p.gate + 4), br0 = __ldg(p.gate + 5), wr1 = __ldg(p.gate + 6), br1 = __ldg(p.gate + 7);
      const float wh0 = __ldg(p.gate + 8), bh0 = __ldg(p.gate + 9), wh1 = __ldg(p.gate + 10), bh1 = __ldg(p.gate + 11);
      const float sp = s_sp[i];
      const float z = sigmoidf_((wz0 * sa + bz0) + (wz1 * sp + bz1));
      const float r = sigmoidf_((wr0 * sa + br0) + (wr1 * sp + br1));
      const float h = tanhf((wh0 * sa + bh0) + (wh1 * (r * sp) + bh1));
      const float sc = h * z + sp * (1.0f - z);
      s_score[i] = sc;
      if (p.score) p.score[(int64_t)g * N + i] = sc;
    }
    __syncthreads();
    // ---- top-k by rank counting: thread (node i, slice of 16 candidates), partial ranks added in shared memory;
    //      ties -> lower index first -----------------------------------------------------------------------------
    {
      const int i = tid & (GS_MAX_N - 1);
      const int j0 = (tid >> 7) * (GS_MAX_N / (GS_THREADS / GS_MAX_N));   // 8 slices of 16
      if (i < N && j0 < N) {
        const float si = s_score[i];
        const int j1 = min(N, j0 + GS_MAX_N / (GS_THREADS / GS_MAX_N));
        int rank = 0;
        for (int j = j0; j < j1; ++j) {
          const float sj = s_score[j];
          rank += ((sj > si) || (sj == si && j < i)) ? 1 : 0;
        }
        if (rank) atomicAdd(&s_rank[i], rank);
      }
    }
    __syncthreads();
    if (tid < N) {
      const uint8_t kp = s_rank[tid] < p.k;
      s_keep[tid] = kp;
      p.keep_out[(int64_t)g * N + tid] = kp;
    }
    __syncthreads();
  } else {
    for (int c = 0; c < nchunks; ++c) gs_mbar_wait(&bars[1 + c], 0);
    __syncthreads();
  }

  // ---- out[i,:] = sum_e w[i][e] * x[idx[i][e],:]  (one warp per output row, everything from shared memory) ---
  const bool masked = FUSED || (p.keep_in != nullptr);
  const char* sFb = reinterpret_cast<const char*>(sF) + lane * 16;
  const float inv_rowbytes = 1.0f / (float)(H * 4);
  for (int i = warp; i < N; i += GS_WARPS) {
    float4 acc[NQ];
#pragma unroll
    for (int u = 0; u < NQ; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    float2* lr = sL + (size_t)i * N;
    const int cnt = s_cnt[i];
    if (masked && s_keep[i] == 0) {   // a dropped node keeps only its edges to kept nodes (wrapper.py:221-225)
      for (int e = lane; e < cnt; e += 32) {
        const float2 en = lr[e];
        const int j = __float2int_rn(__int2float_rn(__float_as_int(en.x)) * inv_rowbytes);
        if (!s_keep[j]) lr[e].y = 0.f;
      }
      __syncwarp();
    }
#pragma unroll 2
    for (int e = 0; e < cnt; ++e) {
      const float2 en = lr[e];
      const int off = __float_as_int(en.x);
      const float wj = en.y;
      const float4* row = reinterpret_cast<const float4*>(sFb + off);
#pragma unroll
      for (int u = 0; u < NQ; ++u) {
        if (u < NQ - 1 || last_ok) {
          const float4 f = row[u * 32];
          acc[u].x = fmaf(wj, f.x, acc[u].x); acc[u].y = fmaf(wj, f.y, acc[u].y);
          acc[u].z = fmaf(wj, f.z, acc[u].z); acc[u].w = fmaf(wj, f.w, acc[u].w);
        }
      }
    }
    float4* orow = reinterpret_cast<float4*>(gout + (int64_t)i * H) + lane;
    __nv_bfloat16* prow = p.out_p ? p.out_p + ((int64_t)g * N + i) * p.ld_p : nullptr;
#pragma unroll
    for (int u = 0; u < NQ; ++u) {
      if (u < NQ - 1 || last_ok) {
        float4 v = acc[u];
        if (p.accumulate) {
          const float4 o = orow[u * 32];
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        if (gout) orow[u * 32] = v;
        if (prow) {
          const float vv[4] = {v.x, v.y, v.z, v.w};
          planes_store4(prow + (lane + u * 32) * 4, p.ps_p, p.np_p, vv);
        }
      }
    }
    if (prow) {   // padding quads up to a multiple of 8 columns (room for the ones column when pad_one)
      const int npq = ((((H + (p.pad_one ? 1 : 0)) + 7) & ~7) - H) >> 2;
      if (lane < npq) {
        const float vv[4] = {(p.pad_one && lane == 0) ? 1.0f : 0.0f, 0.f, 0.f, 0.f};
        planes_store4(prow + H + lane * 4, p.ps_p, p.np_p, vv);
      }
    }
  }
}

// s_p[m] = dropout_s(F[m,:]) . wp  -- the scorer projection for the stand-alone entry points (in the model it is a by-product
// of the epilogue of the GEMM that writes F); one warp per row
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ F, const float* __restrict__ wp, int64_t M, int H,
                                                     uint32_t thr, float scale, uint32_t seed, const uint32_t* __restrict__ salt,
                                                     float* __restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  const uint32_t sd = seed + (thr ? __ldg(salt) : 0u);
  float acc = 0.f;
  for (int q = lane; q < (H >> 2); q += 32) {
    float4 f = __ldg(reinterpret_cast<const float4*>(F + m * H) + q);
    const float4 w = __ldg(reinterpret_cast<const float4*>(wp) + q);
    if (thr) drop_apply4(sd, (uint64_t)m * (uint64_t)H + (uint64_t)q * 4, thr, scale, f);
    acc = fmaf(f.x, w.x, acc); acc = fmaf(f.y, w.y, acc); acc = fmaf(f.z, w.z, acc); acc = fmaf(f.w, w.w, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[m] = acc;
}

typedef void (*GraphSmemFn)(const GraphParams);
static GraphSmemFn graph_smem_fn(bool fused, int nq) {
  switch (nq) {
    case 1: return fused ? graph_smem_kernel<true, 1> : graph_smem_kernel<false, 1>;
    case 2: return fused ? graph_smem_kernel<true, 2> : graph_smem_kernel<false, 2>;
    case 3: return fused ? graph_smem_kernel<true, 3> : graph_smem_kernel<false, 3>;
    default: return fused ? graph_smem_kernel<true, 4> : graph_smem_kernel<false, 4>;
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
  GETB_REQUIRE(p.out || p.out_p, "%s: no output", name);
  GETB_REQUIRE(!p.accumulate || p.out, "%s: accumulate needs the fp32 output", name);
  if (p.out_p)
    GETB_REQUIRE((p.H % 4) == 0 && aligned16(p.x) && (!p.out || aligned16(p.out)) && (((uintptr_t)p.out_p) & 7u) == 0 &&
                     (p.ld_p % 4) == 0 && (p.ps_p % 4) == 0 && p.np_p >= 1 && p.np_p <= 3 && p.ld_p >= ((p.H + (p.pad_one ? 1 : 0) + 7) & ~7),
                 "%s: plane output needs H %% 4 == 0 and aligned tensors", name);
  if (p.G == 0) return 0;
  {
    // single-CTA path: the whole graph (features + adjacency) staged in shared memory by TMA bulk copies
    const size_t need = ((size_t)p.N * p.H + 2 * (size_t)p.N * p.N + 4 * (size_t)p.N) * sizeof(float) +
                        (((size_t)p.N + 15) & ~(size_t)15) + (GS_CHUNKS + 1) * sizeof(uint64_t);
    const bool ok = (p.H % 4) == 0 && p.H <= 128 * GS_MAX_QUADS && (p.N % 2) == 0 && p.N <= GS_MAX_N && aligned16(p.adj) &&
                    aligned16(p.x) && (!p.out || aligned16(p.out)) && (!fused || aligned16(p.wp)) && need <= GS_SMEM_LIMIT;
    if (ok) {
      const int nq = (p.H / 4 + 31) / 32;
      GraphSmemFn fn = graph_smem_fn(fused, nq);
      static bool attr_done[2][GS_MAX_QUADS + 1] = {};
      if (!attr_done[fused ? 1 : 0][nq]) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS_SMEM_LIMIT) != cudaSuccess) {
          set_error("%s: cannot opt in to %d bytes of shared memory", name, (int)GS_SMEM_LIMIT);
          (void)cudaGetLastError();
          return -2;
        }
        attr_done[fused ? 1 : 0][nq] = true;
      }
      fn<<<p.G, GS_THREADS, need, st>>>(p);
      GETB_CHECK_LAUNCH(name);
      return 0;
    }
  }
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
  p.seed_s = seed_scorer; p.seed_2 = seed_layer2; p.salt = dropout_salt_ptr();
  p.score = score; p.keep_out = keep;
  return launch_graph(p, true, (cudaStream_t)stream, "get_gsl_fused_f32");
}

extern "C" int get_graph_aggregate_bp(const float* adj, const float* x, const uint8_t* keep, float* out, void* planes,
                                      int64_t ld_p, int64_t plane_stride, int nplanes, int pad_one, int G, int N, int H,
                                      int transpose, int accumulate, void* stream) {
  GETB_REQUIRE(adj && x && (out || planes), "get_graph_aggregate_bp: null pointer");
  GraphParams p;
  memset(&p, 0, sizeof(p));
  p.adj = adj; p.x = x; p.keep_in = keep; p.out = out;
  p.out_p = reinterpret_cast<__nv_bfloat16*>(planes); p.ld_p = ld_p; p.ps_p = plane_stride; p.np_p = nplanes; p.pad_one = pad_one;
  p.G = G; p.N = N; p.H = H; p.transpose = transpose; p.accumulate = accumulate;
  return launch_graph(p, false, (cudaStream_t)stream, "get_graph_aggregate_bp");
}

extern "C" int get_gsl_fused_bp(const float* adj, const float* F, const float* wp, const float* gate, int G, int N, int H,
                                int k, float drop_p, uint32_t seed_scorer, uint32_t seed_layer2, float* score, uint8_t* keep,
                                float* out, void* planes, int64_t ld_p, int64_t plane_stride, int nplanes, void* stream) {
  GETB_REQUIRE(adj && F && wp && gate && keep && (out || planes), "get_gsl_fused_bp: null pointer");
  GETB_REQUIRE(k >= 0 && k <= N, "get_gsl_fused_bp: k=%d out of [0,%d]", k, N);
  GETB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "get_gsl_fused_bp: dropout probability must be in [0,1)");
  GETB_REQUIRE(aligned16(wp) || (H & 3), "get_gsl_fused_bp: wp must be 16-byte aligned");
  GraphParams p;
  memset(&p, 0, sizeof(p));
  p.adj = adj; p.x = F; p.out = out; p.G = G; p.N = N; p.H = H;
  p.out_p = reinterpret_cast<__nv_bfloat16*>(planes); p.ld_p = ld_p; p.ps_p = plane_stride; p.np_p = nplanes;
  p.wp = wp; p.gate = gate; p.k = k;
  p.thr = drop_p > 0.f ? drop_threshold(drop_p) : 0;
  p.scale = 1.0f / (1.0f - drop_p);
  p.seed_s = seed_scorer; p.seed_2 = seed_layer2; p.salt = dropout_salt_ptr();
  p.score = score; p.keep_out = keep;
  return launch_graph(p, true, (cudaStream_t)stream, "get_gsl_fused_bp");
}

extern "C" int get_rowdot_f32(const float* F, const float* w, int64_t M, int H, float drop_p, uint32_t seed, float* out,
                              void* stream) {
  GETB_REQUIRE(F && w && out && M >= 0 && H >= 4 && (H % 4) == 0 && aligned16(F) && aligned16(w), "get_rowdot_f32: bad arguments");
  GETB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "get_rowdot_f32: dropout probability must be in [0,1)");
  if (M == 0) return 0;
  rowdot_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(F, w, M, H, drop_p > 0.f ? drop_threshold(drop_p) : 0u,
                                                                 1.0f / (1.0f - drop_p), seed, dropout_salt_ptr(), out);
  GETB_CHECK_LAUNCH("get_rowdot_f32");
  return 0;
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
