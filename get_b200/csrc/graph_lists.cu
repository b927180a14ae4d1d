// Graph kernels on packed neighbour lists (the default path of the model).
//
// The reference multiplies by the dense normalised adjacency four times per step and graph (wrapper.py:192 in both GGNN
// layers, and their autograd transposes) although it is 4-13 % dense. Here the dense (G,N,N)
// adjacency is read ONCE per step by build_neighbor_lists_kernel, which packs, per node, {neighbour index, weight} lists of
// the adjacency and of its transpose; every aggregation of the step then runs on the lists:
//   * gather_kernel<FUSED=false>: out[g,i,:] (+)= sum_e w_e * x[g, j_e, :]  (with the GSL keep mask applied per edge);
//   * gather_kernel<FUSED=true> : the fused GSL kernel -- scorer SpMV + scalar GRU gates + top-k (wrapper.py:167,215-219),
//     then the refined aggregation of feat_prop2's dropped-out input (wrapper.py:221-225 + :189-192): the scorer
//     projection s_p arrives as a by-product of the GEMM that wrote the features (get_gemm_bp rowdot_out) or from
//     get_rowdot_f32, the layer-2 dropout draw is a per-graph bit mask in shared memory (one hash per two elements,
//     computed once, instead of once per gathered element).
// Work decomposition: one 256-thread CTA per (graph, slice of <= 64 feature columns). The slice's feature tile is staged
// in shared memory with asynchronous copies issued first; the cheap per-graph scoring runs in the shadow of that load
// (recomputed per slice, so slices are independent CTAs and a 32-claim batch is ~1100 work items, ~7 co-resident per SM).
// HBM traffic per graph = features read once + lists (~4 KB, the repeats hit L2) + output rows written once (fp32 and /
// or bf16 planes for the next tensor-core contraction).
#include "common.cuh"
#include "tcgen05.cuh"
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <set>

namespace getb {

constexpr int GL_MAX_N = 232;   // the builder stages one dense N x (N+1) fp32 tile in shared memory

struct GatherParams {
  const float2* nbr;      // (G, N, N) {neighbour index as int bits, weight}; row i of graph g at (g*N + i)*N
  const int32_t* cnt;     // (G, N) entries per row
  const int32_t* used;    // (G) feature rows any list of the graph refers to (rows >= used are never gathered), or null
  const float* x;         // (G, N, H)
  const uint8_t* keep_in; // (G, N) or null (non-fused)
  float* out;             // (G, N, H) or null
  __nv_bfloat16* out_p; int64_t ld_p, ps_p; int np_p, pad_one;
  int G, N, H, accumulate;
  // fused part
  const float* sp_parts; int n_sp;
  const float* gate;
  int k, np2_shift;
  int tile_rows;          // whole-graph kernel: feature rows the shared-memory tile holds (the rest is gathered from global)
  int nsplit, qs;         // column slices per graph, float4 quads per slice (<= 16)
  uint32_t thr; float scale; uint32_t seed_2; const uint32_t* salt;
  float* score; uint8_t* keep_out;
};

// ---- adjacency -> lists (both orientations), one CTA per graph ---------------------------------------------------------
__global__ void __launch_bounds__(256) build_neighbor_lists_kernel(const float* __restrict__ adj, int N, float2* __restrict__ nbr,
                                                                   int32_t* __restrict__ cnt, float2* __restrict__ nbr_t,
                                                                   int32_t* __restrict__ cnt_t, int32_t* __restrict__ used) {
  extern __shared__ float tile[];                    // N x (N + 1): the dense adjacency of this graph
  __shared__ int s_used[3];                          // [0] rows used by nbr, [1] by nbr_t, [2] asymmetric graph
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 3) s_used[tid] = 0;
  const float* a = adj + (int64_t)g * N * N;
  const int P = N + 1;
  if (((N * N) & 3) == 0) {       // 128-bit loads, several in flight per thread (the graph base stays 16-byte aligned)
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const int nq = (N * N) >> 2;
#pragma unroll 4
    for (int q = tid; q < nq; q += 256) {
      const float4 v = __ldg(a4 + q);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      const int e0 = q * 4;
      int r = e0 / N, c = e0 - r * N;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        tile[r * P + c] = vv[u];
        if (++c == N) { c = 0; ++r; }
      }
    }
  } else {
#pragma unroll 4
    for (int q = tid; q < N * N; q += 256) tile[(q / N) * P + (q % N)] = __ldg(a + q);
  }
  __syncthreads();
  // rows of adj. A symmetric adjacency (every GET graph: D^-1/2 (A + I) D^-1/2) has identical lists for adj^T, written in
  // the same pass; symmetry is verified per entry and the column pass below only runs for graphs that fail it.
  const int64_t gN = (int64_t)g * N;
  bool asym = false;
  for (int i = warp; i < N; i += 8) {
    float2* lr = nbr + (gN + i) * N;
    float2* lt = nbr_t + (gN + i) * N;
    int pos = 0, top = 0;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const float w = j < N ? tile[i * P + j] : 0.f;
      const unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
      if (nz) {                                         // warp-uniform
        if (w != 0.f) {
          const float2 e = make_float2(__int_as_float(j), w);
          const int at = pos + __popc(nz & ((1u << lane) - 1u));
          lr[at] = e;
          lt[at] = e;
          asym |= tile[j * P + i] != w;
        }
        pos += __popc(nz);
        top = j0 + 32 - __clz(nz);                      // 1 + highest neighbour index so far
      }
    }
    if (lane == 0) {
      cnt[gN + i] = pos;
      cnt_t[gN + i] = pos;
      if (top) atomicMax(&s_used[0], top);
    }
  }
  if (__any_sync(0xffffffffu, asym) && lane == 0) s_used[2] = 1;
  __syncthreads();
  if (s_used[2]) {
    for (int i = warp; i < N; i += 8) {                 // rows of adj^T = columns of adj
      float2* lt = nbr_t + (gN + i) * N;
      int pos = 0, top = 0;
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        const float w = j < N ? tile[j * P + i] : 0.f;
        const unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
        if (w != 0.f) lt[pos + __popc(nz & ((1u << lane) - 1u))] = make_float2(__int_as_float(j), w);
        pos += __popc(nz);
        if (nz) top = j0 + 32 - __clz(nz);
      }
      if (lane == 0) {
        cnt_t[gN + i] = pos;
        if (top) atomicMax(&s_used[1], top);
      }
    }
    __syncthreads();
  } else if (tid == 0) {
    s_used[1] = s_used[0];
  }
  __syncthreads();
  // rows 0..used-1 are the only feature rows any list of this graph refers to (pad nodes sit at the end of a text)
  if (tid < 2) used[(int64_t)tid * gridDim.x + g] = s_used[tid];
}

// Output of one row: fp32 and / or NP bf16 planes. RowOut holds the row's base pointers (64-bit arithmetic once per row).
struct RowOut {
  float4* f32;                 // or null
  __nv_bfloat16* pl[3];
};
__device__ __forceinline__ RowOut row_out(const GatherParams& p, int64_t row) {
  RowOut o;
  o.f32 = p.out ? reinterpret_cast<float4*>(p.out + row * p.H) : nullptr;
  __nv_bfloat16* b = p.out_p ? p.out_p + row * p.ld_p : nullptr;
  o.pl[0] = b; o.pl[1] = b + p.ps_p; o.pl[2] = b + 2 * p.ps_p;
  return o;
}
// bf16 planes of 4 values: packed conversions (cvt.rn.bf16x2.f32); the residuals v - bf16(v) are exact in fp32
template <int NP>
__device__ __forceinline__ void store_quad(const RowOut& o, int q, const float4& v) {
  if (o.f32) o.f32[q] = v;
  if (NP > 0) {
    float r0 = v.x, r1 = v.y, r2 = v.z, r3 = v.w;
#pragma unroll
    for (int pl = 0; pl < NP; ++pl) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(r0, r1), b = __floats2bfloat162_rn(r2, r3);
      uint2 w;
      w.x = *reinterpret_cast<const uint32_t*>(&a);
      w.y = *reinterpret_cast<const uint32_t*>(&b);
      *reinterpret_cast<uint2*>(o.pl[pl] + q * 4) = w;
      if (pl + 1 < NP) {
        r0 -= __uint_as_float(w.x << 16); r1 -= __uint_as_float(w.x & 0xFFFF0000u);
        r2 -= __uint_as_float(w.y << 16); r3 -= __uint_as_float(w.y & 0xFFFF0000u);
      }
    }
  }
}
// padding quads of a plane row (columns H .. round_up(H + pad_one, 8)): constants -- 1.0 in column H of plane 0 when pad_one
template <int NP>
__device__ __forceinline__ void store_pad_quad(const RowOut& o, int H, int k, bool pad_one) {
#pragma unroll
  for (int pl = 0; pl < NP; ++pl)
    *reinterpret_cast<uint2*>(o.pl[pl] + H + k * 4) = make_uint2((pl == 0 && k == 0 && pad_one) ? 0x00003F80u : 0u, 0u);
}

// =====================================================================================================
// Whole-graph variant (the default whenever one graph's feature tile fits shared memory: Snopes / PolitiFact dims).
// The graph kernels are INSTRUCTION-ISSUE bound, not bandwidth bound (ncu: issue slots busy, DRAM < 50 %), so this kernel
// is organised around instructions per edge:
//   * one 1024-thread CTA per graph; the N x H tile arrives by TMA bulk copies (no per-thread copy instructions) while the
//     scoring phases run; the tile of the graph that the same SM will see next is prefetched into L2;
//   * every warp owns rows warp, warp+32, ...: lane e loads entry e of each of its rows' lists at kernel start (one
//     coalesced load per row, long before use) and edges are broadcast by two shuffles -- no list traffic later;
//   * a full warp per row, NQ column quads per lane: per edge 2 SHFL + 1 IMAD + NQ x (LDS.128 + 4 FFMA);
//   * per-graph scoring once per graph (not per slice).
// smem: [tile N*H f32][sp N][score N][rank N i32][keep N u8 (padded)][mbarrier]
// =====================================================================================================
constexpr int GR_THREADS = 512;
constexpr int GR_WARPS = GR_THREADS / 32;
constexpr size_t GR_SMEM_LIMIT = 224 * 1024;
constexpr size_t GR_SMEM_HALF = 113 * 1024;       // two CTAs per SM (228 KB per SM, 1 KB reserved per CTA)

__device__ __forceinline__ uint32_t gr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void gr_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = gr_smem_u32(bar);
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

// shared scoring state, 16-byte aligned arrays padded to a multiple of 4 nodes: [sp][score][rank][keep][mbarrier]
struct GraphSmem {
  float* sp; float* score; int* rank; uint8_t* keep; uint64_t* bar;
};
__host__ __device__ __forceinline__ size_t graph_smem_small(int N) { return (size_t)((N + 3) & ~3) * 12 + (size_t)((N + 15) & ~15) + 16; }
__device__ __forceinline__ GraphSmem graph_smem(void* base, int N) {
  const int Np = (N + 3) & ~3;
  GraphSmem m;
  m.sp = reinterpret_cast<float*>(base);
  m.score = m.sp + Np;
  m.rank = reinterpret_cast<int*>(m.score + Np);
  m.keep = reinterpret_cast<uint8_t*>(m.rank + Np);
  m.bar = reinterpret_cast<uint64_t*>(m.keep + ((N + 15) & ~15));      // offset is a multiple of 16 from a 16-byte aligned base
  return m;
}

constexpr int GS_PRE = 6;    // list entries per node the scorer prefetches into registers (longer lists continue from global)

// The per-graph scoring: s_p (sum of the partial vectors) -> s_a = adj @ s_p over the lists -> scalar GRU gates -> score ->
// top-k by rank counting (ties to the lower index) -> keep. Every global load the chain needs (partial sums, list length,
// first GS_PRE list entries) is issued up front, so the chain costs ONE memory latency; the rank counting reads four
// candidate scores per shared load. Ends with the keep flags visible to the whole CTA (barrier inside).
template <int THREADS>
__device__ __forceinline__ void score_topk(const GatherParams& p, const GraphSmem& m, int64_t row0, int tid, bool write_out) {
  const int N = p.N, Np = (N + 3) & ~3;
  float2 pe[GS_PRE];
  int cnt = 0;
  float spv = 0.f;
  if (tid < N) {
    const float2* lr = p.nbr + (row0 + tid) * N;
    cnt = __ldg(p.cnt + row0 + tid);
#pragma unroll
    for (int e = 0; e < GS_PRE; ++e) pe[e] = __ldg(lr + min(e, N - 1));      // beyond cnt: allocated, unread garbage
#pragma unroll 8
    for (int q = 0; q < p.n_sp; ++q) spv += __ldg(p.sp_parts + (int64_t)q * p.G * N + row0 + tid);   // fixed order
    m.sp[tid] = spv;
    m.rank[tid] = 0;
  } else if (tid < Np) {
    m.score[tid] = -INFINITY;          // padding candidates never outrank anything
  }
  __syncthreads();
  if (tid < N) {
    float sa = 0.f;
#pragma unroll
    for (int e = 0; e < GS_PRE; ++e)
      if (e < cnt) sa = fmaf(pe[e].y, m.sp[__float_as_int(pe[e].x)], sa);
    for (int e = GS_PRE; e < cnt; ++e) {
      const float2 en = __ldg(p.nbr + (row0 + tid) * N + e);
      sa = fmaf(en.y, m.sp[__float_as_int(en.x)], sa);
    }
    const float wz0 = __ldg(p.gate + 0), bz0 = __ldg(p.gate + 1), wz1 = __ldg(p.gate + 2), bz1 = __ldg(p.gate + 3);
    const float wr0 = __ldg(p.gate + 4), br0 = __ldg(p.gate + 5), wr1 = __ldg(p.gate + 6), br1 = __ldg(p.gate + 7);
    const float wh0 = __ldg(p.gate + 8), bh0 = __ldg(p.gate + 9), wh1 = __ldg(p.gate + 10), bh1 = __ldg(p.gate + 11);
    const float z = sigmoidf_((wz0 * sa + bz0) + (wz1 * spv + bz1));
    const float r = sigmoidf_((wr0 * sa + br0) + (wr1 * spv + br1));
    const float h = tanhf((wh0 * sa + bh0) + (wh1 * (r * spv) + bh1));
    const float sc = h * z + spv * (1.0f - z);
    m.score[tid] = sc;
    if (write_out && p.score) p.score[row0 + tid] = sc;
  }
  __syncthreads();
  {
    // thread (node i, slice of candidate quads); N <= 2^np2_shift <= THREADS
    const int nsl = THREADS >> p.np2_shift;
    const int nquads = Np >> 2, per = (nquads + nsl - 1) / nsl;
    const int i = tid & ((1 << p.np2_shift) - 1), qa = (tid >> p.np2_shift) * per, qb = min(nquads, qa + per);
    if (i < N && qa < qb) {
      const float si = m.score[i];
      const float4* s4 = reinterpret_cast<const float4*>(m.score);
      int rank = 0;
      for (int q = qa; q < qb; ++q) {
        const float4 c = s4[q];
        const int j = q * 4;
        rank += ((c.x > si) || (c.x == si && j + 0 < i)) ? 1 : 0;
        rank += ((c.y > si) || (c.y == si && j + 1 < i)) ? 1 : 0;
        rank += ((c.z > si) || (c.z == si && j + 2 < i)) ? 1 : 0;
        rank += ((c.w > si) || (c.w == si && j + 3 < i)) ? 1 : 0;
      }
      if (rank) atomicAdd(&m.rank[i], rank);
    }
  }
  __syncthreads();
  if (tid < N) {
    const uint8_t kp = m.rank[tid] < p.k;
    m.keep[tid] = kp;
    if (write_out) p.keep_out[row0 + tid] = kp;
  }
  __syncthreads();
}

// feat_prop2's nn.Dropout draw applied in place to staged feature quads (element index = position in the (G*N, H) tensor; one
// hash per aligned pair, common.cuh). Quad (r, q) of the tile is global quad (row0 + r, q0 + q); 32-bit index arithmetic
// whenever the graph does not straddle a 2^32 boundary of pair indices.
template <int THREADS>
__device__ __forceinline__ void dropout_tile(const GatherParams& p, float4* tile, int pitch_q, int rows, int WQ, int q0, int64_t row0, int tid) {
  const int N = p.N, H = p.H, HQ = H >> 2;
  const uint32_t seed = p.seed_2 + __ldg(p.salt);
  const int total = rows * WQ;
  const uint64_t wbase = ((uint64_t)row0 * (uint64_t)H) >> 1;
  const uint32_t thr = p.thr;
  const float scale = p.scale;
  if ((uint32_t)wbase + 2u * (uint32_t)(N * HQ) >= (uint32_t)wbase) {        // no carry into the high word
    const uint32_t k0 = (uint32_t)(wbase >> 32) * 0x632be5abU + seed * 0x85EBCA6BU + 0x6A09E667U + (uint32_t)wbase * 0x9E3779B1U;
    if (pitch_q == WQ && WQ == HQ) {                  // whole rows staged back to back: tile quad index = global quad index
#pragma unroll 4
      for (int idx = tid; idx < total; idx += THREADS) {
        float4 f = tile[idx];
        const uint32_t h0 = (uint32_t)(2 * idx) * 0x9E3779B1U + k0;
        const uint32_t b0 = mix32(h0), b1 = mix32(h0 + 0x9E3779B1U);
        f.x = (b0 & 0xFFFFu) >= thr ? f.x * scale : 0.f;
        f.y = (b0 >> 16) >= thr ? f.y * scale : 0.f;
        f.z = (b1 & 0xFFFFu) >= thr ? f.z * scale : 0.f;
        f.w = (b1 >> 16) >= thr ? f.w * scale : 0.f;
        tile[idx] = f;
      }
      return;
    }
#pragma unroll 2
    for (int idx = tid; idx < total; idx += THREADS) {
      const int r = idx / WQ, q = idx - r * WQ;
      float4 f = tile[r * pitch_q + q];
      const uint32_t h0 = (uint32_t)(2 * (r * HQ + q0 + q)) * 0x9E3779B1U + k0;
      const uint32_t b0 = mix32(h0), b1 = mix32(h0 + 0x9E3779B1U);
      f.x = (b0 & 0xFFFFu) >= thr ? f.x * scale : 0.f;
      f.y = (b0 >> 16) >= thr ? f.y * scale : 0.f;
      f.z = (b1 & 0xFFFFu) >= thr ? f.z * scale : 0.f;
      f.w = (b1 >> 16) >= thr ? f.w * scale : 0.f;
      tile[r * pitch_q + q] = f;
    }
  } else {
    for (int idx = tid; idx < total; idx += THREADS) {
      const int r = idx / WQ, q = idx - r * WQ;
      float4 f = tile[r * pitch_q + q];
      drop_apply4(seed, (uint64_t)(row0 + r) * (uint64_t)H + (uint64_t)(q0 + q) * 4, thr, scale, f);
      tile[r * pitch_q + q] = f;
    }
  }
}

template <bool FUSED, int NQ, int NP, bool SPILL>
__global__ void __launch_bounds__(GR_THREADS, 2) gather_row_kernel(const __grid_constant__ GatherParams p) {
  extern __shared__ __align__(128) float4 gr_tile[];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, H = p.H, HQ = H >> 2;
  const int64_t row0 = (int64_t)g * N;
  const int cap = SPILL ? p.tile_rows : N;                       // feature rows the shared-memory tile can hold
  const GraphSmem m = graph_smem(gr_tile + (size_t)cap * HQ, N);
  const bool drop = FUSED && p.thr != 0;
  const int n_used = p.used ? min(N, __ldg(p.used + g)) : N;     // feature rows that can be gathered at all
  const int n_tile = min(n_used, cap);
#ifdef GETB_GRAPH_TIMELINE
  __shared__ long long trow[8];
  const long long tr0 = clock64();
#define GR_T(i) do { if (tid == 0) trow[i] = clock64() - tr0; } while (0)
#else
#define GR_T(i) do { } while (0)
#endif

  // ---- this warp's first list, and (fused) the scorer's loads, go out BEFORE the bulk tile traffic ---------------------
  const int lcap = lane < N ? lane : N - 1;           // entries beyond cnt are allocated, unread garbage
  float2 nxt_e = make_float2(0.f, 0.f);
  int nxt_c = 0;
  if (warp < N) {
    nxt_e = __ldg(p.nbr + (row0 + warp) * N + lcap);
    nxt_c = __ldg(p.cnt + row0 + warp);
  }
  if (tid == GR_THREADS - 32) {                       // a warp with no scorer work issues the copies
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gr_smem_u32(m.bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t total = (uint32_t)n_tile * (uint32_t)H * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gr_smem_u32(m.bar)), "r"(total) : "memory");
    const char* src = reinterpret_cast<const char*>(p.x + row0 * H);
    char* dst = reinterpret_cast<char*>(gr_tile);
    for (uint32_t off = 0; off < total; off += 32768u) {
      const uint32_t n = min(32768u, total - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(gr_smem_u32(dst + off)), "l"(src + off), "r"(n), "r"(gr_smem_u32(m.bar)) : "memory");
    }
  }

  if (FUSED) {
    score_topk<GR_THREADS>(p, m, row0, tid, true);
  } else {
    if (p.keep_in && tid < N) m.keep[tid] = p.keep_in[row0 + tid];
    __syncthreads();          // the mbarrier init is visible to every waiter
  }
  const bool masked = FUSED || (p.keep_in != nullptr);
  GR_T(2);
  if (n_tile) gr_mbar_wait(m.bar, 0);
  GR_T(3);
  if (drop) dropout_tile<GR_THREADS>(p, gr_tile, HQ, n_tile, HQ, 0, row0, tid);
  __syncthreads();
  GR_T(4);

  // ---- out[i,:] = sum_e w_e * x[j_e,:] ---------------------------------------------------------------------------------
  const bool last_ok = (lane + (NQ - 1) * 32) < HQ;
  const int npq = NP ? ((((H + (p.pad_one ? 1 : 0)) + 7) & ~7) - H) >> 2 : 0;
  const uint32_t tl = gr_smem_u32(gr_tile) + (uint32_t)lane * 16u;   // explicit shared address: no per-edge base recomputation
  const uint32_t pitch = (uint32_t)HQ * 16u;
  const float4* xg = reinterpret_cast<const float4*>(p.x + row0 * H) + lane;
  const uint32_t seed = drop ? p.seed_2 + __ldg(p.salt) : 0u;
  // rows beyond the tile's capacity (texts with more than `cap` distinct words, SPILL only) come straight from global
  // memory, the layer-2 dropout applied per gathered quad
#define GR_EDGE(J, W)                                                                                          \
  {                                                                                                            \
    if (!SPILL || (J) < cap) {                                                                                 \
      const uint32_t ra = tl + (uint32_t)(J) * pitch;                                                          \
      _Pragma("unroll") for (int u = 0; u < NQ; ++u) {                                                         \
        if (u < NQ - 1 || last_ok) {                                                                           \
          const float4 f = lds128f(ra + u * 512);                                                              \
          acc[u].x = fmaf(W, f.x, acc[u].x); acc[u].y = fmaf(W, f.y, acc[u].y);                                \
          acc[u].z = fmaf(W, f.z, acc[u].z); acc[u].w = fmaf(W, f.w, acc[u].w);                                \
        }                                                                                                      \
      }                                                                                                        \
    } else {                                                                                                   \
      _Pragma("unroll") for (int u = 0; u < NQ; ++u) {                                                         \
        if (u < NQ - 1 || last_ok) {                                                                           \
          float4 f = __ldg(xg + (int64_t)(J) * HQ + u * 32);                                                   \
          if (drop) drop_apply4(seed, (uint64_t)(row0 + (J)) * (uint64_t)H + (uint64_t)(lane + u * 32) * 4, p.thr, p.scale, f); \
          acc[u].x = fmaf(W, f.x, acc[u].x); acc[u].y = fmaf(W, f.y, acc[u].y);                                \
          acc[u].z = fmaf(W, f.z, acc[u].z); acc[u].w = fmaf(W, f.w, acc[u].w);                                \
        }                                                                                                      \
      }                                                                                                        \
    }                                                                                                          \
  }
  for (int i = warp; i < N; i += GR_WARPS) {           // warp-uniform
    const RowOut ro = row_out(p, row0 + i);
    float4 acc[NQ];
#pragma unroll
    for (int u = 0; u < NQ; ++u) {
      acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.accumulate && (u < NQ - 1 || last_ok)) acc[u] = ro.f32[lane + u * 32];   // in flight during the edge loop
    }
    const int cnt = nxt_c;
    float2 my = nxt_e;
    if (i + GR_WARPS < N) {                            // next row's list, one row ahead
      nxt_e = __ldg(p.nbr + (row0 + i + GR_WARPS) * N + lcap);
      nxt_c = __ldg(p.cnt + row0 + i + GR_WARPS);
    }
    const bool dropped = masked && m.keep[i] == 0;     // a dropped node keeps only its edges to kept nodes (wrapper.py:221-225)
    for (int e0 = 0; e0 < cnt; e0 += 32) {
      if (e0) my = (e0 + lane < cnt) ? __ldg(p.nbr + (row0 + i) * N + e0 + lane) : make_float2(0.f, 0.f);
      const int ne = min(32, cnt - e0);
      const int jm = __float_as_int(my.x);
      if (!dropped) {
#pragma unroll 2
        for (int e = 0; e < ne; ++e) {
          const int j = __shfl_sync(0xffffffffu, jm, e);
          const float w = __shfl_sync(0xffffffffu, my.y, e);
          GR_EDGE(j, w)
        }
      } else {
        // edges to dropped neighbours fall away: compact the surviving entries' lane numbers first
        unsigned live = __ballot_sync(0xffffffffu, lane < ne && m.keep[jm] != 0);
        while (live) {
          const int e = __ffs(live) - 1;
          live &= live - 1;
          const int j = __shfl_sync(0xffffffffu, jm, e);
          const float w = __shfl_sync(0xffffffffu, my.y, e);
          GR_EDGE(j, w)
        }
      }
    }
#pragma unroll
    for (int u = 0; u < NQ; ++u)
      if (u < NQ - 1 || last_ok) store_quad<NP>(ro, lane + u * 32, acc[u]);
    if (NP && lane < npq) store_pad_quad<NP>(ro, H, lane, p.pad_one != 0);
  }
#ifdef GETB_GRAPH_TIMELINE
  if (lane == 0 && warp == GR_WARPS - 1) trow[6] = clock64() - tr0;
  __syncthreads();
  GR_T(5);
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 100 || blockIdx.x == 219 || blockIdx.x == 3000))
    printf("GRDBG cta %d fused %d used %d: scoring %lld tile_landed %lld dropout %lld lastwarp_done %lld all_done %lld\n", blockIdx.x,
           (int)FUSED, n_used, trow[2], trow[3], trow[4], trow[6], trow[5]);
#endif
}
#undef GR_EDGE

typedef void (*GatherRowFn)(const GatherParams);
template <bool FUSED, int NP, bool SPILL>
static GatherRowFn gather_row_fn_nq(int nq) {
  switch (nq) {
    case 1: return gather_row_kernel<FUSED, 1, NP, SPILL>;
    case 2: return gather_row_kernel<FUSED, 2, NP, SPILL>;
    case 3: return gather_row_kernel<FUSED, 3, NP, SPILL>;
    default: return gather_row_kernel<FUSED, 4, NP, SPILL>;
  }
}
template <bool FUSED, bool SPILL>
static GatherRowFn gather_row_fn_np(int nq, int np) {
  switch (np) {
    case 0: return gather_row_fn_nq<FUSED, 0, SPILL>(nq);
    case 1: return gather_row_fn_nq<FUSED, 1, SPILL>(nq);
    case 2: return gather_row_fn_nq<FUSED, 2, SPILL>(nq);
    default: return gather_row_fn_nq<FUSED, 3, SPILL>(nq);
  }
}
static GatherRowFn gather_row_fn(bool fused, int nq, int np, bool spill) {
  if (spill) return fused ? gather_row_fn_np<true, true>(nq, np) : gather_row_fn_np<false, true>(nq, np);
  return fused ? gather_row_fn_np<true, false>(nq, np) : gather_row_fn_np<false, false>(nq, np);
}

// =====================================================================================================
// Column-half variant: one 256-thread CTA per (graph, slice of <= 48 column quads). Same instruction economy as the
// whole-graph kernel (half a warp per output row, up to three quads per lane, edges broadcast by shuffles, entries
// fetched one row ahead), but a Snopes tile is 61 KB instead of 120 KB, so THREE CTAs are co-resident per SM: the
// latency-bound scoring phases of one CTA (a chain of dependent global loads with 100 active threads) and its tile load
// overlap the issue-bound aggregation of the others, and a 32-claim batch (~220 graphs = 440 CTAs) is one co-resident
// wave. The per-graph scoring is recomputed by each slice (few instructions; it is latency, not issue).
// Tile rows arrive by one TMA bulk copy per row (a row slice is contiguous), all on one mbarrier.
// smem: [tile N*qs float4][sp N][score N][rank N i32][keep N u8 (padded)][mbarrier]
// =====================================================================================================
constexpr int GC_THREADS = 256;
constexpr int GC_HALVES = GC_THREADS / 16;

template <bool FUSED, int NQ, int NP>
__global__ void __launch_bounds__(GC_THREADS) gather_cols_kernel(const __grid_constant__ GatherParams p) {
  extern __shared__ __align__(128) float4 gc_tile[];
  const int g = blockIdx.x / p.nsplit, slice = blockIdx.x - g * p.nsplit;
  const int tid = threadIdx.x, lane = tid & 31;
  const int N = p.N, H = p.H, HQ = H >> 2;
  const int q0 = slice * p.qs;
  const int WQ = min(p.qs, HQ - q0);
  const int64_t row0 = (int64_t)g * N;
  const GraphSmem m = graph_smem(gc_tile + (size_t)N * p.qs, N);
  const bool drop = FUSED && p.thr != 0;
  const int n_used = p.used ? min(N, __ldg(p.used + g)) : N;     // feature rows that can be gathered at all

  // ---- tile rows: one bulk copy each, issued by warp 0 -----------------------------------------------------------------
  if (tid < 32) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gr_smem_u32(m.bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gr_smem_u32(m.bar)), "r"((uint32_t)n_used * (uint32_t)WQ * 16u) : "memory");
    }
    __syncwarp();
    const char* src = reinterpret_cast<const char*>(p.x + row0 * H) + (size_t)q0 * 16;
    for (int r = tid; r < n_used; r += 32)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(gr_smem_u32(gc_tile + (size_t)r * p.qs)), "l"(src + (size_t)r * H * 4), "r"((uint32_t)WQ * 16u), "r"(gr_smem_u32(m.bar)) : "memory");
  }

  // ---- first row's list entries for every half warp, long before use ---------------------------------------------------
  const int half = tid >> 4, hl = tid & 15, hsh = lane & 16;
  const int lcap = hl < N ? hl : N - 1;               // entries beyond cnt are allocated, unread garbage
  float2 nxt_e = make_float2(0.f, 0.f);
  int nxt_c = 0;
  if (half < N) {
    nxt_e = __ldg(p.nbr + (row0 + half) * N + lcap);
    nxt_c = __ldg(p.cnt + row0 + half);
  }

  if (FUSED) {
    score_topk<GC_THREADS>(p, m, row0, tid, slice == 0);
  } else {
    if (p.keep_in && tid < N) m.keep[tid] = p.keep_in[row0 + tid];
    __syncthreads();          // the mbarrier init is visible to every waiter
  }
  const bool masked = FUSED || (p.keep_in != nullptr);
  if (n_used) gr_mbar_wait(m.bar, 0);
  if (drop) dropout_tile<GC_THREADS>(p, gc_tile, p.qs, n_used, WQ, q0, row0, tid);
  __syncthreads();

  // ---- out[i, slice] = sum_e w_e * x[j_e, slice]: half a warp per output row, lanes own quads hl, hl+16, hl+32 -------
  const bool ok1 = NQ > 1 && (hl + 16) < WQ, ok2 = NQ > 2 && (hl + 32) < WQ, ok0 = hl < WQ;
  const int npq = (NP && slice == p.nsplit - 1) ? ((((H + (p.pad_one ? 1 : 0)) + 7) & ~7) - H) >> 2 : 0;
  const uint32_t tl = gr_smem_u32(gc_tile) + (uint32_t)hl * 16u;
  const uint32_t pitch = (uint32_t)p.qs * 16u;
  const int rounds = (N + GC_HALVES - 1) / GC_HALVES;
  for (int rd = 0; rd < rounds; ++rd) {               // warp-uniform trip count (shuffles inside)
    const int i = half + rd * GC_HALVES;
    const bool row_ok = i < N;
    const int cnt = row_ok ? nxt_c : 0;
    float2 my = nxt_e;
    if (i + GC_HALVES < N) {
      nxt_e = __ldg(p.nbr + (row0 + i + GC_HALVES) * N + lcap);
      nxt_c = __ldg(p.cnt + row0 + i + GC_HALVES);
    }
    const bool dropped = masked && row_ok && m.keep[i] == 0;   // a dropped node keeps only its edges to kept nodes (wrapper.py:221-225)
    const RowOut ro = row_out(p, row0 + (row_ok ? i : 0));
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
    if (p.accumulate && row_ok) {                     // in flight during the edge loop
      if (ok0) a0 = ro.f32[q0 + hl];
      if (ok1) a1 = ro.f32[q0 + hl + 16];
      if (ok2) a2 = ro.f32[q0 + hl + 32];
    }
    const int cmax = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, 16));
    for (int e0 = 0; e0 < cmax; e0 += 16) {
      if (e0) my = (e0 + hl < cnt) ? __ldg(p.nbr + (row0 + i) * N + e0 + hl) : make_float2(0.f, 0.f);
      const int jm = __float_as_int(my.x);
      const bool valid = (e0 + hl < cnt) && !(dropped && !m.keep[jm]);
      const unsigned both = __ballot_sync(0xffffffffu, valid);
      unsigned live = (both >> hsh) & 0xFFFFu;
      const int nmax = max(__popc(both & 0xFFFFu), __popc(both >> 16));
#pragma unroll 2
      for (int t = 0; t < nmax; ++t) {
        const bool on = live != 0;
        const int src = on ? __ffs(live) - 1 : 0;
        live &= live - 1;
        int j = __shfl_sync(0xffffffffu, jm, src + hsh);
        float w = __shfl_sync(0xffffffffu, my.y, src + hsh);
        if (!on) { j = 0; w = 0.f; }
        const uint32_t ra = tl + (uint32_t)j * pitch;
        if (ok0) {
          const float4 f = lds128f(ra);
          a0.x = fmaf(w, f.x, a0.x); a0.y = fmaf(w, f.y, a0.y); a0.z = fmaf(w, f.z, a0.z); a0.w = fmaf(w, f.w, a0.w);
        }
        if (ok1) {
          const float4 f = lds128f(ra + 256);
          a1.x = fmaf(w, f.x, a1.x); a1.y = fmaf(w, f.y, a1.y); a1.z = fmaf(w, f.z, a1.z); a1.w = fmaf(w, f.w, a1.w);
        }
        if (ok2) {
          const float4 f = lds128f(ra + 512);
          a2.x = fmaf(w, f.x, a2.x); a2.y = fmaf(w, f.y, a2.y); a2.z = fmaf(w, f.z, a2.z); a2.w = fmaf(w, f.w, a2.w);
        }
      }
    }
    if (row_ok) {
      if (ok0) store_quad<NP>(ro, q0 + hl, a0);
      if (ok1) store_quad<NP>(ro, q0 + hl + 16, a1);
      if (ok2) store_quad<NP>(ro, q0 + hl + 32, a2);
      if (NP && hl < npq) store_pad_quad<NP>(ro, H, hl, p.pad_one != 0);   // last slice only
    }
  }
}

typedef void (*GatherColsFn)(const GatherParams);
template <bool FUSED, int NP>
static GatherColsFn gather_cols_fn_nq(int nq) {
  switch (nq) {
    case 1: return gather_cols_kernel<FUSED, 1, NP>;
    case 2: return gather_cols_kernel<FUSED, 2, NP>;
    default: return gather_cols_kernel<FUSED, 3, NP>;
  }
}
static GatherColsFn gather_cols_fn(bool fused, int nq, int np) {
  switch (np) {
    case 0: return fused ? gather_cols_fn_nq<true, 0>(nq) : gather_cols_fn_nq<false, 0>(nq);
    case 1: return fused ? gather_cols_fn_nq<true, 1>(nq) : gather_cols_fn_nq<false, 1>(nq);
    case 2: return fused ? gather_cols_fn_nq<true, 2>(nq) : gather_cols_fn_nq<false, 2>(nq);
    default: return fused ? gather_cols_fn_nq<true, 3>(nq) : gather_cols_fn_nq<false, 3>(nq);
  }
}

static int launch_gather(GatherParams& p, bool fused, cudaStream_t st, const char* name) {
  GETB_REQUIRE(p.G >= 0 && p.N >= 1 && p.N <= GL_MAX_N && p.H >= 4 && (p.H % 4) == 0,
               "%s: need 1 <= N <= %d and H %% 4 == 0 (N=%d H=%d)", name, GL_MAX_N, p.N, p.H);
  GETB_REQUIRE(p.nbr && p.cnt && p.x && (p.out || p.out_p) && aligned16(p.x) && (!p.out || aligned16(p.out)), "%s: null / misaligned pointer", name);
  GETB_REQUIRE(!p.accumulate || p.out, "%s: accumulate needs the fp32 output", name);
  if (p.out_p)
    GETB_REQUIRE((((uintptr_t)p.out_p) & 7u) == 0 && (p.ld_p % 4) == 0 && (p.ps_p % 4) == 0 && p.np_p >= 1 && p.np_p <= 3 &&
                     p.ld_p >= ((p.H + (p.pad_one ? 1 : 0) + 7) & ~7),
                 "%s: plane output needs an 8-byte aligned tensor with room for the padding", name);
  if (p.G == 0) return 0;
  const int np = p.out_p ? p.np_p : 0;
  static int which = -1;     // 0 auto: whole graph when the tile fits, else column slices; 1 column slices always
  if (which < 0) {
    const char* e = getenv("GET_B200_GRAPH_KERNEL");
    which = (e && !strcmp(e, "cols")) ? 1 : 0;
  }
  static std::set<const void*> opted;
  auto opt_in = [&](const void* fn, size_t bytes) -> bool {
    if (bytes <= 48 * 1024 || opted.count(fn)) return true;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GR_SMEM_LIMIT) != cudaSuccess) {
      (void)cudaGetLastError();
      getb::set_error("graph gather: cannot opt in to large shared memory");
      return false;
    }
    (void)cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    opted.insert(fn);
    return true;
  };
  const size_t smem_small = graph_smem_small(p.N) + 16;
  // whole-graph kernel, two CTAs per SM: the tile holds as many feature rows as fit 113 KB (94 of 100 at Snopes dims; a text
  // uses `used` <= N rows, typically ~75); worthwhile while at least 3/4 of the rows fit
  const int cap_rows = (int)std::min<size_t>((size_t)p.N, (GR_SMEM_HALF - smem_small) / ((size_t)p.H * 4));
  if (which == 0 && cap_rows * 4 >= p.N * 3 && p.H <= 512) {
    const size_t smem_row = (size_t)cap_rows * p.H * 4 + smem_small;
    const int nq = (p.H / 4 + 31) / 32;
    p.tile_rows = cap_rows;
    GatherRowFn fn = gather_row_fn(fused, nq, np, cap_rows < p.N);
    if (!opt_in((const void*)fn, smem_row)) return -2;
    fn<<<p.G, GR_THREADS, smem_row, st>>>(p);
    GETB_CHECK_LAUNCH(name);
    return 0;
  }
  // column slices of at most 48 quads, as even as possible (H = 512: 3 slices of 43 / 43 / 42 quads)
  const int HQ = p.H / 4;
  p.nsplit = (HQ + 47) / 48;
  p.qs = (HQ + p.nsplit - 1) / p.nsplit;
  GETB_REQUIRE((int64_t)p.G * p.nsplit < (1LL << 31), "%s: too many work items", name);
  const size_t smem = (size_t)p.N * p.qs * 16 + smem_small;
  GETB_REQUIRE(smem <= GR_SMEM_LIMIT, "%s: %zu bytes of shared memory", name, smem);
  GatherColsFn fn = gather_cols_fn(fused, (p.qs + 15) / 16, np);
  if (!opt_in((const void*)fn, smem)) return -2;
  fn<<<p.G * p.nsplit, GC_THREADS, smem, st>>>(p);
  GETB_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace getb

using namespace getb;

extern "C" int get_build_neighbor_lists(const float* adj, int G, int N, void* nbr, int32_t* cnt, void* nbr_t, int32_t* cnt_t,
                                        int32_t* used, void* stream) {
  GETB_REQUIRE(adj && nbr && cnt && nbr_t && cnt_t && used && G >= 0 && N >= 1 && N <= GL_MAX_N, "get_build_neighbor_lists: bad arguments (N <= %d)", GL_MAX_N);
  GETB_REQUIRE((((uintptr_t)nbr) & 7u) == 0 && (((uintptr_t)nbr_t) & 7u) == 0, "get_build_neighbor_lists: lists must be 8-byte aligned");
  if (G == 0) return 0;
  const size_t smem = (size_t)N * (N + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    static bool done = false;
    if (!done) {
      if (cudaFuncSetAttribute(build_neighbor_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
        (void)cudaGetLastError();
        getb::set_error("get_build_neighbor_lists: cannot opt in to large shared memory");
        return -2;
      }
      done = true;
    }
  }
  GETB_REQUIRE(smem <= 220 * 1024, "get_build_neighbor_lists: adjacency tile of %zu bytes does not fit shared memory", smem);  // N <= 236
  build_neighbor_lists_kernel<<<G, 256, smem, (cudaStream_t)stream>>>(adj, N, reinterpret_cast<float2*>(nbr), cnt,
                                                                      reinterpret_cast<float2*>(nbr_t), cnt_t, used);
  GETB_CHECK_LAUNCH("get_build_neighbor_lists");
  return 0;
}

extern "C" int get_graph_gather(const void* nbr, const int32_t* cnt, const int32_t* used, const float* x, const uint8_t* keep, float* out, void* planes,
                                int64_t ld_p, int64_t plane_stride, int nplanes, int pad_one, int G, int N, int H, int accumulate,
                                void* stream) {
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.nbr = reinterpret_cast<const float2*>(nbr); p.cnt = cnt; p.used = used; p.x = x; p.keep_in = keep; p.out = out;
  p.out_p = reinterpret_cast<__nv_bfloat16*>(planes); p.ld_p = ld_p; p.ps_p = plane_stride; p.np_p = nplanes; p.pad_one = pad_one;
  p.G = G; p.N = N; p.H = H; p.accumulate = accumulate;
  return launch_gather(p, false, (cudaStream_t)stream, "get_graph_gather");
}

extern "C" int get_gsl_gather(const void* nbr, const int32_t* cnt, const int32_t* used, const float* F, const float* sp_parts, int n_sp, const float* gate,
                              int G, int N, int H, int k, float drop_p, uint32_t seed_layer2, float* score, uint8_t* keep, float* out,
                              void* planes, int64_t ld_p, int64_t plane_stride, int nplanes, void* stream) {
  GETB_REQUIRE(sp_parts && n_sp >= 1 && gate && keep, "get_gsl_gather: null pointer");
  GETB_REQUIRE(k >= 0 && k <= N, "get_gsl_gather: k=%d out of [0,%d]", k, N);
  GETB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "get_gsl_gather: dropout probability must be in [0,1)");
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.nbr = reinterpret_cast<const float2*>(nbr); p.cnt = cnt; p.used = used; p.x = F; p.out = out;
  p.out_p = reinterpret_cast<__nv_bfloat16*>(planes); p.ld_p = ld_p; p.ps_p = plane_stride; p.np_p = nplanes;
  p.G = G; p.N = N; p.H = H;
  p.sp_parts = sp_parts; p.n_sp = n_sp; p.gate = gate; p.k = k;
  p.np2_shift = 0;
  while ((1 << p.np2_shift) < N) ++p.np2_shift;
  p.thr = drop_p > 0.f ? drop_threshold(drop_p) : 0;
  p.scale = 1.0f / (1.0f - drop_p);
  p.seed_2 = seed_layer2; p.salt = dropout_salt_ptr();
  p.score = score; p.keep_out = keep;
  return launch_gather(p, true, (cudaStream_t)stream, "get_gsl_gather");
}
