// Graph kernels on packed neighbour lists (the default path of the model).
//
// The reference multiplies by the dense normalised adjacency four times per step and graph (wrapper.py:192 in both GGNN
// layers, and their autograd transposes) although it is 4-13 % dense. Here the dense (G,N,N)
// adjacency is read ONCE per step by build_neighbor_lists_kernel, which packs it -- and its transpose -- into per-graph CSR
// records (row pointers + {neighbour index, weight} entries, contiguous per graph); every aggregation of the step then
// walks the lists:
//   * gather_*_kernel<FUSED=false>: out[g,i,:] (+)= sum_e w_e * x[g, j_e, :]  (with the GSL keep mask applied per edge);
//   * gather_*_kernel<FUSED=true> : the fused GSL kernel -- scorer SpMV + scalar GRU gates + top-k (wrapper.py:167,215-219),
//     then the refined aggregation of feat_prop2's dropped-out input (wrapper.py:221-225 + :189-192). The scorer
//     projection s_p arrives as a by-product of the GEMM that wrote the features (get_gemm_bp rowdot_out) or from
//     get_rowdot_f32; the layer-2 dropout draw is applied once to the staged feature tile.
// What bounds these kernels on B200 is not bandwidth but (a) chains of dependent memory latencies and (b) instruction
// issue (measured: profiles/r2_graph_*). Hence:
//   * a graph's lists are ONE contiguous record, copied to shared memory by the TMA engine together with the feature tile
//     at kernel start: after two memory latencies (row count, then everything else) the kernel only touches shared memory;
//   * edges are fetched with one broadcast shared load; a full warp owns an output row with up to four column quads per
//     lane (per edge: LDS.64 + IMAD + NQ x (LDS.128 + 4 FFMA));
//   * two CTAs are co-resident per SM (tile + lists <= 113 KB), so one graph's load / scoring phases overlap the other's
//     aggregation; rows that do not fit the tile (texts with more distinct words than it holds) are gathered from global.
// HBM traffic per graph = features read once + the list record (~4 KB) + output rows written once (fp32 and / or bf16
// planes for the next tensor-core contraction).
#include "common.cuh"
#include "tcgen05.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <set>

namespace getb {

constexpr int GL_MAX_N = 232;   // the builder stages one dense N x (N+1) fp32 tile in shared memory

__host__ __device__ __forceinline__ int gl_rowptr_pitch(int N) { return (N + 1 + 3) & ~3; }   // ints per graph, 16-byte rows
__host__ __device__ __forceinline__ int gl_ent_cap(int N) { return (N * N + 1) & ~1; }        // entries per graph block (16-byte blocks)

struct GatherParams {
  const float2* ent;       // (G, gl_ent_cap(N)) {neighbour index as int bits, weight}, compact from the start of each graph's block
  const int32_t* rowptr;   // (G, pitch) row pointers [0..N], pitch = gl_rowptr_pitch(N)
  const int32_t* used;     // (G) feature rows any list of the graph refers to (rows >= used are never gathered), or null
  const float* x;          // (G, N, H)
  const uint8_t* keep_in;  // (G, N) or null (non-fused)
  float* out;              // (G, N, H) or null
  __nv_bfloat16* out_p; int64_t ld_p, ps_p; int np_p, pad_one;
  int G, N, H, accumulate;
  // fused part
  const float* sp_parts; int n_sp;
  const float* gate;
  int k, np2_shift;
  int tile_rows;           // whole-graph kernel: feature rows the shared-memory tile holds (the rest is gathered from global)
  int lcap, e1;            // list entries held in shared memory; entries fetched by the first (size-independent) copy
  int nsplit, qs;          // column-slice kernel: slices per graph, float4 quads per slice (<= 48)
  uint32_t thr; float scale; uint32_t seed_2; const uint32_t* salt;
  float* score; uint8_t* keep_out;
};

// =====================================================================================================
// adjacency -> CSR records (both orientations), one CTA per graph
// =====================================================================================================
// One orientation: counts per row (ballots over the staged tile) -> exclusive scan -> entries. A symmetric adjacency
// (every GET graph: D^-1/2 (A + I) D^-1/2) has identical records for adj^T: they are written in the same pass (DUAL);
// symmetry is verified per entry and the transposed pass only runs for graphs that fail it.
template <bool TR, bool DUAL>
__device__ __forceinline__ void build_orientation(const float* tile, int N, int P, int* s_ptr, int* s_flags, float2* ent0, float2* ent1,
                                                  int32_t* rp0, int32_t* rp1, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  bool asym = false;
  int top = 0;
  for (int i = warp; i < N; i += 8) {
    int c = 0;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const float w = j < N ? (TR ? tile[j * P + i] : tile[i * P + j]) : 0.f;
      const unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
      c += __popc(nz);
      if (nz) top = max(top, j0 + 32 - __clz(nz));          // 1 + highest neighbour index so far
      if (DUAL && w != 0.f) asym |= tile[j * P + i] != w;
    }
    if (lane == 0) s_ptr[i + 1] = c;
  }
  if (lane == 0 && top) atomicMax(&s_flags[TR ? 1 : 0], top);
  if (DUAL && __any_sync(0xffffffffu, asym) && lane == 0) s_flags[2] = 1;
  __syncthreads();
  if (warp == 0) {                                           // exclusive scan of <= 232 counts by one warp
    const int per = (N + 31) / 32, lo = lane * per, hi = min(N, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += s_ptr[i + 1];
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += v;
    }
    int run = inc - sum;
    if (lane == 0) s_ptr[0] = 0;
    for (int i = lo; i < hi; ++i) {
      run += s_ptr[i + 1];
      s_ptr[i + 1] = run;
    }
  }
  __syncthreads();
  const bool both = DUAL && s_flags[2] == 0;
  for (int i = tid; i <= N; i += 256) {
    rp0[i] = s_ptr[i];
    if (both) rp1[i] = s_ptr[i];
  }
  for (int i = warp; i < N; i += 8) {
    int pos = s_ptr[i];
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const float w = j < N ? (TR ? tile[j * P + i] : tile[i * P + j]) : 0.f;
      const unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
      if (w != 0.f) {
        const float2 e = make_float2(__int_as_float(j), w);
        const int at = pos + __popc(nz & ((1u << lane) - 1u));
        ent0[at] = e;
        if (both) ent1[at] = e;
      }
      pos += __popc(nz);
    }
  }
}

__global__ void __launch_bounds__(256) build_neighbor_lists_kernel(const float* __restrict__ adj, int N, float2* __restrict__ ent,
                                                                   int32_t* __restrict__ rowptr, int32_t* __restrict__ used) {
  extern __shared__ float tile[];                    // N x (N + 1): the dense adjacency of this graph, then the row pointers
  __shared__ int s_flags[3];                         // [0] rows used by adj's lists, [1] by adj^T's, [2] asymmetric graph
  const int g = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
  const int P = N + 1;
  int* s_ptr = reinterpret_cast<int*>(tile + (size_t)N * P);
  if (tid < 3) s_flags[tid] = 0;
  const float* a = adj + (int64_t)g * N * N;
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
  const int64_t cap = gl_ent_cap(N);
  const int pitch = gl_rowptr_pitch(N);
  float2* e0p = ent + (int64_t)g * cap;
  float2* e1p = ent + ((int64_t)G + g) * cap;
  int32_t* r0p = rowptr + (int64_t)g * pitch;
  int32_t* r1p = rowptr + ((int64_t)G + g) * pitch;
  build_orientation<false, true>(tile, N, P, s_ptr, s_flags, e0p, e1p, r0p, r1p, tid);
  __syncthreads();
  if (s_flags[2]) {
    build_orientation<true, false>(tile, N, P, s_ptr, s_flags, e1p, nullptr, r1p, nullptr, tid);
    __syncthreads();
  } else if (tid == 0) {
    s_flags[1] = s_flags[0];
  }
  __syncthreads();
  // rows 0..used-1 are the only feature rows any list of this graph refers to (pad nodes sit at the end of a text)
  if (tid < 2) used[(int64_t)tid * G + g] = s_flags[tid];
}

// =====================================================================================================
// shared helpers of the gather kernels
// =====================================================================================================
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
__device__ __forceinline__ void gr_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gr_smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA engine, no tensor map needed for a contiguous block); 16-byte aligned, bytes % 16 == 0,
// split in pieces of <= 32 KB
__device__ __forceinline__ void gr_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = min(32768u, bytes - off);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gr_smem_u32(reinterpret_cast<char*>(dst) + off)), "l"(reinterpret_cast<const char*>(src) + off), "r"(n),
                   "r"(gr_smem_u32(bar)) : "memory");
  }
}

// per-graph scratch behind the feature tile, every array 16-byte aligned:
//   [ent lcap float2][rowptr pitch i32][sp Np f32][score Np f32][rank Np i32][keep pad16(N) u8][3 mbarriers]
struct GraphSmem {
  float2* ent; int* rowptr; float* sp; float* score; int* rank; uint8_t* keep; uint64_t* bar;
};
__host__ __device__ __forceinline__ size_t graph_smem_bytes(int N, int lcap) {
  return (size_t)lcap * 8 + (size_t)gl_rowptr_pitch(N) * 4 + (size_t)((N + 3) & ~3) * 12 + (size_t)((N + 15) & ~15) + 32;
}
__device__ __forceinline__ GraphSmem graph_smem(void* base, int N, int lcap) {
  const int Np = (N + 3) & ~3;
  GraphSmem m;
  m.ent = reinterpret_cast<float2*>(base);
  m.rowptr = reinterpret_cast<int*>(m.ent + lcap);           // lcap is even
  m.sp = reinterpret_cast<float*>(m.rowptr + gl_rowptr_pitch(N));
  m.score = m.sp + Np;
  m.rank = reinterpret_cast<int*>(m.score + Np);
  m.keep = reinterpret_cast<uint8_t*>(m.rank + Np);
  m.bar = reinterpret_cast<uint64_t*>(m.keep + ((N + 15) & ~15));
  return m;
}
// list entry e of the graph: from shared memory, or (graphs with more than lcap edges) from the global record
__device__ __forceinline__ float2 list_entry(const GatherParams& p, const GraphSmem& m, const float2* gent, int e) {
  return e < p.lcap ? m.ent[e] : __ldg(gent + e);
}

// Start of every gather kernel, executed by ONE thread: the list record (row pointers + the first e1 entries: sizes that do
// not depend on the graph) on bar[0]; the caller then puts the feature rows on bar[1].
__device__ __forceinline__ void issue_list_copy(const GatherParams& p, const GraphSmem& m, int g) {
  const int N = p.N, pitch = gl_rowptr_pitch(N);
#pragma unroll
  for (int b = 0; b < 3; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gr_smem_u32(m.bar + b)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  gr_expect_tx(m.bar, (uint32_t)pitch * 4u + (uint32_t)p.e1 * 8u);
  gr_bulk_g2s(m.rowptr, p.rowptr + (int64_t)g * pitch, (uint32_t)pitch * 4u, m.bar);
  gr_bulk_g2s(m.ent, p.ent + (int64_t)g * gl_ent_cap(N), (uint32_t)p.e1 * 8u, m.bar);
}
// After bar[0]: graphs with more than e1 edges fetch the rest (up to lcap) on bar[2]. Called by every thread; `issuer` is
// the thread that issues the copy.
__device__ __forceinline__ void finish_list_copy(const GatherParams& p, const GraphSmem& m, int g, bool issuer) {
  gr_mbar_wait(m.bar, 0);
  const int nnz = m.rowptr[p.N];
  if (nnz > p.e1) {
    if (issuer) {
      const uint32_t bytes = (uint32_t)(((min(nnz, p.lcap) - p.e1) + 1) & ~1) * 8u;      // e1, lcap even: stays inside the record
      gr_expect_tx(m.bar + 2, bytes);
      gr_bulk_g2s(m.ent + p.e1, p.ent + (int64_t)g * gl_ent_cap(p.N) + p.e1, bytes, m.bar + 2);
    }
    gr_mbar_wait(m.bar + 2, 0);
  }
}

// The per-graph scoring: s_p (sum of the partial vectors) -> s_a = adj @ s_p over the lists -> scalar GRU gates -> score ->
// top-k by rank counting (ties to the lower index) -> keep. `spv` = this thread's s_p (loaded before the lists arrived).
// The rank counting reads four candidate scores per shared load. Ends with the keep flags visible to the whole CTA.
template <int THREADS>
__device__ __forceinline__ void score_topk(const GatherParams& p, const GraphSmem& m, const float2* gent, int64_t row0, int tid, float spv,
                                           bool write_out) {
  const int N = p.N, Np = (N + 3) & ~3;
  if (tid < N) {
    m.sp[tid] = spv;
    m.rank[tid] = 0;
  } else if (tid < Np) {
    m.score[tid] = -INFINITY;          // padding candidates never outrank anything
  }
  __syncthreads();
  if (tid < N) {
    float sa = 0.f;
    const int e1 = m.rowptr[tid + 1];
    for (int e = m.rowptr[tid]; e < e1; ++e) {
      const float2 en = list_entry(p, m, gent, e);
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
// s_p of node tid: the partial vectors summed in a fixed order, all loads in flight together
__device__ __forceinline__ float load_sp(const GatherParams& p, int64_t row0, int tid) {
  float v[8];
  const float* src = p.sp_parts + row0 + tid;
  const int64_t stride = (int64_t)p.G * p.N;
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = q < p.n_sp ? __ldg(src + q * stride) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < p.n_sp) s += v[q];
  for (int q = 8; q < p.n_sp; ++q) s += __ldg(src + q * stride);
  return s;
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

// =====================================================================================================
// Whole-graph kernel (default whenever >= 3/4 of a graph's feature rows fit the tile: Snopes / PolitiFact dims)
// smem: [tile cap*H f32][GraphSmem]
// =====================================================================================================
constexpr int GR_THREADS = 512;
constexpr int GR_WARPS = GR_THREADS / 32;
constexpr size_t GR_SMEM_LIMIT = 224 * 1024;
constexpr size_t GR_SMEM_HALF = 113 * 1024;       // two CTAs per SM (228 KB per SM, 1 KB reserved per CTA)

template <bool FUSED, int NQ, int NP, bool SPILL>
__global__ void __launch_bounds__(GR_THREADS, 2) gather_row_kernel(const __grid_constant__ GatherParams p) {
  extern __shared__ __align__(128) float4 gr_tile[];
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, H = p.H, HQ = H >> 2;
  const int64_t row0 = (int64_t)g * N;
  const int cap = SPILL ? p.tile_rows : N;                       // feature rows the shared-memory tile can hold
  const GraphSmem m = graph_smem(gr_tile + (size_t)cap * HQ, N, p.lcap);
  const float2* gent = p.ent + (int64_t)g * gl_ent_cap(N);
  const bool drop = FUSED && p.thr != 0;
  const bool issuer = tid == GR_THREADS - 32;                    // a warp without scorer work issues the copies
  const int n_used = p.used ? min(N, __ldg(p.used + g)) : N;     // feature rows that can be gathered at all
  const int n_tile = min(n_used, cap);
#ifdef GETB_GRAPH_TIMELINE
  __shared__ long long trow[8];
  const long long tr0 = clock64();
#define GR_T(i) do { if (tid == 0) trow[i] = clock64() - tr0; } while (0)
#else
#define GR_T(i) do { } while (0)
#endif
  float spv = 0.f;
  if (FUSED && tid < N) spv = load_sp(p, row0, tid);             // in flight while the copies are set up
  if (issuer) {
    issue_list_copy(p, m, g);
    gr_expect_tx(m.bar + 1, (uint32_t)n_tile * (uint32_t)H * 4u);
    gr_bulk_g2s(gr_tile, p.x + row0 * H, (uint32_t)n_tile * (uint32_t)H * 4u, m.bar + 1);
  }
  __syncthreads();            // the mbarrier inits are visible to every waiter
  finish_list_copy(p, m, g, issuer);
  GR_T(1);
  if (FUSED) {
    score_topk<GR_THREADS>(p, m, gent, row0, tid, spv, true);
  } else if (p.keep_in) {
    if (tid < N) m.keep[tid] = p.keep_in[row0 + tid];
    __syncthreads();
  }
  const bool masked = FUSED || (p.keep_in != nullptr);
  GR_T(2);
  if (n_tile) gr_mbar_wait(m.bar + 1, 0);
  GR_T(3);
  if (drop) {
    dropout_tile<GR_THREADS>(p, gr_tile, HQ, n_tile, HQ, 0, row0, tid);
    __syncthreads();
  }
  GR_T(4);

  // ---- out[i,:] = sum_e w_e * x[j_e,:]: a warp per output row, lanes own column quads lane, lane+32, ... ----------------
  const bool last_ok = (lane + (NQ - 1) * 32) < HQ;
  const int npq = NP ? ((((H + (p.pad_one ? 1 : 0)) + 7) & ~7) - H) >> 2 : 0;
  const uint32_t tl = gr_smem_u32(gr_tile) + (uint32_t)lane * 16u;   // explicit shared address: no per-edge base recomputation
  const uint32_t pitch = (uint32_t)HQ * 16u;
  const float4* xg = reinterpret_cast<const float4*>(p.x + row0 * H) + lane;
  const uint32_t seed = (SPILL && drop) ? p.seed_2 + __ldg(p.salt) : 0u;
  // The loop body is instruction-issue bound, so the common case is specialised: (a) every edge of the graph sits in shared
  // memory (nnz <= lcap: always at window 3 / 5), (b) the row is kept, so no per-edge mask test, (c) rows without edges
  // (the pad nodes of a text, ~25 % of the rows) store constants without any conversion; row pointers advance by constant
  // strides instead of being recomputed.
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
      /* a row beyond the tile's capacity (a text with more than `cap` distinct words): straight from global */ \
      /* memory, the layer-2 dropout applied per gathered quad */                                              \
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
  // CTA-uniform: every list entry is in shared memory and every row a list can refer to is in the tile
  const bool all_in_smem = m.rowptr[N] <= p.lcap && (!SPILL || n_used <= cap);
  RowOut ro = row_out(p, row0 + warp);
  const int64_t step_pl = (int64_t)GR_WARPS * p.ld_p;
  for (int i = warp; i < N; i += GR_WARPS) {           // warp-uniform
    const int ea = m.rowptr[i], eb = m.rowptr[i + 1];
    if (ea == eb && !p.accumulate) {                   // no neighbours: the row is zero
#pragma unroll
      for (int u = 0; u < NQ; ++u) {
        if (u < NQ - 1 || last_ok) {
          const int q = lane + u * 32;
          if (ro.f32) ro.f32[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) *reinterpret_cast<uint2*>(ro.pl[pl] + q * 4) = make_uint2(0u, 0u);
        }
      }
    } else {
      float4 acc[NQ];
#pragma unroll
      for (int u = 0; u < NQ; ++u) {
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.accumulate && (u < NQ - 1 || last_ok)) acc[u] = ro.f32[lane + u * 32];   // in flight during the edge loop
      }
      const bool dropped = masked && m.keep[i] == 0;   // a dropped node keeps only its edges to kept nodes (wrapper.py:221-225)
      if (all_in_smem && !dropped) {
#pragma unroll 2
        for (int e = ea; e < eb; ++e) {
          const float2 en = m.ent[e];                  // one broadcast shared load
          const uint32_t ra = tl + (uint32_t)__float_as_int(en.x) * pitch;
#pragma unroll
          for (int u = 0; u < NQ; ++u) {
            if (u < NQ - 1 || last_ok) {
              const float4 f = lds128f(ra + u * 512);
              acc[u].x = fmaf(en.y, f.x, acc[u].x); acc[u].y = fmaf(en.y, f.y, acc[u].y);
              acc[u].z = fmaf(en.y, f.z, acc[u].z); acc[u].w = fmaf(en.y, f.w, acc[u].w);
            }
          }
        }
      } else {
        for (int e = ea; e < eb; ++e) {
          const float2 en = list_entry(p, m, gent, e);
          const int j = __float_as_int(en.x);
          if (dropped && !m.keep[j]) continue;         // warp-uniform
          GR_EDGE(j, en.y)
        }
      }
#pragma unroll
      for (int u = 0; u < NQ; ++u)
        if (u < NQ - 1 || last_ok) store_quad<NP>(ro, lane + u * 32, acc[u]);
    }
    if (NP && lane < npq) store_pad_quad<NP>(ro, H, lane, p.pad_one != 0);
    if (ro.f32) ro.f32 += GR_WARPS * HQ;               // next row of this warp
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) ro.pl[pl] += step_pl;
  }
#undef GR_EDGE
#ifdef GETB_GRAPH_TIMELINE
  if (lane == 0 && warp == GR_WARPS - 1) trow[6] = clock64() - tr0;
  __syncthreads();
  GR_T(5);
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 100 || blockIdx.x == 219 || blockIdx.x == 3000))
    printf("GRDBG cta %d fused %d used %d: lists %lld scoring %lld tile_landed %lld dropout %lld lastwarp_done %lld all_done %lld\n", blockIdx.x,
           (int)FUSED, n_used, trow[1], trow[2], trow[3], trow[4], trow[6], trow[5]);
#endif
}

typedef void (*GatherFn)(const GatherParams);
template <bool FUSED, int NP, bool SPILL>
static GatherFn gather_row_fn_nq(int nq) {
  switch (nq) {
    case 1: return gather_row_kernel<FUSED, 1, NP, SPILL>;
    case 2: return gather_row_kernel<FUSED, 2, NP, SPILL>;
    case 3: return gather_row_kernel<FUSED, 3, NP, SPILL>;
    default: return gather_row_kernel<FUSED, 4, NP, SPILL>;
  }
}
template <bool FUSED, bool SPILL>
static GatherFn gather_row_fn_np(int nq, int np) {
  switch (np) {
    case 0: return gather_row_fn_nq<FUSED, 0, SPILL>(nq);
    case 1: return gather_row_fn_nq<FUSED, 1, SPILL>(nq);
    case 2: return gather_row_fn_nq<FUSED, 2, SPILL>(nq);
    default: return gather_row_fn_nq<FUSED, 3, SPILL>(nq);
  }
}
static GatherFn gather_row_fn(bool fused, int nq, int np, bool spill) {
  if (spill) return fused ? gather_row_fn_np<true, true>(nq, np) : gather_row_fn_np<false, true>(nq, np);
  return fused ? gather_row_fn_np<true, false>(nq, np) : gather_row_fn_np<false, false>(nq, np);
}

// =====================================================================================================
// Column-slice kernel (graphs whose feature tile does not fit: R=200, H=512): one 256-thread CTA per (graph, slice of
// <= 48 column quads); half a warp per output row, up to three quads per lane; the per-graph scoring is recomputed by each
// slice. Tile rows arrive by one bulk copy per row (a row slice is contiguous).
// smem: [tile N*qs float4][GraphSmem]
// =====================================================================================================
constexpr int GC_THREADS = 256;
constexpr int GC_HALVES = GC_THREADS / 16;

template <bool FUSED, int NQ, int NP>
__global__ void __launch_bounds__(GC_THREADS) gather_cols_kernel(const __grid_constant__ GatherParams p) {
  extern __shared__ __align__(128) float4 gc_tile[];
  const int g = blockIdx.x / p.nsplit, slice = blockIdx.x - g * p.nsplit;
  const int tid = threadIdx.x;
  const int N = p.N, H = p.H, HQ = H >> 2;
  const int q0 = slice * p.qs;
  const int WQ = min(p.qs, HQ - q0);
  const int64_t row0 = (int64_t)g * N;
  const GraphSmem m = graph_smem(gc_tile + (size_t)N * p.qs, N, p.lcap);
  const float2* gent = p.ent + (int64_t)g * gl_ent_cap(N);
  const bool drop = FUSED && p.thr != 0;
  const int n_used = p.used ? min(N, __ldg(p.used + g)) : N;     // feature rows that can be gathered at all
  const bool issuer = tid == GC_THREADS - 32;

  float spv = 0.f;
  if (FUSED && tid < N) spv = load_sp(p, row0, tid);
  if (issuer) {
    issue_list_copy(p, m, g);
    gr_expect_tx(m.bar + 1, (uint32_t)n_used * (uint32_t)WQ * 16u);
  }
  __syncthreads();            // the mbarrier inits are visible
  if (tid >= GC_THREADS - 32) {                                  // tile rows: one bulk copy each, issued by the last warp
    const char* src = reinterpret_cast<const char*>(p.x + row0 * H) + (size_t)q0 * 16;
    for (int r = tid - (GC_THREADS - 32); r < n_used; r += 32)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(gr_smem_u32(gc_tile + (size_t)r * p.qs)), "l"(src + (size_t)r * H * 4), "r"((uint32_t)WQ * 16u), "r"(gr_smem_u32(m.bar + 1)) : "memory");
  }
  finish_list_copy(p, m, g, issuer);
  if (FUSED) {
    score_topk<GC_THREADS>(p, m, gent, row0, tid, spv, slice == 0);
  } else if (p.keep_in) {
    if (tid < N) m.keep[tid] = p.keep_in[row0 + tid];
    __syncthreads();
  }
  const bool masked = FUSED || (p.keep_in != nullptr);
  if (n_used) gr_mbar_wait(m.bar + 1, 0);
  if (drop) {
    dropout_tile<GC_THREADS>(p, gc_tile, p.qs, n_used, WQ, q0, row0, tid);
    __syncthreads();
  }

  // ---- out[i, slice] = sum_e w_e * x[j_e, slice]: half a warp per output row, lanes own quads hl, hl+16, hl+32 -------
  const int half = tid >> 4, hl = tid & 15;
  const bool ok1 = NQ > 1 && (hl + 16) < WQ, ok2 = NQ > 2 && (hl + 32) < WQ, ok0 = hl < WQ;
  const int npq = (NP && slice == p.nsplit - 1) ? ((((H + (p.pad_one ? 1 : 0)) + 7) & ~7) - H) >> 2 : 0;
  const uint32_t tl = gr_smem_u32(gc_tile) + (uint32_t)hl * 16u;
  const uint32_t pitch = (uint32_t)p.qs * 16u;
  for (int i = half; i < N; i += GC_HALVES) {
    const RowOut ro = row_out(p, row0 + i);
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
    if (p.accumulate) {                               // in flight during the edge loop
      if (ok0) a0 = ro.f32[q0 + hl];
      if (ok1) a1 = ro.f32[q0 + hl + 16];
      if (ok2) a2 = ro.f32[q0 + hl + 32];
    }
    const bool dropped = masked && m.keep[i] == 0;    // a dropped node keeps only its edges to kept nodes (wrapper.py:221-225)
    const int eb = m.rowptr[i + 1];
    for (int e = m.rowptr[i]; e < eb; ++e) {
      const float2 en = list_entry(p, m, gent, e);
      const int j = __float_as_int(en.x);
      const float w = en.y;
      if (dropped && !m.keep[j]) continue;
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
    if (ok0) store_quad<NP>(ro, q0 + hl, a0);
    if (ok1) store_quad<NP>(ro, q0 + hl + 16, a1);
    if (ok2) store_quad<NP>(ro, q0 + hl + 32, a2);
    if (NP && hl < npq) store_pad_quad<NP>(ro, H, hl, p.pad_one != 0);   // last slice only
  }
}

template <bool FUSED, int NP>
static GatherFn gather_cols_fn_nq(int nq) {
  switch (nq) {
    case 1: return gather_cols_kernel<FUSED, 1, NP>;
    case 2: return gather_cols_kernel<FUSED, 2, NP>;
    default: return gather_cols_kernel<FUSED, 3, NP>;
  }
}
static GatherFn gather_cols_fn(bool fused, int nq, int np) {
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
  GETB_REQUIRE(p.ent && p.rowptr && p.x && (p.out || p.out_p) && aligned16(p.x) && (!p.out || aligned16(p.out)) && aligned16(p.ent) &&
                   aligned16(p.rowptr), "%s: null / misaligned pointer", name);
  GETB_REQUIRE(!p.accumulate || p.out, "%s: accumulate needs the fp32 output", name);
  if (p.out_p)
    GETB_REQUIRE((((uintptr_t)p.out_p) & 7u) == 0 && (p.ld_p % 4) == 0 && (p.ps_p % 4) == 0 && p.np_p >= 1 && p.np_p <= 3 &&
                     p.ld_p >= ((p.H + (p.pad_one ? 1 : 0) + 7) & ~7),
                 "%s: plane output needs an 8-byte aligned tensor with room for the padding", name);
  if (p.G == 0) return 0;
  const int np = p.out_p ? p.np_p : 0;
  static int which = -1;     // 0 auto: whole graph when most rows fit the tile, else column slices; 1 column slices always
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
  // list entries in shared memory: 16 per node (a window-3 graph has ~4.5, window 9 ~14), the whole record when smaller;
  // the first copy fetches 6 per node whatever the graph holds (never past the graph's own block)
  p.lcap = std::min(gl_ent_cap(p.N), std::max(16 * p.N, 64));
  p.e1 = std::min(p.lcap, std::max(6 * p.N, 64));
  const size_t smem_small = graph_smem_bytes(p.N, p.lcap) + 16;
  // whole-graph kernel, two CTAs per SM: the tile holds as many feature rows as fit 113 KB (83 of 100 at Snopes dims; a text
  // uses `used` <= N rows, typically ~75); worthwhile while at least 3/4 of the rows fit
  const int cap_rows = smem_small < GR_SMEM_HALF ? (int)std::min<size_t>((size_t)p.N, (GR_SMEM_HALF - smem_small) / ((size_t)p.H * 4)) : 0;
  if (which == 0 && cap_rows * 4 >= p.N * 3 && p.H <= 512) {
    const size_t smem_row = (size_t)cap_rows * p.H * 4 + smem_small;
    const int nq = (p.H / 4 + 31) / 32;
    p.tile_rows = cap_rows;
    GatherFn fn = gather_row_fn(fused, nq, np, cap_rows < p.N);
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
  GatherFn fn = gather_cols_fn(fused, (p.qs + 15) / 16, np);
  if (!opt_in((const void*)fn, smem)) return -2;
  fn<<<p.G * p.nsplit, GC_THREADS, smem, st>>>(p);
  GETB_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace getb

using namespace getb;

extern "C" int get_neighbor_lists_rowptr_pitch(int N) { return gl_rowptr_pitch(N); }
extern "C" int get_neighbor_lists_entry_capacity(int N) { return gl_ent_cap(N); }

extern "C" int get_build_neighbor_lists(const float* adj, int G, int N, void* ent, int32_t* rowptr, int32_t* used, void* stream) {
  GETB_REQUIRE(adj && ent && rowptr && used && G >= 0 && N >= 1 && N <= GL_MAX_N, "get_build_neighbor_lists: bad arguments (N <= %d)", GL_MAX_N);
  GETB_REQUIRE(aligned16(ent) && aligned16(rowptr), "get_build_neighbor_lists: records must be 16-byte aligned");
  if (G == 0) return 0;
  const size_t smem = (size_t)N * (N + 1) * sizeof(float) + (size_t)(N + 2) * sizeof(int);
  if (smem > 48 * 1024) {
    static bool done = false;
    if (!done) {
      if (cudaFuncSetAttribute(build_neighbor_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024) != cudaSuccess) {
        (void)cudaGetLastError();
        getb::set_error("get_build_neighbor_lists: cannot opt in to large shared memory");
        return -2;
      }
      done = true;
    }
  }
  GETB_REQUIRE(smem <= 224 * 1024, "get_build_neighbor_lists: adjacency tile of %zu bytes does not fit shared memory", smem);
  build_neighbor_lists_kernel<<<G, 256, smem, (cudaStream_t)stream>>>(adj, N, reinterpret_cast<float2*>(ent), rowptr, used);
  GETB_CHECK_LAUNCH("get_build_neighbor_lists");
  return 0;
}

extern "C" int get_graph_gather(const void* ent, const int32_t* rowptr, const int32_t* used, const float* x, const uint8_t* keep, float* out,
                                void* planes, int64_t ld_p, int64_t plane_stride, int nplanes, int pad_one, int G, int N, int H,
                                int accumulate, void* stream) {
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.ent = reinterpret_cast<const float2*>(ent); p.rowptr = rowptr; p.used = used; p.x = x; p.keep_in = keep; p.out = out;
  p.out_p = reinterpret_cast<__nv_bfloat16*>(planes); p.ld_p = ld_p; p.ps_p = plane_stride; p.np_p = nplanes; p.pad_one = pad_one;
  p.G = G; p.N = N; p.H = H; p.accumulate = accumulate;
  return launch_gather(p, false, (cudaStream_t)stream, "get_graph_gather");
}

extern "C" int get_gsl_gather(const void* ent, const int32_t* rowptr, const int32_t* used, const float* F, const float* sp_parts, int n_sp,
                              const float* gate, int G, int N, int H, int k, float drop_p, uint32_t seed_layer2, float* score,
                              uint8_t* keep, float* out, void* planes, int64_t ld_p, int64_t plane_stride, int nplanes, void* stream) {
  GETB_REQUIRE(sp_parts && n_sp >= 1 && gate && keep, "get_gsl_gather: null pointer");
  GETB_REQUIRE(k >= 0 && k <= N, "get_gsl_gather: k=%d out of [0,%d]", k, N);
  GETB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "get_gsl_gather: dropout probability must be in [0,1)");
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.ent = reinterpret_cast<const float2*>(ent); p.rowptr = rowptr; p.used = used; p.x = F; p.out = out;
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
