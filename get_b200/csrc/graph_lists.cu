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

namespace getb {

constexpr int GL_THREADS = 256;
constexpr int GL_MAX_N = 232;   // the builder stages one dense N x (N+1) fp32 tile in shared memory

struct GatherParams {
  const float2* nbr;      // (G, N, N) {neighbour index as int bits, weight}; row i of graph g at (g*N + i)*N
  const int32_t* cnt;     // (G, N) entries per row
  const float* x;         // (G, N, H)
  const uint8_t* keep_in; // (G, N) or null (non-fused)
  float* out;             // (G, N, H) or null
  __nv_bfloat16* out_p; int64_t ld_p, ps_p; int np_p, pad_one;
  int G, N, H, accumulate;
  // fused part
  const float* sp_parts; int n_sp;
  const float* gate;
  int k, np2_shift;
  int nsplit, qs;         // column slices per graph, float4 quads per slice (<= 16)
  uint32_t thr; float scale; uint32_t seed_2; const uint32_t* salt;
  float* score; uint8_t* keep_out;
};

// ---- adjacency -> lists (both orientations), one CTA per graph ---------------------------------------------------------
__global__ void __launch_bounds__(256) build_neighbor_lists_kernel(const float* __restrict__ adj, int N, float2* __restrict__ nbr,
                                                                   int32_t* __restrict__ cnt, float2* __restrict__ nbr_t,
                                                                   int32_t* __restrict__ cnt_t) {
  extern __shared__ float tile[];                    // N x (N + 1): the dense adjacency of this graph
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
  for (int r = warp; r < 2 * N; r += 8) {            // rows of adj, then rows of adj^T
    const bool tr = r >= N;
    const int i = tr ? r - N : r;
    float2* lr = (tr ? nbr_t : nbr) + ((int64_t)g * N + i) * N;
    int pos = 0;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const float w = j < N ? (tr ? tile[j * P + i] : tile[i * P + j]) : 0.f;
      const unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
      if (w != 0.f) lr[pos + __popc(nz & ((1u << lane) - 1u))] = make_float2(__int_as_float(j), w);
      pos += __popc(nz);
    }
    if (lane == 0) (tr ? cnt_t : cnt)[(int64_t)g * N + i] = pos;
  }
}

__device__ __forceinline__ void gl_store_quad(const GatherParams& p, int64_t row, int q, const float4& v) {
  if (p.out) *(reinterpret_cast<float4*>(p.out + row * p.H) + q) = v;
  if (p.out_p) {
    // two / three bf16 planes of 4 values: packed conversions (cvt.rn.bf16x2.f32), residuals exact in fp32
    __nv_bfloat16* d = p.out_p + row * p.ld_p + q * 4;
    float r0 = v.x, r1 = v.y, r2 = v.z, r3 = v.w;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      if (pl < p.np_p) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(r0, r1), b = __floats2bfloat162_rn(r2, r3);
        uint2 w;
        w.x = *reinterpret_cast<const uint32_t*>(&a);
        w.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(d + (int64_t)pl * p.ps_p) = w;
        r0 -= __uint_as_float(w.x << 16); r1 -= __uint_as_float(w.x & 0xFFFF0000u);
        r2 -= __uint_as_float(w.y << 16); r3 -= __uint_as_float(w.y & 0xFFFF0000u);
      }
    }
  }
}

__device__ __forceinline__ void gl_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// One CTA = (graph g, column slice): the slice's feature tile (N rows x WQ quads, WQ <= 16) is staged in shared memory by
// asynchronous 16-byte copies issued FIRST; the per-graph scoring (fused kernel) runs in the shadow of that load, recomputed
// by every slice of the graph (a few hundred instructions) so that slices stay independent CTAs: a 220-graph launch is
// 1100 work items, ~7 co-resident per SM, instead of 1.5 whole graphs per SM.
// smem: [tile N*WQ float4][sp N f32][score N f32][rank N i32][keep N u8 (padded)]
template <bool FUSED>
__global__ void __launch_bounds__(GL_THREADS) gather_kernel(const __grid_constant__ GatherParams p) {
  extern __shared__ __align__(16) float4 gl_tile[];
  const int g = blockIdx.x / p.nsplit, slice = blockIdx.x - g * p.nsplit;
  const int tid = threadIdx.x, lane = tid & 31;
  const int N = p.N, H = p.H, HQ = H >> 2;
  const int q0 = slice * p.qs;                 // first global quad of this slice
  const int WQ = min(p.qs, HQ - q0);           // quads owned (<= 16)
  const int64_t row0 = (int64_t)g * N;
  float* s_sp = reinterpret_cast<float*>(gl_tile + (size_t)N * p.qs);
  float* s_score = s_sp + N;
  int* s_rank = reinterpret_cast<int*>(s_score + N);
  uint8_t* s_keep = reinterpret_cast<uint8_t*>(s_rank + N);
  const bool drop = FUSED && p.thr != 0;
#ifdef GETB_GRAPH_TIMELINE
  __shared__ long long tl[8];
  const long long t0 = clock64();
#define GL_T(i) do { if (tid == 0) tl[i] = clock64() - t0; } while (0)
#else
#define GL_T(i) do { } while (0)
#endif

  // ---- feature tile: all copies in flight at once ---------------------------------------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(p.x + row0 * H) + q0;
    const int total = N * WQ;
    for (int idx = tid; idx < total; idx += GL_THREADS) {
      const int r = idx / WQ, q = idx - r * WQ;
      gl_cp_async16(gl_tile + r * p.qs + q, src + (int64_t)r * HQ + q);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  if (FUSED) {
    if (tid < N) {
      float v = 0.f;
#pragma unroll 8
      for (int q = 0; q < p.n_sp; ++q) v += __ldg(p.sp_parts + (int64_t)q * p.G * N + row0 + tid);   // fixed order
      s_sp[tid] = v;
      s_rank[tid] = 0;
    }
    __syncthreads();
    GL_T(0);
    // ---- s_a = adj @ s_p over the lists + scalar GRU gates (GGNN with out_features = 1), one thread per node ----
    if (tid < N) {
      const float2* lr = p.nbr + (row0 + tid) * N;
      const int cnt = __ldg(p.cnt + row0 + tid);
      float sa = 0.f;
      for (int e = 0; e < cnt; ++e) {
        const float2 en = __ldg(lr + e);
        sa = fmaf(en.y, s_sp[__float_as_int(en.x)], sa);
      }
      const float wz0 = __ldg(p.gate + 0), bz0 = __ldg(p.gate + 1), wz1 = __ldg(p.gate + 2), bz1 = __ldg(p.gate + 3);
      const float wr0 = __ldg(p.gate + 4), br0 = __ldg(p.gate + 5), wr1 = __ldg(p.gate + 6), br1 = __ldg(p.gate + 7);
      const float wh0 = __ldg(p.gate + 8), bh0 = __ldg(p.gate + 9), wh1 = __ldg(p.gate + 10), bh1 = __ldg(p.gate + 11);
      const float spv = s_sp[tid];
      const float z = sigmoidf_((wz0 * sa + bz0) + (wz1 * spv + bz1));
      const float r = sigmoidf_((wr0 * sa + br0) + (wr1 * spv + br1));
      const float h = tanhf((wh0 * sa + bh0) + (wh1 * (r * spv) + bh1));
      const float sc = h * z + spv * (1.0f - z);
      s_score[tid] = sc;
      if (slice == 0 && p.score) p.score[row0 + tid] = sc;
    }
    __syncthreads();
    GL_T(1);
    // ---- top-k by rank counting: thread (node i, slice of candidates); ties -> lower index first ------------------
    {
      const int nsl = GL_THREADS >> p.np2_shift;         // candidate slices per node (N <= 2^np2_shift <= GL_THREADS)
      const int per = (N + nsl - 1) / nsl;
      const int i = tid & ((1 << p.np2_shift) - 1), j0 = (tid >> p.np2_shift) * per;
      if (i < N && j0 < N) {
        const float si = s_score[i];
        const int j1 = min(N, j0 + per);
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
      if (slice == 0) p.keep_out[row0 + tid] = kp;
    }
    GL_T(2);
  } else if (p.keep_in) {
    if (tid < N) s_keep[tid] = p.keep_in[row0 + tid];
  }
  const bool masked = FUSED || (p.keep_in != nullptr);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  GL_T(3);

  if (drop) {
    // feat_prop2's nn.Dropout draw, applied once to the staged tile (element index = position in the (G*N, H) tensor)
    const uint32_t seed = p.seed_2 + __ldg(p.salt);
    const int total = N * WQ;
    for (int idx = tid; idx < total; idx += GL_THREADS) {
      const int r = idx / WQ, q = idx - r * WQ;
      float4 f = gl_tile[r * p.qs + q];
      drop_apply4(seed, (uint64_t)(row0 + r) * (uint64_t)H + (uint64_t)(q0 + q) * 4, p.thr, p.scale, f);
      gl_tile[r * p.qs + q] = f;
    }
    __syncthreads();
  }
  GL_T(4);

  // ---- out[i, slice] = sum_e w_e * x[j_e, slice]: half a warp per output row, one quad per lane; lane e of the half holds
  // entry e of the row's list (fetched one row ahead), broadcast by shuffles
  const int half = (tid >> 4), hl = tid & 15, hsh = lane & 16;        // 16 rows in flight per CTA
  const int npq = (p.out_p && slice == p.nsplit - 1) ? ((((H + (p.pad_one ? 1 : 0)) + 7) & ~7) - H) >> 2 : 0;
  const int lcap = hl < N ? hl : N - 1;             // entries beyond cnt are allocated, unread garbage
  constexpr int GL_HALVES = GL_THREADS / 16;
  float2 nxt_e = make_float2(0.f, 0.f);
  int nxt_c = 0;
  if (half < N) {
    nxt_e = __ldg(p.nbr + (row0 + half) * N + lcap);
    nxt_c = __ldg(p.cnt + row0 + half);
  }
  const int rounds = (N + GL_HALVES - 1) / GL_HALVES;
  for (int rd = 0; rd < rounds; ++rd) {             // warp-uniform trip count (shuffles inside)
    const int i = half + rd * GL_HALVES;
    const bool row_ok = i < N;
    const int cnt = row_ok ? nxt_c : 0;
    float2 my = nxt_e;
    if (i + GL_HALVES < N) {
      nxt_e = __ldg(p.nbr + (row0 + i + GL_HALVES) * N + lcap);
      nxt_c = __ldg(p.cnt + row0 + i + GL_HALVES);
    }
    const bool dropped = masked && row_ok && s_keep[i] == 0;   // a dropped node keeps only its edges to kept nodes (wrapper.py:221-225)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int cmax = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, 16));
    for (int e0 = 0; e0 < cmax; e0 += 16) {
      if (e0) my = (e0 + hl < cnt) ? __ldg(p.nbr + (row0 + i) * N + e0 + hl) : make_float2(0.f, 0.f);
      const int jm = __float_as_int(my.x);
      const bool valid = (e0 + hl < cnt) && !(dropped && !s_keep[jm]);
      const unsigned both = __ballot_sync(0xffffffffu, valid);
      unsigned live = (both >> hsh) & 0xFFFFu;
      int n = __popc(live);
      const int nmax = max(__popc(both & 0xFFFFu), __popc(both >> 16));
      for (int t = 0; t < nmax; t += 2) {
        int s0 = 0, s1 = 0;
        float w0 = 0.f, w1 = 0.f;
        if (live) { s0 = __ffs(live) - 1; live &= live - 1; w0 = 1.f; }
        if (live) { s1 = __ffs(live) - 1; live &= live - 1; w1 = 1.f; }
        const int ja = __shfl_sync(0xffffffffu, jm, s0 + hsh), jb = __shfl_sync(0xffffffffu, jm, s1 + hsh);
        const float wa = __shfl_sync(0xffffffffu, my.y, s0 + hsh), wb = __shfl_sync(0xffffffffu, my.y, s1 + hsh);
        if (hl < WQ) {
          if (w0 != 0.f) {
            const float4 f = gl_tile[ja * p.qs + hl];
            acc.x = fmaf(wa, f.x, acc.x); acc.y = fmaf(wa, f.y, acc.y); acc.z = fmaf(wa, f.z, acc.z); acc.w = fmaf(wa, f.w, acc.w);
          }
          if (w1 != 0.f) {
            const float4 f = gl_tile[jb * p.qs + hl];
            acc.x = fmaf(wb, f.x, acc.x); acc.y = fmaf(wb, f.y, acc.y); acc.z = fmaf(wb, f.z, acc.z); acc.w = fmaf(wb, f.w, acc.w);
          }
        }
      }
      (void)n;
    }
    if (row_ok) {
      if (hl < WQ) {
        const int q = q0 + hl;
        if (p.accumulate) {
          const float4 o = *(reinterpret_cast<const float4*>(p.out + (row0 + i) * H) + q);
          acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        gl_store_quad(p, row0 + i, q, acc);
      }
      if (hl < npq) {     // padding quads of the plane row (room for the ones column when pad_one): last slice only
        const float vv[4] = {(p.pad_one && hl == 0) ? 1.0f : 0.0f, 0.f, 0.f, 0.f};
        planes_store4(p.out_p + (row0 + i) * p.ld_p + H + hl * 4, p.ps_p, p.np_p, vv);
      }
    }
  }
#ifdef GETB_GRAPH_TIMELINE
  __syncthreads();
  GL_T(5);
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 500 || blockIdx.x == 1099 || blockIdx.x == 15000)) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    printf("GLDBG cta %d sm %u fused %d: sp %lld spmv %lld topk %lld tile_landed %lld dropout %lld aggregate+store %lld\n", blockIdx.x, smid,
           (int)FUSED, tl[0], tl[1], tl[2], tl[3], tl[4], tl[5]);
  }
#endif
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
  // column slices of at most 16 quads, as even as possible (H = 300: 5 slices of 15 quads)
  const int HQ = p.H / 4;
  p.nsplit = (HQ + 15) / 16;
  p.qs = (HQ + p.nsplit - 1) / p.nsplit;
  GETB_REQUIRE((int64_t)p.G * p.nsplit < (1LL << 31), "%s: too many work items", name);
  const size_t smem = (size_t)p.N * p.qs * 16 + (size_t)3 * p.N * 4 + (((size_t)p.N + 15) & ~(size_t)15) + 16;
  auto fn = fused ? gather_kernel<true> : gather_kernel<false>;
  if (smem > 48 * 1024) {
    static bool done[2] = {false, false};
    if (!done[fused]) {
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
        (void)cudaGetLastError();
        getb::set_error("graph gather: cannot opt in to large shared memory");
        return -2;
      }
      done[fused] = true;
    }
  }
  GETB_REQUIRE(smem <= 100 * 1024, "%s: %zu bytes of shared memory", name, smem);
  fn<<<p.G * p.nsplit, GL_THREADS, smem, st>>>(p);
  GETB_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace getb

using namespace getb;

extern "C" int get_build_neighbor_lists(const float* adj, int G, int N, void* nbr, int32_t* cnt, void* nbr_t, int32_t* cnt_t,
                                        void* stream) {
  GETB_REQUIRE(adj && nbr && cnt && nbr_t && cnt_t && G >= 0 && N >= 1 && N <= GL_MAX_N, "get_build_neighbor_lists: bad arguments (N <= %d)", GL_MAX_N);
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
                                                                      reinterpret_cast<float2*>(nbr_t), cnt_t);
  GETB_CHECK_LAUNCH("get_build_neighbor_lists");
  return 0;
}

extern "C" int get_graph_gather(const void* nbr, const int32_t* cnt, const float* x, const uint8_t* keep, float* out, void* planes,
                                int64_t ld_p, int64_t plane_stride, int nplanes, int pad_one, int G, int N, int H, int accumulate,
                                void* stream) {
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.nbr = reinterpret_cast<const float2*>(nbr); p.cnt = cnt; p.x = x; p.keep_in = keep; p.out = out;
  p.out_p = reinterpret_cast<__nv_bfloat16*>(planes); p.ld_p = ld_p; p.ps_p = plane_stride; p.np_p = nplanes; p.pad_one = pad_one;
  p.G = G; p.N = N; p.H = H; p.accumulate = accumulate;
  return launch_gather(p, false, (cudaStream_t)stream, "get_graph_gather");
}

extern "C" int get_gsl_gather(const void* nbr, const int32_t* cnt, const float* F, const float* sp_parts, int n_sp, const float* gate,
                              int G, int N, int H, int k, float drop_p, uint32_t seed_layer2, float* score, uint8_t* keep, float* out,
                              void* planes, int64_t ld_p, int64_t plane_stride, int nplanes, void* stream) {
  GETB_REQUIRE(sp_parts && n_sp >= 1 && gate && keep, "get_gsl_gather: null pointer");
  GETB_REQUIRE(k >= 0 && k <= N, "get_gsl_gather: k=%d out of [0,%d]", k, N);
  GETB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "get_gsl_gather: dropout probability must be in [0,1)");
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.nbr = reinterpret_cast<const float2*>(nbr); p.cnt = cnt; p.x = F; p.out = out;
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
