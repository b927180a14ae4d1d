// tcgen05 (5th-gen tensor core) contraction with fp32-level accuracy: error-compensated 3xTF32
//   A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi        (hi = tf32 round-to-nearest, lo = x - hi)
// for the dense `Linear`s of the GET hot path (reference Models/BiDAF/wrapper.py:191,194-204;
// thirdparty/two_branches_attention.py:140) and their dX backward. The top-k cliff of the GSL scorer
// (SURVEY.md section 7 "hard parts") forbids plain TF32/BF16 on this chain; 3xTF32 keeps ~2^-22 relative error.
//
// One CTA computes a 128 x BN (BN <= 256) output tile, accumulators live in TMEM:
//   warps 0-3 : A producers -- coalesced 128-bit global loads (optional row gather = embedding lookup, optional
//               counter-based dropout), hi/lo split, stores into 128B-swizzled K-major shared-memory tiles; after
//               the main loop the same warps run the epilogue (tcgen05.ld -> fused epilogue -> global)
//   warp 4    : TMEM allocation + single-thread tcgen05.mma issue (kind::tf32, M=128, N=BN, K=8 per instruction)
//   warp 5    : TMA producer for the pre-split weight tiles B_hi / B_lo (cp.async.bulk.tensor, SWIZZLE_128B)
// Full/empty mbarrier ring of `stages` slots; tcgen05.commit releases a slot when the MMAs that read it retire.
#include <cuda.h>
#include <string.h>

#include "gemm_common.cuh"

namespace getb {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                       // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_THREADS = 192;
constexpr int TC_A_TILE_BYTES = TC_BM * 128;    // 16 KB
constexpr int TC_MAX_STAGES = 4;
constexpr uint32_t TC_SPIN_LIMIT = 1u << 28;

struct TcMaps {
  CUtensorMap hi[GET_GEMM_MAX_SEG];
  CUtensorMap lo[GET_GEMM_MAX_SEG];
};

struct TcCfg {
  int BN;          // output-tile columns, multiple of 16, <= 256
  int stages;
  int tmem_cols;   // power of two >= BN
  int kblocks[GET_GEMM_MAX_SEG];
  int kblocks_total;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
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
    if (++spins > TC_SPIN_LIMIT) __trap();   // a protocol bug must fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled shared-memory operand descriptor (rows at a 128-byte pitch, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t f32_to_tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// ---- kernel -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ TcCfg cfg, const __grid_constant__ TcMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_fullA[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_fullB[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_accum;
  __shared__ uint32_t tmem_holder;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = cfg.BN;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * TC_BM;
  const uint32_t b_tile_bytes = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = 2u * TC_A_TILE_BYTES + 2u * b_tile_bytes;
  // 1024-byte aligned base (SWIZZLE_128B atoms)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  if (tid == 0) {
    for (int s = 0; s < cfg.stages; ++s) {
      mbar_init(&bar_fullA[s], 128);
      mbar_init(&bar_fullB[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_accum, 1);
    fence_barrier_init();
  }
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&maps.hi[s]);
      tma_prefetch_desc(&maps.lo[s]);
    }
  }
  if (warp == 4) tmem_alloc(&tmem_holder, (uint32_t)cfg.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const int nkb = cfg.kblocks_total;

  if (warp < 4) {
    // =============================== A producers ===============================
    const int c = tid & 7;            // 16-byte chunk within the 128-byte row
    const int r0 = tid >> 3;          // rows r0 + 16*i
    int stage = 0;
    uint32_t phase = 0;
    int seg = 0, kb_in_seg = 0;
    const float* rowptr[8];
    bool rowok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rowptr[i] = nullptr;
    int cur_seg = -1;
    for (int kb = 0; kb < nkb; ++kb) {
      if (seg != cur_seg) {           // (re)compute the row base pointers of this segment
        const GemmOp& op = p.A[seg];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m0 + r0 + 16 * i;
          rowok[i] = m < p.M;
          int64_t srow = m;
          if (rowok[i] && op.rowidx) srow = op.rowidx[m];
          rowptr[i] = op.ptr + srow * op.ld;
        }
        cur_seg = seg;
      }
      const int k = kb_in_seg * TC_BK + c * 4;
      const bool kok = k < p.K[seg];
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kok && rowok[i]) v[i] = __ldg(reinterpret_cast<const float4*>(rowptr[i] + k));
      }
      if (seg == 0 && p.drop_thr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint64_t base = (uint64_t)(m0 + r0 + 16 * i) * (uint64_t)p.drop_cols + (uint64_t)k;
          drop_apply4(p.drop_seed + __ldg(p.salt), base, p.drop_thr, p.drop_scale, v[i]);
        }
      }
      mbar_wait(&bar_empty[stage], phase ^ 1);
      uint8_t* a_hi = smem + (size_t)stage * stage_bytes;
      uint8_t* a_lo = a_hi + TC_A_TILE_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + 16 * i;
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
        uint4 hi, lo;
        hi.x = f32_to_tf32_rn(v[i].x); hi.y = f32_to_tf32_rn(v[i].y);
        hi.z = f32_to_tf32_rn(v[i].z); hi.w = f32_to_tf32_rn(v[i].w);
        lo.x = __float_as_uint(v[i].x - __uint_as_float(hi.x));
        lo.y = __float_as_uint(v[i].y - __uint_as_float(hi.y));
        lo.z = __float_as_uint(v[i].z - __uint_as_float(hi.z));
        lo.w = __float_as_uint(v[i].w - __uint_as_float(hi.w));
        *reinterpret_cast<uint4*>(a_hi + off) = hi;
        *reinterpret_cast<uint4*>(a_lo + off) = lo;
      }
      fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&bar_fullA[stage]);
      if (++stage == cfg.stages) { stage = 0; phase ^= 1; }
      if (++kb_in_seg == cfg.kblocks[seg]) { kb_in_seg = 0; ++seg; }
    }
    // =============================== epilogue ==================================
    mbar_wait(&bar_accum, 0);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;          // TMEM lane == output row within the tile
    const uint32_t tbase = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int col = 0; col < BN; col += 16) {
      float acc[16];
      tmem_ld16(tbase + (uint32_t)col, acc);      // warp-collective: executed by every lane
      if (m < p.M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = n0 + col + q * 4;
          if (n < p.N) epilogue4(p, m, n, &acc[q * 4]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =============================== MMA issuer ================================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&bar_fullA[stage], phase);
        mbar_wait(&bar_fullB[stage], phase);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t a_lo = a_hi + TC_A_TILE_BYTES;
        const uint32_t b_hi = a_lo + TC_A_TILE_BYTES;
        const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ++ks) {
          const uint64_t dah = smem_desc_sw128(a_hi + ks * 32), dal = smem_desc_sw128(a_lo + ks * 32);
          const uint64_t dbh = smem_desc_sw128(b_hi + ks * 32), dbl = smem_desc_sw128(b_lo + ks * 32);
          umma_tf32(tmem_base, dal, dbh, idesc, (kb | ks) ? 1u : 0u);   // small terms first
          umma_tf32(tmem_base, dah, dbl, idesc, 1u);
          umma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        umma_commit(&bar_empty[stage]);           // slot is free once these MMAs have read it
        if (kb == nkb - 1) umma_commit(&bar_accum);
        if (++stage == cfg.stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =============================== TMA producer for B =========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int seg = 0, kb_in_seg = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&bar_empty[stage], phase ^ 1);
        uint8_t* b_hi = smem + (size_t)stage * stage_bytes + 2 * TC_A_TILE_BYTES;
        uint8_t* b_lo = b_hi + b_tile_bytes;
        mbar_arrive_expect_tx(&bar_fullB[stage], 2u * b_tile_bytes);
        tma_load_2d(b_hi, &maps.hi[seg], kb_in_seg * TC_BK, n0, &bar_fullB[stage]);
        tma_load_2d(b_lo, &maps.lo[seg], kb_in_seg * TC_BK, n0, &bar_fullB[stage]);
        if (++stage == cfg.stages) { stage = 0; phase ^= 1; }
        if (++kb_in_seg == cfg.kblocks[seg]) { kb_in_seg = 0; ++seg; }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)cfg.tmem_cols);
  }
}

// hi/lo split of a weight matrix (optionally through a transposed view)
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ src, int64_t ld_r, int64_t ld_c,
                                                        int rows, int cols, float* __restrict__ hi,
                                                        float* __restrict__ lo, int64_t ld_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  const float x = src[(int64_t)r * ld_r + (int64_t)c * ld_c];
  const float h = __uint_as_float(f32_to_tf32_rn(x));
  hi[(int64_t)r * ld_out + c] = h;
  lo[(int64_t)r * ld_out + c] = __uint_as_float(f32_to_tf32_rn(x - h));
}

// every cached weight of the model in ONE launch: block -> job by binary search over the jobs' first blocks
__global__ void __launch_bounds__(256) split_tf32_multi_kernel(const get_split_job* __restrict__ jobs, int n_jobs) {
  int lo_j = 0, hi_j = n_jobs - 1;
  const int64_t b = blockIdx.x;
  while (lo_j < hi_j) {
    const int mid = (lo_j + hi_j + 1) >> 1;
    if (jobs[mid].first_block <= b) lo_j = mid; else hi_j = mid - 1;
  }
  const get_split_job j = jobs[lo_j];
  const int64_t i = (b - j.first_block) * blockDim.x + threadIdx.x;
  if (i >= (int64_t)j.rows * j.cols) return;
  const int r = (int)(i / j.cols), c = (int)(i % j.cols);
  const float x = j.src[(int64_t)r * j.ld_r + (int64_t)c * j.ld_c];
  const float h = __uint_as_float(f32_to_tf32_rn(x));
  j.hi[(int64_t)r * j.ld_out + c] = h;
  j.lo[(int64_t)r * j.ld_out + c] = __uint_as_float(f32_to_tf32_rn(x - h));
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static bool make_b_map(CUtensorMap* map, const float* ptr, int N, int K, int64_t ld, int BN) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static int pad16(int n) { return (n + 15) / 16 * 16; }

// 0 = eligible (cfg filled), 1 = not eligible
static int tc_plan(const get_gemm_desc* d, const GemmParams& p, TcCfg& cfg) {
  if (d->tc_mode != 1 || p.split_k > 1) return 1;
  if (p.M < 1 || p.N < 8) return 1;
  memset(&cfg, 0, sizeof(cfg));
  for (int s = 0; s < p.nseg; ++s) {
    if (p.A[s].trans != 0 || !p.A[s].vec) return 1;
    if (d->B_hi[s] == nullptr || d->B_lo[s] == nullptr) return 1;
    if (!aligned16(d->B_hi[s]) || !aligned16(d->B_lo[s]) || (d->ld_split[s] % 4) != 0 || d->ld_split[s] < p.K[s]) return 1;
    if (p.K[s] < TC_BK) return 1;
    if (s > 0 && p.A[s].rowidx) return 1;
    cfg.kblocks[s] = (p.K[s] + TC_BK - 1) / TC_BK;
    cfg.kblocks_total += cfg.kblocks[s];
  }
  const int m_tiles = (p.M + TC_BM - 1) / TC_BM;
  const int nt_min = (p.N + 255) / 256;
  int best_nt = nt_min;
  if (d->tc_n_tiles > 0) {
    best_nt = d->tc_n_tiles < nt_min ? nt_min : d->tc_n_tiles;
  } else {
    double best_cost = 1e30;
    for (int nt = nt_min; nt <= nt_min + 3; ++nt) {
      const int bn = pad16((p.N + nt - 1) / nt);
      if (bn < 16) break;
      const int64_t ctas = (int64_t)m_tiles * nt;
      const double rounds = (double)((ctas + 147) / 148);
      const double cost = rounds * (bn + 64.0);     // 64 "columns" of fixed per-CTA cost (prologue, A split, epilogue)
      if (cost < best_cost) { best_cost = cost; best_nt = nt; }
    }
  }
  cfg.BN = pad16((p.N + best_nt - 1) / best_nt);
  if (cfg.BN > 256 || cfg.BN < 16) return 1;
  int tc = 32;
  while (tc < cfg.BN) tc <<= 1;
  cfg.tmem_cols = tc;
  const int stage_bytes = 2 * TC_A_TILE_BYTES + 2 * cfg.BN * 128;
  int stages = (225 * 1024 - 2048) / stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > cfg.kblocks_total) stages = cfg.kblocks_total < 2 ? 2 : cfg.kblocks_total;
  if (stages < 2) return 1;
  cfg.stages = stages;
  return 0;
}

int gemm_tc_launch(const get_gemm_desc* d, const GemmParams& p, cudaStream_t st) {
  TcCfg cfg;
  if (tc_plan(d, p, cfg) != 0) return 1;
  TcMaps maps;          // filled per launch; passed by value as a kernel parameter
  memset(&maps, 0, sizeof(maps));
  for (int s = 0; s < p.nseg; ++s) {
    if (!make_b_map(&maps.hi[s], d->B_hi[s], p.N, p.K[s], d->ld_split[s], cfg.BN)) return 1;
    if (!make_b_map(&maps.lo[s], d->B_lo[s], p.N, p.K[s], d->ld_split[s], cfg.BN)) return 1;
  }
  const size_t smem = (size_t)cfg.stages * (2 * TC_A_TILE_BYTES + 2 * cfg.BN * 128) + 1024;
  static int max_dyn = -1;   // opt-in limit (227 KB per CTA) minus this kernel's static shared memory
  if (max_dyn < 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, gemm_tc_kernel);
    if (e == cudaSuccess) {
      const int want = 227 * 1024 - (int)((fa.sharedSizeBytes + 1023) / 1024 * 1024);
      e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
      if (e == cudaSuccess) max_dyn = want;
    }
    if (e != cudaSuccess) {
      set_error("gemm_tc_kernel: cannot opt in to large shared memory: %s", cudaGetErrorString(e));
      (void)cudaGetLastError();
      return -(int)e - 1000;
    }
  }
  if ((int)smem > max_dyn) return 1;
  dim3 grid((unsigned)((p.N + cfg.BN - 1) / cfg.BN), (unsigned)((p.M + TC_BM - 1) / TC_BM), 1);
  if (grid.y > 65535) return 1;
  gemm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(p, cfg, maps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gemm_tc_kernel: launch failed: %s", cudaGetErrorString(e));
    return -(int)e - 1000;
  }
  count_launch();
  return 0;
}

}  // namespace getb

using namespace getb;

int getb::gemm_tc_eligible(const get_gemm_desc* d, const GemmParams& p) {
  TcCfg cfg;
  return tc_plan(d, p, cfg) == 0 ? 1 : 0;
}

extern "C" int get_gemm_f32_uses_tc(const get_gemm_desc* d) {
  GemmParams p;
  if (gemm_build_params(d, p) != 0) return -1;
  if (p.M == 0 || p.N == 0) return 0;
  if (gemm_tc2_plan_splits(d, p) > 0) return 2;
  return gemm_tc_eligible(d, p);
}

extern "C" int get_split_tf32_f32(const float* src, int64_t ld_r, int64_t ld_c, int rows, int cols, float* hi,
                                  float* lo, int64_t ld_out, void* stream) {
  GETB_REQUIRE(src && hi && lo && rows > 0 && cols > 0, "get_split_tf32_f32: bad arguments");
  const int64_t n = (int64_t)rows * cols;
  split_tf32_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(src, ld_r, ld_c, rows, cols, hi, lo, ld_out);
  GETB_CHECK_LAUNCH("get_split_tf32_f32");
  return 0;
}

extern "C" int get_split_tf32_multi_f32(const get_split_job* jobs_dev, int n_jobs, int64_t total_blocks, void* stream) {
  GETB_REQUIRE(jobs_dev && n_jobs > 0 && total_blocks > 0 && total_blocks < 2147483647, "get_split_tf32_multi_f32: bad arguments");
  split_tf32_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs_dev, n_jobs);
  GETB_CHECK_LAUNCH("get_split_tf32_multi_f32");
  return 0;
}
