// Persistent, warp-specialised tcgen05 contraction with fp32-level accuracy (error-compensated 3xTF32) for the dense
// `Linear`s of the GET hot path and their backward passes (reference Models/BiDAF/wrapper.py:191,194-204;
// thirdparty/two_branches_attention.py:140; autograd of both).
//
//   acc = A_hi.B_hi  (main accumulator)  +  A_hi.B_lo + A_lo.B_hi  (separate "small" accumulator, summed in the epilogue
//   so the tensor core's truncating fp32 accumulation only ever sees terms of like magnitude)
//
// Every operand tile is brought in by TMA (cp.async.bulk.tensor, SWIZZLE_128B) as raw fp32:
//   * K-major operands (activations (M,K) row-major; pre-split weights (N,K)): one box {32 k, rows};
//   * MN-major operands (the weight-gradient contraction dW = dG^T . X, both operands (K,MN) row-major): boxes
//     {32 mn, 32 k} (TMA swizzle 128B_ATOM_32B), consumed through MN-major UMMA descriptors (SWIZZLE_128B_BASE32B, the
//     only MN-major layout fp32 operands have) -- no transposition pass anywhere.
// Warp roles (448 threads, one persistent CTA per SM, static round-robin over work items):
//   warp 0      TMA producer                         warp 1      tcgen05.mma issuer, owns the TMEM allocation
//   warps 2-5   splitter: raw tile -> hi (tf32 round-to-nearest, in place) + lo (second buffer), smem -> smem
//   warps 6-13  epilogue: tcgen05.ld main+small -> fused epilogue (gemm_common.cuh) -> global; with two TMEM
//               accumulator sets the epilogue of item i overlaps the main loop of item i+1
// Work item = (m tile, n tile, k split); split-K items store raw partial tiles that gemm_splitk_reduce_kernel sums in
// a fixed order (deterministic weight gradients).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "gemm_common.cuh"

namespace getb {

constexpr int T2_BM = 128;
constexpr int T2_BK = 32;
constexpr int T2_THREADS = 448;                      // 14 warps: TMA, MMA, 4 splitter, 8 epilogue
constexpr int T2_EPI_WARPS = 8;
constexpr int T2_A_TILE = T2_BM * 128;   // bytes of one A tile (hi or lo)
constexpr int T2_MAX_STAGES = 6;
constexpr uint32_t T2_SPIN_LIMIT = 1u << 28;
constexpr int T2_STG_LD = 36;                        // floats per row of the epilogue staging tile (32 + 4: conflict-free)
constexpr int T2_STG_BYTES = T2_EPI_WARPS * 32 * T2_STG_LD * 4;

struct Tc2Maps {
  CUtensorMap a[GET_GEMM_MAX_SEG];
  CUtensorMap bh[GET_GEMM_MAX_SEG];
  CUtensorMap bl[GET_GEMM_MAX_SEG];
};

struct Tc2Cfg {
  int BN, stages, acc_bufs, tmem_cols;
  int kblocks[GET_GEMM_MAX_SEG];
  int kblocks_total;
  int a_mn, b_mn;        // operand is MN-major (stored (K, MN) row-major)
  int split_b;           // B arrives raw and is split in the kernel (otherwise B_hi / B_lo are loaded pre-split)
  int ntm, ntn, splits, kb_per_split, items;
  uint32_t b_tile, stage_bytes;
  int a_tmem;            // A hi/lo tiles live in TMEM (tcgen05.st by the splitter, TS-mode MMAs): K-major A only
  int tstages;           // depth of the TMEM ring of A tiles (64 columns each)
  int a_tmem_col;        // first TMEM column of that ring
  uint32_t a_stage;      // bytes of the A part of one shared-memory stage (raw tile, plus the lo tile in SS mode)
  int single;            // reduced-precision mode (desc->tc_mode == 2): ONE tf32 MMA per k step on the raw operands, no hi/lo
                         // split, no small-term accumulator (relative error ~1e-3: for the 1e-2 parity class only)
  int debug;             // GET_B200_T2_DEBUG=9 with -DGETB_T2_TIMELINE: print the per-role timeline of CTA 0
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
namespace t2 {
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
    if (++spins > T2_SPIN_LIMIT) __trap();   // a protocol bug must fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory operand descriptor.
// K-major, SWIZZLE_128B (layout type 2): rows at a 128-byte pitch, 8-row groups `sbo` = 1024 B apart, LBO unused.
// MN-major fp32/tf32 operands only exist as SWIZZLE_128B_BASE32B (layout type 1; the TMA mode is 128B_ATOM_32B):
// 128-byte rows of 32 MN elements, 4-k groups `sbo` = 512 B apart, 32-element MN groups `lbo` bytes apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
__device__ __forceinline__ uint32_t f32_to_tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// explicit shared-space accesses (the 1024-byte realignment of the dynamic buffer hides the address space from nvcc)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// hi = tf32 round-to-nearest of v, lo = tf32 round-to-nearest of the remainder
__device__ __forceinline__ void split4(const float4& v, uint4& hi, uint4& lo) {
  hi.x = f32_to_tf32_rn(v.x); hi.y = f32_to_tf32_rn(v.y); hi.z = f32_to_tf32_rn(v.z); hi.w = f32_to_tf32_rn(v.w);
  lo.x = f32_to_tf32_rn(v.x - __uint_as_float(hi.x)); lo.y = f32_to_tf32_rn(v.y - __uint_as_float(hi.y));
  lo.z = f32_to_tf32_rn(v.z - __uint_as_float(hi.z)); lo.w = f32_to_tf32_rn(v.w - __uint_as_float(hi.w));
}
}  // namespace t2

// locate k-block `kb` (global index over the segments): segment and k offset inside it
__device__ __forceinline__ void t2_locate(const Tc2Cfg& cfg, int nseg, int kb, int& seg, int& kin) {
  seg = 0;
  while (seg + 1 < nseg && kb >= cfg.kblocks[seg]) { kb -= cfg.kblocks[seg]; ++seg; }
  kin = kb * T2_BK;
}

template <int EPI>
__global__ void __launch_bounds__(T2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ Tc2Cfg cfg, const __grid_constant__ Tc2Maps maps) {
  using namespace t2;
  extern __shared__ __align__(1024) uint8_t smem_raw2[];
  __shared__ __align__(8) uint64_t bar_raw[T2_MAX_STAGES];     // TMA landed (tx bytes)
  __shared__ __align__(8) uint64_t bar_ready[T2_MAX_STAGES];   // hi/lo tiles complete (128 splitter arrivals)
  __shared__ __align__(8) uint64_t bar_empty[T2_MAX_STAGES];   // MMAs that read the stage retired (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_tfree[T2_MAX_STAGES];   // TS mode: MMAs that read a TMEM A stage retired
  __shared__ __align__(8) uint64_t bar_accf[2];                // accumulator set complete (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_acce[2];                // accumulator set drained (all epilogue threads arrive)
  __shared__ uint32_t tmem_holder;
#ifdef GETB_T2_TIMELINE   // compile with -DGETB_T2_TIMELINE and run with GET_B200_T2_DEBUG=9: per-role timeline of CTA 0
  __shared__ long long dbg_ts[4][24];
  const long long dbg_t0 = clock64();
#define T2_DBG(role, idx) do { if (cfg.debug == 9 && blockIdx.x == 0 && (idx) < 24) dbg_ts[role][idx] = clock64() - dbg_t0; } while (0)
#else
#define T2_DBG(role, idx) do { } while (0)
#endif

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = cfg.BN;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 0) {
    for (int s = 0; s < T2_MAX_STAGES; ++s) {
      mbar_init(&bar_raw[s], 1);
      mbar_init(&bar_ready[s], 128);
      mbar_init(&bar_empty[s], 1);
      mbar_init(&bar_tfree[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_accf[b], 1);
      mbar_init(&bar_acce[b], 32 * T2_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)cfg.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  const int n_my = ((int)blockIdx.x < cfg.items) ? (cfg.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = T2_A_TILE;
      const uint32_t tx = a_bytes + cfg.b_tile * ((cfg.split_b || cfg.single) ? 1u : 2u);
      for (int it = 0; it < n_my; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int nt = item % cfg.ntn, mt = (item / cfg.ntn) % cfg.ntm, z = item / (cfg.ntn * cfg.ntm);
        const int m0 = mt * T2_BM, n0 = nt * BN;
        const int kb0 = z * cfg.kb_per_split, kb1 = min(cfg.kblocks_total, kb0 + cfg.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          int seg, kin;
          t2_locate(cfg, p.nseg, kb, seg, kin);
          mbar_wait(&bar_empty[stage], phase ^ 1);
          if (kb == kb0) T2_DBG(0, it * 2); if (kb == kb1 - 1) T2_DBG(0, it * 2 + 1);
          if (it == 0 && kb - kb0 >= 4 && kb - kb0 < 7) T2_DBG(0, 8 + (kb - kb0 - 4));
          const uint32_t sa = smem_base + (uint32_t)stage * cfg.stage_bytes;
          const uint32_t sbh = sa + cfg.a_stage, sbl = sbh + cfg.b_tile;
          mbar_arrive_expect_tx(&bar_raw[stage], tx);
          if (cfg.a_mn) {
#pragma unroll
            for (int a = 0; a < T2_BM / 32; ++a) tma_load_2d(sa + a * 4096u, &maps.a[seg], m0 + a * 32, kin, &bar_raw[stage]);
          } else {
            tma_load_2d(sa, &maps.a[seg], kin, m0, &bar_raw[stage]);
          }
          if (cfg.b_mn) {
            for (int a = 0; a < BN / 32; ++a) tma_load_2d(sbh + a * 4096u, &maps.bh[seg], n0 + a * 32, kin, &bar_raw[stage]);
          } else {
            tma_load_2d(sbh, &maps.bh[seg], kin, n0, &bar_raw[stage]);
            if (!cfg.split_b && !cfg.single) tma_load_2d(sbl, &maps.bl[seg], kin, n0, &bar_raw[stage]);
          }
          if (++stage == cfg.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================================== MMA issuer =============================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(cfg.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(cfg.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(T2_BM >> 4) << 24);
      const uint32_t a_step = cfg.a_mn ? 1024u : 32u;     // bytes per 8-k step
      const uint32_t b_step = cfg.b_mn ? 1024u : 32u;
      const uint32_t a_lbo = cfg.a_mn ? 4096u : 16u, b_lbo = cfg.b_mn ? 4096u : 16u;
      const uint32_t a_sbo = cfg.a_mn ? 512u : 1024u, b_sbo = cfg.b_mn ? 512u : 1024u;
      const uint32_t a_lt = cfg.a_mn ? 1u : 2u, b_lt = cfg.b_mn ? 1u : 2u;
      int stage = 0, acc = 0, tstage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int it = 0; it < n_my; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int z = item / (cfg.ntn * cfg.ntm);
        const int kb0 = z * cfg.kb_per_split, kb1 = min(cfg.kblocks_total, kb0 + cfg.kb_per_split);
        mbar_wait(&bar_acce[acc], acc_phase ^ 1);
        tc_fence_after();
        T2_DBG(1, it * 3);
        const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * BN);
        const uint32_t d_small = d_main + (uint32_t)BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bar_raw[stage], phase);
          if (!cfg.single) mbar_wait(&bar_ready[stage], phase);
          tc_fence_after();
          if (kb == kb0) T2_DBG(1, it * 3 + 1);
          if (it == 0 && kb - kb0 >= 4 && kb - kb0 < 7) T2_DBG(1, 12 + (kb - kb0 - 4) * 2);
          const uint32_t a_hi = smem_base + (uint32_t)stage * cfg.stage_bytes;
          const uint32_t a_lo = a_hi + T2_A_TILE;
          const uint32_t b_hi = a_hi + cfg.a_stage;
          const uint32_t b_lo = b_hi + cfg.b_tile;
          if (cfg.single) {
#pragma unroll
            for (int ks = 0; ks < T2_BK / 8; ++ks)
              umma_tf32(d_main, smem_desc(a_hi + ks * a_step, a_lbo, a_sbo, a_lt), smem_desc(b_hi + ks * b_step, b_lbo, b_sbo, b_lt),
                        idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
          } else if (cfg.a_tmem) {
            const uint32_t ta_hi = tmem_base + (uint32_t)(cfg.a_tmem_col + tstage * 64);
            const uint32_t ta_lo = ta_hi + 32u;
#pragma unroll
            for (int ks = 0; ks < T2_BK / 8; ++ks) {
              const uint64_t dbh = smem_desc(b_hi + ks * b_step, b_lbo, b_sbo, b_lt), dbl = smem_desc(b_lo + ks * b_step, b_lbo, b_sbo, b_lt);
              const uint32_t first = (kb > kb0 || ks > 0) ? 1u : 0u;
              umma_tf32_ts(d_main, ta_hi + ks * 8u, dbh, idesc, first);
              umma_tf32_ts(d_small, ta_hi + ks * 8u, dbl, idesc, first);
              umma_tf32_ts(d_small, ta_lo + ks * 8u, dbh, idesc, 1u);
            }
            umma_commit(&bar_tfree[tstage]);
            if (++tstage == cfg.tstages) tstage = 0;
          } else {
#pragma unroll
          for (int ks = 0; ks < T2_BK / 8; ++ks) {
            const uint64_t dah = smem_desc(a_hi + ks * a_step, a_lbo, a_sbo, a_lt), dal = smem_desc(a_lo + ks * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t dbh = smem_desc(b_hi + ks * b_step, b_lbo, b_sbo, b_lt), dbl = smem_desc(b_lo + ks * b_step, b_lbo, b_sbo, b_lt);
            const uint32_t first = (kb > kb0 || ks > 0) ? 1u : 0u;
            umma_tf32(d_main, dah, dbh, idesc, first);
            umma_tf32(d_small, dah, dbl, idesc, first);
            umma_tf32(d_small, dal, dbh, idesc, 1u);
          }
          }
          umma_commit(&bar_empty[stage]);
          if (it == 0 && kb - kb0 >= 4 && kb - kb0 < 7) T2_DBG(1, 13 + (kb - kb0 - 4) * 2);
          if (++stage == cfg.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bar_accf[acc]);
        T2_DBG(1, it * 3 + 2);
        if (cfg.acc_bufs == 2) {
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        } else {
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // =========================================== splitter ===============================================
    const int t = tid - 64;   // 0..127
    int stage = 0, tstage = 0;
    uint32_t phase = 0, tphase = 0;
    if (cfg.single) goto t2_role_done;                                // raw operands go straight to the tensor core
    {
    constexpr int a_chunks = T2_A_TILE / 16 / 128;                   // 16-byte chunks per thread in the A tile
    const int b_chunks = cfg.split_b ? (int)(cfg.b_tile / 16) : 0;   // total chunks of the B tile
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int z = item / (cfg.ntn * cfg.ntm);
      const int kb0 = z * cfg.kb_per_split, kb1 = min(cfg.kblocks_total, kb0 + cfg.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&bar_raw[stage], phase);
        if (t == 0 && kb == kb0) T2_DBG(3, it * 2); if (t == 0 && kb == kb1 - 1) T2_DBG(3, it * 2 + 1);
        if (t == 0 && it == 0 && kb - kb0 >= 4 && kb - kb0 < 7) T2_DBG(3, 8 + (kb - kb0 - 4) * 4);
        const uint32_t a_hi = smem_base + (uint32_t)stage * cfg.stage_bytes;
        const uint32_t a_lo = a_hi + T2_A_TILE;
        if (cfg.a_tmem) {
          // this thread owns row `trow` of the tile: 128 bytes of the K-major SW128 tile -> hi / lo -> its TMEM lane
          const int trow = (warp & 3) * 32 + lane;
          const uint32_t rbase = a_hi + (uint32_t)trow * 128u;
          const uint32_t sw = (uint32_t)(trow & 7);
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cfg.a_tmem_col + tstage * 64);
          mbar_wait(&bar_tfree[tstage], tphase ^ 1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = lds128(rbase + (((uint32_t)(h * 4 + c) ^ sw) << 4));
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 h4, l4;
              split4(v[c], h4, l4);
              hi[c * 4 + 0] = h4.x; hi[c * 4 + 1] = h4.y; hi[c * 4 + 2] = h4.z; hi[c * 4 + 3] = h4.w;
              lo[c * 4 + 0] = l4.x; lo[c * 4 + 1] = l4.y; lo[c * 4 + 2] = l4.z; lo[c * 4 + 3] = l4.w;
            }
            tmem_st16(ta + (uint32_t)(h * 16), hi);
            tmem_st16(ta + 32u + (uint32_t)(h * 16), lo);
          }
          tmem_st_wait();
          tc_fence_before();
          if (++tstage == cfg.tstages) { tstage = 0; tphase ^= 1; }
        } else {
          float4 v[a_chunks];
#pragma unroll
          for (int i = 0; i < a_chunks; ++i) v[i] = lds128(a_hi + (uint32_t)(t + i * 128) * 16u);
#pragma unroll
          for (int i = 0; i < a_chunks; ++i) {
            const uint32_t off = (uint32_t)(t + i * 128) * 16u;
            uint4 hi, lo;
            split4(v[i], hi, lo);
            sts128(a_hi + off, hi);
            sts128(a_lo + off, lo);
          }
        }
        if (cfg.split_b) {
          const uint32_t b_hi = a_hi + cfg.a_stage;
          const uint32_t b_lo = b_hi + cfg.b_tile;
          for (int c = t; c < b_chunks; c += 128) {
            const uint32_t off = (uint32_t)c * 16u;
            const float4 v = lds128(b_hi + off);
            uint4 hi, lo;
            split4(v, hi, lo);
            sts128(b_hi + off, hi);
            sts128(b_lo + off, lo);
          }
        }
        if (t == 0 && it == 0 && kb - kb0 >= 4 && kb - kb0 < 7) T2_DBG(3, 9 + (kb - kb0 - 4) * 4);
        fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
        if (t == 0 && it == 0 && kb - kb0 >= 4 && kb - kb0 < 7) T2_DBG(3, 10 + (kb - kb0 - 4) * 4);
        mbar_arrive(&bar_ready[stage]);
        if (++stage == cfg.stages) { stage = 0; phase ^= 1; }
      }
    }
    }
  } else {
    // =========================================== epilogue ===============================================
    // TMEM hands every thread one output ROW; a per-warp 32 x 32 staging tile in shared memory turns that into 8 lanes
    // per row x 4 rows per instruction, so every global access of the fused epilogue is a full 128-byte line.
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int chalf = (warp - 6) >> 2;            // two warps per quarter: even / odd 32-column chunks
    const uint32_t stg = smem_base + (uint32_t)cfg.stages * cfg.stage_bytes + (uint32_t)(warp - 6) * (32u * T2_STG_LD * 4u);
    const int rrow = lane >> 3, rq = lane & 7;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int nt = item % cfg.ntn, mt = (item / cfg.ntn) % cfg.ntm, z = item / (cfg.ntn * cfg.ntm);
      const int m_base = mt * T2_BM + quarter * 32;
      const int n0 = nt * BN;
      mbar_wait(&bar_accf[acc], acc_phase);
      tc_fence_after();
      if (warp == 6 && lane == 0) T2_DBG(2, it * 2);
      const uint32_t t_main = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * BN);
      const uint32_t t_small = t_main + (uint32_t)BN;
      const int nch = (BN + 31) / 32;     // 32-column chunks; this warp takes chunks chalf, chalf + 2, ...
      const int last_col = (nch - 1 >= chalf) ? ((nch - 1 - chalf) / 2 * 2 + chalf) * 32 : -1;
      if (last_col < 0) {                 // single-chunk tiles: the odd warps have nothing to read
        tc_fence_before();
        mbar_arrive(&bar_acce[acc]);
      }
      for (int col = chalf * 32; col < BN; col += 64) {
        const bool two = col + 16 < BN;
        uint32_t rm[16], rs[16], rm2[16], rs2[16];
        tmem_ld16_nowait(t_main + (uint32_t)col, rm);
        if (!cfg.single) tmem_ld16_nowait(t_small + (uint32_t)col, rs);
        if (two) {
          tmem_ld16_nowait(t_main + (uint32_t)col + 16u, rm2);
          if (!cfg.single) tmem_ld16_nowait(t_small + (uint32_t)col + 16u, rs2);
        }
        tmem_ld_wait();
        if (cfg.single) {
#pragma unroll
          for (int e = 0; e < 16; ++e) { rs[e] = 0u; rs2[e] = 0u; }
        }
        if (col == last_col) {           // last chunk of this accumulator set for this warp: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&bar_acce[acc]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = __float_as_uint(__uint_as_float(rm[q * 4 + 0]) + __uint_as_float(rs[q * 4 + 0]));
          v.y = __float_as_uint(__uint_as_float(rm[q * 4 + 1]) + __uint_as_float(rs[q * 4 + 1]));
          v.z = __float_as_uint(__uint_as_float(rm[q * 4 + 2]) + __uint_as_float(rs[q * 4 + 2]));
          v.w = __float_as_uint(__uint_as_float(rm[q * 4 + 3]) + __uint_as_float(rs[q * 4 + 3]));
          sts128(stg + (uint32_t)(lane * T2_STG_LD + q * 4) * 4u, v);
        }
        if (two) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = __float_as_uint(__uint_as_float(rm2[q * 4 + 0]) + __uint_as_float(rs2[q * 4 + 0]));
            v.y = __float_as_uint(__uint_as_float(rm2[q * 4 + 1]) + __uint_as_float(rs2[q * 4 + 1]));
            v.z = __float_as_uint(__uint_as_float(rm2[q * 4 + 2]) + __uint_as_float(rs2[q * 4 + 2]));
            v.w = __float_as_uint(__uint_as_float(rm2[q * 4 + 3]) + __uint_as_float(rs2[q * 4 + 3]));
            sts128(stg + (uint32_t)(lane * T2_STG_LD + 16 + q * 4) * 4u, v);
          }
        }
        __syncwarp();
        const int n = n0 + col + rq * 4;
        if ((two || rq < 4) && n < p.N) {
          if (cfg.splits > 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + rrow;
              const int m = m_base + row;
              if (m < p.M) {
                const float4 f = lds128(stg + (uint32_t)(row * T2_STG_LD + rq * 4) * 4u);
                const float v[4] = {f.x, f.y, f.z, f.w};
                store4(p.workspace + ((int64_t)z * p.M + m) * p.N + n, true, 4, v);
              }
            }
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {          // two batches of four rows: all loads first, then math + stores
              EpiIn in[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int m = m_base + (h * 4 + i) * 4 + rrow;
                if (m < p.M) epilogue4_load_t<EPI, true>(p, m, n, in[i]);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int row = (h * 4 + i) * 4 + rrow;
                const int m = m_base + row;
                if (m < p.M) {
                  const float4 f = lds128(stg + (uint32_t)(row * T2_STG_LD + rq * 4) * 4u);
                  const float v[4] = {f.x, f.y, f.z, f.w};
                  epilogue4_apply_t<EPI, true>(p, m, n, v, in[i]);
                }
              }
            }
          }
        }
        __syncwarp();
      }
      if (warp == 6 && lane == 0) T2_DBG(2, it * 2 + 1);
      if (cfg.acc_bufs == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
  }
t2_role_done:
  tc_fence_before();
  __syncthreads();
#ifdef GETB_T2_TIMELINE
  if (cfg.debug == 9 && blockIdx.x == 0 && tid == 0) {
    for (int it = 0; it < n_my && it < 4; ++it)
      printf("T2DBG it=%d prod[%lld %lld] split[%lld %lld] mma[acce %lld ready %lld commit %lld] epi[%lld %lld]\n", it,
             dbg_ts[0][it * 2], dbg_ts[0][it * 2 + 1], dbg_ts[3][it * 2], dbg_ts[3][it * 2 + 1], dbg_ts[1][it * 3], dbg_ts[1][it * 3 + 1],
             dbg_ts[1][it * 3 + 2], dbg_ts[2][it * 2], dbg_ts[2][it * 2 + 1]);
    for (int j = 0; j < 3; ++j)
      printf("T2DBG kb=%d tma_issue %lld | split: raw_seen %lld stored %lld fenced %lld | mma: ready_seen %lld issued+commit %lld\n", 4 + j,
             dbg_ts[0][8 + j], dbg_ts[3][8 + j * 4], dbg_ts[3][9 + j * 4], dbg_ts[3][10 + j * 4], dbg_ts[1][12 + j * 2], dbg_ts[1][13 + j * 2]);
    printf("T2DBG end %lld\n", clock64() - dbg_t0);
  }
#endif
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg.tmem_cols) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 t2_encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor map: `inner` contiguous elements per row, `outer` rows `ld` elements apart; box {box_inner, box_outer}
struct T2MapKey {
  const float* ptr;
  int64_t inner, outer, ld;
  int bi, bo, mn;
  bool operator==(const T2MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && bi == o.bi && bo == o.bo && mn == o.mn;
  }
};
struct T2MapKeyHash {
  size_t operator()(const T2MapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)k.inner * 0xC2B2AE3D27D4EB4Full + (uint64_t)k.outer * 0x165667B19E3779F9ull + (uint64_t)k.ld * 31 +
         (uint64_t)k.bi * 7 + (uint64_t)k.bo + (uint64_t)k.mn * 1315423911ull;
    return (size_t)(h ^ (h >> 29));
  }
};

static bool t2_encode(CUtensorMap* map, const float* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                      int box_outer, int mn);

// Tensor maps depend only on (address, extents, stride, box): encoding is memoised because the allocator hands the same
// activation addresses back step after step.
static bool t2_make_map(CUtensorMap* map, const float* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                        int box_outer, int mn) {
  static std::mutex mu;
  static std::unordered_map<T2MapKey, CUtensorMap, T2MapKeyHash> cache;
  const T2MapKey key{ptr, inner, outer, ld, box_inner, box_outer, mn};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *map = it->second;
    return true;
  }
  if (!t2_encode(map, ptr, inner, outer, ld, box_inner, box_outer, mn)) return false;
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *map);
  return true;
}

static bool t2_encode(CUtensorMap* map, const float* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                      int box_outer, int mn) {
  EncodeTiledFn2 enc = t2_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int t2_pad(int n, int q) { return (n + q - 1) / q * q; }

// 0 = eligible (cfg filled), 1 = not eligible
static int tc2_plan(const get_gemm_desc* d, const GemmParams& p, Tc2Cfg& cfg) {
  if (d->tc_mode < 1) return 1;
  if (p.M < 1 || p.N < 8) return 1;
  const int single = d->tc_mode == 2 ? 1 : 0;
  if (p.drop_thr) return 1;                                   // A-operand dropout: register-staged kernel (gemm_tc.cu)
  if (!p.vec_epi || (p.N & 3)) return 1;                      // the epilogue works on aligned float4 quads only
  memset(&cfg, 0, sizeof(cfg));
  const bool presplit = d->B_hi[0] != nullptr;
  cfg.a_mn = p.A[0].trans;
  cfg.b_mn = presplit ? 0 : p.B[0].trans;
  cfg.split_b = presplit ? 0 : 1;
  cfg.single = single;
  for (int s = 0; s < p.nseg; ++s) {
    if (p.A[s].rowidx) return 1;                              // gathered A: register-staged kernel
    if (p.A[s].trans != cfg.a_mn || !aligned16(p.A[s].ptr) || (p.A[s].ld % 4) != 0) return 1;
    if (presplit) {
      if (!d->B_hi[s] || !d->B_lo[s] || !aligned16(d->B_hi[s]) || !aligned16(d->B_lo[s]) || (d->ld_split[s] % 4) != 0 ||
          d->ld_split[s] < p.K[s])
        return 1;
    } else {
      if (d->B_hi[s] || p.B[s].trans != cfg.b_mn || !aligned16(p.B[s].ptr) || (p.B[s].ld % 4) != 0) return 1;
    }
    if (p.K[s] < 8) return 1;
    cfg.kblocks[s] = (p.K[s] + T2_BK - 1) / T2_BK;
    cfg.kblocks_total += cfg.kblocks[s];
  }
  cfg.ntm = (p.M + T2_BM - 1) / T2_BM;
  const int q = cfg.b_mn ? 32 : 16;                           // MN-major B tiles are built from 32-wide boxes
  // split-K (host decides through desc->split_k; p.split_k was clamped against 16-wide SIMT k tiles, redo it here)
  int splits = d->split_k > 1 ? d->split_k : 1;
  if (splits > cfg.kblocks_total) splits = cfg.kblocks_total;
  cfg.kb_per_split = (cfg.kblocks_total + splits - 1) / splits;
  cfg.splits = (cfg.kblocks_total + cfg.kb_per_split - 1) / cfg.kb_per_split;
  if (cfg.splits > 1 && !d->workspace) return 1;
  // Optional: A hi/lo tiles in TMEM (TS-mode MMAs), K-major A with pre-split B only.
  // Measured on B200 (profiles/): a TS-mode tf32 MMA costs ~128 cycles whatever N is, an SS-mode one 128*N/256, so at
  // the N <= 160 tiles of this model SS mode wins; TS mode stays available for experiments (GET_B200_T2_TS=1).
  static int want_ts = -1;
  if (want_ts < 0) {
    const char* e = getenv("GET_B200_T2_TS");
    want_ts = e ? atoi(e) : 0;
  }
  cfg.a_tmem = (want_ts && !cfg.a_mn && !cfg.split_b) ? 1 : 0;
  // n tiling. TMEM budget (512 columns): accumulator sets (main + small = 2*BN columns each; two sets when the CTA runs
  // several items so that the epilogue overlaps the next main loop) + in TS mode >= 2 A stages of 64 columns.
  int best_nt = 0;
  double best_cost = 1e30;
  for (int pass = 0; pass < 2 && best_nt == 0; ++pass) {
    if (pass == 1) cfg.a_tmem = 0;                          // no TS tiling fits: fall back to SS mode
    for (int nt = 1; nt <= 8; ++nt) {
      const int bn = t2_pad((p.N + nt - 1) / nt, q);
      if (bn > 256 || bn < 16) continue;
      if ((nt - 1) * bn >= p.N) continue;
      const int64_t items = (int64_t)cfg.ntm * nt * cfg.splits;
      const int64_t rounds = (items + 147) / 148;
      const int bufs = rounds > 1 ? 2 : 1;
      const int cols = bufs * 2 * bn + (cfg.a_tmem ? 2 * 64 : 0);
      if (cols > 512) continue;
      const double cost = (double)rounds * (bn + 48.0);
      if (cost < best_cost) { best_cost = cost; best_nt = nt; }
    }
  }
  if (best_nt == 0) return 1;
  if (d->tc_n_tiles > 0) {
    const int bn = t2_pad((p.N + d->tc_n_tiles - 1) / d->tc_n_tiles, q);
    const int64_t rounds = ((int64_t)cfg.ntm * d->tc_n_tiles * cfg.splits + 147) / 148;
    const int cols = (rounds > 1 ? 2 : 1) * 2 * bn + (cfg.a_tmem ? 2 * 64 : 0);
    if (bn <= 256 && bn >= 16 && (d->tc_n_tiles - 1) * bn < p.N && cols <= 512) best_nt = d->tc_n_tiles;
  }
  cfg.ntn = best_nt;
  cfg.BN = t2_pad((p.N + best_nt - 1) / best_nt, q);
  cfg.items = cfg.ntm * cfg.ntn * cfg.splits;
  cfg.acc_bufs = (cfg.items > 148) ? 2 : 1;
  const int acc_cols = cfg.acc_bufs * 2 * cfg.BN;
  if (cfg.a_tmem) {
    cfg.tstages = (512 - acc_cols) / 64;
    if (cfg.tstages > T2_MAX_STAGES) cfg.tstages = T2_MAX_STAGES;
    if (cfg.tstages < 2) return 1;
    cfg.a_tmem_col = acc_cols;
    cfg.tmem_cols = 512;
  } else {
    int tc = 32;
    while (tc < acc_cols) tc <<= 1;
    if (tc > 512) return 1;
    cfg.tmem_cols = tc;
  }
  cfg.b_tile = (uint32_t)cfg.BN * 128u;
  cfg.a_stage = (cfg.a_tmem || cfg.single) ? (uint32_t)T2_A_TILE : 2u * T2_A_TILE;
  cfg.stage_bytes = cfg.a_stage + (cfg.single ? 1u : 2u) * cfg.b_tile;
  int stages = (224 * 1024 - 2048 - T2_STG_BYTES) / (int)cfg.stage_bytes;
  if (stages > T2_MAX_STAGES) stages = T2_MAX_STAGES;
  if (stages < 2) return 1;
  cfg.stages = stages;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("GET_B200_T2_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  cfg.debug = dbg;
  return 0;
}

int gemm_tc2_launch(const get_gemm_desc* d, GemmParams& p, cudaStream_t st) {
  Tc2Cfg cfg;
  if (tc2_plan(d, p, cfg) != 0) return 1;
  p.split_k = cfg.splits;
  Tc2Maps maps;
  memset(&maps, 0, sizeof(maps));
  for (int s = 0; s < p.nseg; ++s) {
    bool ok;
    if (cfg.a_mn) ok = t2_make_map(&maps.a[s], p.A[s].ptr, p.M, p.K[s], p.A[s].ld, 32, T2_BK, 1);
    else ok = t2_make_map(&maps.a[s], p.A[s].ptr, p.K[s], p.M, p.A[s].ld, T2_BK, T2_BM, 0);
    if (!ok) return 1;
    if (!cfg.split_b) {
      if (!t2_make_map(&maps.bh[s], d->B_hi[s], p.K[s], p.N, d->ld_split[s], T2_BK, cfg.BN, 0)) return 1;
      if (!t2_make_map(&maps.bl[s], d->B_lo[s], p.K[s], p.N, d->ld_split[s], T2_BK, cfg.BN, 0)) return 1;
    } else if (cfg.b_mn) {
      if (!t2_make_map(&maps.bh[s], p.B[s].ptr, p.N, p.K[s], p.B[s].ld, 32, T2_BK, 1)) return 1;
    } else {
      if (!t2_make_map(&maps.bh[s], p.B[s].ptr, p.K[s], p.N, p.B[s].ld, T2_BK, cfg.BN, 0)) return 1;
    }
  }
  const size_t smem = (size_t)cfg.stages * cfg.stage_bytes + T2_STG_BYTES + 1024;
  typedef void (*KernelFn)(const GemmParams, const Tc2Cfg, const Tc2Maps);
  static const KernelFn kernels[7] = {gemm_tc2_kernel<0>, gemm_tc2_kernel<1>, gemm_tc2_kernel<2>, gemm_tc2_kernel<3>,
                                      gemm_tc2_kernel<4>, gemm_tc2_kernel<5>, gemm_tc2_kernel<6>};
  const int epi = cfg.splits > 1 ? 0 : p.epilogue;
  if (epi < 0 || epi > 6) return 1;
  KernelFn fn = kernels[epi];
  static int max_dyn[7] = {-1, -1, -1, -1, -1, -1, -1};
  if (max_dyn[epi] < 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, fn);
    if (e == cudaSuccess) {
      const int want = 227 * 1024 - (int)((fa.sharedSizeBytes + 1023) / 1024 * 1024);
      e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
      if (e == cudaSuccess) max_dyn[epi] = want;
    }
    if (e != cudaSuccess) {
      set_error("gemm_tc2_kernel: cannot opt in to large shared memory: %s", cudaGetErrorString(e));
      (void)cudaGetLastError();
      return -(int)e - 1000;
    }
  }
  if ((int)smem > max_dyn[epi]) return 1;
  const int grid = cfg.items < 148 ? cfg.items : 148;
  fn<<<grid, T2_THREADS, smem, st>>>(p, cfg, maps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gemm_tc2_kernel: launch failed: %s", cudaGetErrorString(e));
    return -(int)e - 1000;
  }
  count_launch();
  return cfg.splits > 1 ? 2 : 0;   // 2: caller must run the split-K reduction with p.split_k = cfg.splits
}

int gemm_tc2_plan_splits(const get_gemm_desc* d, const GemmParams& p) {
  Tc2Cfg cfg;
  if (tc2_plan(d, p, cfg) != 0) return -1;
  return cfg.splits;
}

}  // namespace getb
