// Persistent, warp-specialised tcgen05 contraction on bf16 PLANES (see include/get_b200.h, get_gemm_bp): the dense
// `Linear`s of the GET hot path and their backward passes (reference Models/BiDAF/wrapper.py:191,194-204;
// thirdparty/two_branches_attention.py:140; autograd of both).
//
// Every fp32 operand arrives pre-split into bf16 planes (p0 + p1 + p2 = v) written by the kernel that produced it, so
// the main loop is nothing but TMA -> tcgen05.mma.kind::f16 -> TMEM:
//   mode 1:  a0.b0                                   plain bf16
//   mode 2:  a0.b0 + a0.b1 + a1.b0                   16-bit operands, one accumulator
//   mode 3:  a0.b0 | a0.b1 + a1.b0 + a1.b1 + a0.b2 + a2.b0    fp32-exact class: the small terms have their own TMEM
//            accumulator (the tensor core's fp32 accumulation truncates; like magnitudes stay together)
// Operand tiles come in by ONE 3-D TMA box per operand and stage (all planes at once):
//   * K-major (activations (M,K) row-major; packed weights (N,K)): box {kb k, rows, planes}, SWIZZLE_64B (kb = 32) or
//     SWIZZLE_128B (kb = 64);
//   * MN-major (weight gradients dW = dG^T X, both operands stored (K, MN) row-major): boxes {64 mn, kb k, planes},
//     SWIZZLE_128B, consumed through MN-major UMMA descriptors -- no transposition pass anywhere.
// Warp roles (320 threads, one persistent CTA per SM, static round-robin over work items):
//   warp 0   TMA producer          warp 1   tcgen05.mma issuer, owns the TMEM allocation
//   warps 2-9  epilogue: tcgen05.ld -> per-warp 32x32 staging tile -> 8 lanes per row, so every global access of the
//              fused epilogue is a full 128-byte line; two TMEM accumulator sets overlap the epilogue of item i with
//              the main loop of item i+1. The epilogue also writes the bf16 planes of its output for the next GEMM.
// Work item = (m tile, n tile, k split); split-K items store raw partial tiles (get_bp_splitk_reduce sums them in a
// fixed order: deterministic weight gradients).
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tcgen05.cuh"

namespace getb {

constexpr int BP_BM = 128;
constexpr int BP_THREADS = 320;
constexpr int BP_EPI_WARPS = 8;
constexpr int BP_MAX_STAGES = 8;
// epilogue staging: per warp two sets (used alternately: the TMA stores drain asynchronously) of fp32 tiles of 2 KB
// (C, out1) and bf16 plane tiles of 1 KB; the set size depends on the outputs of the launch (cfg.stg_set, <= 7 KB)
constexpr int BP_STG_SET_MAX = 7168;
constexpr int BP_SMS = 148;

struct BpMaps {
  CUtensorMap a[GET_GEMM_MAX_SEG];
  CUtensorMap b[GET_GEMM_MAX_SEG];
  CUtensorMap c;       // fp32 output (or the split-K workspace), box {16 cols, 32 rows[, 1]}
  CUtensorMap o1;      // second fp32 output
  CUtensorMap pl;      // bf16 output planes {cols, rows, planes}, box {16, 32, 1}
};

struct BpCfg {
  int BN, stages, acc_bufs, acc_cols, tmem_cols;
  int kblocks[GET_GEMM_MAX_SEG];
  int kblocks_total, nseg;
  int a_mn, b_mn;
  int np;                 // planes per operand used by the mode (1, 2, 3)
  int mode;
  int kb;                 // k elements per stage
  int ntm, ntn, splits, kb_per_split, items;
  uint32_t a_plane, b_plane;       // byte offset between planes inside a stage
  uint32_t a_bytes, b_bytes, stage_bytes;
  uint32_t a_lbo, b_lbo, a_sbo, b_sbo, a_kstep, b_kstep, a_lt, b_lt;
  int a_boxes, b_boxes;
  uint32_t a_box_bytes, b_box_bytes;
  uint32_t stg_set;       // bytes of one epilogue staging set
  int debug;              // GET_B200_BP_DEBUG=9 with -DGETB_BP_TIMELINE: print the per-role timeline of CTA 0
};

struct BpParams {
  int M, N, Npad;
  int epilogue, accumulate;
  const float* C; int64_t ldc;          // read only for accumulate (the stores go through the tensor maps)
  const float* out1; int64_t ld_out1;   // read only by DGATE_R (dx += ...)
  int has_c, has_o1, has_pl;
  const float* bias;
  const float* aux0; int64_t ld_aux0;
  const float* aux1; int64_t ld_aux1;
  int nplanes, pad_one;
  int group_rows, zr_gs, zr_cols, zr_cols_pad;
  uint32_t drop_thr, drop_seed; float drop_scale;
  const uint32_t* salt;
  // optional row-dot by-product of the TANH_BLEND epilogue (the scorer projection of the GSL block, wrapper.py:158,167):
  // rd_out[(n_tile*2 + half)*M + m] = sum over this warp's columns of dropout_s(out[m,n]) * rd_w[n]
  const float* rd_w; float* rd_out;
  uint32_t rd_thr, rd_seed; float rd_scale;
};

__device__ __forceinline__ void bp_locate(const BpCfg& cfg, int kb, int& seg, int& kin) {
  seg = 0;
  while (seg + 1 < cfg.nseg && kb >= cfg.kblocks[seg]) { kb -= cfg.kblocks[seg]; ++seg; }
  kin = kb * cfg.kb;
}

namespace tc {
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; }
}  // namespace tc

// 16 fp32 values of one row -> staging tile of 32 rows x 16 columns (64-byte rows, TMA SWIZZLE_64B pattern)
__device__ __forceinline__ void bp_stage_f32(uint32_t tile, int lane, const float (&v)[16]) {
  const uint32_t row = tile + (uint32_t)lane * 64u, sw = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
  for (int q = 0; q < 4; ++q) tc::sts128f(row + (((uint32_t)q ^ sw) << 4), v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
}
// 16 values -> nplanes bf16 staging tiles of 32 rows x 16 columns (32-byte rows, TMA SWIZZLE_32B pattern), 1 KB apart
__device__ __forceinline__ void bp_stage_planes(uint32_t tile, int lane, int nplanes, float (&r)[16]) {
  const uint32_t sw = (uint32_t)(lane >> 2) & 1u;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    if (p < nplanes) {
      uint32_t w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const __nv_bfloat16 q0 = __float2bfloat16_rn(r[2 * e]), q1 = __float2bfloat16_rn(r[2 * e + 1]);
        w[e] = (uint32_t)__bfloat16_as_ushort(q0) | ((uint32_t)__bfloat16_as_ushort(q1) << 16);
        r[2 * e] -= __bfloat162float(q0);
        r[2 * e + 1] -= __bfloat162float(q1);
      }
      const uint32_t row = tile + (uint32_t)p * 1024u + (uint32_t)lane * 32u;
      tc::sts128(row + ((0u ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
      tc::sts128(row + ((1u ^ sw) << 4), make_uint4(w[4], w[5], w[6], w[7]));
    }
  }
}
// 16 consecutive fp32 of one row (quads at or beyond `nvalid` columns read as zero)
__device__ __forceinline__ void bp_ld16(const float* p, int nvalid, bool row_ok, float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_ok && q * 4 < nvalid) t = *reinterpret_cast<const float4*>(p + q * 4);
    v[q * 4 + 0] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
  }
}

template <int EPI>
__global__ void __launch_bounds__(BP_THREADS, 1)
gemm_bp_kernel(const __grid_constant__ BpParams p, const __grid_constant__ BpCfg cfg, const __grid_constant__ BpMaps maps) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t bp_smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[BP_MAX_STAGES];    // TMA landed (tx bytes)
  __shared__ __align__(8) uint64_t bar_empty[BP_MAX_STAGES];   // MMAs that read the stage retired (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_accf[2];                // accumulator set complete (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_acce[2];                // accumulator set drained (all epilogue threads arrive)
  __shared__ uint32_t tmem_holder;
#ifdef GETB_BP_TIMELINE   // compile with -DGETB_BP_TIMELINE and run with GET_B200_BP_DEBUG=9: per-role timeline of CTA 0
  __shared__ long long dbg_ts[4][32];
  const long long dbg_t0 = clock64();
#define BP_DBG(role, idx) do { if (cfg.debug == 9 && blockIdx.x == 0 && (idx) < 32) dbg_ts[role][idx] = clock64() - dbg_t0; } while (0)
#else
#define BP_DBG(role, idx) do { } while (0)
#endif

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = cfg.BN;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bp_smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 0) {
    for (int s = 0; s < BP_MAX_STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_accf[b], 1);
      mbar_init(&bar_acce[b], 32 * BP_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)cfg.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = tmem_holder;

  const int n_my = ((int)blockIdx.x < cfg.items) ? (cfg.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int ntn = cfg.ntn, ntm = cfg.ntm, kb_per_split = cfg.kb_per_split, kblocks_total = cfg.kblocks_total;
  const int nstages = cfg.stages;
  const uint32_t stage_bytes = cfg.stage_bytes;

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      for (int s = 0; s < cfg.nseg; ++s) { prefetch_tmap(&maps.a[s]); prefetch_tmap(&maps.b[s]); }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = cfg.a_bytes + cfg.b_bytes, a_bytes = cfg.a_bytes;
      const int a_mn = cfg.a_mn, b_mn = cfg.b_mn, a_boxes = cfg.a_boxes, b_boxes = cfg.b_boxes;
      const uint32_t a_box_bytes = cfg.a_box_bytes, b_box_bytes = cfg.b_box_bytes;
      for (int it = 0; it < n_my; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int nt = item % ntn, mt = (item / ntn) % ntm, z = item / (ntn * ntm);
        const int m0 = mt * BP_BM, n0 = nt * BN;
        const int kb0 = z * kb_per_split, kb1 = min(kblocks_total, kb0 + kb_per_split);
        int seg, kin;
        bp_locate(cfg, kb0, seg, kin);
        int seg_left = cfg.kblocks[seg] - kin / cfg.kb;      // k blocks left in the current segment
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bar_empty[stage], phase ^ 1);
          if (kb == kb0) BP_DBG(0, it * 2);
          if (kb == kb1 - 1) BP_DBG(0, it * 2 + 1);
          if (it == 1 && kb - kb0 < 8) BP_DBG(0, 16 + (kb - kb0));
          const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
          const uint32_t sb = sa + a_bytes;
          mbar_arrive_expect_tx(&bar_full[stage], tx);
          if (a_mn) {
            for (int b = 0; b < a_boxes; ++b) tma_load_3d(sa + (uint32_t)b * a_box_bytes, &maps.a[seg], m0 + b * 64, kin, 0, &bar_full[stage]);
          } else {
            tma_load_3d(sa, &maps.a[seg], kin, m0, 0, &bar_full[stage]);
          }
          if (b_mn) {
            for (int b = 0; b < b_boxes; ++b) tma_load_3d(sb + (uint32_t)b * b_box_bytes, &maps.b[seg], n0 + b * 64, kin, 0, &bar_full[stage]);
          } else {
            tma_load_3d(sb, &maps.b[seg], kin, n0, 0, &bar_full[stage]);
          }
          if (++stage == nstages) { stage = 0; phase ^= 1; }
          kin += cfg.kb;
          if (--seg_left == 0 && seg + 1 < cfg.nseg) { ++seg; kin = 0; seg_left = cfg.kblocks[seg]; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================================== MMA issuer =============================================
    // One thread issues everything, so its instruction count per tcgen05.mma is what bounds the tensor pipe at these
    // tile sizes: the descriptors' constant halves are built once, per MMA only the 14-bit start-address field moves.
    if (lane == 0) {
      // kind::f16 instruction descriptor: D = f32, A = B = bf16, majors, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(cfg.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(cfg.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BP_BM >> 4) << 24);
      const int ksteps = cfg.kb >> 4, mode = cfg.mode, acc_cols = cfg.acc_cols, acc_bufs = cfg.acc_bufs;
      const uint64_t a_d0 = smem_desc(smem_base, cfg.a_lbo, cfg.a_sbo, cfg.a_lt);
      const uint64_t b_d0 = smem_desc(smem_base + cfg.a_bytes, cfg.b_lbo, cfg.b_sbo, cfg.b_lt);
      const uint32_t a_hi = (uint32_t)(a_d0 >> 32), b_hi = (uint32_t)(b_d0 >> 32);
      const uint32_t a_lo0 = (uint32_t)a_d0, b_lo0 = (uint32_t)b_d0;
      const uint32_t a_ks = cfg.a_kstep >> 4, b_ks = cfg.b_kstep >> 4, a_pl = cfg.a_plane >> 4, b_pl = cfg.b_plane >> 4;
      const uint32_t stage16 = stage_bytes >> 4;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int it = 0; it < n_my; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int z = item / (ntn * ntm);
        const int kb0 = z * kb_per_split, kb1 = min(kblocks_total, kb0 + kb_per_split);
        mbar_wait(&bar_acce[acc], acc_phase ^ 1);
        fence_after();
        BP_DBG(1, it * 3);
        const uint32_t d_main = tmem_base + (uint32_t)(acc * acc_cols);
        const uint32_t d_small = d_main + (uint32_t)BN;
        uint32_t first = 0u;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bar_full[stage], phase);
          fence_after();
          if (kb == kb0) BP_DBG(1, it * 3 + 1);
          if (it == 1 && kb - kb0 < 8) BP_DBG(3, (kb - kb0) * 2);
          const uint32_t sa = a_lo0 + (uint32_t)stage * stage16, sb = b_lo0 + (uint32_t)stage * stage16;
          if (mode == 1) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < ksteps) {
                umma_bf16(d_main, desc64(sa + ks * a_ks, a_hi), desc64(sb + ks * b_ks, b_hi), idesc, first);
                first = 1u;
              }
            }
          } else if (mode == 2) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < ksteps) {
                const uint32_t a0 = sa + ks * a_ks, b0 = sb + ks * b_ks;
                umma_bf16(d_main, desc64(a0, a_hi), desc64(b0, b_hi), idesc, first);
                umma_bf16(d_main, desc64(a0, a_hi), desc64(b0 + b_pl, b_hi), idesc, 1u);
                umma_bf16(d_main, desc64(a0 + a_pl, a_hi), desc64(b0, b_hi), idesc, 1u);
                first = 1u;
              }
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < ksteps) {
                const uint32_t a0 = sa + ks * a_ks, b0 = sb + ks * b_ks;
                umma_bf16(d_main, desc64(a0, a_hi), desc64(b0, b_hi), idesc, first);
                umma_bf16(d_small, desc64(a0, a_hi), desc64(b0 + b_pl, b_hi), idesc, first);
                umma_bf16(d_small, desc64(a0 + a_pl, a_hi), desc64(b0, b_hi), idesc, 1u);
                umma_bf16(d_small, desc64(a0 + a_pl, a_hi), desc64(b0 + b_pl, b_hi), idesc, 1u);
                umma_bf16(d_small, desc64(a0, a_hi), desc64(b0 + 2u * b_pl, b_hi), idesc, 1u);
                umma_bf16(d_small, desc64(a0 + 2u * a_pl, a_hi), desc64(b0, b_hi), idesc, 1u);
                first = 1u;
              }
            }
          }
          umma_commit(&bar_empty[stage]);
          if (it == 1 && kb - kb0 < 8) BP_DBG(3, (kb - kb0) * 2 + 1);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bar_accf[acc]);
        BP_DBG(1, it * 3 + 2);
        if (acc_bufs == 2) {
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        } else {
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // =========================================== epilogue ===============================================
    // TMEM hands every thread one output ROW. The fused math runs in that layout on 16-column pieces: the auxiliary
    // operands are read row-wise (one 64-byte piece per thread and instruction: 2 KB in flight per warp instruction),
    // the results are staged in swizzled shared-memory tiles and leave through TMA stores (coalesced, asynchronous,
    // clipped at the tensor edges by the hardware), fp32 outputs and bf16 planes alike.
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int chalf = (warp - 2) >> 2;            // two warps per quarter: even / odd 16-column pieces
    const uint32_t stg_set = cfg.stg_set;
    const uint32_t stg0 = smem_base + (uint32_t)nstages * stage_bytes + (uint32_t)(warp - 2) * 2u * stg_set;
    // tile offsets inside a staging set (fixed per launch): [C 2 KB][out1 2 KB][planes 1 KB each]; the fused z|r epilogue
    // writes either z (C) or r (out1) + planes per piece, so its out1 tile shares the C slot
    const uint32_t off_o1 = (EPI != GET_BPE_ZR && (p.has_c || cfg.splits > 1)) ? 2048u : 0u;
    const uint32_t off_p = off_o1 + (p.has_o1 ? 2048u : 0u);
    uint32_t set = 0;
    const bool small = cfg.mode == 3;
    const int npieces = BN >> 4;
    const int last_piece = (npieces - 1 >= chalf) ? ((npieces - 1 - chalf) / 2 * 2 + chalf) : -1;
    const int splits = cfg.splits, acc_cols = cfg.acc_cols, acc_bufs = cfg.acc_bufs;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int nt = item % ntn, mt = (item / ntn) % ntm, z = item / (ntn * ntm);
      const int m_base = mt * BP_BM + quarter * 32;
      const int m = m_base + lane;
      const bool row_ok = m < p.M;
      const int n0 = nt * BN;
      mbar_wait(&bar_accf[acc], acc_phase);
      fence_after();
      if (warp == 2 && lane == 0) BP_DBG(2, it * 2);
      const uint32_t t_main = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols);
      const uint32_t t_small = t_main + (uint32_t)BN;
      if (last_piece < 0) {               // 16-column tiles: the odd warps have nothing to read
        fence_before();
        mbar_arrive(&bar_acce[acc]);
      }
      float rowdot = 0.f;
      for (int pc = chalf; pc < npieces; pc += 2) {
        const int col = pc << 4;
        const int n = n0 + col;
        uint32_t rm[16], rs[16];
        tmem_ld16_nowait(t_main + (uint32_t)col, rm);
        if (small) tmem_ld16_nowait(t_small + (uint32_t)col, rs);
        // ---- where does this piece go? c = column inside the destination tensors, nv = valid columns of the piece
        int grp = 0, c = n, nv = p.N - n, nvp = p.Npad - n;
        if (EPI == GET_BPE_ZR) {
          grp = n / p.zr_gs;
          c = n - grp * p.zr_gs;
          nv = grp > 1 ? 0 : p.zr_cols - c;
          nvp = grp == 1 ? p.zr_cols_pad - c : 0;
        }
        nv = max(0, min(16, nv));
        nvp = max(0, min(16, nvp));
        // ---- auxiliary operands, row-wise (issued before the accumulator wait: the latencies overlap)
        float a0[16], a1[16], o[16];
        if (splits == 1) {
          if (EPI == GET_BPE_STORE) {
            if (p.accumulate) bp_ld16(p.C + (int64_t)m * p.ldc + c, nv, row_ok, o);
          } else if (EPI == GET_BPE_ZR) {
            if (grp == 1) bp_ld16(p.aux0 + (int64_t)m * p.ld_aux0 + c, nv, row_ok, a0);
          } else if (EPI == GET_BPE_TANH_BLEND) {
            bp_ld16(p.aux0 + (int64_t)m * p.ld_aux0 + c, nv, row_ok, a0);
            bp_ld16(p.aux1 + (int64_t)m * p.ld_aux1 + c, nv, row_ok, a1);
          } else if (EPI == GET_BPE_TANH_ROWGROUP) {
            bp_ld16(p.aux0 + (int64_t)(m / p.group_rows) * p.ld_aux0 + c, nv, row_ok, a0);
          } else if (EPI == GET_BPE_DGATE_R) {
            bp_ld16(p.aux0 + (int64_t)m * p.ld_aux0 + c, nv, row_ok, a0);
            bp_ld16(p.aux1 + (int64_t)m * p.ld_aux1 + c, nv, row_ok, a1);
            bp_ld16(p.out1 + (int64_t)m * p.ld_out1 + c, nv, row_ok, o);
          }
        }
        tmem_ld_wait();
        if (pc == last_piece) {           // last piece of this accumulator set for this warp: hand it back to the MMA warp
          fence_before();
          mbar_arrive(&bar_acce[acc]);
        }
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = small ? __uint_as_float(rm[e]) + __uint_as_float(rs[e]) : __uint_as_float(rm[e]);
        // the stores issued from this staging set (two pieces ago) must have finished READING it before it is overwritten
        const uint32_t stg_c = stg0 + set * stg_set, stg_o1 = stg_c + off_o1, stg_p = stg_c + off_p;
        set ^= 1u;
        if (lane == 0) bulk_wait_read1();
        __syncwarp();
        if (splits > 1) {
          bp_stage_f32(stg_c, lane, v);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&maps.c, stg_c, n, m_base, z);
            bulk_commit();
          }
          continue;
        }
        if (nv == 0 && nvp == 0) {
          if (lane == 0) bulk_commit();    // keep one bulk group per piece: the wait above counts groups
          continue;
        }
        if (p.bias) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q * 4 < nv) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n) + q);
              v[q * 4 + 0] += b.x; v[q * 4 + 1] += b.y; v[q * 4 + 2] += b.z; v[q * 4 + 3] += b.w;
            }
          }
        }
        bool st_c = false, st_o1 = false, st_p = false;
        float w[16];                       // values for the planes
        if (EPI == GET_BPE_STORE) {
          if (p.drop_thr) {
            const uint32_t sd = p.drop_seed + __ldg(p.salt);
#pragma unroll
            for (int e = 0; e < 16; ++e)
              v[e] = drop_keep(sd, (uint64_t)m * (uint64_t)p.N + (uint64_t)(n + e), p.drop_thr) ? v[e] * p.drop_scale : 0.f;
          }
          if (p.accumulate) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += o[e];
          }
          st_c = p.has_c && nv > 0;
          if (st_c) bp_stage_f32(stg_c, lane, v);
#pragma unroll
          for (int e = 0; e < 16; ++e) w[e] = v[e];
          st_p = p.has_pl;
        } else if (EPI == GET_BPE_ZR) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = sigmoid_fast(v[e]);
          if (grp == 0) {
            st_c = nv > 0;
            if (st_c) bp_stage_f32(stg_c, lane, v);
          } else {
            st_o1 = nv > 0;
            if (st_o1) bp_stage_f32(stg_o1, lane, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) w[e] = v[e] * a0[e];
            st_p = p.has_pl;
          }
        } else if (EPI == GET_BPE_TANH_BLEND) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            v[e] = tanh_fast(v[e]);
            w[e] = v[e] * a0[e] + a1[e] * (1.0f - a0[e]);
          }
          st_o1 = p.has_o1 && nv > 0;
          if (st_o1) bp_stage_f32(stg_o1, lane, v);
          st_c = p.has_c && nv > 0;
          if (st_c) bp_stage_f32(stg_c, lane, w);
          st_p = p.has_pl;
          if (p.rd_out) {
            const uint32_t sd = p.rd_seed + (p.rd_thr ? __ldg(p.salt) : 0u);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q * 4 < nv) {
                const float4 rw = __ldg(reinterpret_cast<const float4*>(p.rd_w + n) + q);
                float4 f = make_float4(w[q * 4 + 0], w[q * 4 + 1], w[q * 4 + 2], w[q * 4 + 3]);
                if (p.rd_thr) drop_apply4(sd, (uint64_t)m * (uint64_t)p.N + (uint64_t)(n + q * 4), p.rd_thr, p.rd_scale, f);
                rowdot = fmaf(f.x, rw.x, rowdot); rowdot = fmaf(f.y, rw.y, rowdot);
                rowdot = fmaf(f.z, rw.z, rowdot); rowdot = fmaf(f.w, rw.w, rowdot);
              }
            }
          }
        } else if (EPI == GET_BPE_TANH_ROWGROUP) {
#pragma unroll
          for (int e = 0; e < 16; ++e) { v[e] = tanh_fast(v[e] + a0[e]); w[e] = v[e]; }
          st_c = nv > 0;
          if (st_c) bp_stage_f32(stg_c, lane, v);
          st_p = p.has_pl;
        } else if (EPI == GET_BPE_DGATE_R) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            w[e] = v[e] * a0[e] * a1[e] * (1.0f - a1[e]);
            o[e] += v[e] * a1[e];
          }
          st_c = p.has_c && nv > 0;
          if (st_c) bp_stage_f32(stg_c, lane, w);
          st_o1 = nv > 0;
          if (st_o1) bp_stage_f32(stg_o1, lane, o);
          st_p = p.has_pl;
        } else {   // GET_BPE_TANH
#pragma unroll
          for (int e = 0; e < 16; ++e) { v[e] = tanh_fast(v[e]); w[e] = v[e]; }
          st_c = nv > 0;
          if (st_c) bp_stage_f32(stg_c, lane, v);
          st_p = p.has_pl;
        }
        st_p = st_p && nvp > 0;
        if (st_p) {
          // columns at or beyond the logical width are padding: zeros, 1.0 in the first pad column when pad_one
          if (nv < 16) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (e >= nv) w[e] = (p.pad_one && e == nv) ? 1.0f : 0.0f;
          }
          bp_stage_planes(stg_p, lane, p.nplanes, w);
        }
        fence_proxy_async();               // generic-proxy stores -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          if (st_c) tma_store_2d(&maps.c, stg_c, c, m_base);
          if (st_o1) tma_store_2d(&maps.o1, stg_o1, c, m_base);
          if (st_p) {
            for (int pl = 0; pl < p.nplanes; ++pl) tma_store_3d(&maps.pl, stg_p + (uint32_t)pl * 1024u, c, m_base, pl);
          }
          bulk_commit();
        }
      }
      if (EPI == GET_BPE_TANH_BLEND && p.rd_out && row_ok && splits == 1) p.rd_out[(int64_t)(nt * 2 + chalf) * p.M + m] = rowdot;
      if (warp == 2 && lane == 0) BP_DBG(2, it * 2 + 1);
      if (acc_bufs == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
    if (lane == 0) bulk_wait_read0();     // shared memory must outlive the last stores' reads
    __syncwarp();
  }
  fence_before();
  __syncthreads();
#ifdef GETB_BP_TIMELINE
  if (cfg.debug == 9 && blockIdx.x == 0 && tid == 0) {
    printf("BPDBG cfg BN=%d kb=%d stages=%d mode=%d items=%d kblocks=%d a_mn=%d stage_bytes=%u\n", cfg.BN, cfg.kb, cfg.stages, cfg.mode,
           cfg.items, cfg.kblocks_total, cfg.a_mn, cfg.stage_bytes);
    for (int it = 0; it < n_my && it < 5; ++it)
      printf("BPDBG it=%d prod[%lld %lld] mma[acce %lld first_full %lld last_commit %lld] epi[%lld %lld]\n", it, dbg_ts[0][it * 2],
             dbg_ts[0][it * 2 + 1], dbg_ts[1][it * 3], dbg_ts[1][it * 3 + 1], dbg_ts[1][it * 3 + 2], dbg_ts[2][it * 2], dbg_ts[2][it * 2 + 1]);
    for (int j = 0; j < 8; ++j)
      printf("BPDBG it=1 kb=%d tma_issue %lld | mma full_seen %lld issued %lld\n", j, dbg_ts[0][16 + j], dbg_ts[3][j * 2], dbg_ts[3][j * 2 + 1]);
    printf("BPDBG end %lld\n", clock64() - dbg_t0);
  }
#endif
  if (warp == 1) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg.tmem_cols) : "memory");
  }
}

// ---- split-K reduction into gradient blocks -----------------------------------------------------------------------
struct BpDstList {
  get_bp_dst d[GET_BP_MAX_DST];
  int n;
};

__global__ void __launch_bounds__(256) bp_splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int64_t ws_ld,
                                                               const __grid_constant__ BpDstList L, int accumulate) {
  // Fixed-order sum over the split-K partials (deterministic), scattered into the destination blocks. A thread owns four
  // consecutive columns of a row: 128-bit loads from the workspace (its pitch and the block's first column are multiples
  // of 4 floats in every weight-gradient call; anything else takes the scalar path), all splits of a quad in flight.
  const int64_t zs = (int64_t)M * ws_ld;
  for (int b = 0; b < L.n; ++b) {
    const get_bp_dst& d = L.d[b];
    const bool vec = (d.col0 & 3) == 0 && (ws_ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(ws) & 15u) == 0);
    const unsigned nq = vec ? (unsigned)((d.ncols + 3) >> 2) : (unsigned)d.ncols;
    const unsigned total = (unsigned)d.nrows * nq;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const unsigned r = e / nq, q = e - r * nq;
      if (vec) {
        const int c = (int)q * 4;
        const float* src = ws + (int64_t)(d.row0 + r) * ws_ld + (d.col0 + c);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int z = 0; z < splits; ++z) {
          const float4 v = *reinterpret_cast<const float4*>(src + (int64_t)z * zs);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float* o = d.dst + (int64_t)r * d.ld + c;
        const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c + u < d.ncols) o[u] = accumulate ? o[u] + a4[u] : a4[u];
      } else {
        const float* src = ws + (int64_t)(d.row0 + r) * ws_ld + (d.col0 + (int)q);
        float acc = 0.f;
        for (int z = 0; z < splits; ++z) acc += src[(int64_t)z * zs];
        float* o = d.dst + (int64_t)r * d.ld + q;
        *o = accumulate ? *o + acc : acc;
      }
    }
  }
}

// ---- fp32 -> planes ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) to_planes_kernel(const float* __restrict__ src, int64_t ld_src, int rows, int cols,
                                                        __nv_bfloat16* __restrict__ dst, int64_t ld_out, int64_t plane_stride,
                                                        int nplanes, int pad_one, int vec, int width) {
  const int nq = width >> 2;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (int64_t)rows * nq) return;
  const int r = (int)(q / nq), c = (int)(q % nq) * 4;
  float v[4];
  if (vec && c + 4 <= cols) {
    const float4 t = *reinterpret_cast<const float4*>(src + (int64_t)r * ld_src + c);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (c + e < cols) ? src[(int64_t)r * ld_src + c + e] : ((pad_one && c + e == cols) ? 1.0f : 0.0f);
  }
  planes_store4(dst + (int64_t)r * ld_out + c, plane_stride, nplanes, v);
}

// ---- weight packing: many (strided) matrices -> 3 planes each, plus fused bias vectors, one launch -------------------
// One block per 32 x 32 tile of a job (kind 0) or per 1024 elements of a bias vector (kind 1). Transposed views
// (ld_r == 1) are read along their contiguous index and turned through shared memory, so reads and writes are coalesced
// whatever the orientation of the packed weight.
__global__ void __launch_bounds__(256) pack_planes_multi_kernel(const get_pack_job* __restrict__ jobs, int n_jobs) {
  __shared__ float tile[32][33];
  int lo = 0, hi = n_jobs - 1;
  const int64_t blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const get_pack_job j = jobs[lo];
  const int64_t lb = blk - j.first_block;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (j.kind == 1) {
    for (int e = (int)lb * 1024 + threadIdx.x; e < min(j.rows, (int)(lb + 1) * 1024); e += 256)
      reinterpret_cast<float*>(j.dst)[e] = j.src[e] + (j.src2 ? j.src2[e] : 0.f);
    return;
  }
  const int tiles_c = (j.cols + 31) >> 5;
  const int r0 = (int)(lb / tiles_c) * 32, c0 = (int)(lb % tiles_c) * 32;
  const bool by_rows = j.ld_r == 1 && j.ld_c != 1;      // the source is contiguous along r: read with threads along r
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = ty + i * 8;
    const int r = by_rows ? r0 + tx : r0 + a, c = by_rows ? c0 + a : c0 + tx;
    float v = 0.f;
    if (r < j.rows && c < j.cols) v = j.src[(int64_t)r * j.ld_r + (int64_t)c * j.ld_c];
    if (by_rows) tile[tx][a] = v; else tile[a][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    if (r < j.rows && c < j.cols) {
      float v = tile[ty + i * 8][tx];
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(j.dst) + (int64_t)r * j.ld_out + c;
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        const __nv_bfloat16 q = __float2bfloat16_rn(v);
        d[(int64_t)p * j.plane_stride] = q;
        v -= __bfloat162float(q);
      }
    }
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnBp)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnBp bp_encode_fn() {
  static EncodeTiledFnBp fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnBp>(ptr);
  });
  return fn;
}

struct BpMapKey {
  const void* ptr;
  int64_t d0, d1, d2, ld, ps;
  int b0, b1, b2, sw, f32, rank;
  bool operator==(const BpMapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld == o.ld && ps == o.ps && b0 == o.b0 && b1 == o.b1 &&
           b2 == o.b2 && sw == o.sw && f32 == o.f32 && rank == o.rank;
  }
};
struct BpMapKeyHash {
  size_t operator()(const BpMapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)k.d0 * 0xC2B2AE3D27D4EB4Full + (uint64_t)k.d1 * 0x165667B19E3779F9ull + (uint64_t)k.d2 * 0x27D4EB2F165667C5ull;
    h ^= (uint64_t)k.ld * 31 + (uint64_t)k.ps * 131 + (uint64_t)k.b0 * 7 + (uint64_t)k.b1 * 1315423911ull + (uint64_t)k.b2 * 2654435761ull +
         (uint64_t)k.sw * 97 + (uint64_t)k.f32 * 1009 + (uint64_t)k.rank * 7919;
    return (size_t)(h ^ (h >> 29));
  }
};

// Tensor map over {d0 inner contiguous, d1 rows `ld` elements apart[, d2 slices `ps` elements apart]}, box {b0, b1[, b2]};
// elements bf16 (f32 = 0) or fp32 (f32 = 1); sw = swizzle span in bytes (32 / 64 / 128, 0 = none). Encoding is memoised: the
// allocator hands the same activation addresses back step after step.
bool make_tensor_map(void* map_out, const void* ptr, int f32, int rank, int64_t d0, int64_t d1, int64_t d2, int64_t ld,
                     int64_t ps, int b0, int b1, int b2, int sw) {
  CUtensorMap* map = reinterpret_cast<CUtensorMap*>(map_out);
  static std::mutex mu;
  static std::unordered_map<BpMapKey, CUtensorMap, BpMapKeyHash> cache;
  const BpMapKey key{ptr, d0, d1, d2, ld, ps, b0, b1, b2, sw, f32, rank};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *map = it->second;
    return true;
  }
  EncodeTiledFnBp enc = bp_encode_fn();
  if (!enc) {
    set_error("get_gemm_bp: cuTensorMapEncodeTiled is not available");
    return false;
  }
  const uint64_t es = f32 ? 4u : 2u;
  cuuint64_t gdim[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * es, (cuuint64_t)(d2 > 1 ? ps : ld * d1) * es};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle swz = sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (sw == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
  const CUresult rc = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                          const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("get_gemm_bp: cuTensorMapEncodeTiled failed (%d): rank %d f32 %d dims %lld %lld %lld ld %lld ps %lld box %d %d %d", (int)rc,
              rank, f32, (long long)d0, (long long)d1, (long long)d2, (long long)ld, (long long)ps, b0, b1, b2);
    return false;
  }
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *map);
  return true;
}

static int bp_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static int bp_round_up(int v, int q) { return (v + q - 1) / q * q; }

// Cost model for the N tile (cycles per k element and work item; L2 -> SM delivers ~40 B/cycle/SM when all SMs pull):
static int bp_choose_bn(int M, int Npad, int mode, int splits) {
  const int np = mode;
  const int nmma = mode == 1 ? 1 : (mode == 2 ? 3 : 6);
  const int ntm = (M + BP_BM - 1) / BP_BM;
  int best = 0;
  double best_cost = 1e30;
  for (int bn = 16; bn <= 256; bn += 16) {
    const int acc_cols = bn * (mode == 3 ? 2 : 1);
    if (acc_cols > 512) continue;
    const int nt = (Npad + bn - 1) / bn;
    if ((nt - 1) * bn >= Npad) continue;
    const int64_t items = (int64_t)ntm * nt * splits;
    const int64_t rounds = (items + BP_SMS - 1) / BP_SMS;
    const bool dbl = 2 * acc_cols <= 512;
    const uint32_t stage = (uint32_t)(BP_BM + bn) * np * 64u;
    const int stg_bytes = BP_EPI_WARPS * 2 * (mode == 3 ? 5120 : 6144);   // largest staging the mode is used with
    if ((226 * 1024 - stg_bytes - 1024) / (int)stage < 3) continue;
    const double l2 = (double)(BP_BM + bn) * np * 2.0 / 40.0;
    const double mma = (double)nmma * bn / 32.0;
    const double epi = 3.0 + bn / 16.0;            // in units comparable to one k element of a 300-deep contraction
    double per_item = (l2 > mma ? l2 : mma) + ((rounds > 1 && !dbl) ? 2.0 * epi : 0.25 * epi);
    const double cost = (double)rounds * per_item;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

static int bp_plan(const get_gemm_bp_desc* d, BpCfg& cfg, BpParams& p) {
  GETB_REQUIRE(d != nullptr, "get_gemm_bp: null descriptor");
  GETB_REQUIRE(d->nseg >= 1 && d->nseg <= GET_GEMM_MAX_SEG, "get_gemm_bp: nseg=%d out of range", d->nseg);
  GETB_REQUIRE(d->M >= 1 && d->N >= 4 && (d->N % 4) == 0, "get_gemm_bp: need M >= 1, N >= 4, N %% 4 == 0 (M=%d N=%d)", d->M, d->N);
  GETB_REQUIRE(d->mode >= 1 && d->mode <= 3, "get_gemm_bp: mode=%d out of range", d->mode);
  GETB_REQUIRE(d->epilogue >= GET_BPE_STORE && d->epilogue <= GET_BPE_TANH, "get_gemm_bp: bad epilogue %d", d->epilogue);
  memset(&cfg, 0, sizeof(cfg));
  memset(&p, 0, sizeof(p));
  cfg.mode = d->mode;
  cfg.np = d->mode;
  cfg.nseg = d->nseg;
  cfg.a_mn = d->A[0].trans ? 1 : 0;
  cfg.b_mn = d->B[0].trans ? 1 : 0;
  int kb = d->kblock ? d->kblock : bp_env_int("GET_B200_BP_KB", 32);
  GETB_REQUIRE(kb == 32 || kb == 64, "get_gemm_bp: kblock must be 32 or 64");
  cfg.kb = kb;
  for (int s = 0; s < d->nseg; ++s) {
    const get_bp_tensor &A = d->A[s], &B = d->B[s];
    GETB_REQUIRE(A.ptr && B.ptr && d->K[s] >= 1, "get_gemm_bp: null operand / K in segment %d", s);
    GETB_REQUIRE((A.trans ? 1 : 0) == cfg.a_mn && (B.trans ? 1 : 0) == cfg.b_mn, "get_gemm_bp: segments must share trans");
    GETB_REQUIRE(A.planes >= cfg.np && B.planes >= cfg.np, "get_gemm_bp: mode %d needs %d planes per operand", d->mode, cfg.np);
    GETB_REQUIRE(aligned16(A.ptr) && aligned16(B.ptr) && (A.ld % 8) == 0 && (B.ld % 8) == 0 &&
                     (cfg.np == 1 || ((A.plane_stride % 8) == 0 && (B.plane_stride % 8) == 0)),
                 "get_gemm_bp: operands must be 16-byte aligned with ld %% 8 == 0 and plane_stride %% 8 == 0");
    cfg.kblocks[s] = (d->K[s] + kb - 1) / kb;
    cfg.kblocks_total += cfg.kblocks[s];
  }
  p.M = d->M; p.N = d->N;
  // columns written to planes_out: the logical ones plus padding up to a multiple of 8 (zeros; with pad_one column N holds
  // 1.0, so the extent always contains it)
  p.Npad = d->planes_out ? bp_round_up(d->N + (d->planes_out_pad_one ? 1 : 0), 8) : d->N;
  p.epilogue = d->epilogue; p.accumulate = d->accumulate;
  p.C = d->C; p.ldc = d->ldc; p.out1 = d->out1; p.ld_out1 = d->ld_out1;
  p.has_c = d->C != nullptr; p.has_o1 = d->out1 != nullptr; p.has_pl = d->planes_out != nullptr;
  p.bias = d->bias; p.aux0 = d->aux0; p.ld_aux0 = d->ld_aux0; p.aux1 = d->aux1; p.ld_aux1 = d->ld_aux1;
  p.nplanes = d->planes_out_n; p.pad_one = d->planes_out_pad_one;
  p.group_rows = d->group_rows; p.zr_gs = d->zr_group_stride; p.zr_cols = d->zr_cols;
  p.zr_cols_pad = bp_round_up(d->zr_cols, 8);
  p.salt = dropout_salt_ptr();
  if (d->rowdot_out) {
    GETB_REQUIRE(d->epilogue == GET_BPE_TANH_BLEND && d->rowdot_w && aligned16(d->rowdot_w) && d->rowdot_p >= 0.f && d->rowdot_p < 1.f,
                 "get_gemm_bp: the row-dot by-product belongs to TANH_BLEND and needs a 16-byte aligned weight vector");
    p.rd_w = d->rowdot_w; p.rd_out = d->rowdot_out;
    p.rd_thr = d->rowdot_p > 0.f ? drop_threshold(d->rowdot_p) : 0u;
    p.rd_seed = d->rowdot_seed; p.rd_scale = 1.0f / (1.0f - d->rowdot_p);
  }
  {
    const int npl = p.has_pl ? p.nplanes : 0;
    int kb_set = d->epilogue == GET_BPE_ZR ? 2 + npl : 2 * p.has_c + 2 * p.has_o1 + npl;
    if (d->split_k > 1 || kb_set < 2) kb_set = 2;
    GETB_REQUIRE(kb_set * 1024 <= BP_STG_SET_MAX, "get_gemm_bp: C + out1 + planes_out exceed the epilogue staging set (drop one output)");
    cfg.stg_set = (uint32_t)kb_set * 1024u;
  }
  GETB_REQUIRE(d->drop_out_p >= 0.f && d->drop_out_p < 1.f, "get_gemm_bp: dropout probability must be in [0,1)");
  if (d->drop_out_p > 0.f) {
    GETB_REQUIRE(d->epilogue == GET_BPE_STORE, "get_gemm_bp: dropout-out belongs to the STORE epilogue");
    p.drop_thr = drop_threshold(d->drop_out_p);
    p.drop_seed = d->drop_out_seed;
    p.drop_scale = 1.0f / (1.0f - d->drop_out_p);
  }
  // pointer / alignment contract of the vectorised epilogue
  auto ok4 = [](const void* q, int64_t ld) { return q == nullptr || (aligned16(q) && (ld % 4) == 0); };
  GETB_REQUIRE(ok4(p.C, p.ldc) && ok4(p.out1, p.ld_out1) && ok4(p.aux0, p.ld_aux0) && ok4(p.aux1, p.ld_aux1) &&
                   (p.bias == nullptr || aligned16(p.bias)),
               "get_gemm_bp: fp32 epilogue tensors must be 16-byte aligned with ld %% 4 == 0");
  if (d->planes_out) {
    GETB_REQUIRE(d->epilogue == GET_BPE_ZR || d->ld_planes_out >= p.Npad,
                 "get_gemm_bp: planes_out row pitch %lld is smaller than the written width %d (logical width + ones column, padded to 8)",
                 (long long)d->ld_planes_out, p.Npad);
    GETB_REQUIRE(aligned16(d->planes_out) && (d->ld_planes_out % 8) == 0 && (p.nplanes == 1 || (d->planes_out_stride % 8) == 0) &&
                     p.nplanes >= 1 && p.nplanes <= 3,
                 "get_gemm_bp: planes_out must be 16-byte aligned with ld %% 8 == 0, plane stride %% 8 == 0, 1..3 planes");
  }
  switch (d->epilogue) {
    case GET_BPE_STORE: GETB_REQUIRE(p.C || p.has_pl || d->split_k > 1, "get_gemm_bp: STORE needs C or planes_out"); GETB_REQUIRE(!p.accumulate || p.C, "get_gemm_bp: accumulate needs C"); break;
    case GET_BPE_ZR:
      GETB_REQUIRE(p.C && p.out1 && p.zr_gs > 0 && p.zr_cols > 0 && (p.zr_cols % 4) == 0 && p.zr_cols <= p.zr_gs && (!p.has_pl || p.aux0),
                   "get_gemm_bp: ZR needs C (z), out1 (r), zr_group_stride >= zr_cols, aux0 (x) with planes_out");
      break;
    case GET_BPE_TANH_BLEND: GETB_REQUIRE(p.aux0 && p.aux1, "get_gemm_bp: TANH_BLEND needs aux0 (z) and aux1 (x)"); break;
    case GET_BPE_TANH_ROWGROUP: GETB_REQUIRE(p.C && p.aux0 && p.group_rows > 0, "get_gemm_bp: TANH_ROWGROUP needs C, aux0, group_rows"); break;
    case GET_BPE_DGATE_R: GETB_REQUIRE(p.aux0 && p.aux1 && p.out1, "get_gemm_bp: DGATE_R needs aux0 (x), aux1 (r), out1 (dx)"); break;
    case GET_BPE_TANH: GETB_REQUIRE(p.C != nullptr, "get_gemm_bp: TANH needs C"); break;
    default: break;
  }
  cfg.ntm = (d->M + BP_BM - 1) / BP_BM;
  int splits = d->split_k > 1 ? d->split_k : 1;
  if (splits > cfg.kblocks_total) splits = cfg.kblocks_total;
  cfg.kb_per_split = (cfg.kblocks_total + splits - 1) / splits;
  cfg.splits = (cfg.kblocks_total + cfg.kb_per_split - 1) / cfg.kb_per_split;
  int bn = d->tile_n;
  if (bn <= 0) bn = bp_choose_bn(d->M, p.Npad, d->mode, cfg.splits);
  GETB_REQUIRE(bn >= 16 && bn <= 256 && (bn % 16) == 0, "get_gemm_bp: tile_n=%d must be a multiple of 16 in [16,256]", bn);
  GETB_REQUIRE(d->epilogue != GET_BPE_ZR || (p.zr_gs % bn) == 0, "get_gemm_bp: zr_group_stride must be a multiple of the N tile (%d)", bn);
  cfg.BN = bn;
  cfg.ntn = (p.Npad + bn - 1) / bn;
  cfg.items = cfg.ntm * cfg.ntn * cfg.splits;
  cfg.acc_cols = bn * (d->mode == 3 ? 2 : 1);
  GETB_REQUIRE(cfg.acc_cols <= 512, "get_gemm_bp: tile_n=%d does not fit TMEM in mode %d", bn, d->mode);
  cfg.acc_bufs = (cfg.items > BP_SMS && 2 * cfg.acc_cols <= 512) ? 2 : 1;
  int tc = 32;
  while (tc < cfg.acc_bufs * cfg.acc_cols) tc <<= 1;
  cfg.tmem_cols = tc;
  if (cfg.splits > 1) {
    const int64_t ws_ld = (int64_t)cfg.ntn * bn;
    GETB_REQUIRE(d->workspace && d->workspace_floats >= (int64_t)cfg.splits * d->M * ws_ld,
                 "get_gemm_bp: split-K workspace too small (%lld floats needed)", (long long)((int64_t)cfg.splits * d->M * ws_ld));
    GETB_REQUIRE(aligned16(d->workspace), "get_gemm_bp: workspace must be 16-byte aligned");
  }
  // shared-memory geometry
  const uint32_t rb = (uint32_t)kb * 2u;                 // bytes of one K-major tile row
  if (cfg.a_mn) {
    cfg.a_boxes = BP_BM / 64;
    cfg.a_box_bytes = (uint32_t)cfg.np * kb * 128u;
    cfg.a_bytes = cfg.a_boxes * cfg.a_box_bytes;
    cfg.a_plane = (uint32_t)kb * 128u;
    cfg.a_lbo = cfg.a_box_bytes; cfg.a_sbo = 1024u; cfg.a_kstep = 2048u; cfg.a_lt = 2u;
  } else {
    cfg.a_plane = (uint32_t)BP_BM * rb;
    cfg.a_bytes = cfg.np * cfg.a_plane;
    cfg.a_lbo = 16u; cfg.a_sbo = 8u * rb; cfg.a_kstep = 32u; cfg.a_lt = kb == 64 ? 2u : 4u;
  }
  if (cfg.b_mn) {
    cfg.b_boxes = (bn + 63) / 64;
    cfg.b_box_bytes = (uint32_t)cfg.np * kb * 128u;
    cfg.b_bytes = cfg.b_boxes * cfg.b_box_bytes;
    cfg.b_plane = (uint32_t)kb * 128u;
    cfg.b_lbo = cfg.b_box_bytes; cfg.b_sbo = 1024u; cfg.b_kstep = 2048u; cfg.b_lt = 2u;
  } else {
    cfg.b_plane = (uint32_t)bn * rb;
    cfg.b_bytes = cfg.np * cfg.b_plane;
    cfg.b_lbo = 16u; cfg.b_sbo = 8u * rb; cfg.b_kstep = 32u; cfg.b_lt = kb == 64 ? 2u : 4u;
  }
  cfg.stage_bytes = cfg.a_bytes + cfg.b_bytes;
  // descriptor experiments from the environment (tests / bring-up only)
  {
    const int v = bp_env_int("GET_B200_BP_MN_SBO", 0);
    if (v > 0) { if (cfg.a_mn) cfg.a_sbo = (uint32_t)v; if (cfg.b_mn) cfg.b_sbo = (uint32_t)v; }
    const int swap = bp_env_int("GET_B200_BP_MN_SWAP", 0);   // swap the roles of LBO / SBO for MN-major operands
    if (swap) {
      if (cfg.a_mn) { const uint32_t t = cfg.a_lbo; cfg.a_lbo = cfg.a_sbo; cfg.a_sbo = t; }
      if (cfg.b_mn) { const uint32_t t = cfg.b_lbo; cfg.b_lbo = cfg.b_sbo; cfg.b_sbo = t; }
    }
  }
  const int stg_bytes = BP_EPI_WARPS * 2 * (int)cfg.stg_set;
  int stages = (226 * 1024 - stg_bytes - 1024) / (int)cfg.stage_bytes;
  if (stages > BP_MAX_STAGES) stages = BP_MAX_STAGES;
  const int cap = bp_env_int("GET_B200_BP_STAGES", 0);
  if (cap > 0 && stages > cap) stages = cap;
  GETB_REQUIRE(stages >= 2, "get_gemm_bp: tile (mode %d, tile_n %d, kblock %d) does not fit shared memory", d->mode, bn, kb);
  cfg.stages = stages;
  cfg.debug = bp_env_int("GET_B200_BP_DEBUG", 0);
  return 0;
}

typedef void (*BpKernelFn)(const BpParams, const BpCfg, const BpMaps);

static int bp_launch(const get_gemm_bp_desc* d, cudaStream_t st) {
  BpCfg cfg;
  BpParams p;
  const int rc = bp_plan(d, cfg, p);
  if (rc != 0) return rc;
  BpMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int kb = cfg.kb, np = cfg.np;
  for (int s = 0; s < d->nseg; ++s) {
    const get_bp_tensor &A = d->A[s], &B = d->B[s];
    bool ok;
    if (cfg.a_mn) ok = make_tensor_map(&maps.a[s], A.ptr, 0, 3, d->M, d->K[s], np, A.ld, A.plane_stride, 64, kb, np, 128);
    else ok = make_tensor_map(&maps.a[s], A.ptr, 0, 3, d->K[s], d->M, np, A.ld, A.plane_stride, kb, BP_BM, np, kb == 64 ? 128 : 64);
    if (!ok) return -3;
    if (cfg.b_mn) ok = make_tensor_map(&maps.b[s], B.ptr, 0, 3, d->N, d->K[s], np, B.ld, B.plane_stride, 64, kb, np, 128);
    else ok = make_tensor_map(&maps.b[s], B.ptr, 0, 3, d->K[s], d->N, np, B.ld, B.plane_stride, kb, cfg.BN, np, kb == 64 ? 128 : 64);
    if (!ok) return -3;
  }
  // output maps: the epilogue leaves through TMA stores of 32-row x 16-column pieces (clipped at the tensor edges)
  const bool zr = d->epilogue == GET_BPE_ZR;
  if (cfg.splits > 1) {
    const int64_t ws_ld = (int64_t)cfg.ntn * cfg.BN;
    if (!make_tensor_map(&maps.c, d->workspace, 1, 3, ws_ld, d->M, cfg.splits, ws_ld, (int64_t)d->M * ws_ld, 16, 32, 1, 64)) return -3;
  } else {
    const int64_t ncols = zr ? d->zr_cols : d->N;
    if (d->C && !make_tensor_map(&maps.c, d->C, 1, 2, ncols, d->M, 1, d->ldc, 0, 16, 32, 1, 64)) return -3;
    if (d->out1 && !make_tensor_map(&maps.o1, d->out1, 1, 2, ncols, d->M, 1, d->ld_out1, 0, 16, 32, 1, 64)) return -3;
    if (d->planes_out &&
        !make_tensor_map(&maps.pl, d->planes_out, 0, 3, zr ? p.zr_cols_pad : p.Npad, d->M, d->planes_out_n, d->ld_planes_out,
                     d->planes_out_stride, 16, 32, 1, 32))
      return -3;
  }
  const size_t smem = (size_t)cfg.stages * cfg.stage_bytes + (size_t)BP_EPI_WARPS * 2 * cfg.stg_set + 1024;
  static const BpKernelFn kernels[6] = {gemm_bp_kernel<0>, gemm_bp_kernel<1>, gemm_bp_kernel<2>,
                                        gemm_bp_kernel<3>, gemm_bp_kernel<4>, gemm_bp_kernel<5>};
  const int epi = cfg.splits > 1 ? 0 : p.epilogue;
  BpKernelFn fn = kernels[epi];
  static int max_dyn[6] = {-1, -1, -1, -1, -1, -1};
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (max_dyn[epi] < 0) {
      cudaFuncAttributes fa;
      cudaError_t e = cudaFuncGetAttributes(&fa, fn);
      if (e == cudaSuccess) {
        const int want = 227 * 1024 - (int)((fa.sharedSizeBytes + 1023) / 1024 * 1024);
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
        if (e == cudaSuccess) max_dyn[epi] = want;
      }
      if (e != cudaSuccess) {
        set_error("gemm_bp_kernel: cannot opt in to large shared memory: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return -4;
      }
    }
  }
  GETB_REQUIRE((int)smem <= max_dyn[epi], "get_gemm_bp: %zu bytes of shared memory exceed the limit %d", smem, max_dyn[epi]);
  const int grid = cfg.items < BP_SMS ? cfg.items : BP_SMS;
  fn<<<grid, BP_THREADS, smem, st>>>(p, cfg, maps);
  GETB_CHECK_LAUNCH("gemm_bp_kernel");
  return 0;
}

}  // namespace getb

using namespace getb;

extern "C" int get_gemm_bp(const get_gemm_bp_desc* desc, void* stream) { return bp_launch(desc, (cudaStream_t)stream); }

extern "C" int get_gemm_bp_tile_n(int M, int N, int mode) {
  if (M < 1 || N < 4 || mode < 1 || mode > 3) return -1;
  return bp_choose_bn(M, bp_round_up(N, 8), mode, 1);
}

extern "C" int64_t get_gemm_bp_ws_ld(const get_gemm_bp_desc* desc) {
  BpCfg cfg;
  BpParams p;
  get_gemm_bp_desc d = *desc;
  d.workspace = reinterpret_cast<float*>(16);      // planning only: any aligned non-null pointer
  d.workspace_floats = INT64_MAX;
  if (bp_plan(&d, cfg, p) != 0) return -1;
  return (int64_t)cfg.ntn * cfg.BN;
}

extern "C" int get_gemm_bp_splits(const get_gemm_bp_desc* desc) {
  BpCfg cfg;
  BpParams p;
  get_gemm_bp_desc d = *desc;
  d.workspace = reinterpret_cast<float*>(16);
  d.workspace_floats = INT64_MAX;
  if (bp_plan(&d, cfg, p) != 0) return -1;
  return cfg.splits;
}

extern "C" int get_gemm_bp_rowdot_parts(const get_gemm_bp_desc* desc) {
  BpCfg cfg;
  BpParams p;
  get_gemm_bp_desc d = *desc;
  d.workspace = reinterpret_cast<float*>(16);
  d.workspace_floats = INT64_MAX;
  d.rowdot_out = nullptr;
  if (bp_plan(&d, cfg, p) != 0) return -1;
  return 2 * cfg.ntn;
}

extern "C" int get_bp_splitk_reduce(const float* workspace, int splits, int M, int64_t ws_ld, const get_bp_dst* dsts, int ndst,
                                    int accumulate, void* stream) {
  GETB_REQUIRE(workspace && dsts && ndst >= 1 && ndst <= GET_BP_MAX_DST && splits >= 1, "get_bp_splitk_reduce: bad arguments");
  BpDstList L;
  memset(&L, 0, sizeof(L));
  int64_t total = 0, maxq = 0;
  for (int b = 0; b < ndst; ++b) {
    const get_bp_dst& d = dsts[b];
    GETB_REQUIRE(d.dst && d.nrows >= 0 && d.ncols >= 0 && d.row0 >= 0 && d.row0 + d.nrows <= M && d.col0 >= 0 && d.col0 + d.ncols <= ws_ld,
                 "get_bp_splitk_reduce: destination block %d out of range", b);
    L.d[b] = d;
    total += (int64_t)d.nrows * d.ncols;
    maxq = std::max<int64_t>(maxq, (int64_t)d.nrows * ((d.ncols + 3) / 4));
  }
  L.n = ndst;
  if (total == 0) return 0;
  int grid = (int)std::min<int64_t>(ceil_div(maxq, (int64_t)256), 4 * BP_SMS);   // every destination block is swept by the whole grid
  if (grid < 1) grid = 1;
  bp_splitk_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(workspace, splits, M, ws_ld, L, accumulate);
  GETB_CHECK_LAUNCH("get_bp_splitk_reduce");
  return 0;
}

extern "C" int get_to_planes_bf16(const float* src, int64_t ld_src, int rows, int cols, void* planes, int64_t ld_out,
                                  int64_t plane_stride, int nplanes, int pad_one, void* stream) {
  GETB_REQUIRE(src && planes && rows >= 0 && cols >= 1 && ld_out >= cols && (ld_out % 4) == 0 && nplanes >= 1 && nplanes <= 3,
               "get_to_planes_bf16: bad arguments");
  GETB_REQUIRE((((uintptr_t)planes) & 7u) == 0 && (plane_stride % 4) == 0, "get_to_planes_bf16: planes must be 8-byte aligned");
  if (rows == 0) return 0;
  const int vec = aligned16(src) && (ld_src % 4) == 0;
  int width = bp_round_up(cols + (pad_one ? 1 : 0), 8);   // written columns: logical ones (+ ones column) padded to a multiple of 8
  if (width > ld_out) width = (int)ld_out;
  const int64_t nq = (int64_t)rows * (width / 4);
  to_planes_kernel<<<ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows, cols, reinterpret_cast<__nv_bfloat16*>(planes),
                                                                       ld_out, plane_stride, nplanes, pad_one, vec, width);
  GETB_CHECK_LAUNCH("get_to_planes_bf16");
  return 0;
}

extern "C" int get_pack_planes_multi(const get_pack_job* jobs, int n_jobs, int64_t total_blocks, void* stream) {
  GETB_REQUIRE(jobs && n_jobs >= 1 && total_blocks >= 1 && total_blocks < 2147483647, "get_pack_planes_multi: bad arguments");
  pack_planes_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs, n_jobs);
  GETB_CHECK_LAUNCH("get_pack_planes_multi");
  return 0;
}
