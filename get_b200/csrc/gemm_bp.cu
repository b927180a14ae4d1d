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
constexpr int BP_STG_LD = 36;                          // floats per row of the epilogue staging tile (32 + 4: conflict-free)
constexpr int BP_STG_BYTES = BP_EPI_WARPS * 32 * BP_STG_LD * 4;
constexpr int BP_SMS = 148;

struct BpMaps {
  CUtensorMap a[GET_GEMM_MAX_SEG];
  CUtensorMap b[GET_GEMM_MAX_SEG];
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
};

struct BpParams {
  int M, N, Npad;
  int epilogue, accumulate;
  float* C; int64_t ldc;
  float* out1; int64_t ld_out1;
  const float* bias;
  const float* aux0; int64_t ld_aux0;
  const float* aux1; int64_t ld_aux1;
  __nv_bfloat16* planes; int64_t ld_p, plane_stride;
  int nplanes, pad_one;
  int group_rows, zr_gs, zr_cols, zr_cols_pad;
  uint32_t drop_thr, drop_seed; float drop_scale;
  const uint32_t* salt;
  float* workspace; int64_t ws_ld;
};

// ---- fused epilogues on 4 consecutive columns -----------------------------------------------------------------------
struct BpEpiIn {
  float a0[4], a1[4], o[4];
};

__device__ __forceinline__ void bp_ld4(const float* p, float v[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void bp_st4(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// classification of a quad starting at tile-global column n: 0 = nothing to do, 1 = real columns, 2 = plane padding only
template <int EPI>
__device__ __forceinline__ int bp_quad_kind(const BpParams& p, int n, int& c, int& grp) {
  if (EPI == GET_BPE_ZR) {
    grp = n / p.zr_gs;
    c = n - grp * p.zr_gs;
    if (grp > 1) return 0;
    if (c < p.zr_cols) return 1;
    return (grp == 1 && c < p.zr_cols_pad) ? 2 : 0;
  }
  c = n; grp = 0;
  if (n < p.N) return 1;
  return (p.planes && n < p.Npad) ? 2 : 0;
}

template <int EPI>
__device__ __forceinline__ void bp_epi_load(const BpParams& p, int m, int c, int grp, BpEpiIn& in) {
  switch (EPI) {
    case GET_BPE_STORE:
      if (p.accumulate) bp_ld4(p.C + (int64_t)m * p.ldc + c, in.o);
      break;
    case GET_BPE_ZR:
      if (grp == 1) bp_ld4(p.aux0 + (int64_t)m * p.ld_aux0 + c, in.a0);          // x
      break;
    case GET_BPE_TANH_BLEND:
      bp_ld4(p.aux0 + (int64_t)m * p.ld_aux0 + c, in.a0);                          // z
      bp_ld4(p.aux1 + (int64_t)m * p.ld_aux1 + c, in.a1);                          // x
      break;
    case GET_BPE_TANH_ROWGROUP:
      bp_ld4(p.aux0 + (int64_t)(m / p.group_rows) * p.ld_aux0 + c, in.a0);
      break;
    case GET_BPE_DGATE_R:
      bp_ld4(p.aux0 + (int64_t)m * p.ld_aux0 + c, in.a0);                          // x
      bp_ld4(p.aux1 + (int64_t)m * p.ld_aux1 + c, in.a1);                          // r
      bp_ld4(p.out1 + (int64_t)m * p.ld_out1 + c, in.o);                           // dx
      break;
    default: break;
  }
}

template <int EPI>
__device__ __forceinline__ void bp_epi_apply(const BpParams& p, int m, int n, int c, int grp, const float acc[4], const BpEpiIn& in) {
  float v[4], o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = acc[e];
  if (p.bias) {
    float b[4];
    bp_ld4(p.bias + n, b);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] += b[e];
  }
  __nv_bfloat16* prow = p.planes ? p.planes + (int64_t)m * p.ld_p + c : nullptr;
  switch (EPI) {
    case GET_BPE_STORE: {
      if (p.drop_thr) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool keep = drop_keep(p.drop_seed + __ldg(p.salt), (uint64_t)m * (uint64_t)p.N + (uint64_t)(n + e), p.drop_thr);
          v[e] = keep ? v[e] * p.drop_scale : 0.f;
        }
      }
      if (p.accumulate) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] += in.o[e];
      }
      if (p.C) bp_st4(p.C + (int64_t)m * p.ldc + c, v);
      if (prow) planes_store4(prow, p.plane_stride, p.nplanes, v);
    } break;
    case GET_BPE_ZR: {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = sigmoid_fast(v[e]);
      if (grp == 0) {
        bp_st4(p.C + (int64_t)m * p.ldc + c, v);
      } else {
        bp_st4(p.out1 + (int64_t)m * p.ld_out1 + c, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = v[e] * in.a0[e];
        if (prow) planes_store4(prow, p.plane_stride, p.nplanes, o);
      }
    } break;
    case GET_BPE_TANH_BLEND: {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = tanh_fast(v[e]);
        o[e] = v[e] * in.a0[e] + in.a1[e] * (1.0f - in.a0[e]);
      }
      if (p.out1) bp_st4(p.out1 + (int64_t)m * p.ld_out1 + c, v);
      if (p.C) bp_st4(p.C + (int64_t)m * p.ldc + c, o);
      if (prow) planes_store4(prow, p.plane_stride, p.nplanes, o);
    } break;
    case GET_BPE_TANH_ROWGROUP: {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = tanh_fast(v[e] + in.a0[e]);
      bp_st4(p.C + (int64_t)m * p.ldc + c, v);
      if (prow) planes_store4(prow, p.plane_stride, p.nplanes, v);
    } break;
    case GET_BPE_DGATE_R: {
      float g[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        g[e] = v[e] * in.a0[e] * in.a1[e] * (1.0f - in.a1[e]);
        o[e] = in.o[e] + v[e] * in.a1[e];
      }
      if (p.C) bp_st4(p.C + (int64_t)m * p.ldc + c, g);
      if (prow) planes_store4(prow, p.plane_stride, p.nplanes, g);
      bp_st4(p.out1 + (int64_t)m * p.ld_out1 + c, o);
    } break;
    case GET_BPE_TANH: {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = tanh_fast(v[e]);
      bp_st4(p.C + (int64_t)m * p.ldc + c, v);
      if (prow) planes_store4(prow, p.plane_stride, p.nplanes, v);
    } break;
    default: break;
  }
}

// padding columns of the output planes: zeros, 1.0 in the first pad column when pad_one
__device__ __forceinline__ void bp_epi_pad(const BpParams& p, int m, int c, int first_pad) {
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = (p.pad_one && c + e == first_pad) ? 1.0f : 0.0f;
  planes_store4(p.planes + (int64_t)m * p.ld_p + c, p.plane_stride, p.nplanes, v);
}

__device__ __forceinline__ void bp_locate(const BpCfg& cfg, int kb, int& seg, int& kin) {
  seg = 0;
  while (seg + 1 < cfg.nseg && kb >= cfg.kblocks[seg]) { kb -= cfg.kblocks[seg]; ++seg; }
  kin = kb * cfg.kb;
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

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      for (int s = 0; s < cfg.nseg; ++s) { prefetch_tmap(&maps.a[s]); prefetch_tmap(&maps.b[s]); }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = cfg.a_bytes + cfg.b_bytes;
      for (int it = 0; it < n_my; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int nt = item % cfg.ntn, mt = (item / cfg.ntn) % cfg.ntm, z = item / (cfg.ntn * cfg.ntm);
        const int m0 = mt * BP_BM, n0 = nt * BN;
        const int kb0 = z * cfg.kb_per_split, kb1 = min(cfg.kblocks_total, kb0 + cfg.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          int seg, kin;
          bp_locate(cfg, kb, seg, kin);
          mbar_wait(&bar_empty[stage], phase ^ 1);
          const uint32_t sa = smem_base + (uint32_t)stage * cfg.stage_bytes;
          const uint32_t sb = sa + cfg.a_bytes;
          mbar_arrive_expect_tx(&bar_full[stage], tx);
          if (cfg.a_mn) {
            for (int b = 0; b < cfg.a_boxes; ++b) tma_load_3d(sa + (uint32_t)b * cfg.a_box_bytes, &maps.a[seg], m0 + b * 64, kin, 0, &bar_full[stage]);
          } else {
            tma_load_3d(sa, &maps.a[seg], kin, m0, 0, &bar_full[stage]);
          }
          if (cfg.b_mn) {
            for (int b = 0; b < cfg.b_boxes; ++b) tma_load_3d(sb + (uint32_t)b * cfg.b_box_bytes, &maps.b[seg], n0 + b * 64, kin, 0, &bar_full[stage]);
          } else {
            tma_load_3d(sb, &maps.b[seg], kin, n0, 0, &bar_full[stage]);
          }
          if (++stage == cfg.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================================== MMA issuer =============================================
    if (lane == 0) {
      // kind::f16 instruction descriptor: D = f32, A = B = bf16, majors, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(cfg.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(cfg.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BP_BM >> 4) << 24);
      const int ksteps = cfg.kb / 16;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int it = 0; it < n_my; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int z = item / (cfg.ntn * cfg.ntm);
        const int kb0 = z * cfg.kb_per_split, kb1 = min(cfg.kblocks_total, kb0 + cfg.kb_per_split);
        mbar_wait(&bar_acce[acc], acc_phase ^ 1);
        fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(acc * cfg.acc_cols);
        const uint32_t d_small = d_main + (uint32_t)BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bar_full[stage], phase);
          fence_after();
          const uint32_t sa = smem_base + (uint32_t)stage * cfg.stage_bytes;
          const uint32_t sb = sa + cfg.a_bytes;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t a0 = sa + (uint32_t)ks * cfg.a_kstep, b0 = sb + (uint32_t)ks * cfg.b_kstep;
            const uint64_t da0 = smem_desc(a0, cfg.a_lbo, cfg.a_sbo, cfg.a_lt);
            const uint64_t db0 = smem_desc(b0, cfg.b_lbo, cfg.b_sbo, cfg.b_lt);
            const uint32_t first = (kb > kb0 || ks > 0) ? 1u : 0u;
            umma_bf16(d_main, da0, db0, idesc, first);
            if (cfg.mode >= 2) {
              const uint64_t da1 = smem_desc(a0 + cfg.a_plane, cfg.a_lbo, cfg.a_sbo, cfg.a_lt);
              const uint64_t db1 = smem_desc(b0 + cfg.b_plane, cfg.b_lbo, cfg.b_sbo, cfg.b_lt);
              if (cfg.mode == 2) {
                umma_bf16(d_main, da0, db1, idesc, 1u);
                umma_bf16(d_main, da1, db0, idesc, 1u);
              } else {
                const uint64_t da2 = smem_desc(a0 + 2u * cfg.a_plane, cfg.a_lbo, cfg.a_sbo, cfg.a_lt);
                const uint64_t db2 = smem_desc(b0 + 2u * cfg.b_plane, cfg.b_lbo, cfg.b_sbo, cfg.b_lt);
                umma_bf16(d_small, da0, db1, idesc, first);
                umma_bf16(d_small, da1, db0, idesc, 1u);
                umma_bf16(d_small, da1, db1, idesc, 1u);
                umma_bf16(d_small, da0, db2, idesc, 1u);
                umma_bf16(d_small, da2, db0, idesc, 1u);
              }
            }
          }
          umma_commit(&bar_empty[stage]);
          if (++stage == cfg.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bar_accf[acc]);
        if (cfg.acc_bufs == 2) {
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
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int chalf = (warp - 2) >> 2;            // two warps per quarter: even / odd 32-column chunks
    const uint32_t stg = smem_base + (uint32_t)cfg.stages * cfg.stage_bytes + (uint32_t)(warp - 2) * (32u * BP_STG_LD * 4u);
    const int rrow = lane >> 3, rq = lane & 7;
    const bool small = cfg.mode == 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = 0; it < n_my; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int nt = item % cfg.ntn, mt = (item / cfg.ntn) % cfg.ntm, z = item / (cfg.ntn * cfg.ntm);
      const int m_base = mt * BP_BM + quarter * 32;
      const int n0 = nt * BN;
      mbar_wait(&bar_accf[acc], acc_phase);
      fence_after();
      const uint32_t t_main = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * cfg.acc_cols);
      const uint32_t t_small = t_main + (uint32_t)BN;
      const int nch = (BN + 31) / 32;
      const int last_col = (nch - 1 >= chalf) ? ((nch - 1 - chalf) / 2 * 2 + chalf) * 32 : -1;
      if (last_col < 0) {                 // single-chunk tiles: the odd warps have nothing to read
        fence_before();
        mbar_arrive(&bar_acce[acc]);
      }
      for (int col = chalf * 32; col < BN; col += 64) {
        const bool two = col + 16 < BN;
        uint32_t rm[16], rs[16], rm2[16], rs2[16];
        tmem_ld16_nowait(t_main + (uint32_t)col, rm);
        if (small) tmem_ld16_nowait(t_small + (uint32_t)col, rs);
        if (two) {
          tmem_ld16_nowait(t_main + (uint32_t)col + 16u, rm2);
          if (small) tmem_ld16_nowait(t_small + (uint32_t)col + 16u, rs2);
        }
        tmem_ld_wait();
        if (!small) {
#pragma unroll
          for (int e = 0; e < 16; ++e) { rs[e] = 0u; rs2[e] = 0u; }
        }
        if (col == last_col) {           // last chunk of this accumulator set for this warp: hand it back to the MMA warp
          fence_before();
          mbar_arrive(&bar_acce[acc]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = __float_as_uint(__uint_as_float(rm[q * 4 + 0]) + __uint_as_float(rs[q * 4 + 0]));
          v.y = __float_as_uint(__uint_as_float(rm[q * 4 + 1]) + __uint_as_float(rs[q * 4 + 1]));
          v.z = __float_as_uint(__uint_as_float(rm[q * 4 + 2]) + __uint_as_float(rs[q * 4 + 2]));
          v.w = __float_as_uint(__uint_as_float(rm[q * 4 + 3]) + __uint_as_float(rs[q * 4 + 3]));
          sts128(stg + (uint32_t)(lane * BP_STG_LD + q * 4) * 4u, v);
        }
        if (two) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = __float_as_uint(__uint_as_float(rm2[q * 4 + 0]) + __uint_as_float(rs2[q * 4 + 0]));
            v.y = __float_as_uint(__uint_as_float(rm2[q * 4 + 1]) + __uint_as_float(rs2[q * 4 + 1]));
            v.z = __float_as_uint(__uint_as_float(rm2[q * 4 + 2]) + __uint_as_float(rs2[q * 4 + 2]));
            v.w = __float_as_uint(__uint_as_float(rm2[q * 4 + 3]) + __uint_as_float(rs2[q * 4 + 3]));
            sts128(stg + (uint32_t)(lane * BP_STG_LD + 16 + q * 4) * 4u, v);
          }
        }
        __syncwarp();
        const int n = n0 + col + rq * 4;
        if (two || rq < 4) {
          if (cfg.splits > 1) {
            if (n < (int)p.ws_ld) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = i * 4 + rrow;
                const int m = m_base + row;
                if (m < p.M) {
                  const float4 f = lds128(stg + (uint32_t)(row * BP_STG_LD + rq * 4) * 4u);
                  *reinterpret_cast<float4*>(p.workspace + ((int64_t)z * p.M + m) * p.ws_ld + n) = f;
                }
              }
            }
          } else {
            int c, grp;
            const int kind = bp_quad_kind<EPI>(p, n, c, grp);
            if (kind == 1) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {          // two batches of four rows: all loads first, then math + stores
                BpEpiIn in[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int m = m_base + (h * 4 + i) * 4 + rrow;
                  if (m < p.M) bp_epi_load<EPI>(p, m, c, grp, in[i]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int row = (h * 4 + i) * 4 + rrow;
                  const int m = m_base + row;
                  if (m < p.M) {
                    const float4 f = lds128(stg + (uint32_t)(row * BP_STG_LD + rq * 4) * 4u);
                    const float v[4] = {f.x, f.y, f.z, f.w};
                    bp_epi_apply<EPI>(p, m, n, c, grp, v, in[i]);
                  }
                }
              }
            } else if (kind == 2) {
              const int first_pad = (EPI == GET_BPE_ZR) ? p.zr_cols : p.N;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = m_base + i * 4 + rrow;
                if (m < p.M) bp_epi_pad(p, m, c, first_pad);
              }
            }
          }
        }
        __syncwarp();
      }
      if (cfg.acc_bufs == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
  }
  fence_before();
  __syncthreads();
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
  // block (x over columns, y over destination blocks x row chunks)
  for (int b = 0; b < L.n; ++b) {
    const get_bp_dst& d = L.d[b];
    const int64_t total = (int64_t)d.nrows * d.ncols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(e / d.ncols), c = (int)(e % d.ncols);
      const float* src = ws + (int64_t)(d.row0 + r) * ws_ld + (d.col0 + c);
      float acc = 0.f;
      for (int z = 0; z < splits; ++z) acc += src[(int64_t)z * M * ws_ld];
      float* o = d.dst + (int64_t)r * d.ld + c;
      *o = accumulate ? *o + acc : acc;
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
    bp_ld4(src + (int64_t)r * ld_src + c, v);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (c + e < cols) ? src[(int64_t)r * ld_src + c + e] : ((pad_one && c + e == cols) ? 1.0f : 0.0f);
  }
  planes_store4(dst + (int64_t)r * ld_out + c, plane_stride, nplanes, v);
}

// ---- weight packing: many (strided) matrices -> 3 planes each, plus fused bias vectors, one launch -------------------
__global__ void __launch_bounds__(256) pack_planes_multi_kernel(const get_pack_job* __restrict__ jobs, int n_jobs) {
  // binary search for the job covering this block
  int lo = 0, hi = n_jobs - 1;
  const int64_t blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const get_pack_job j = jobs[lo];
  const int64_t e = (blk - j.first_block) * 256 + threadIdx.x;
  if (j.kind == 1) {
    if (e < j.rows) reinterpret_cast<float*>(j.dst)[e] = j.src[e] + (j.src2 ? j.src2[e] : 0.f);
    return;
  }
  if (e >= (int64_t)j.rows * j.cols) return;
  const int r = (int)(e / j.cols), c = (int)(e % j.cols);
  float v = j.src[(int64_t)r * j.ld_r + (int64_t)c * j.ld_c];
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(j.dst) + (int64_t)r * j.ld_out + c;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const __nv_bfloat16 q = __float2bfloat16_rn(v);
    d[(int64_t)p * j.plane_stride] = q;
    v -= __bfloat162float(q);
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
  int b0, b1, b2, sw;
  bool operator==(const BpMapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld == o.ld && ps == o.ps && b0 == o.b0 && b1 == o.b1 &&
           b2 == o.b2 && sw == o.sw;
  }
};
struct BpMapKeyHash {
  size_t operator()(const BpMapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)k.d0 * 0xC2B2AE3D27D4EB4Full + (uint64_t)k.d1 * 0x165667B19E3779F9ull + (uint64_t)k.d2 * 0x27D4EB2F165667C5ull;
    h ^= (uint64_t)k.ld * 31 + (uint64_t)k.ps * 131 + (uint64_t)k.b0 * 7 + (uint64_t)k.b1 * 1315423911ull + (uint64_t)k.b2 * 2654435761ull +
         (uint64_t)k.sw * 97;
    return (size_t)(h ^ (h >> 29));
  }
};

// 3-D bf16 tensor map {d0 inner contiguous, d1 rows `ld` apart, d2 planes `ps` apart}, box {b0, b1, b2}; sw: 64 or 128
static bool bp_make_map(CUtensorMap* map, const void* ptr, int64_t d0, int64_t d1, int64_t d2, int64_t ld, int64_t ps, int b0,
                        int b1, int b2, int sw) {
  static std::mutex mu;
  static std::unordered_map<BpMapKey, CUtensorMap, BpMapKeyHash> cache;
  const BpMapKey key{ptr, d0, d1, d2, ld, ps, b0, b1, b2, sw};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *map = it->second;
    return true;
  }
  EncodeTiledFnBp enc = bp_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2u, (cuuint64_t)(d2 > 1 ? ps : ld * d1) * 2u};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("get_gemm_bp: cuTensorMapEncodeTiled failed (%d): dims %lld %lld %lld ld %lld ps %lld box %d %d %d", (int)rc,
              (long long)d0, (long long)d1, (long long)d2, (long long)ld, (long long)ps, b0, b1, b2);
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
    if ((224 * 1024 - BP_STG_BYTES - 2048) / (int)stage < 3) continue;
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
  p.Npad = d->planes_out ? bp_round_up(d->N, 8) : d->N;
  p.epilogue = d->epilogue; p.accumulate = d->accumulate;
  p.C = d->C; p.ldc = d->ldc; p.out1 = d->out1; p.ld_out1 = d->ld_out1;
  p.bias = d->bias; p.aux0 = d->aux0; p.ld_aux0 = d->ld_aux0; p.aux1 = d->aux1; p.ld_aux1 = d->ld_aux1;
  p.planes = reinterpret_cast<__nv_bfloat16*>(d->planes_out);
  p.ld_p = d->ld_planes_out; p.plane_stride = d->planes_out_stride;
  p.nplanes = d->planes_out_n; p.pad_one = d->planes_out_pad_one;
  p.group_rows = d->group_rows; p.zr_gs = d->zr_group_stride; p.zr_cols = d->zr_cols;
  p.zr_cols_pad = bp_round_up(d->zr_cols, 8);
  p.salt = dropout_salt_ptr();
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
  if (p.planes) {
    GETB_REQUIRE((((uintptr_t)p.planes) & 7u) == 0 && (p.ld_p % 4) == 0 && (p.plane_stride % 4) == 0 && p.nplanes >= 1 && p.nplanes <= 3,
                 "get_gemm_bp: planes_out must be 8-byte aligned, ld %% 4 == 0, 1..3 planes");
  }
  switch (d->epilogue) {
    case GET_BPE_STORE: GETB_REQUIRE(p.C || p.planes || d->split_k > 1, "get_gemm_bp: STORE needs C or planes_out"); GETB_REQUIRE(!p.accumulate || p.C, "get_gemm_bp: accumulate needs C"); break;
    case GET_BPE_ZR:
      GETB_REQUIRE(p.C && p.out1 && p.zr_gs > 0 && p.zr_cols > 0 && (p.zr_cols % 4) == 0 && p.zr_cols <= p.zr_gs && (!p.planes || p.aux0),
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
    p.ws_ld = (int64_t)cfg.ntn * bn;
    p.workspace = d->workspace;
    GETB_REQUIRE(d->workspace && d->workspace_floats >= (int64_t)cfg.splits * d->M * p.ws_ld,
                 "get_gemm_bp: split-K workspace too small (%lld floats needed)", (long long)((int64_t)cfg.splits * d->M * p.ws_ld));
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
  int stages = (225 * 1024 - BP_STG_BYTES - 1024) / (int)cfg.stage_bytes;
  if (stages > BP_MAX_STAGES) stages = BP_MAX_STAGES;
  const int cap = bp_env_int("GET_B200_BP_STAGES", 0);
  if (cap > 0 && stages > cap) stages = cap;
  GETB_REQUIRE(stages >= 2, "get_gemm_bp: tile (mode %d, tile_n %d, kblock %d) does not fit shared memory", d->mode, bn, kb);
  cfg.stages = stages;
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
    if (cfg.a_mn) ok = bp_make_map(&maps.a[s], A.ptr, d->M, d->K[s], np, A.ld, A.plane_stride, 64, kb, np, 128);
    else ok = bp_make_map(&maps.a[s], A.ptr, d->K[s], d->M, np, A.ld, A.plane_stride, kb, BP_BM, np, kb == 64 ? 128 : 64);
    if (!ok) return -3;
    if (cfg.b_mn) ok = bp_make_map(&maps.b[s], B.ptr, d->N, d->K[s], np, B.ld, B.plane_stride, 64, kb, np, 128);
    else ok = bp_make_map(&maps.b[s], B.ptr, d->K[s], d->N, np, B.ld, B.plane_stride, kb, cfg.BN, np, kb == 64 ? 128 : 64);
    if (!ok) return -3;
  }
  const size_t smem = (size_t)cfg.stages * cfg.stage_bytes + BP_STG_BYTES + 1024;
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

extern "C" int get_bp_splitk_reduce(const float* workspace, int splits, int M, int64_t ws_ld, const get_bp_dst* dsts, int ndst,
                                    int accumulate, void* stream) {
  GETB_REQUIRE(workspace && dsts && ndst >= 1 && ndst <= GET_BP_MAX_DST && splits >= 1, "get_bp_splitk_reduce: bad arguments");
  BpDstList L;
  memset(&L, 0, sizeof(L));
  int64_t total = 0;
  for (int b = 0; b < ndst; ++b) {
    const get_bp_dst& d = dsts[b];
    GETB_REQUIRE(d.dst && d.nrows >= 0 && d.ncols >= 0 && d.row0 >= 0 && d.row0 + d.nrows <= M && d.col0 >= 0 && d.col0 + d.ncols <= ws_ld,
                 "get_bp_splitk_reduce: destination block %d out of range", b);
    L.d[b] = d;
    total += (int64_t)d.nrows * d.ncols;
  }
  L.n = ndst;
  if (total == 0) return 0;
  int grid = ceil_div(total / ndst + 1, 256);
  if (grid > 4 * BP_SMS) grid = 4 * BP_SMS;
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
  int width = bp_round_up(cols, 8);          // written columns: the logical ones plus padding up to a multiple of 8
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
