// Element-wise / segment kernels of the GET hot path and the library bookkeeping.
#include <stdarg.h>
#include <atomic>
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace getb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

const uint32_t* dropout_salt_ptr() {
  // one word per DEVICE (normally one process per GPU, but the device current at library-load time need not be the one the
  // process computes on); allocated on first use on that device, never freed
  static uint32_t* salt[64] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    (void)cudaGetLastError();
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (!salt[dev]) {
    uint32_t* ptr = nullptr;
    if (cudaMalloc(&ptr, 256) == cudaSuccess && cudaMemset(ptr, 0, 256) == cudaSuccess) salt[dev] = ptr;
    else (void)cudaGetLastError();
  }
  return salt[dev];
}

// ---- GGNN backward, element-wise stage (reference forward Models/BiDAF/wrapper.py:194-206)
__global__ void __launch_bounds__(256) ggnn_gate_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ z,
                                                            const float* __restrict__ h, const float* __restrict__ x,
                                                            int M, int H, int64_t ld_g, float* __restrict__ dhp,
                                                            float* __restrict__ dzp, float* __restrict__ dx, int vec) {
  // inputs and dx are (M, H) contiguous; dhp / dzp are (M, H) views with row stride ld_g (columns of the [dz'|dr'|dh'] buffer)
  const int hq = (H + 3) >> 2;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (int64_t)M * hq) return;
  const int m = (int)(q / hq), c = (int)(q % hq) * 4;
  const int64_t i = (int64_t)m * H + c, o = (int64_t)m * ld_g + c;
  float d[4], zz[4], hh[4], xx[4], o0[4], o1[4], o2[4];
  const int nvalid = min(4, H - c);
  if (vec) {
    float4 a = *reinterpret_cast<const float4*>(dout + i); d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
    a = *reinterpret_cast<const float4*>(z + i); zz[0] = a.x; zz[1] = a.y; zz[2] = a.z; zz[3] = a.w;
    a = *reinterpret_cast<const float4*>(h + i); hh[0] = a.x; hh[1] = a.y; hh[2] = a.z; hh[3] = a.w;
    a = *reinterpret_cast<const float4*>(x + i); xx[0] = a.x; xx[1] = a.y; xx[2] = a.z; xx[3] = a.w;
  } else {
    for (int e = 0; e < 4; ++e) {
      const bool ok = e < nvalid;
      d[e] = ok ? dout[i + e] : 0.f; zz[e] = ok ? z[i + e] : 0.f; hh[e] = ok ? h[i + e] : 0.f; xx[e] = ok ? x[i + e] : 0.f;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    o0[e] = d[e] * zz[e] * (1.0f - hh[e] * hh[e]);
    o1[e] = d[e] * (hh[e] - xx[e]) * zz[e] * (1.0f - zz[e]);
    o2[e] = d[e] * (1.0f - zz[e]);
  }
  if (vec) {
    *reinterpret_cast<float4*>(dhp + o) = make_float4(o0[0], o0[1], o0[2], o0[3]);
    *reinterpret_cast<float4*>(dzp + o) = make_float4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<float4*>(dx + i) = make_float4(o2[0], o2[1], o2[2], o2[3]);
  } else {
    for (int e = 0; e < nvalid; ++e) { dhp[o + e] = o0[e]; dzp[o + e] = o1[e]; dx[i + e] = o2[e]; }
  }
}

// ---- column sums (bias gradients): stage 1 = per-chunk partial sums, stage 2 = fixed-order reduction
constexpr int COLSUM_ROWS = 256;  // rows per stage-1 block
__global__ void __launch_bounds__(256) colsum_stage1_kernel(const float* __restrict__ a, int64_t ld, int M, int N,
                                                            float* __restrict__ part) {
  // block (bx, by): columns bx*32..+31, rows by*COLSUM_ROWS..; 8 warps each stride the rows, lanes = columns
  __shared__ float s[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * COLSUM_ROWS;
  const int r1 = min(M, r0 + COLSUM_ROWS);
  float acc = 0.f;
  if (n < N)
    for (int r = r0 + warp; r < r1; r += 8) acc += __ldg(a + (int64_t)r * ld + n);
  s[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && n < N) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s[w][lane];
    part[(int64_t)blockIdx.y * N + n] = v;
  }
}
__global__ void __launch_bounds__(256) colsum_stage2_kernel(const float* __restrict__ part, int nparts, int N,
                                                            float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float v = 0.f;
  for (int q = 0; q < nparts; ++q) v += part[(int64_t)q * N + n];
  out[n] = v;
}

// ---- row gather / scatter / segment sum (reference basic_fc_model.py:80-121 without the Python loops)
__global__ void __launch_bounds__(256) rows_copy_kernel(const float* __restrict__ src, int64_t ld_src,
                                                        const int32_t* __restrict__ idx, int R, int W,
                                                        float* __restrict__ out, int64_t ld_out, int scatter) {
  const int r = blockIdx.x;
  const int64_t srow = scatter ? r : idx[r];
  const int64_t drow = scatter ? idx[r] : r;
  const float* s = src + srow * ld_src;
  float* d = out + drow * ld_out;
  for (int c = threadIdx.x; c < W; c += blockDim.x) d[c] = s[c];
}

__global__ void __launch_bounds__(256) segment_sum_kernel(const float* __restrict__ src, int64_t ld_src,
                                                          const int32_t* __restrict__ offsets, int W,
                                                          float* __restrict__ out, int64_t ld_out) {
  const int sgm = blockIdx.x;
  const int r0 = offsets[sgm], r1 = offsets[sgm + 1];
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float v = 0.f;
    for (int r = r0; r < r1; ++r) v += src[(int64_t)r * ld_src + c];
    out[(int64_t)sgm * ld_out + c] = v;
  }
}

// ---- claim read-out: masked mean over nodes (reference graph_based_semantic_structure.py:145-153)
__global__ void __launch_bounds__(256) masked_mean_fwd_kernel(const float* __restrict__ h, const int64_t* __restrict__ ids,
                                                              const int64_t* __restrict__ lens, int N, int H,
                                                              float* __restrict__ out) {
  const int g = blockIdx.x;
  const float len = (float)lens[g];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v = 0.f;
    for (int i = 0; i < N; ++i)
      if (ids[(int64_t)g * N + i] > 0) v += h[((int64_t)g * N + i) * H + c];
    out[(int64_t)g * H + c] = v / len;
  }
}
__global__ void __launch_bounds__(256) masked_mean_bwd_kernel(const float* __restrict__ dout, const int64_t* __restrict__ ids,
                                                              const int64_t* __restrict__ lens, int N, int H,
                                                              float* __restrict__ dh) {
  const int g = blockIdx.x / N, i = blockIdx.x % N;
  const float len = (float)lens[g];
  const bool on = ids[(int64_t)g * N + i] > 0;
  for (int c = threadIdx.x; c < H; c += blockDim.x)
    dh[((int64_t)g * N + i) * H + c] = on ? dout[(int64_t)g * H + c] / len : 0.f;
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(float* __restrict__ out, int64_t numel, uint32_t thr,
                                                           float scale, uint32_t seed, const uint32_t* __restrict__ salt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < numel) out[i] = drop_keep(seed + __ldg(salt), (uint64_t)i, thr) ? scale : 0.f;
}

__global__ void salt_kernel(uint32_t* salt, uint32_t value, int advance) {
  *salt = advance ? (*salt) * 1664525u + 1013904223u : value;
}

// out[r, :] = dropout(src[idx ? idx[r] : r, :]), mask index r*W + c (the A operand of the input projection of a GGNN:
// embedding gather gbss.py:100,150 + nn.Dropout wrapper.py:189-190, materialised once for forward AND weight gradient)
__global__ void __launch_bounds__(256) rows_gather_dropout_kernel(const float* __restrict__ src, int64_t ld_src,
                                                                  const int64_t* __restrict__ idx, int R, int W,
                                                                  uint32_t thr, float scale, uint32_t seed,
                                                                  const uint32_t* __restrict__ salt,
                                                                  float* __restrict__ out, int64_t ld_out, int vec) {
  const uint32_t sd = seed + (thr ? __ldg(salt) : 0u);
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const float* s = src + (idx ? idx[r] : (int64_t)r) * ld_src;
    float* d = out + (int64_t)r * ld_out;
    if (vec) {
      for (int q = threadIdx.x; q < (W >> 2); q += blockDim.x) {
        float4 f = __ldg(reinterpret_cast<const float4*>(s) + q);
        if (thr) drop_apply4(sd, (uint64_t)r * (uint64_t)W + (uint64_t)q * 4, thr, scale, f);
        reinterpret_cast<float4*>(d)[q] = f;
      }
    } else {
      for (int c = threadIdx.x; c < W; c += blockDim.x) {
        float f = __ldg(s + c);
        if (thr) f = drop_keep(sd, (uint64_t)r * (uint64_t)W + c, thr) ? f * scale : 0.f;
        d[c] = f;
      }
    }
  }
}

// ---- plane-emitting variants (operands of the bf16-plane tensor-core GEMM, see get_gemm_bp) ----------------------------
// GGNN backward element-wise stage writing the gate gradients as bf16 planes: dz' -> column block col_z, dh' -> col_h of
// the (M, ld_g) gate-gradient plane buffer [dz' | dr' | dh'] (pad columns up to a multiple of 8 written as zeros);
// dx (fp32) = dout * (1 - z).
__global__ void __launch_bounds__(256) ggnn_gate_bwd_bp_kernel(const float* __restrict__ dout, const float* __restrict__ z,
                                                               const float* __restrict__ h, const float* __restrict__ x,
                                                               int M, int H, __nv_bfloat16* __restrict__ dg, int64_t ld_g,
                                                               int64_t plane_stride, int nplanes, int col_z, int col_h,
                                                               float* __restrict__ dx) {
  const int hq = ((H + 7) & ~7) >> 2;        // quads per row including the padding quad(s)
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (int64_t)M * hq) return;
  const int m = (int)(q / hq), c = (int)(q % hq) * 4;
  float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < H) {
    const int64_t i = (int64_t)m * H + c;
    const float4 d = *reinterpret_cast<const float4*>(dout + i), zz = *reinterpret_cast<const float4*>(z + i);
    const float4 hh = *reinterpret_cast<const float4*>(h + i), xx = *reinterpret_cast<const float4*>(x + i);
    o0[0] = d.x * zz.x * (1.0f - hh.x * hh.x); o0[1] = d.y * zz.y * (1.0f - hh.y * hh.y);
    o0[2] = d.z * zz.z * (1.0f - hh.z * hh.z); o0[3] = d.w * zz.w * (1.0f - hh.w * hh.w);
    o1[0] = d.x * (hh.x - xx.x) * zz.x * (1.0f - zz.x); o1[1] = d.y * (hh.y - xx.y) * zz.y * (1.0f - zz.y);
    o1[2] = d.z * (hh.z - xx.z) * zz.z * (1.0f - zz.z); o1[3] = d.w * (hh.w - xx.w) * zz.w * (1.0f - zz.w);
    *reinterpret_cast<float4*>(dx + i) = make_float4(d.x * (1.0f - zz.x), d.y * (1.0f - zz.y), d.z * (1.0f - zz.z), d.w * (1.0f - zz.w));
  }
  planes_store4(dg + (int64_t)m * ld_g + col_h + c, plane_stride, nplanes, o0);
  planes_store4(dg + (int64_t)m * ld_g + col_z + c, plane_stride, nplanes, o1);
}

// out planes[r, :] = dropout(src[idx ? idx[r] : r, :]) (embedding gather gbss.py:100,150 + nn.Dropout wrapper.py:189-190),
// mask index r*W + c; pad columns W..round_up(W,8)-1 written as zeros
__global__ void __launch_bounds__(128) rows_gather_dropout_bp_kernel(const float* __restrict__ src, int64_t ld_src,
                                                                     const int64_t* __restrict__ idx, int R, int W,
                                                                     uint32_t thr, float scale, uint32_t seed,
                                                                     const uint32_t* __restrict__ salt,
                                                                     __nv_bfloat16* __restrict__ out, int64_t ld_out,
                                                                     int64_t plane_stride, int nplanes, int src_rows) {
  const uint32_t sd = seed + (thr ? __ldg(salt) : 0u);
  const int wq = ((W + 7) & ~7) >> 2;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const int64_t sr = idx ? idx[r] : (int64_t)r;
    if (src_rows > 0 && (sr < 0 || sr >= src_rows)) __trap();   // out-of-vocabulary id: fail like nn.Embedding's device assert
    const float* s = src + sr * ld_src;
    __nv_bfloat16* d = out + (int64_t)r * ld_out;
    for (int q = threadIdx.x; q < wq; q += blockDim.x) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (q * 4 < W) {
        float4 f = __ldg(reinterpret_cast<const float4*>(s) + q);
        if (thr) drop_apply4(sd, (uint64_t)r * (uint64_t)W + (uint64_t)q * 4, thr, scale, f);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
      }
      planes_store4(d + q * 4, plane_stride, nplanes, v);
    }
  }
}

// ---- segment index arithmetic (reference basic_fc_model.py:80-121: the per-claim Python loops) in ONE launch -------------
// evd_cnt (B,) -> offsets (B+1,) = exclusive prefix sums, seg_of_row (B1,) = claim of every flattened evidence,
// slot_of_row (B1,) = claim * n + position inside the claim. Single block; B1 is the host-known row count.
__global__ void __launch_bounds__(1024) segments_kernel(const int64_t* __restrict__ cnt64, const int32_t* __restrict__ cnt32,
                                                        int B, int B1, int n, int32_t* __restrict__ seg,
                                                        int32_t* __restrict__ slot, int32_t* __restrict__ offsets) {
  extern __shared__ int s_off[];                      // B + 1
  for (int c = threadIdx.x; c <= B; c += blockDim.x) {
    int acc = 0;
    for (int q = 0; q < c; ++q) acc += cnt64 ? (int)cnt64[q] : cnt32[q];
    s_off[c] = acc;
    offsets[c] = acc;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < B1; r += blockDim.x) {
    int lo = 0, hi = B - 1;                           // last claim whose offset is <= r
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_off[mid] <= r) lo = mid; else hi = mid - 1;
    }
    seg[r] = lo;
    slot[r] = lo * n + (r - s_off[lo]);
  }
}

// mask[i] = ids[i] >= 1 (word level, gbss.py:98) / mask[r] = sum_j ids[r, j] >= 1 (evidence level, gbss.py:215)
__global__ void __launch_bounds__(256) ids_mask_kernel(const int64_t* __restrict__ ids64, const int32_t* __restrict__ ids32,
                                                       int64_t rows, int W, uint8_t* __restrict__ mask) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  long long acc = 0;
  for (int j = lane; j < W; j += 32) acc += ids64 ? (long long)ids64[r * W + j] : (long long)ids32[r * W + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) mask[r] = acc >= 1 ? 1 : 0;
}

// ---- trainable source embeddings (reference base_model.py:184-188, gbss.py:157-171): out[r,:] = table[max(idx[r],0),:]
// (the -1 padding id is looked up as row 0, gbss.py:166-168) and its dense, DETERMINISTIC gradient: one block per table
// row sums the matching output-gradient rows in index order.
__global__ void __launch_bounds__(128) embedding_rows_fwd_kernel(const float* __restrict__ table, int E, const int64_t* __restrict__ idx,
                                                                 int R, float* __restrict__ out, int64_t ld_out) {
  const int r = blockIdx.x;
  int64_t v = idx[r];
  if (v < 0) v = 0;
  for (int c = threadIdx.x; c < E; c += blockDim.x) out[(int64_t)r * ld_out + c] = __ldg(table + v * E + c);
}
constexpr int EB_THREADS = 512, EB_GROUPS = 4, EB_COLS = EB_THREADS / EB_GROUPS;
__global__ void __launch_bounds__(EB_THREADS) embedding_rows_bwd_kernel(const float* __restrict__ g, int64_t ld_g, const int64_t* __restrict__ idx,
                                                                        int R, int V, int E, float* __restrict__ dtable) {
  // block = one looked-up row r0. Only the FIRST occurrence of a table id does work: it lists every occurrence of that id
  // (index order), sums their gradient rows -- four thread groups own consecutive quarters of the list, eight independent
  // loads in flight per thread, partial sums combined in a fixed order (deterministic) -- and adds the total to the table
  // row once. The padding id (-1 -> row 0) occurs in most slots of a batch: its list is what sets the kernel's time.
  extern __shared__ int s_idx[];                      // [R] looked-up ids (-1 -> 0), then [R] the occurrence list
  __shared__ int s_dup, s_n;
  __shared__ float s_part[EB_GROUPS][EB_COLS];
  int* s_list = s_idx + R;
  const int r0 = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) s_dup = 0;
  for (int r = tid; r < R; r += EB_THREADS) {
    const int64_t id = idx[r];
    s_idx[r] = id < 0 ? 0 : (int)id;
  }
  __syncthreads();
  const int v = s_idx[r0];
  if (v >= V) __trap();                               // an id outside the table must fail loudly
  bool dup = false;
  for (int r = tid; r < r0; r += EB_THREADS) dup |= s_idx[r] == v;
  if (dup) s_dup = 1;
  __syncthreads();
  if (s_dup) return;
  if (tid < 32) {                                     // occurrence list in index order (one warp, ballot compaction)
    int n = 0;
    for (int base = r0; base < R; base += 32) {
      const int r = base + lane;
      const bool hit = r < R && s_idx[r] == v;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) s_list[n + __popc(m & ((1u << lane) - 1u))] = r;
      n += __popc(m);
    }
    if (lane == 0) s_n = n;
  }
  __syncthreads();
  const int n = s_n, grp = tid / EB_COLS, col = tid - grp * EB_COLS;
  const int per = (n + EB_GROUPS - 1) / EB_GROUPS, lo = min(n, grp * per), hi = min(n, lo + per);
  for (int c0 = 0; c0 < E; c0 += EB_COLS) {
    const int c = c0 + col;
    float acc = 0.f;
    if (c < E) {
      int k = lo;
      for (; k + 8 <= hi; k += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = g[(int64_t)s_list[k + u] * ld_g + c];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += t[u];
      }
      for (; k < hi; ++k) acc += g[(int64_t)s_list[k] * ld_g + c];
    }
    s_part[grp][col] = acc;
    __syncthreads();
    if (grp == 0 && c < E) {
      float tot = s_part[0][col];
#pragma unroll
      for (int q = 1; q < EB_GROUPS; ++q) tot += s_part[q][col];
      dtable[(int64_t)v * E + c] += tot;
    }
    __syncthreads();
  }
}

// ---- Adam with L2 weight decay over flat buffers (the reference fitter's torch.optim.Adam(lr, weight_decay),
// Fitting/FittingFC/declare_fitter.py:57-61): g += wd*p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps). The step counter t lives on the device (CUDA-graph replays).
__global__ void adam_step_advance_kernel(float* step) { *step += 1.0f; }
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n4, int64_t n, float lr, float b1, float b2,
                                                        float eps, float wd, const float* __restrict__ step) {
  const float t = __ldg(step);
  const float bc1 = 1.0f - powf(b1, t), bc2_sqrt = sqrtf(1.0f - powf(b2, t));
  const float step_size = lr / bc1;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    float pp[4], gg[4], mm[4], vv[4];
    const int64_t i = q * 4;
    const int nv = (int)min((int64_t)4, n - i);
    if (nv == 4) {
      const float4 a = *reinterpret_cast<const float4*>(p + i), b = *reinterpret_cast<const float4*>(g + i);
      const float4 c = *reinterpret_cast<const float4*>(m + i), d = *reinterpret_cast<const float4*>(v + i);
      pp[0] = a.x; pp[1] = a.y; pp[2] = a.z; pp[3] = a.w; gg[0] = b.x; gg[1] = b.y; gg[2] = b.z; gg[3] = b.w;
      mm[0] = c.x; mm[1] = c.y; mm[2] = c.z; mm[3] = c.w; vv[0] = d.x; vv[1] = d.y; vv[2] = d.z; vv[3] = d.w;
    } else {
      for (int e = 0; e < 4; ++e) { const bool ok = e < nv; pp[e] = ok ? p[i + e] : 0.f; gg[e] = ok ? g[i + e] : 0.f; mm[e] = ok ? m[i + e] : 0.f; vv[e] = ok ? v[i + e] : 0.f; }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = gg[e] + wd * pp[e];
      mm[e] = b1 * mm[e] + (1.0f - b1) * gr;
      vv[e] = b2 * vv[e] + (1.0f - b2) * gr * gr;
      pp[e] -= step_size * mm[e] / (sqrtf(vv[e]) / bc2_sqrt + eps);
    }
    if (nv == 4) {
      *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
      for (int e = 0; e < nv; ++e) { p[i + e] = pp[e]; m[i + e] = mm[e]; v[i + e] = vv[e]; }
    }
  }
}

// ---- mean cross-entropy + dlogits (reference losses.py:29-32); one warp, B is a few hundred at most
__global__ void __launch_bounds__(32) cross_entropy_kernel(const float* __restrict__ logits,
                                                           const int64_t* __restrict__ labels, int B, int C,
                                                           float* __restrict__ loss, float* __restrict__ dlogits) {
  const int lane = threadIdx.x;
  float total = 0.f;
  const float invB = 1.0f / (float)B;
  for (int b = lane; b < B; b += 32) {
    const float* row = logits + (int64_t)b * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, row[c]);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(row[c] - mx);
    const float lse = mx + logf(sum);
    const int y = (int)labels[b];
    total += lse - row[y];
    if (dlogits)
      for (int c = 0; c < C; ++c) dlogits[(int64_t)b * C + c] = (expf(row[c] - lse) - (c == y ? 1.f : 0.f)) * invB;
  }
  // fixed-order reduction over lanes for run-to-run determinism
  __shared__ float s[32];
  s[lane] = total;
  __syncwarp();
  if (lane == 0) {
    float v = 0.f;
    for (int l = 0; l < 32; ++l) v += s[l];
    *loss = v * invB;
  }
}

}  // namespace getb

using namespace getb;

extern "C" int get_ggnn_gate_bwd_f32(const float* dout, const float* z, const float* h, const float* x, int M, int H,
                                     int64_t ld_g, float* dhp, float* dzp, float* dx, void* stream) {
  GETB_REQUIRE(dout && z && h && x && dhp && dzp && dx, "get_ggnn_gate_bwd_f32: null pointer");
  GETB_REQUIRE(ld_g >= H, "get_ggnn_gate_bwd_f32: ld_g must be >= H");
  if (M <= 0 || H <= 0) return 0;
  const int vec = (H % 4 == 0) && (ld_g % 4 == 0) && aligned16(dout) && aligned16(z) && aligned16(h) && aligned16(x) &&
                  aligned16(dhp) && aligned16(dzp) && aligned16(dx);
  const int64_t nq = (int64_t)M * ((H + 3) / 4);
  ggnn_gate_bwd_kernel<<<ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(dout, z, h, x, M, H, ld_g, dhp, dzp, dx, vec);
  GETB_CHECK_LAUNCH("get_ggnn_gate_bwd_f32");
  return 0;
}

extern "C" int64_t get_colsum_workspace_floats(int M, int N) {
  return (int64_t)ceil_div(M > 0 ? M : 1, COLSUM_ROWS) * (N > 0 ? N : 1);
}

extern "C" int get_colsum_f32(const float* a, int64_t ld, int M, int N, float* out, float* workspace, void* stream) {
  GETB_REQUIRE(a && out && workspace, "get_colsum_f32: null pointer");
  GETB_REQUIRE(M > 0 && N > 0, "get_colsum_f32: bad sizes");
  const int nparts = ceil_div(M, COLSUM_ROWS);
  dim3 grid(ceil_div(N, 32), nparts);
  colsum_stage1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, ld, M, N, workspace);
  GETB_CHECK_LAUNCH("get_colsum_f32/stage1");
  colsum_stage2_kernel<<<ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(workspace, nparts, N, out);
  GETB_CHECK_LAUNCH("get_colsum_f32/stage2");
  return 0;
}

extern "C" int get_rows_gather_f32(const float* src, int64_t ld_src, const int32_t* idx, int R, int W, float* out,
                                   int64_t ld_out, void* stream) {
  GETB_REQUIRE(src && idx && out, "get_rows_gather_f32: null pointer");
  if (R <= 0 || W <= 0) return 0;
  rows_copy_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(src, ld_src, idx, R, W, out, ld_out, 0);
  GETB_CHECK_LAUNCH("get_rows_gather_f32");
  return 0;
}

extern "C" int get_rows_scatter_f32(const float* src, int64_t ld_src, const int32_t* idx, int R, int W, float* out,
                                    int64_t ld_out, void* stream) {
  GETB_REQUIRE(src && idx && out, "get_rows_scatter_f32: null pointer");
  if (R <= 0 || W <= 0) return 0;
  rows_copy_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(src, ld_src, idx, R, W, out, ld_out, 1);
  GETB_CHECK_LAUNCH("get_rows_scatter_f32");
  return 0;
}

extern "C" int get_segment_sum_f32(const float* src, int64_t ld_src, const int32_t* offsets, int S, int W, float* out,
                                   int64_t ld_out, void* stream) {
  GETB_REQUIRE(src && offsets && out, "get_segment_sum_f32: null pointer");
  if (S <= 0 || W <= 0) return 0;
  segment_sum_kernel<<<S, 256, 0, (cudaStream_t)stream>>>(src, ld_src, offsets, W, out, ld_out);
  GETB_CHECK_LAUNCH("get_segment_sum_f32");
  return 0;
}

extern "C" int get_masked_mean_fwd_f32(const float* h, const int64_t* ids, const int64_t* lens, int G, int N, int H,
                                       float* out, void* stream) {
  GETB_REQUIRE(h && ids && lens && out, "get_masked_mean_fwd_f32: null pointer");
  if (G <= 0) return 0;
  masked_mean_fwd_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(h, ids, lens, N, H, out);
  GETB_CHECK_LAUNCH("get_masked_mean_fwd_f32");
  return 0;
}

extern "C" int get_masked_mean_bwd_f32(const float* dout, const int64_t* ids, const int64_t* lens, int G, int N, int H,
                                       float* dh, void* stream) {
  GETB_REQUIRE(dout && ids && lens && dh, "get_masked_mean_bwd_f32: null pointer");
  if (G <= 0) return 0;
  masked_mean_bwd_kernel<<<G * N, 256, 0, (cudaStream_t)stream>>>(dout, ids, lens, N, H, dh);
  GETB_CHECK_LAUNCH("get_masked_mean_bwd_f32");
  return 0;
}

extern "C" int get_dropout_mask_f32(float* out, int64_t numel, float p, uint32_t seed, void* stream) {
  GETB_REQUIRE(out && p >= 0.f && p < 1.f, "get_dropout_mask_f32: bad arguments");
  if (numel <= 0) return 0;
  dropout_mask_kernel<<<ceil_div(numel, 256), 256, 0, (cudaStream_t)stream>>>(out, numel, drop_threshold(p),
                                                                             1.0f / (1.0f - p), seed, dropout_salt_ptr());
  GETB_CHECK_LAUNCH("get_dropout_mask_f32");
  return 0;
}

extern "C" int get_dropout_salt_set(uint32_t value, void* stream) {
  uint32_t* s = const_cast<uint32_t*>(dropout_salt_ptr());
  GETB_REQUIRE(s != nullptr, "get_dropout_salt_set: cannot allocate the salt word");
  salt_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(s, value, 0);
  GETB_CHECK_LAUNCH("get_dropout_salt_set");
  return 0;
}

extern "C" int get_dropout_salt_advance(void* stream) {
  uint32_t* s = const_cast<uint32_t*>(dropout_salt_ptr());
  GETB_REQUIRE(s != nullptr, "get_dropout_salt_advance: cannot allocate the salt word");
  salt_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(s, 0u, 1);
  GETB_CHECK_LAUNCH("get_dropout_salt_advance");
  return 0;
}

extern "C" int get_dropout_salt_get(uint32_t* host_value) {
  GETB_REQUIRE(host_value != nullptr, "get_dropout_salt_get: null pointer");
  const uint32_t* s = dropout_salt_ptr();
  GETB_REQUIRE(s != nullptr, "get_dropout_salt_get: cannot allocate the salt word");
  cudaError_t e = cudaMemcpy(host_value, s, sizeof(uint32_t), cudaMemcpyDeviceToHost);   // synchronises: tests only
  if (e != cudaSuccess) {
    getb::set_error("get_dropout_salt_get: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

extern "C" int get_rows_gather_dropout_f32(const float* src, int64_t ld_src, const int64_t* idx, int R, int W, float p,
                                           uint32_t seed, float* out, int64_t ld_out, void* stream) {
  GETB_REQUIRE(src && out && p >= 0.f && p < 1.f, "get_rows_gather_dropout_f32: bad arguments");
  if (R <= 0 || W <= 0) return 0;
  const int vec = (W % 4) == 0 && (ld_src % 4) == 0 && (ld_out % 4) == 0 && aligned16(src) && aligned16(out);
  const int grid = R < 148 * 16 ? R : 148 * 16;
  rows_gather_dropout_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(src, ld_src, idx, R, W, p > 0.f ? drop_threshold(p) : 0u,
                                                                    1.0f / (1.0f - p), seed, dropout_salt_ptr(), out, ld_out, vec);
  GETB_CHECK_LAUNCH("get_rows_gather_dropout_f32");
  return 0;
}

extern "C" int get_ggnn_gate_bwd_bp(const float* dout, const float* z, const float* h, const float* x, int M, int H, void* dg,
                                    int64_t ld_g, int64_t plane_stride, int nplanes, int col_z, int col_h, float* dx,
                                    void* stream) {
  GETB_REQUIRE(dout && z && h && x && dg && dx, "get_ggnn_gate_bwd_bp: null pointer");
  GETB_REQUIRE((H % 4) == 0 && (ld_g % 4) == 0 && (plane_stride % 4) == 0 && (col_z % 4) == 0 && (col_h % 4) == 0 &&
                   nplanes >= 1 && nplanes <= 3 && aligned16(dout) && aligned16(z) && aligned16(h) && aligned16(x) &&
                   aligned16(dx) && (((uintptr_t)dg) & 7u) == 0,
               "get_ggnn_gate_bwd_bp: H %% 4, alignment or plane count");
  if (M <= 0 || H <= 0) return 0;
  const int64_t nq = (int64_t)M * (((H + 7) & ~7) / 4);
  ggnn_gate_bwd_bp_kernel<<<ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(dout, z, h, x, M, H, reinterpret_cast<__nv_bfloat16*>(dg),
                                                                              ld_g, plane_stride, nplanes, col_z, col_h, dx);
  GETB_CHECK_LAUNCH("get_ggnn_gate_bwd_bp");
  return 0;
}

extern "C" int get_rows_gather_dropout_bp(const float* src, int64_t ld_src, int src_rows, const int64_t* idx, int R, int W, float p,
                                          uint32_t seed, void* planes, int64_t ld_out, int64_t plane_stride, int nplanes,
                                          void* stream) {
  GETB_REQUIRE(src && planes && p >= 0.f && p < 1.f, "get_rows_gather_dropout_bp: bad arguments");
  GETB_REQUIRE((W % 4) == 0 && (ld_src % 4) == 0 && (ld_out % 4) == 0 && (plane_stride % 4) == 0 && aligned16(src) &&
                   (((uintptr_t)planes) & 7u) == 0 && nplanes >= 1 && nplanes <= 3 && ld_out >= ((W + 7) & ~7),
               "get_rows_gather_dropout_bp: W %% 4, alignment or plane count");
  if (R <= 0 || W <= 0) return 0;
  const int grid = R < 148 * 16 ? R : 148 * 16;
  rows_gather_dropout_bp_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(src, ld_src, idx, R, W, p > 0.f ? drop_threshold(p) : 0u,
                                                                       1.0f / (1.0f - p), seed, dropout_salt_ptr(),
                                                                       reinterpret_cast<__nv_bfloat16*>(planes), ld_out,
                                                                       plane_stride, nplanes, src_rows);
  GETB_CHECK_LAUNCH("get_rows_gather_dropout_bp");
  return 0;
}

extern "C" int get_segments_i32(const void* evd_cnt, int cnt_is_int64, int B, int B1, int n, int32_t* seg_of_row,
                                int32_t* slot_of_row, int32_t* offsets, void* stream) {
  GETB_REQUIRE(evd_cnt && seg_of_row && slot_of_row && offsets && B >= 1 && B1 >= 0 && n >= 1, "get_segments_i32: bad arguments");
  GETB_REQUIRE(B <= 8192, "get_segments_i32: at most 8192 claims per call");
  segments_kernel<<<1, 1024, (size_t)(B + 1) * sizeof(int), (cudaStream_t)stream>>>(
      cnt_is_int64 ? reinterpret_cast<const int64_t*>(evd_cnt) : nullptr, cnt_is_int64 ? nullptr : reinterpret_cast<const int32_t*>(evd_cnt),
      B, B1, n, seg_of_row, slot_of_row, offsets);
  GETB_CHECK_LAUNCH("get_segments_i32");
  return 0;
}

extern "C" int get_ids_mask_u8(const void* ids, int ids_is_int64, int64_t rows, int W, uint8_t* mask, void* stream) {
  GETB_REQUIRE(ids && mask && rows >= 0 && W >= 1, "get_ids_mask_u8: bad arguments");
  if (rows == 0) return 0;
  ids_mask_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(
      ids_is_int64 ? reinterpret_cast<const int64_t*>(ids) : nullptr, ids_is_int64 ? nullptr : reinterpret_cast<const int32_t*>(ids), rows, W, mask);
  GETB_CHECK_LAUNCH("get_ids_mask_u8");
  return 0;
}

extern "C" int get_embedding_rows_fwd_f32(const float* table, int V, int E, const int64_t* idx, int R, float* out, int64_t ld_out,
                                          void* stream) {
  GETB_REQUIRE(table && idx && out && V >= 1 && E >= 1 && ld_out >= E, "get_embedding_rows_fwd_f32: bad arguments");
  if (R <= 0) return 0;
  embedding_rows_fwd_kernel<<<R, 128, 0, (cudaStream_t)stream>>>(table, E, idx, R, out, ld_out);
  GETB_CHECK_LAUNCH("get_embedding_rows_fwd_f32");
  return 0;
}

extern "C" int get_embedding_rows_bwd_f32(const float* g, int64_t ld_g, const int64_t* idx, int R, int V, int E, float* dtable,
                                          int accumulate, void* stream) {
  GETB_REQUIRE(g && idx && dtable && V >= 1 && E >= 1 && R >= 0 && ld_g >= E, "get_embedding_rows_bwd_f32: bad arguments");
  GETB_REQUIRE((size_t)R * 2 * sizeof(int) <= 160 * 1024, "get_embedding_rows_bwd_f32: at most 20480 looked-up rows per call");
  const size_t smem = (size_t)(R > 0 ? R : 1) * 2 * sizeof(int);
  if (smem > 48 * 1024) {
    static bool done = false;
    if (!done) {
      cudaFuncSetAttribute(embedding_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      done = true;
    }
  }
  if (!accumulate) {
    if (cudaMemsetAsync(dtable, 0, (size_t)V * E * sizeof(float), (cudaStream_t)stream) != cudaSuccess) {
      (void)cudaGetLastError();
      getb::set_error("get_embedding_rows_bwd_f32: memset failed");
      return -2;
    }
  }
  if (R == 0) return 0;
  embedding_rows_bwd_kernel<<<R, EB_THREADS, smem, (cudaStream_t)stream>>>(g, ld_g, idx, R, V, E, dtable);
  GETB_CHECK_LAUNCH("get_embedding_rows_bwd_f32");
  return 0;
}

extern "C" int get_adam_flat_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, float* step, void* stream) {
  GETB_REQUIRE(param && grad && exp_avg && exp_avg_sq && step && n >= 0, "get_adam_flat_f32: bad arguments");
  GETB_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), "get_adam_flat_f32: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  adam_step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
  GETB_CHECK_LAUNCH("get_adam_flat_f32/step");
  const int64_t n4 = (n + 3) / 4;
  int grid = ceil_div(n4, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  adam_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n4, n, lr, beta1, beta2, eps, weight_decay, step);
  GETB_CHECK_LAUNCH("get_adam_flat_f32");
  return 0;
}

extern "C" int get_cross_entropy_f32(const float* logits, const int64_t* labels, int B, int C, float* loss,
                                     float* dlogits, void* stream) {
  GETB_REQUIRE(logits && labels && loss && B > 0 && C > 0, "get_cross_entropy_f32: bad arguments");
  cross_entropy_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(logits, labels, B, C, loss, dlogits);
  GETB_CHECK_LAUNCH("get_cross_entropy_f32");
  return 0;
}

extern "C" int get_b200_abi_version(void) { return GET_B200_ABI_VERSION; }
extern "C" const char* get_b200_last_error(void) { return g_err; }
extern "C" int64_t get_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
