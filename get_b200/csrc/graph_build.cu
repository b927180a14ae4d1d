// Word-graph construction on the GPU: the host code of the reference builds, once per
// text, a sliding-window co-occurrence graph over de-duplicated tokens and normalises it symmetrically
// (interactions.py:334-351 `convert_text`, interactions.py:11-18 `_laplacian_normalize`); its dense float64 adjacency is
// what the fitter then ships to the device every step (18.8 MB per 32-claim batch). This kernel builds the same node list
// and the same normalised adjacency from the raw token ids (800 bytes per text) on the device.
//
// One CTA per text. nodes = distinct tokens of the first `length` positions in first-occurrence order; edge (u,v) iff the
// two words occur within `window-1` positions of each other (self loops included); adj = D^-1/2 A D^-1/2 evaluated in
// double precision exactly like the numpy code ((1 * d_j) * d_i) and rounded once to fp32, zero rows/cols for unused slots.
#include "common.cuh"

namespace getb {

constexpr int GB_THREADS = 256;

// smem: tok[T] i64 | first[T] i32 | node_of_pos[T] i32 | is_first[T] i32 | A[N*N] u8 | dis[N] f64
__global__ void __launch_bounds__(GB_THREADS) build_word_graphs_kernel(const int64_t* __restrict__ tokens,
                                                                       const int32_t* __restrict__ lengths, int T, int N,
                                                                       int window, int64_t* __restrict__ nodes,
                                                                       float* __restrict__ adj, int32_t* __restrict__ n_nodes) {
  extern __shared__ __align__(16) unsigned char gb_smem[];
  const int g = blockIdx.x, tid = threadIdx.x;
  double* s_dis = reinterpret_cast<double*>(gb_smem);
  int64_t* s_tok = reinterpret_cast<int64_t*>(s_dis + N);
  int* s_first = reinterpret_cast<int*>(s_tok + T);
  int* s_node = s_first + T;
  int* s_isf = s_node + T;
  unsigned char* s_A = reinterpret_cast<unsigned char*>(s_isf + T);
  __shared__ int s_n;

  int len = lengths[g];
  len = max(0, min(len, min(T, N)));      // the reference keeps the first `fixed_length` tokens (interactions.py:303,321)
  for (int p = tid; p < T; p += GB_THREADS) s_tok[p] = p < len ? tokens[(int64_t)g * T + p] : 0;
  for (int q = tid; q < N * N; q += GB_THREADS) s_A[q] = 0;
  __syncthreads();
  // first occurrence of every position's word
  for (int p = tid; p < len; p += GB_THREADS) {
    const int64_t t = s_tok[p];
    int f = p;
    for (int q = 0; q < p; ++q)
      if (s_tok[q] == t) { f = q; break; }
    s_first[p] = f;
    s_isf[p] = f == p;
  }
  __syncthreads();
  // node id of a first occurrence = number of first occurrences before it
  for (int p = tid; p < len; p += GB_THREADS) {
    if (s_isf[p]) {
      int c = 0;
      for (int q = 0; q < p; ++q) c += s_isf[q];
      s_node[p] = c;
      nodes[(int64_t)g * N + c] = s_tok[p];
    }
  }
  if (tid == 0) {
    int c = 0;
    for (int q = 0; q < len; ++q) c += s_isf[q];
    s_n = c;
    n_nodes[g] = c;
  }
  __syncthreads();
  const int nn = s_n;
  for (int i = nn + tid; i < N; i += GB_THREADS) nodes[(int64_t)g * N + i] = 0;
  for (int p = tid; p < len; p += GB_THREADS)
    if (!s_isf[p]) s_node[p] = s_node[s_first[p]];
  __syncthreads();
  // co-occurrence within the window (both directions, self loops)
  const int span = 2 * window - 1;
  for (int e = tid; e < len * span; e += GB_THREADS) {
    const int p = e / span, q = p + (e % span) - (window - 1);
    if (q >= 0 && q < len) s_A[s_node[p] * N + s_node[q]] = 1;
  }
  __syncthreads();
  for (int i = tid; i < N; i += GB_THREADS) {
    int deg = 0;
    for (int j = 0; j < N; ++j) deg += s_A[i * N + j];
    s_dis[i] = deg > 0 ? pow((double)deg, -0.5) : 0.0;
  }
  __syncthreads();
  float* ag = adj + (int64_t)g * N * N;
  for (int q = tid; q < N * N; q += GB_THREADS) {
    const int i = q / N, j = q % N;
    // numpy: ((A * d[None,:]).T * d[None,:])[i][j] = (A[j][i] * d[i]) * d[j]
    ag[q] = s_A[j * N + i] ? (float)((1.0 * s_dis[i]) * s_dis[j]) : 0.f;
  }
}

}  // namespace getb

using namespace getb;

extern "C" int get_build_word_graphs(const int64_t* tokens, const int32_t* lengths, int G, int T, int N, int window,
                                     int64_t* nodes, float* adj, int32_t* n_nodes, void* stream) {
  GETB_REQUIRE(tokens && lengths && nodes && adj && n_nodes, "get_build_word_graphs: null pointer");
  GETB_REQUIRE(G >= 0 && T > 0 && N > 0 && window >= 1, "get_build_word_graphs: bad sizes");
  if (G == 0) return 0;
  const size_t smem = (size_t)N * sizeof(double) + (size_t)T * (sizeof(int64_t) + 3 * sizeof(int)) + (size_t)N * N;
  GETB_REQUIRE(smem <= 200 * 1024, "get_build_word_graphs: N=%d too large for one CTA", N);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(build_word_graphs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  build_word_graphs_kernel<<<G, GB_THREADS, smem, (cudaStream_t)stream>>>(tokens, lengths, T, N, window, nodes, adj, n_nodes);
  GETB_CHECK_LAUNCH("get_build_word_graphs");
  return 0;
}
