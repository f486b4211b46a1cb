// The reference's candidate stage on the device, for a whole query batch in ONE launch:
//
//   trec_eval_step (mfar/modeling/contrastive.py:676-696), per query:
//     union of the F per-field top-k hit lists            (678-679: Python sets of doc keys)
//     score_batch(union) on every field index             (681-683: dict lookups + memmap fancy-index + matmul, x F)
//     stack * mask -> LinearWeights mixture               (685-694)
//     torch.topk(k)                                       (696)
//
// One CTA per query: the F*k candidate rows are sorted in shared memory (block bitonic sort) and de-duplicated - the
// union, in ascending row order; a warp per candidate gathers the doc's F_d bf16 field vectors straight from the packed
// corpus (16-byte loads, the query vector held in shared memory), adds the sparse fields' stored scores, mixes with the
// query's masked softmax weights and leaves a (score, row) key in shared memory; a second block sort ranks the union
// and the top k go out.  Latency-bound gather (U <= F*k rows of F_d*dim*2 bytes per query), no intermediate in HBM.
#include "common.cuh"
#include "kernels.h"

namespace mfar {

constexpr int kUnionThreads = 512;
constexpr uint32_t kNoRow = 0xFFFFFFFFu;

template <typename T, bool kDescending>
__device__ __forceinline__ void block_bitonic_sort(T* buf, int n) {          // n: power of two
  for (int s = 2; s <= n; s <<= 1) {
    for (int d = s >> 1; d > 0; d >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int lo = ((t & ~(d - 1)) << 1) | (t & (d - 1));
        const int hi = lo | d;
        const bool first_larger = ((lo & s) == 0) == kDescending;            // direction of this bitonic block
        const T a = buf[lo], b = buf[hi];
        if ((a < b) == first_larger) { buf[lo] = b; buf[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kUnionThreads)
union_rescore_kernel(const __nv_bfloat16* __restrict__ corpus, int64_t n_docs, int corpus_fields, int n_dense, int dim,
                     const __nv_bfloat16* __restrict__ q_vecs, const float* __restrict__ w, int w_ld,
                     const void* __restrict__ sparse, int sparse_f16, int64_t sparse_ld, int n_sparse,
                     const int64_t* __restrict__ cand, int L, int Q, int k_in, int k_out, int P,
                     float* __restrict__ out_scores, int64_t* __restrict__ out_rows, int* __restrict__ out_union) {
  extern __shared__ __align__(16) uint8_t smem_u[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_u);                       // [P]
  uint32_t* ids = reinterpret_cast<uint32_t*>(keys + P);                      // [P]
  float* qf = reinterpret_cast<float*>(ids + P);                              // [dim]
  float* wf = qf + dim;                                                       // [n_dense + n_sparse]
  __shared__ int s_part[kUnionThreads / 32];
  __shared__ int s_total;
  const int q = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = kUnionThreads / 32;

  // ---- candidates of this query from every field list; rows outside the shard (and padding) become the sentinel
  for (int i = t; i < P; i += kUnionThreads) {
    uint32_t r = kNoRow;
    if (i < L * k_in) {
      const int64_t row = cand[(int64_t(i / k_in) * Q + q) * k_in + (i % k_in)];
      if (row >= 0 && row < n_docs) r = uint32_t(row);
    }
    ids[i] = r;
  }
  for (int i = t; i < dim; i += kUnionThreads) qf[i] = __bfloat162float(q_vecs[int64_t(q) * dim + i]);
  for (int i = t; i < n_dense + n_sparse; i += kUnionThreads) wf[i] = w[int64_t(q) * w_ld + i];
  block_bitonic_sort<uint32_t, false>(ids, P);                                // ascending; sentinels last

  // ---- union: first occurrence of every row, compacted in order (block scan over per-thread chunk counts)
  const int chunk = (P + kUnionThreads - 1) / kUnionThreads;
  const int c0 = t * chunk, c1 = min(P, c0 + chunk);
  int mine = 0;
  for (int i = c0; i < c1; ++i) mine += (ids[i] != kNoRow && (i == 0 || ids[i] != ids[i - 1])) ? 1 : 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_part[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = lane < nwarp ? s_part[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane < nwarp) s_part[lane] = v;                                       // inclusive over warps
    if (lane == nwarp - 1) s_total = v;
  }
  __syncthreads();
  int pos = incl - mine + (warp ? s_part[warp - 1] : 0);
  const int U = s_total;
  // the compacted rows go to the LOW words of keys[] (ids[] is still being read by neighbours)
  for (int i = c0; i < c1; ++i)
    if (ids[i] != kNoRow && (i == 0 || ids[i] != ids[i - 1])) keys[pos++] = uint64_t(ids[i]);
  __syncthreads();

  // ---- re-score: a warp per candidate, all fields, mixture in fp32
  const int chunks = dim >> 3;                                                // 16-byte chunks per row
  for (int u = warp; u < U; u += nwarp) {
    const uint32_t row = uint32_t(keys[u]);
    const int64_t tile = row / kTileDocs, in_tile = row % kTileDocs;
    float mix = 0.f;
    for (int f = 0; f < n_dense; ++f) {
      const uint4* v = reinterpret_cast<const uint4*>(
          corpus + ((tile * corpus_fields + f) * kTileDocs + in_tile) * int64_t(dim));
      float acc = 0.f;
      for (int c = lane; c < chunks; c += 32) {
        const uint4 x = __ldg(v + c);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&x);
        const float* qq = qf + c * 8;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 d2 = __bfloat1622float2(h[e]);
          acc = fmaf(d2.x, qq[2 * e], acc);
          acc = fmaf(d2.y, qq[2 * e + 1], acc);
        }
      }
      for (int i = (chunks << 3) + lane; i < dim; i += 32)                    // dim % 8 tail (packed dims are multiples of 64)
        acc = fmaf(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(v)[i]), qf[i], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      mix = fmaf(wf[f], acc, mix);
    }
    if (lane == 0) {
      for (int j = 0; j < n_sparse; ++j) {                                    // stored per-field BM25 score of (query, doc)
        const int64_t off = (int64_t(q) * n_sparse + j) * sparse_ld + row;
        const float sv = sparse_f16 ? __half2float(static_cast<const __half*>(sparse)[off])
                                    : static_cast<const float*>(sparse)[off];
        mix = fmaf(wf[n_dense + j], sv, mix);
      }
      keys[u] = make_key(mix, row);
    }
  }
  __syncthreads();
  if (t == 0) out_union[q] = U;
  // ---- rank the union, emit the top k_out (score desc, row asc); fewer than k_out candidates: (-inf, -1) tail
  int P2 = 128;
  while (P2 < U) P2 <<= 1;
  for (int i = U + t; i < P2; i += kUnionThreads) keys[i] = 0ull;
  block_bitonic_sort<uint64_t, true>(keys, P2);
  for (int j = t; j < k_out; j += kUnionThreads) {
    const uint64_t key = j < P2 ? keys[j] : 0ull;
    out_scores[int64_t(q) * k_out + j] = key ? key_score(key) : -INFINITY;
    out_rows[int64_t(q) * k_out + j] = key ? int64_t(key_doc(key)) : int64_t(-1);
  }
}

int launch_union_rescore(const void* corpus, int64_t n_docs, int corpus_fields, int n_dense, int dim, const void* q_vecs,
                         int Q, const float* w, int w_ld, const void* sparse, int sparse_dtype, int64_t sparse_ld,
                         int n_sparse, const int64_t* cand, int L, int k_in, int k_out, float* out_scores,
                         int64_t* out_rows, int* out_union, cudaStream_t st) {
  int P = 128;
  while (P < L * k_in) P <<= 1;
  if (P > 8192) return MFAR_ERR_SHAPE;
  const size_t smem = size_t(P) * 12 + size_t(dim) * 4 + size_t(n_dense + n_sparse) * 4 + 16;
  static PerDeviceOnce attr_once;
  MFAR_CUDA_OK(attr_once.run([&] {
    return cudaFuncSetAttribute(union_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  }));
  if (smem > 128 * 1024) return MFAR_ERR_SHAPE;
  union_rescore_kernel<<<Q, kUnionThreads, smem, st>>>(
      static_cast<const __nv_bfloat16*>(corpus), n_docs, corpus_fields, n_dense, dim,
      static_cast<const __nv_bfloat16*>(q_vecs), w, w_ld, sparse, sparse_dtype == MFAR_F16 ? 1 : 0, sparse_ld, n_sparse,
      cand, L, Q, k_in, k_out, P, out_scores, out_rows, out_union);
  MFAR_CUDA_OK(cudaGetLastError());
  return MFAR_OK;
}

}  // namespace mfar
