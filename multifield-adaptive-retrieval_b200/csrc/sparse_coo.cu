// Producer of the reference's precomputed-BM25 score files: dense per-query score rows -> COO pairs.
//
// Reference: precompute_score_for_field (mfar/commands/precompute_bm25s_scores.py:12-30) maps every train query
// through BM25sSparseIndex.get_scores_sparse (mfar/data/index.py:78-84: a full-corpus get_scores vector, keep the
// entries that are != 0 AND whose doc id is in `safe_docs`), appends (int(qid), int(doc_id)) / np.float16(score) per
// entry - queries in dict order, docs ascending - and saves the int32 [nnz,2] / float16 [nnz] array pair.
//
// Here the score rows already live in HBM (fp32 [Q, ld], written by the BM25 postings scatter of bm25.cu), so the
// filter + compaction runs on the device in two passes over the rows and only the nnz pairs cross PCIe:
//   count : grid (segments of kCooSegDocs docs, Q); a warp owns 512 consecutive docs (16 coalesced 128-byte loads in
//           flight), flag = (score != 0) & safe bit, count = popc(ballot); per-segment counts -> one CTA scans them
//           in (query, segment) order into exclusive offsets (last slot = nnz);
//   write : same walk; slot = segment offset + earlier warps of the segment + earlier ballots of the warp + lanes
//           below -> output order is exactly the reference's (query asc, doc asc), deterministic, no atomics.
// HBM-bound: algorithmic bytes = 2 * Q*N*4 (rows read twice) + nnz*10 written.
#include "common.cuh"
#include "kernels.h"

namespace mfar {

constexpr int kCooThreads = 256;
constexpr int kCooIters = 16;                                   // 32-doc rows per warp
constexpr int kCooWarpDocs = 32 * kCooIters;                    // 512
constexpr int kCooSegDocs = (kCooThreads / 32) * kCooWarpDocs;  // 4096 docs per CTA

long long sparse_coo_segments(long long n_docs) { return (n_docs + kCooSegDocs - 1) / kCooSegDocs; }

// The 16 keep-masks of this warp's 512 docs (and the values).  Score loads and the safe-set bitmap words are all
// issued up front: a row of 32 docs needs 32 consecutive bitmap bits = a funnel shift of two adjacent words, so lanes
// 0..16 fetch the warp's 17 words once and the rows pick theirs with shuffles - no dependent or divergent load sits
// between the score loads and the ballots (the first version tested the bit per kept lane: 16 serialised L1 round
// trips per warp, ncu: 2.0 TB/s with 95 % of the warp slots occupied).
__device__ __forceinline__ void coo_load_flags(const float* __restrict__ row, long long n_docs, long long doc0,
                                               const uint32_t* __restrict__ safe_bits, long long n_words,
                                               long long doc_id_base, int lane, float (&v)[kCooIters],
                                               unsigned (&ballots)[kCooIters]) {
  const unsigned long long g0 = (unsigned long long)(doc_id_base + doc0);   // global id of the warp's first doc
  unsigned w = 0xffffffffu;                                                 // no bitmap: every doc is safe
  if (safe_bits != nullptr) {
    const long long wi = (long long)(g0 >> 5) + lane;
    w = (lane <= kCooIters && wi < n_words) ? __ldg(safe_bits + wi) : 0u;   // index.py:82-83
  }
#pragma unroll
  for (int it = 0; it < kCooIters; ++it) {
    const long long n = doc0 + it * 32 + lane;
    v[it] = (n < n_docs) ? __ldcs(row + n) : 0.f;               // streamed once per pass
  }
  const unsigned sh = unsigned(g0 & 31ull);                     // the same for all rows: doc0 + it*32 keeps g0 mod 32
#pragma unroll
  for (int it = 0; it < kCooIters; ++it) {
    const unsigned lo = __shfl_sync(0xffffffffu, w, it), hi = __shfl_sync(0xffffffffu, w, it + 1);
    const unsigned safe = __funnelshift_r(lo, hi, sh);          // bit l = doc g0 + it*32 + l
    ballots[it] = __ballot_sync(0xffffffffu, v[it] != 0.f) & safe;   // index.py:81 (NaN != 0 is true there too)
  }
}

__global__ void __launch_bounds__(kCooThreads)
sparse_coo_count_kernel(const float* __restrict__ scores, long long ld, long long n_docs,
                        const uint32_t* __restrict__ safe_bits, long long n_words, long long doc_id_base,
                        long long* __restrict__ seg_offsets) {
  __shared__ int s_warp[kCooThreads / 32];
  const int q = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long doc0 = (long long)blockIdx.x * kCooSegDocs + warp * kCooWarpDocs;
  float v[kCooIters];
  unsigned ballots[kCooIters];
  coo_load_flags(scores + (long long)q * ld, n_docs, doc0, safe_bits, n_words, doc_id_base, lane, v, ballots);
  int c = 0;
#pragma unroll
  for (int it = 0; it < kCooIters; ++it) c += __popc(ballots[it]);
  if (lane == 0) s_warp[warp] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kCooThreads / 32; ++w) t += s_warp[w];
    seg_offsets[1 + (long long)q * gridDim.x + blockIdx.x] = t;   // slot 0 stays 0: the scan below is inclusive
  }
}

// in-place inclusive scan of seg_offsets[1 .. m] (one CTA; m = Q * segments is at most a few 100k): every thread owns
// a contiguous chunk (serial sum, block scan of the 1024 chunk sums, serial rewrite)
__global__ void __launch_bounds__(1024) sparse_coo_scan_kernel(long long* __restrict__ seg_offsets, long long m) {
  __shared__ long long s_part[32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long per = (m + 1023) / 1024;
  const long long lo = min(m, (long long)t * per), hi = min(m, lo + per);
  long long sum = 0;
  for (long long i = lo; i < hi; ++i) sum += seg_offsets[1 + i];
  long long x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_part[warp] = x;
  __syncthreads();
  if (warp == 0) {
    long long p = s_part[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, p, o);
      if (lane >= o) p += y;
    }
    s_part[lane] = p;                                            // inclusive over warps
  }
  __syncthreads();
  long long run = x - sum + (warp ? s_part[warp - 1] : 0);       // exclusive prefix of this thread's chunk
  for (long long i = lo; i < hi; ++i) {
    run += seg_offsets[1 + i];
    seg_offsets[1 + i] = run;
  }
  if (t == 0) seg_offsets[0] = 0;
}

template <typename VT>
__global__ void __launch_bounds__(kCooThreads)
sparse_coo_write_kernel(const float* __restrict__ scores, long long ld, long long n_docs,
                        const uint32_t* __restrict__ safe_bits, long long n_words, const int* __restrict__ qids,
                        long long doc_id_base, const long long* __restrict__ seg_offsets, int2* __restrict__ out_keys,
                        VT* __restrict__ out_vals) {
  __shared__ int s_warp[kCooThreads / 32];
  const int q = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long doc0 = (long long)blockIdx.x * kCooSegDocs + warp * kCooWarpDocs;
  float v[kCooIters];
  unsigned ballots[kCooIters];
  coo_load_flags(scores + (long long)q * ld, n_docs, doc0, safe_bits, n_words, doc_id_base, lane, v, ballots);
  int c = 0;
#pragma unroll
  for (int it = 0; it < kCooIters; ++it) c += __popc(ballots[it]);
  if (lane == 0) s_warp[warp] = c;
  __syncthreads();
  long long pos = seg_offsets[(long long)q * gridDim.x + blockIdx.x];
  for (int w = 0; w < warp; ++w) pos += s_warp[w];
  const int qid = qids ? __ldg(qids + q) : q;
  const unsigned below = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < kCooIters; ++it) {
    if ((ballots[it] >> lane) & 1u) {
      const long long o = pos + __popc(ballots[it] & below);
      out_keys[o] = make_int2(qid, int(doc_id_base + doc0 + it * 32 + lane));
      out_vals[o] = VT(v[it]);
    }
    pos += __popc(ballots[it]);
  }
}

// np.float16(score): round to nearest even, overflow -> inf (numpy semantics == __float2half_rn)
struct HalfRn {
  __half h;
  __device__ explicit HalfRn(float f) : h(__float2half_rn(f)) {}
};

int launch_sparse_coo_count(const float* scores, long long ld, int Q, long long n_docs, const uint32_t* safe_bits,
                            long long doc_id_base, long long* seg_offsets, cudaStream_t st) {
  const long long segs = sparse_coo_segments(n_docs);
  if (segs > 0x7fffffffll || Q > 65535) return MFAR_ERR_SHAPE;
  dim3 grid((unsigned)segs, (unsigned)Q);
  const long long n_words = (doc_id_base + n_docs + 31) / 32;   // the bitmap covers every global id of this shard
  sparse_coo_count_kernel<<<grid, kCooThreads, 0, st>>>(scores, ld, n_docs, safe_bits, n_words, doc_id_base,
                                                        seg_offsets);
  MFAR_CUDA_OK(cudaGetLastError());
  sparse_coo_scan_kernel<<<1, 1024, 0, st>>>(seg_offsets, segs * Q);
  MFAR_CUDA_OK(cudaGetLastError());
  return MFAR_OK;
}

int launch_sparse_coo_write(const float* scores, long long ld, int Q, long long n_docs, const uint32_t* safe_bits,
                            const int* qids, long long doc_id_base, const long long* seg_offsets, int* out_keys,
                            void* out_vals, int vals_dtype, cudaStream_t st) {
  const long long segs = sparse_coo_segments(n_docs);
  if (segs > 0x7fffffffll || Q > 65535) return MFAR_ERR_SHAPE;
  dim3 grid((unsigned)segs, (unsigned)Q);
  const long long n_words = (doc_id_base + n_docs + 31) / 32;
  if (vals_dtype == MFAR_F16)
    sparse_coo_write_kernel<HalfRn><<<grid, kCooThreads, 0, st>>>(scores, ld, n_docs, safe_bits, n_words, qids,
                                                                  doc_id_base, seg_offsets,
                                                                  reinterpret_cast<int2*>(out_keys),
                                                                  static_cast<HalfRn*>(out_vals));
  else if (vals_dtype == MFAR_F32)
    sparse_coo_write_kernel<float><<<grid, kCooThreads, 0, st>>>(scores, ld, n_docs, safe_bits, n_words, qids,
                                                                 doc_id_base, seg_offsets,
                                                                 reinterpret_cast<int2*>(out_keys),
                                                                 static_cast<float*>(out_vals));
  else
    return MFAR_ERR_ARG;
  MFAR_CUDA_OK(cudaGetLastError());
  return MFAR_OK;
}

}  // namespace mfar
