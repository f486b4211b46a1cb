// BM25 sparse-field scorer: the device replacement of bm25s.BM25.get_scores as the reference calls it
// (mfar/data/index.py:72-76, 111-118) and of the score-matrix arithmetic of bm25s.BM25.index (index.py:138-140).
//
// Index layout in HBM, per sparse field (what bm25s keeps as a scipy CSC matrix): token-major postings
//   indptr  int64 [V+1]   postings of token t are [indptr[t], indptr[t+1])
//   indices int32 [nnz]   LOCAL doc row of each posting, ascending inside a token
//   data    fp32  [nnz]   idf(t) * tfc(t, doc), precomputed at index time
// A query batch is a flat list of (query row, sparse field, token id) entries, one per query-token occurrence
// (repeated tokens are repeated entries: bm25s adds their postings once per occurrence).
//
// Scoring = scatter-add.  base[q, doc] += w[q, F_d + j] * data[p] for every posting p of every entry (q, j, t).
// The work list is the concatenation of the entries' postings ranges: a one-CTA plan kernel turns the entries into
// an exclusive prefix sum over their postings counts, and the scatter kernel cuts that flat range into equal chunks
// so that a 3-posting rare token and a 300k-posting common token load the SMs evenly.  HBM-bound: 8 B read per
// posting (coalesced, streamed once) + one fp32 RED per posting that resolves in L2.
#include "common.cuh"
#include "kernels.h"

namespace mfar {

constexpr int kPlanThreads = 1024;
constexpr int kScatterThreads = 256;
constexpr int kScatterPerThread = 16;
constexpr int kScatterBatch = 8;                                      // loads in flight per thread
constexpr int kScatterChunk = kScatterThreads * kScatterPerThread;   // postings per CTA iteration

// entries -> (first posting, flat start) ; flat_start[n_entries] = total postings of the batch
__global__ void __launch_bounds__(kPlanThreads)
bm25_plan_kernel(const int* __restrict__ entries, long long n_entries, Bm25Fields f, int n_sparse, int Q,
                 long long* __restrict__ ent_first, long long* __restrict__ flat_start) {
  __shared__ long long warp_sum[kPlanThreads / 32];
  __shared__ long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (long long base = 0; base < n_entries; base += kPlanThreads) {
    const long long e = base + tid;
    long long len = 0, first = 0;
    if (e < n_entries) {
      const int q = entries[3 * e], j = entries[3 * e + 1], t = entries[3 * e + 2];
      if (q >= 0 && q < Q && j >= 0 && j < n_sparse && t >= 0 && t < f.vocab[j]) {
        first = f.indptr[j][t];
        len = f.indptr[j][t + 1] - first;
        if (len < 0) len = 0;
      }
    }
    long long incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      long long s = warp_sum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long v = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += v;
      }
      warp_sum[lane] = s;      // inclusive over warps
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long excl = carry + (wid ? warp_sum[wid - 1] : 0) + incl - len;
    if (e < n_entries) {
      ent_first[e] = first;
      flat_start[e] = excl;
    }
    __syncthreads();
    if (tid == kPlanThreads - 1) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
  if (tid == 0) flat_start[n_entries] = carry_s;
}

// last entry e in [lo, hi] with flat_start[e] <= pos   (flat_start is non-decreasing; zero-length entries are
// skipped because the LAST such entry is the one whose range contains pos)
__device__ __forceinline__ long long find_entry(const long long* __restrict__ flat_start, long long lo, long long hi,
                                                long long pos) {
  while (lo < hi) {
    const long long mid = lo + ((hi - lo + 1) >> 1);
    if (__ldg(flat_start + mid) <= pos) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kScatterThreads)
bm25_scatter_kernel(const int* __restrict__ entries, long long n_entries, const long long* __restrict__ ent_first,
                    const long long* __restrict__ flat_start, Bm25Fields f, const float* __restrict__ w, int w_ld,
                    int w_off, long long n_docs, float* __restrict__ base, long long base_ld) {
  __shared__ long long range_s[2];
  const long long total = __ldg(flat_start + n_entries);
  for (long long c0 = (long long)blockIdx.x * kScatterChunk; c0 < total; c0 += (long long)gridDim.x * kScatterChunk) {
    const long long c1 = min(total, c0 + kScatterChunk);
    if (threadIdx.x < 2)
      range_s[threadIdx.x] = find_entry(flat_start, 0, n_entries - 1, threadIdx.x == 0 ? c0 : c1 - 1);
    __syncthreads();
    long long e = range_s[0];
    const long long e_hi = range_s[1];
    long long e_begin = -1, e_end = -1, p_first = 0;    // cached entry: flat range and first posting
    int q = 0, j = 0;
    float wq = 0.f;
    // Two phases per batch of kScatterBatch postings: resolve (entry, posting address) for the whole batch, issue
    // all its loads back to back, then the REDs - the entry-boundary branch must not sit between the loads or each
    // warp keeps only one pair of loads in flight and the kernel is DRAM-latency bound (ncu: 34 % of DRAM peak).
#pragma unroll 1
    for (int i0 = 0; i0 < kScatterPerThread; i0 += kScatterBatch) {
      const int* ip[kScatterBatch];
      const float* dp[kScatterBatch];
      float* bp[kScatterBatch];
      float wv[kScatterBatch];
#pragma unroll
      for (int u = 0; u < kScatterBatch; ++u) {
        const long long pos = c0 + (long long)(i0 + u) * kScatterThreads + threadIdx.x;
        ip[u] = nullptr;
        if (pos < c1) {
          if (pos >= e_end) {
            e = find_entry(flat_start, e, e_hi, pos);
            e_begin = __ldg(flat_start + e);
            e_end = __ldg(flat_start + e + 1);
            p_first = __ldg(ent_first + e);
            q = __ldg(entries + 3 * e);
            j = __ldg(entries + 3 * e + 1);
            wq = w ? __ldg(w + (long long)q * w_ld + w_off + j) : 1.f;
          }
          const long long p = p_first + (pos - e_begin);
          ip[u] = f.indices[j] + p;
          dp[u] = f.data[j] + p;
          bp[u] = base + (long long)q * base_ld;
          wv[u] = wq;
        }
      }
      int doc[kScatterBatch];
      float val[kScatterBatch];
#pragma unroll
      for (int u = 0; u < kScatterBatch; ++u) {
        doc[u] = -1; val[u] = 0.f;
        if (ip[u] != nullptr) { doc[u] = __ldcs(ip[u]); val[u] = __ldcs(dp[u]); }
      }
#pragma unroll
      for (int u = 0; u < kScatterBatch; ++u)
        if (doc[u] >= 0 && doc[u] < n_docs) atomicAdd(bp[u] + doc[u], wv[u] * val[u]);
    }
    __syncthreads();
  }
}

// score-matrix arithmetic of bm25s.BM25.index (method "lucene"): one posting per thread, float64 like numpy
__global__ void bm25_build_scores_kernel(const int* __restrict__ post_token, const int* __restrict__ post_doc,
                                         const int* __restrict__ post_tf, long long nnz, const int* __restrict__ df,
                                         const int* __restrict__ doc_len, long long n_docs_total, double l_avg,
                                         double k1, double b, float* __restrict__ data) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  // explicit round-to-nearest intrinsics: no FMA contraction, so every step rounds like numpy's float64 ufuncs
  const double dfv = double(df[post_token[p]]);
  const float idf = float(log(__dadd_rn(1.0, __ddiv_rn(double(n_docs_total) - dfv + 0.5, dfv + 0.5))));
  const double tf = double(post_tf[p]);
  const double norm = __dadd_rn(1.0 - b, __ddiv_rn(__dmul_rn(b, double(doc_len[post_doc[p]])), l_avg));
  const double tfc = __ddiv_rn(tf, __dadd_rn(__dmul_rn(k1, norm), tf));
  data[p] = float(__dmul_rn(double(idf), tfc));
}

int launch_bm25_plan(const int* entries, long long n_entries, const Bm25Fields& f, int n_sparse, int Q,
                     long long* ent_first, long long* flat_start, cudaStream_t st) {
  bm25_plan_kernel<<<1, kPlanThreads, 0, st>>>(entries, n_entries, f, n_sparse, Q, ent_first, flat_start);
  MFAR_CUDA_OK(cudaGetLastError());
  return MFAR_OK;
}

int launch_bm25_scatter(const int* entries, long long n_entries, const long long* ent_first,
                        const long long* flat_start, const Bm25Fields& f, const float* w, int w_ld, int w_off,
                        long long n_docs, float* base, long long base_ld, cudaStream_t st) {
  if (n_entries <= 0) return MFAR_OK;
  bm25_scatter_kernel<<<kNumSmsB200 * 8, kScatterThreads, 0, st>>>(entries, n_entries, ent_first, flat_start, f, w,
                                                                  w_ld, w_off, n_docs, base, base_ld);
  MFAR_CUDA_OK(cudaGetLastError());
  return MFAR_OK;
}

int launch_bm25_build_scores(const int* post_token, const int* post_doc, const int* post_tf, long long nnz,
                             const int* df, const int* doc_len, long long n_docs_total, double l_avg, double k1,
                             double b, float* data, cudaStream_t st) {
  if (nnz <= 0) return MFAR_OK;
  bm25_build_scores_kernel<<<unsigned((nnz + 255) / 256), 256, 0, st>>>(post_token, post_doc, post_tf, nnz, df,
                                                                        doc_len, n_docs_total, l_avg, k1, b, data);
  MFAR_CUDA_OK(cudaGetLastError());
  return MFAR_OK;
}

}  // namespace mfar
