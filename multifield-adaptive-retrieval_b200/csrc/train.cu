// Training-time scorer: the per-field query.doc components and the field mixture of the contrastive
// losses, forward and backward.  Replaces, for fp32 tensors already on the device,
//   DecomposedContrastiveLoss.compute_query_doc_field_components   mfar/modeling/losses.py:176-188
//   DecomposedContrastiveLoss.compute_doc_query_scores             mfar/modeling/losses.py:199-202
//   LinearWeights.forward (training mode, with autograd)           mfar/modeling/weighting.py:17-29
// Shapes are training-sized (B = 12..192 queries, N = a few thousand docs, F <= 44, E = 768): tens of MB and
// under a GFLOP per call, so these are latency/HBM-bound CUDA-core kernels (fp32 like the reference's parameters);
// the tensor-core kernels of score_tc.cu / score_qs.cu are for the corpus-sized contraction.
//
// Doc addressing: doc n = (p, s) with p = n / inner, s = n % inner lives at docs + p*stride_p + f*stride_f + s*stride_s,
// which covers d_pos [P,F,E] (inner = 1) and d_neg [P,F,Neg,E] (inner = Neg, doc order p*Neg + s exactly as
// d_neg.permute(0,2,1,3).view(1, P*Neg, F, E) orders them, losses.py:186) without a permuted copy.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace mfar {

constexpr int kTrWarps = 8;
constexpr int kTrThreads = kTrWarps * 32;
constexpr int kTrQTile = 32;          // queries per shared-memory tile (one per lane)
constexpr int kTrMaxChunks = 8;       // E <= 8 * 128

__device__ __forceinline__ const float* doc_row(const float* docs, long long n, int f, const DocLayout& L) {
  const long long p = n / L.inner, s = n - p * L.inner;
  return docs + p * L.stride_p + (long long)f * L.stride_f + s * L.stride_s;
}

__device__ __forceinline__ void load_q_tile(float* q_s, const float* __restrict__ q, int b0, int B, int E) {
  // q_s[32][E]; rows beyond B are zero
  const int n4 = E / 4;
  for (int i = threadIdx.x; i < kTrQTile * n4; i += kTrThreads) {
    const int b = i / n4, c = i - b * n4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b0 + b < B) v = __ldg(reinterpret_cast<const float4*>(q + (long long)(b0 + b) * E) + c);
    reinterpret_cast<float4*>(q_s)[i] = v;
  }
}

// comp[b, n, f] = <q[b], doc[n, f]> / temperature.   grid (ceil(N / 8), ceil(B / 32)); warp = one doc, lane = one query.
__global__ void __launch_bounds__(kTrThreads)
field_components_fwd_kernel(const float* __restrict__ q, int B, int E, const float* __restrict__ docs, long long N,
                            int F, DocLayout L, float temperature, float* __restrict__ comp) {
  extern __shared__ __align__(16) float q_s[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.y * kTrQTile;
  load_q_tile(q_s, q, b0, B, E);
  __syncthreads();
  const long long n = (long long)blockIdx.x * kTrWarps + warp;
  if (n >= N) return;
  const int n_chunks = E / 128, tail4 = (E % 128) / 4;      // lane owns float4 #(c*32 + lane); tail: lanes < tail4
  for (int f = 0; f < F; ++f) {
    const float4* row = reinterpret_cast<const float4*>(doc_row(docs, n, f, L));
    float4 r[kTrMaxChunks + 1];
#pragma unroll
    for (int c = 0; c <= kTrMaxChunks; ++c) {
      const bool live = c < n_chunks || (c == n_chunks && lane < tail4);
      r[c] = live ? __ldg(row + c * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float mine = 0.f;
    for (int b = 0; b < kTrQTile; ++b) {
      const float4* qb = reinterpret_cast<const float4*>(q_s + (long long)b * E);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c <= kTrMaxChunks; ++c) {
        if (c < n_chunks || (c == n_chunks && lane < tail4)) {
          const float4 v = qb[c * 32 + lane];
          acc = fmaf(r[c].x, v.x, acc); acc = fmaf(r[c].y, v.y, acc);
          acc = fmaf(r[c].z, v.z, acc); acc = fmaf(r[c].w, v.w, acc);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == b) mine = acc;
    }
    if (b0 + lane < B) comp[((long long)(b0 + lane) * N + n) * F + f] = mine / temperature;     // losses.py:184
  }
}

// ddocs[n, f, :] = (1/T) * sum_b dcomp[b, n, f] * q[b, :].   warp = one (doc, field) row; loops over query tiles.
__global__ void __launch_bounds__(kTrThreads)
field_components_bwd_docs_kernel(const float* __restrict__ q, int B, int E, long long N, int F, DocLayout L,
                                 float temperature, const float* __restrict__ dcomp, float* __restrict__ ddocs) {
  extern __shared__ __align__(16) float q_s[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row_id = (long long)blockIdx.x * kTrWarps + warp;       // n * F + f
  const bool live_row = row_id < N * F;
  const long long n = live_row ? row_id / F : 0;
  const int f = live_row ? int(row_id - n * F) : 0;
  const int n_chunks = E / 128, tail4 = (E % 128) / 4;
  float4 a[kTrMaxChunks + 1];
#pragma unroll
  for (int c = 0; c <= kTrMaxChunks; ++c) a[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b0 = 0; b0 < B; b0 += kTrQTile) {
    __syncthreads();
    load_q_tile(q_s, q, b0, B, E);
    __syncthreads();
    float coef = 0.f;
    if (live_row && b0 + lane < B) coef = __ldg(dcomp + ((long long)(b0 + lane) * N + n) * F + f) / temperature;
    for (int b = 0; b < kTrQTile; ++b) {
      const float cb = __shfl_sync(0xffffffffu, coef, b);
      if (cb == 0.f) continue;                                              // warp-uniform
      const float4* qb = reinterpret_cast<const float4*>(q_s + (long long)b * E);
#pragma unroll
      for (int c = 0; c <= kTrMaxChunks; ++c) {
        if (c < n_chunks || (c == n_chunks && lane < tail4)) {
          const float4 v = qb[c * 32 + lane];
          a[c].x = fmaf(cb, v.x, a[c].x); a[c].y = fmaf(cb, v.y, a[c].y);
          a[c].z = fmaf(cb, v.z, a[c].z); a[c].w = fmaf(cb, v.w, a[c].w);
        }
      }
    }
  }
  if (!live_row) return;
  float4* out = reinterpret_cast<float4*>(const_cast<float*>(doc_row(ddocs, n, f, L)));
#pragma unroll
  for (int c = 0; c <= kTrMaxChunks; ++c)
    if (c < n_chunks || (c == n_chunks && lane < tail4)) out[c * 32 + lane] = a[c];
}

// dq[b, :] += (1/T) * sum_{n,f} dcomp[b, n, f] * doc[n, f, :].   grid (row slices, query tiles): a thread owns one
// float4 column of E for the 32 queries of its tile and walks its slice of (doc, field) rows; partial sums are added
// to dq with fp32 atomics (dq is zeroed by the launcher).
constexpr int kDqRows = 16;           // rows whose coefficients are staged per step
__global__ void __launch_bounds__(kTrThreads)
field_components_bwd_q_kernel(const float* __restrict__ docs, int B, int E, long long N, int F, DocLayout L,
                              float temperature, const float* __restrict__ dcomp, long long rows_per_cta,
                              float* __restrict__ dq) {
  __shared__ float c_s[kDqRows][kTrQTile];
  const int b0 = blockIdx.y * kTrQTile;
  const long long rows = N * F;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  const int n4 = E / 4;
  const bool live_col = int(threadIdx.x) < n4;                // E <= 1024: one float4 column per thread
  float4 acc[kTrQTile];
#pragma unroll
  for (int b = 0; b < kTrQTile; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long rb = r0; rb < r1; rb += kDqRows) {
    __syncthreads();
    for (int i = threadIdx.x; i < kDqRows * kTrQTile; i += kTrThreads) {
      const int b = i / kDqRows, rr = i - b * kDqRows;         // consecutive threads: consecutive rows of one query
      float v = 0.f;
      if (rb + rr < r1 && b0 + b < B) v = __ldg(dcomp + (long long)(b0 + b) * rows + rb + rr) / temperature;
      c_s[rr][b] = v;
    }
    __syncthreads();
    const int nr = int(min((long long)kDqRows, r1 - rb));
    for (int rr = 0; rr < nr; ++rr) {
      const long long row_id = rb + rr;
      const long long n = row_id / F;
      const int f = int(row_id - n * F);
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live_col) d = __ldg(reinterpret_cast<const float4*>(doc_row(docs, n, f, L)) + threadIdx.x);
#pragma unroll
      for (int b = 0; b < kTrQTile; ++b) {
        const float cb = c_s[rr][b];
        acc[b].x = fmaf(cb, d.x, acc[b].x); acc[b].y = fmaf(cb, d.y, acc[b].y);
        acc[b].z = fmaf(cb, d.z, acc[b].z); acc[b].w = fmaf(cb, d.w, acc[b].w);
      }
    }
  }
  if (!live_col) return;
#pragma unroll
  for (int b = 0; b < kTrQTile; ++b) {
    if (b0 + b < B) {
      float* o = dq + (long long)(b0 + b) * E + threadIdx.x * 4;
      atomicAdd(o, acc[b].x); atomicAdd(o + 1, acc[b].y); atomicAdd(o + 2, acc[b].z); atomicAdd(o + 3, acc[b].w);
    }
  }
}

// Mixture backward, step 1: one CTA per query b.
//   dx[b,s,f] = g[b,s] * w[b,f];  dwt[f] = sum_s g[b,s] * x[b,s,f];  dlogit[b,f] = w[b,f] * (dwt[f] - sum_f' w[b,f'] dwt[f'])
__global__ void __launch_bounds__(256)
mixture_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int w_rows, const float* __restrict__ g,
                   int S, int F, float* __restrict__ dx, float* __restrict__ dlogit) {
  __shared__ float w_s[MFAR_MAX_FIELDS];
  __shared__ float red[MFAR_MAX_FIELDS][8];
  __shared__ float dwt_s[MFAR_MAX_FIELDS];
  const int b = blockIdx.x;
  const float* wb = w + (w_rows == 1 ? 0 : (long long)b * F);
  if (threadIdx.x < F) w_s[threadIdx.x] = wb[threadIdx.x];
  __syncthreads();
  // thread t handles field f = t % F' of rows s = t / F', F' = F rounded so that 256 % F' == 0 is not needed:
  // flat index i = s*F + f walks the row block contiguously (coalesced), partial sums land in per-field bins
  const long long base = (long long)b * S * F;
  const long long total = (long long)S * F;
  // each thread keeps the fields constant along its walk when the stride is a multiple of F
  const int stride = (256 / F) * F;                                   // <= 256, multiple of F
  float part = 0.f;
  const int f_mine = threadIdx.x % F;
  if (int(threadIdx.x) < stride) {
    for (long long i = threadIdx.x; i < total; i += stride) {
      const long long s = i / F;
      const float gv = __ldg(g + (long long)b * S + s);
      const float xv = __ldg(x + base + i);
      if (dx) dx[base + i] = gv * w_s[f_mine];
      part = fmaf(gv, xv, part);
    }
  }
  // reduce the partial sums of the threads that share a field
  for (int f = 0; f < F; ++f) {
    float v = (f_mine == f && int(threadIdx.x) < stride) ? part : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[f][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < F) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[threadIdx.x][k];
    dwt_s[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x < F) {
    float dot = 0.f;
    for (int f = 0; f < F; ++f) dot = fmaf(w_s[f], dwt_s[f], dot);
    dlogit[(long long)b * F + threadIdx.x] = w_s[threadIdx.x] * (dwt_s[threadIdx.x] - dot);
  }
}

// Mixture backward, step 2.  query_cond: dW[e,f] = sum_b q[b,e] dlogit[b,f];  dq[b,e] = sum_f dlogit[b,f] W[e,f].
// One thread per embedding column e.  Not query_cond (W is [F,1]): dW[f] = sum_b dlogit[b,f].
__global__ void mixture_bwd_params_kernel(const float* __restrict__ q, const float* __restrict__ W,
                                          const float* __restrict__ dlogit, int B, int E, int F, int query_cond,
                                          float* __restrict__ dW, float* __restrict__ dq) {
  if (!query_cond) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float v = 0.f;
    for (int b = 0; b < B; ++b) v += dlogit[(long long)b * F + f];
    dW[f] = v;
    return;
  }
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float we[MFAR_MAX_FIELDS], dwe[MFAR_MAX_FIELDS];
  for (int f = 0; f < F; ++f) { we[f] = W[(long long)e * F + f]; dwe[f] = 0.f; }
  for (int b = 0; b < B; ++b) {
    const float qv = q[(long long)b * E + e];
    float dqv = 0.f;
    for (int f = 0; f < F; ++f) {
      const float dl = __ldg(dlogit + (long long)b * F + f);
      dwe[f] = fmaf(qv, dl, dwe[f]);
      dqv = fmaf(dl, we[f], dqv);
    }
    if (dq) dq[(long long)b * E + e] = dqv;
  }
  for (int f = 0; f < F; ++f) dW[(long long)e * F + f] = dwe[f];
}

static size_t q_tile_bytes(int E) { return size_t(kTrQTile) * E * sizeof(float); }

int launch_field_components_fwd(const float* q, int B, int E, const float* docs, long long N, int F,
                                const DocLayout& L, float temperature, float* comp, cudaStream_t st) {
  const size_t smem = q_tile_bytes(E);
  MFAR_CUDA_OK(cudaFuncSetAttribute(field_components_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  dim3 grid(unsigned((N + kTrWarps - 1) / kTrWarps), unsigned((B + kTrQTile - 1) / kTrQTile));
  field_components_fwd_kernel<<<grid, kTrThreads, smem, st>>>(q, B, E, docs, N, F, L, temperature, comp);
  MFAR_CUDA_OK(cudaGetLastError());
  return MFAR_OK;
}

int launch_field_components_bwd(const float* q, int B, int E, const float* docs, long long N, int F,
                                const DocLayout& L, float temperature, const float* dcomp, float* dq, float* ddocs,
                                cudaStream_t st) {
  if (ddocs) {
    const size_t smem = q_tile_bytes(E);
    MFAR_CUDA_OK(cudaFuncSetAttribute(field_components_bwd_docs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(smem)));
    const unsigned blocks = unsigned((N * F + kTrWarps - 1) / kTrWarps);
    field_components_bwd_docs_kernel<<<blocks, kTrThreads, smem, st>>>(q, B, E, N, F, L, temperature, dcomp, ddocs);
    MFAR_CUDA_OK(cudaGetLastError());
  }
  if (dq) {
    MFAR_CUDA_OK(cudaMemsetAsync(dq, 0, size_t(B) * E * sizeof(float), st));
    const long long rows = N * F;
    long long slices = std::min<long long>(2LL * kNumSmsB200, (rows + kDqRows - 1) / kDqRows);
    if (slices < 1) slices = 1;
    long long rows_per_cta = (rows + slices - 1) / slices;
    rows_per_cta = (rows_per_cta + kDqRows - 1) / kDqRows * kDqRows;
    dim3 grid(unsigned((rows + rows_per_cta - 1) / rows_per_cta), unsigned((B + kTrQTile - 1) / kTrQTile));
    field_components_bwd_q_kernel<<<grid, kTrThreads, 0, st>>>(docs, B, E, N, F, L, temperature, dcomp, rows_per_cta, dq);
    MFAR_CUDA_OK(cudaGetLastError());
  }
  return MFAR_OK;
}

int launch_mixture_bwd(const float* x, const float* q, const float* W, const float* w, int w_rows, const float* g,
                       int B, int S, int E, int F, int query_cond, float* dx, float* dW, float* dq, float* dlogit,
                       cudaStream_t st) {
  mixture_bwd_kernel<<<B, 256, 0, st>>>(x, w, w_rows, g, S, F, dx, dlogit);
  MFAR_CUDA_OK(cudaGetLastError());
  const int n = query_cond ? E : F;
  mixture_bwd_params_kernel<<<(n + 127) / 128, 128, 0, st>>>(q, W, dlogit, B, E, F, query_cond, dW, dq);
  MFAR_CUDA_OK(cudaGetLastError());
  return MFAR_OK;
}

}  // namespace mfar
