// vrgdg_histmatch.cuh — histogram / CDF colour transfer (BASELINE.json north_star: "two-pass per-channel histogram + monotone-CDF
// LUT mapping", configs[2] "histogram CDF transfer").  LABELLED EXTENSION: the reference's ColorMatchToReference is a LAB mean/std
// transfer (nodes.py:97-115) and contains no histogram anywhere, so the specification below is this repository's and
// parity is "unpinned" (a NumPy statement of the same formulas lives in the test suite).
//
//   bins      : per channel (R, G, B) 256 bins over [0,1]:  bin(v) = min(floor(clamp01(v) * 256), 255)
//   pass 1    : per-frame counts h_f[c][k] and reference counts h_r[c][k] (exact integers; the reference's rows can be sharded over
//               ranks and the partial counts added: one all-gather of 3 x 256 uint32 per rank)
//   tables    : both CDFs are piecewise linear over the bin EDGES (g_i = C_r[i-1] / N_r at edge i / 256, C = inclusive cumulative counts,
//               C[-1] = 0); T[k] = G_r^-1(q) with q = C_f[k-1] / N_f for the 257 source edges k = 0..256, all comparisons on exact
//               integers (C_r * N_f vs C_f * N_r):
//                 q strictly inside bin j's rising segment: T[k] = (j + (q - g_j) / (g_{j+1} - g_j)) / 256, fp64, one rounding per op;
//                 q equal to edge values g_lo .. g_hi (a plateau of empty bins, or a single edge): T[k] = clamp(k, lo, hi) / 256, the
//                 inverse closest to the source edge - a frame matched to itself is mapped to itself exactly.
//               Monotone non-decreasing, T[0] >= 0, T[256] = 1, rounded once to fp32.
//   pass 2    : u = clamp01(v) * 256, k = min(floor(u), 255), w = u - k;  m = T[k] + w * (T[k+1] - T[k])  (one FMA);
//               out = clamp01(v * (1 - t) + m * t) with the strength t, one rounding per operation.
#pragma once
#include "vrgdg_kernels.cuh"

namespace vrgdg {

constexpr int HIST_BINS = 256;

__device__ __forceinline__ int hist_bin(float v) {
  const float u = clamp01(v) * 256.0f;
  const int k = (int)u;                       // u >= 0: truncation = floor
  return k < 255 ? k : 255;
}

// counts[frame][c][k] += ... (uint32, zeroed by the caller of the launcher); one privatised histogram per WARP in shared memory
// (8 x 3 x 256 counters = 24 KB), merged into global memory with one atomic per non-empty bin and block.
template <typename T>
__global__ void __launch_bounds__(256)
k_hist_counts(const T* __restrict__ in, int W, int64_t hw, int row0, int rows, uint32_t* __restrict__ counts) {
  constexpr bool BGR = Io<T>::BGR;
  __shared__ uint32_t sh[8][3][HIST_BINS];
  for (int i = threadIdx.x; i < 8 * 3 * HIST_BINS; i += 256) (&sh[0][0][0])[i] = 0u;
  __syncthreads();
  const int frame = blockIdx.y, wid = threadIdx.x >> 5;
  const T* fbase = in + (int64_t)frame * hw * 3;
  const int64_t pbeg = (int64_t)row0 * W, n = (int64_t)rows * W;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const T* s = fbase + (pbeg + i) * 3;
    const float r = Elem<T>::ld(s[BGR ? 2 : 0]), g = Elem<T>::ld(s[1]), b = Elem<T>::ld(s[BGR ? 0 : 2]);
    atomicAdd(&sh[wid][0][hist_bin(r)], 1u);
    atomicAdd(&sh[wid][1][hist_bin(g)], 1u);
    atomicAdd(&sh[wid][2][hist_bin(b)], 1u);
  }
  __syncthreads();
  uint32_t* dst = counts + (int64_t)frame * 3 * HIST_BINS;
  for (int i = threadIdx.x; i < 3 * HIST_BINS; i += 256) {
    uint32_t v = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += (&sh[w][0][0])[i];
    if (v) atomicAdd(dst + i, v);
  }
}

// tables[frame][c][k] = {T[k], T[k+1] - T[k]} for k = 0..255 (float2): one 8-byte shared-memory load per channel in pass 2.
// One block per (frame, channel), 256 threads; exact integer CDFs, fp64 arithmetic with one rounding per operation.
static __global__ void __launch_bounds__(256)
k_hist_tables(const uint32_t* __restrict__ fcounts, const uint32_t* __restrict__ rcounts, int n_ref, float2* __restrict__ tables) {
  __shared__ unsigned long long cf[HIST_BINS], cr[HIST_BINS];
  __shared__ float tv[HIST_BINS + 1];
  const int frame = blockIdx.x / 3, c = blockIdx.x - frame * 3, k = threadIdx.x;
  const uint32_t* hf = fcounts + ((int64_t)frame * 3 + c) * HIST_BINS;
  const uint32_t* hr = rcounts + ((int64_t)(n_ref == 1 ? 0 : frame) * 3 + c) * HIST_BINS;
  if (k == 0) {                                        // 256-term inclusive scans: serial, exact, negligible
    unsigned long long a = 0, b = 0;
    for (int i = 0; i < HIST_BINS; ++i) { a += hf[i]; b += hr[i]; cf[i] = a; cr[i] = b; }
  }
  __syncthreads();
  const unsigned long long nf = cf[HIST_BINS - 1], nr = cr[HIST_BINS - 1];
  for (int e = k; e <= HIST_BINS; e += 256) {          // edge e: cumulative source mass below it
    const unsigned long long cq = (e == 0) ? 0ull : cf[e - 1];
    float t;
    if (nf == 0 || nr == 0) {
      t = (float)((double)e / 256.0);                  // empty frame or reference: identity
    } else {
      // G(i/256) = g_i = C_r[i-1] / N_r at the 257 edges i (g_0 = 0, g_256 = 1); all comparisons exact: g_i >= q <=> C_r[i-1] * N_f >= cq * N_r
      const unsigned long long rhs = cq * nr;
      auto g = [&](int i) -> unsigned long long { return (i == 0 ? 0ull : cr[i - 1]) * nf; };
      int lo = 0, hi = HIST_BINS;                      // i_lo = smallest edge with g >= q   (g_256 * nf = nr * nf >= rhs)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (g(mid) >= rhs) hi = mid; else lo = mid + 1;
      }
      const int ilo = lo;
      if (g(ilo) == rhs) {
        // q is taken ON edges: every t in the plateau [i_lo, i_hi] / 256 inverts it; take the one closest to the source edge, so that a
        // frame matched to itself maps every edge to itself (identity) and the map stays monotone
        int a = ilo, b = HIST_BINS;                    // i_hi = largest edge with g <= q
        while (a < b) {
          const int mid = (a + b + 1) >> 1;
          if (g(mid) <= rhs) a = mid; else b = mid - 1;
        }
        const int ihi = a;
        const int ec = e < ilo ? ilo : (e > ihi ? ihi : e);
        t = (float)((double)ec / 256.0);
      } else {                                         // g_{ilo-1} < q < g_ilo: inside the rising segment of bin ilo - 1
        const int j = ilo - 1;
        const double q = __ddiv_rn((double)cq, (double)nf);
        const double prev = (j > 0) ? __ddiv_rn((double)cr[j - 1], (double)nr) : 0.0;
        const double cur = __ddiv_rn((double)cr[j], (double)nr);
        const double frac = __ddiv_rn(__dsub_rn(q, prev), __dsub_rn(cur, prev));
        t = (float)__ddiv_rn(__dadd_rn((double)j, frac), 256.0);
      }
    }
    tv[e] = t;
  }
  __syncthreads();
  tables[((int64_t)frame * 3 + c) * HIST_BINS + k] = make_float2(tv[k], __fsub_rn(tv[k + 1], tv[k]));
}

__device__ __forceinline__ float hist_map(const float2* __restrict__ tab, float v, float t, float omt) {
  const float u = clamp01(v) * 256.0f;
  int k = (int)u;
  k = k < 255 ? k : 255;
  const float w = __fsub_rn(u, (float)k);
  const float2 e = tab[k];
  const float m = __fmaf_rn(w, e.y, e.x);
  return clamp01(__fadd_rn(__fmul_rn(v, omt), __fmul_rn(m, t)));
}

// pass 2: streaming; the frame's three tables (6 KB) are staged in shared memory per block
template <typename T>
__global__ void __launch_bounds__(256)
k_histmatch_apply(const T* __restrict__ in, T* __restrict__ out, int64_t hw, const float2* __restrict__ tables, float t, float omt,
                  int blocks_per_frame) {
  constexpr bool BGR = Io<T>::BGR;
  __shared__ float2 tab[3][HIST_BINS];
  const int frame = blockIdx.x / blocks_per_frame, bif = blockIdx.x - frame * blocks_per_frame;
  const float2* src = tables + (int64_t)frame * 3 * HIST_BINS;
  for (int i = threadIdx.x; i < 3 * HIST_BINS; i += 256) (&tab[0][0])[i] = src[i];
  __syncthreads();
  const T* fin = in + (int64_t)frame * hw * 3;
  T* fout = out + (int64_t)frame * hw * 3;
  for (int64_t p = (int64_t)bif * 256 + threadIdx.x; p < hw; p += (int64_t)blocks_per_frame * 256) {
    const T* s = fin + p * 3;
    T* d = fout + p * 3;
    const float r = Elem<T>::ld(s[BGR ? 2 : 0]), g = Elem<T>::ld(s[1]), b = Elem<T>::ld(s[BGR ? 0 : 2]);
    d[BGR ? 2 : 0] = Elem<T>::st(hist_map(tab[0], r, t, omt));
    d[1] = Elem<T>::st(hist_map(tab[1], g, t, omt));
    d[BGR ? 0 : 2] = Elem<T>::st(hist_map(tab[2], b, t, omt));
  }
}

template <typename T>
cudaError_t launch_hist_counts(const void* in, int B, int H, int W, int row0, int rows, uint32_t* counts, const LaunchCtx& ctx) {
  if (B == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)B * 3 * HIST_BINS * sizeof(uint32_t), ctx.stream);
  if (e != cudaSuccess) return e;
  const int64_t n = (int64_t)rows * W;
  int bx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)std::max(1, ctx.sms * 8 / std::max(1, std::min(B, 8))));
  if (bx < 1) bx = 1;
  dim3 grid(bx, B);
  k_hist_counts<T><<<grid, 256, 0, ctx.stream>>>(reinterpret_cast<const T*>(in), W, (int64_t)H * W, row0, rows, counts);
  count_launch();
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_histmatch_apply(const void* in, void* out, int B, int64_t hw, const float2* tables, float t, float omt, const LaunchCtx& ctx) {
  if (B == 0 || hw == 0) return cudaSuccess;
  int bpf = (int)std::min<int64_t>((hw + 255) / 256, (int64_t)std::max(1, ctx.sms * 16 / std::max(1, std::min(B, 16))));
  if (bpf < 1) bpf = 1;
  const int64_t blocks = (int64_t)bpf * B;
  if (blocks >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
  k_histmatch_apply<T><<<(unsigned)blocks, 256, 0, ctx.stream>>>(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out), hw, tables, t, omt, bpf);
  count_launch();
  return cudaGetLastError();
}

}  // namespace vrgdg
