// vrgdg_temporal.cuh — 3-frame temporal unsharp (BASELINE.json configs[4]).  LABELLED EXTENSION: the reference has no temporal
// operator (VRGDG_VideoEnhanceNodes.py holds no sharpen / blur / stencil), so the specification is this repository's:
//     out[t] = clamp( x[t] + s * ( x[t] - (x[t-1] + x[t] + x[t+1]) / 3 ), 0, 1 ),   x[-1] := x[0],  x[T] := x[T-1]
// evaluated in fp32 with one rounding per operation in exactly this order (the test suite holds a NumPy statement of the same
// formula; parity "unpinned": there is no reference output to be identical to).
// Streaming kernel, 16 bytes per thread and frame; blocks walk the clip in frame order, so the three reads of a frame (as next,
// current and previous frame of consecutive outputs) are served by L2 and HBM traffic stays at the algorithmic 1 read + 1 write.
#pragma once
#include "vrgdg_kernels.cuh"

namespace vrgdg {

struct TemporalParams {
  int B;
  int64_t frame_elems;        // H * W * 3
  float strength;
  const void* prev;           // frame before in[0] (halo from the previous shard) or null: replicate
  const void* next;           // frame after in[B-1] or null: replicate
};

template <typename T>
__device__ __forceinline__ float temporal_px(T p, T c, T n, float s) {
  const float fp = Elem<T>::ld(p), fc = Elem<T>::ld(c), fn = Elem<T>::ld(n);
  const float mean = divx(addx(addx(fp, fc), fn), 3.0f);
  return clamp01(addx(fc, mulx(s, subx(fc, mean))));
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
k_temporal3(const T* __restrict__ in, T* __restrict__ out, TemporalParams P, int chunks_per_frame, int64_t total_chunks) {
  constexpr int NE = VEC ? (int)(16 / sizeof(T)) : 1;
  const T* halo_prev = reinterpret_cast<const T*>(P.prev);
  const T* halo_next = reinterpret_cast<const T*>(P.next);
  for (int64_t vb = blockIdx.x; vb < total_chunks; vb += gridDim.x) {      // frame-major order: concurrent blocks share frames in L2
    const int t = (int)(vb / chunks_per_frame);
    const int64_t e0 = ((vb - (int64_t)t * chunks_per_frame) * 256 + threadIdx.x) * NE;
    if (e0 >= P.frame_elems) continue;
    const T* cur = in + (int64_t)t * P.frame_elems + e0;
    const T* prv = (t > 0) ? cur - P.frame_elems : (halo_prev ? halo_prev + e0 : cur);
    const T* nxt = (t + 1 < P.B) ? cur + P.frame_elems : (halo_next ? halo_next + e0 : cur);
    T* dst = out + (int64_t)t * P.frame_elems + e0;
    if (VEC) {
      union V { uint4 q; T e[NE]; } a, b, c, o;
      a.q = __ldg(reinterpret_cast<const uint4*>(prv));
      b.q = __ldg(reinterpret_cast<const uint4*>(cur));
      c.q = __ldg(reinterpret_cast<const uint4*>(nxt));
#pragma unroll
      for (int i = 0; i < NE; ++i) o.e[i] = Elem<T>::st(temporal_px<T>(a.e[i], b.e[i], c.e[i], P.strength));
      *reinterpret_cast<uint4*>(dst) = o.q;
    } else {
      dst[0] = Elem<T>::st(temporal_px<T>(prv[0], cur[0], nxt[0], P.strength));
    }
  }
}

template <typename T>
cudaError_t launch_temporal(const void* in, void* out, const TemporalParams& P, const LaunchCtx& ctx) {
  if (P.B == 0 || P.frame_elems == 0) return cudaSuccess;
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec = (P.frame_elems * (int64_t)sizeof(T)) % 16 == 0 && al(in) && al(out) && al(P.prev) && al(P.next);
  const int ne = vec ? (int)(16 / sizeof(T)) : 1;
  const int cpf = (int)((P.frame_elems + (int64_t)256 * ne - 1) / ((int64_t)256 * ne));
  const int64_t total = (int64_t)cpf * P.B;
  const int grid = (int)std::min<int64_t>(total, (int64_t)ctx.sms * 16);
  if (vec) k_temporal3<T, true><<<grid, 256, 0, ctx.stream>>>(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out), P, cpf, total);
  else k_temporal3<T, false><<<grid, 256, 0, ctx.stream>>>(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out), P, cpf, total);
  count_launch();
  return cudaGetLastError();
}

}  // namespace vrgdg
