// vrgdg_lanczos.cuh — the enhancer's resize of uint8 frames (_resize_frames, VRGDG_StandaloneVideoEnhancerNodes.py:213-230 =
// cv2.resize(..., INTER_LANCZOS4)).  OpenCV's 8-bit path is fixed point: 8 taps per axis, weights rounded to shorts (x 2048),
// a horizontal pass into int32, a vertical pass, then (v + 2^21) >> 22 saturated to a byte; taps beyond the frame replicate the
// border.  The weight tables come from the host (vrgdg_lanczos4_tables: OpenCV's float/double recipe incl. libm sin/cos, which
// device code could not reproduce bit for bit); the kernels are pure integer work, so the result is bit-identical to cv2's.
//
//   k_lanczos_h: [B,Hs,Ws,3] u8  -> [B,Hs,Wd,3] int32 (scratch)     one thread per intermediate pixel, 24 byte loads (L1 hits)
//   k_lanczos_v: scratch         -> [B,Hd,Wd,3] u8                  one thread per 4 output bytes, 8 x 128-bit loads
// Algorithmic bytes per output pixel: 3 written + 3 * (Hs*Ws)/(Hd*Wd) read; the int32 intermediate adds 12 * Hs/Hd written and
// read back (mostly through L2).  Bound: HBM / L2 on the intermediate.
#pragma once
#include "vrgdg_kernels.cuh"

namespace vrgdg {

struct LanczosParams {
  int B, Hs, Ws, Hd, Wd;
  const int32_t* xofs;      // [Wd]     source column of tap 3
  const int16_t* xcoef;     // [Wd][8]
  const int32_t* yofs;      // [Hd]
  const int16_t* ycoef;     // [Hd][8]
};

cudaError_t launch_lanczos4(const uint8_t* in, uint8_t* out, int32_t* mid, const LanczosParams& L, const LaunchCtx& ctx);

#ifdef VRGDG_LANCZOS_IMPL   // defined by the one translation unit that owns these kernels (vrgdg_u8.cu)
static __global__ void __launch_bounds__(256) k_lanczos_h(const uint8_t* __restrict__ in, int32_t* __restrict__ mid, const LanczosParams L) {
  const int64_t total = (int64_t)L.B * L.Hs * L.Wd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int dx = (int)(i % L.Wd);
    const int64_t row = i / L.Wd;                                   // b * Hs + y
    const uint8_t* src = in + row * (int64_t)L.Ws * 3;
    const int sx = __ldg(L.xofs + dx);
    const int4 cw = __ldg(reinterpret_cast<const int4*>(L.xcoef) + dx);
    const int w[4] = {cw.x, cw.y, cw.z, cw.w};
    int a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = (int)(short)((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xFFFF));
      const int xi = max(0, min(sx - 3 + k, L.Ws - 1));
      const uint8_t* p = src + xi * 3;
      a0 += (int)__ldg(p) * c; a1 += (int)__ldg(p + 1) * c; a2 += (int)__ldg(p + 2) * c;
    }
    int32_t* d = mid + i * 3;
    d[0] = a0; d[1] = a1; d[2] = a2;
  }
}

__device__ __forceinline__ int lanczos_fix(int v) {                 // FixedPtCast<int, uchar, 22>
  return max(0, min((v + (1 << 21)) >> 22, 255));
}
__device__ __forceinline__ int mac_wrap(int acc, int v, int c) {    // OpenCV accumulates in int: keep its wrap-around defined
  return (int)((unsigned)acc + (unsigned)v * (unsigned)c);
}

template <int VEC>
static __global__ void __launch_bounds__(256) k_lanczos_v(const int32_t* __restrict__ mid, uint8_t* __restrict__ out, const LanczosParams L) {
  const int RE = L.Wd * 3;                                           // elements per row
  const int RV = (RE + VEC - 1) / VEC;
  const int64_t total = (int64_t)L.B * L.Hd * RV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ev = (int)(i % RV);
    const int64_t q = i / RV;
    const int dy = (int)(q % L.Hd);
    const int b = (int)(q / L.Hd);
    const int sy = __ldg(L.yofs + dy);
    const int4 cw = __ldg(reinterpret_cast<const int4*>(L.ycoef) + dy);
    const int w[4] = {cw.x, cw.y, cw.z, cw.w};
    const int32_t* base = mid + (int64_t)b * L.Hs * RE + (int64_t)ev * VEC;
    int acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = (int)(short)((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xFFFF));
      const int yi = max(0, min(sy - 3 + k, L.Hs - 1));
      const int32_t* p = base + (int64_t)yi * RE;
      if (VEC == 4) {
        const int4 v = *reinterpret_cast<const int4*>(p);
        acc[0] = mac_wrap(acc[0], v.x, c); acc[1] = mac_wrap(acc[1], v.y, c);
        acc[2] = mac_wrap(acc[2], v.z, c); acc[3] = mac_wrap(acc[3], v.w, c);
      } else {
        acc[0] = mac_wrap(acc[0], *p, c);
      }
    }
    uint8_t* d = out + ((int64_t)b * L.Hd + dy) * RE + (int64_t)ev * VEC;
    if (VEC == 4) {
      const uint32_t word = (uint32_t)lanczos_fix(acc[0]) | ((uint32_t)lanczos_fix(acc[1]) << 8) | ((uint32_t)lanczos_fix(acc[2]) << 16) |
                            ((uint32_t)lanczos_fix(acc[3]) << 24);
      *reinterpret_cast<uint32_t*>(d) = word;
    } else {
      *d = (uint8_t)lanczos_fix(acc[0]);
    }
  }
}

cudaError_t launch_lanczos4(const uint8_t* in, uint8_t* out, int32_t* mid, const LanczosParams& L, const LaunchCtx& ctx) {
  const int64_t nh = (int64_t)L.B * L.Hs * L.Wd;
  if (nh == 0 || (int64_t)L.B * L.Hd * L.Wd == 0) return cudaSuccess;
  const int gh = (int)std::min<int64_t>((nh + 255) / 256, (int64_t)ctx.sms * 32);
  k_lanczos_h<<<gh, 256, 0, ctx.stream>>>(in, mid, L);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const bool vec = ((L.Wd * 3) & 3) == 0;                           // rows of the intermediate stay 16-byte aligned, output rows 4-byte
  const int64_t nv = (int64_t)L.B * L.Hd * (vec ? (L.Wd * 3) / 4 : L.Wd * 3);
  const int gv = (int)std::min<int64_t>((nv + 255) / 256, (int64_t)ctx.sms * 32);
  if (vec) k_lanczos_v<4><<<gv, 256, 0, ctx.stream>>>(mid, out, L);
  else k_lanczos_v<1><<<gv, 256, 0, ctx.stream>>>(mid, out, L);
  count_launch();
  return cudaGetLastError();
}
#endif  // VRGDG_LANCZOS_IMPL

}  // namespace vrgdg
