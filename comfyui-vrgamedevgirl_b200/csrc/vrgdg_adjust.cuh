// vrgdg_adjust.cuh — the Builder UI's "adjust" pass (_apply_adjust_tensor, VRGDG_LUTVideoTools.py:307-391):
// temperature/tint offset, exposure, contrast, saturation, highlight/shadow/white/black masks   (stage A, per pixel)
// clarity  = x + (x - box_k(x, reflect pad)) * clarity * 1.55 * (0.35 + midtone * 0.65), k = min(9, odd(H), odd(W))   (stage C)
// sharpen  = x + (x - box_3(x, replicate pad)) * sharpen * 5                                                          (stage S)
// fade, vignette, clamp                                                                                               (stage D)
//
// Every operation is evaluated in the reference's order with one fp32 rounding per op (scalars are Python doubles rounded to
// fp32 where torch rounds them), and both box sums run sequentially in avg_pool2d's row-major window order, so the result is
// bit-identical to the reference's CPU tensor path for fp32 (and, through the exact codecs, uint8) frames.  One exception: the
// vignette mask contains a torch.sqrt, which on CPU is MKL VML's "< 1 ulp" routine rather than IEEE sqrt (0.7 % of mask values
// differ by 1 ulp), so with vignette > 0 parity is 2e-7, not equality.
//
// Kernels: k_adjust_point (A [+D]) streaming; k_adjust_box<MODE> (C or S [+D]) shared-memory tiles of an fp32 scratch frame.
// Bound: k_adjust_point HBM; k_adjust_box FADD issue (k*k sequential adds per element: 81 for clarity).
#pragma once
#include "vrgdg_kernels.cuh"

namespace vrgdg {

struct AdjustParams {
  int B, H, W;
  // stage A (all already rounded to fp32 the way torch rounds Python scalars)
  float off[3];
  float exposure, contrast, saturation;
  float hl, sh, wh, bl;               // highlights/220, shadows/220, whites/240, blacks/240
  // stage C / S
  int clarity_on, sharpen_on, kbox;   // kbox = blur kernel size (odd, <= 9; < 3 means "blur = x")
  float clarity, sharpen;
  // stage D
  int fade_on, vignette_on;
  float fade_mul, fade_add;           // (1 - fade*0.35), fade*0.18
  float vignette;
  const float* xx;                    // torch.linspace(-1, 1, W) [W]
  const float* yy;                    // torch.linspace(-1, 1, H) [H]
};

__device__ __forceinline__ float adj_luma(float r, float g, float b) {
  return addx(addx(mulx(r, 0.2126f), mulx(g, 0.7152f)), mulx(b, 0.0722f));
}

// stage A on one RGB pixel (input already clamped by the caller)
__device__ __forceinline__ void adjust_stage_a(const AdjustParams& A, float& r, float& g, float& b) {
  float v[3] = {r, g, b};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float o = addx(v[c], A.off[c]);                              // out + tensor([...])
    o = mulx(o, A.exposure);                                     // out * exposure
    v[c] = addx(mulx(subx(o, 0.5f), A.contrast), 0.5f);          // (out - 0.5) * contrast + 0.5
  }
  const float l1 = adj_luma(v[0], v[1], v[2]);
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = addx(l1, mulx(subx(v[c], l1), A.saturation));   // gray + (out - gray) * saturation
  const float l2 = adj_luma(v[0], v[1], v[2]);
  const float hm = clamp01(divx(subx(l2, 0.55f), 0.45f));
  const float sm = clamp01(divx(subx(0.45f, l2), 0.45f));
  const float wm = clamp01(mulx(subx(l2, 0.75f), 4.0f));        // / 0.25: a power of two, the product is the exact quotient
  const float bm = clamp01(mulx(subx(0.25f, l2), 4.0f));
  const float t1 = mulx(hm, A.hl), t2 = mulx(sm, A.sh), t3 = mulx(wm, A.wh), t4 = mulx(bm, A.bl);
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = addx(addx(addx(addx(v[c], t1), t2), t3), t4);
  r = v[0]; g = v[1]; b = v[2];
}

// stage D on one element at pixel (x, y)
__device__ __forceinline__ float adjust_stage_d(const AdjustParams& A, float v, float vmask) {
  if (A.fade_on) v = addx(mulx(v, A.fade_mul), A.fade_add);      // out * (1 - fade*0.35) + fade*0.18
  if (A.vignette_on) v = mulx(v, vmask);
  return clamp01(v);
}
__device__ __forceinline__ float vignette_mask(const AdjustParams& A, int x, int y) {
  if (!A.vignette_on) return 1.0f;
  const float xv = __ldg(A.xx + x), yv = __ldg(A.yy + y);
  const float d = sqrtx(addx(mulx(xv, xv), mulx(yv, yv)));        // sqrt(xx*xx + yy*yy)
  const float c = clamp01(divx(subx(d, 0.35f), 1.05f));
  return subx(1.0f, mulx(mulx(c, A.vignette), 0.75f));           // 1 - clamp(...) * vignette * 0.75
}

// ---- stage A (+ D when nothing follows): one pixel per thread ------------------------------------------------------------
template <typename T, bool TO_SCRATCH>
__global__ void __launch_bounds__(256)
k_adjust_point(const T* __restrict__ in, T* __restrict__ out, float* __restrict__ scratch, AdjustParams A, int enabled) {
  constexpr bool BGR = Io<T>::BGR;
  const int64_t hw = (int64_t)A.H * A.W, total = (int64_t)A.B * hw;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < total; p += (int64_t)gridDim.x * 256) {
    const T* s = in + p * 3;
    float r = clamp01(Elem<T>::ld(s[BGR ? 2 : 0])), g = clamp01(Elem<T>::ld(s[1])), b = clamp01(Elem<T>::ld(s[BGR ? 0 : 2]));   // source.clamp(0, 1)
    if (enabled) {
      adjust_stage_a(A, r, g, b);
      if (!TO_SCRATCH) {
        const int64_t pif = p % hw;
        const int y = (int)(pif / A.W), x = (int)(pif - (int64_t)y * A.W);
        const float m = vignette_mask(A, x, y);
        r = adjust_stage_d(A, r, m); g = adjust_stage_d(A, g, m); b = adjust_stage_d(A, b, m);
      }
    }
    if (TO_SCRATCH) {
      float* d = scratch + p * 3;
      d[0] = r; d[1] = g; d[2] = b;                              // RGB order in the scratch frame
    } else {
      T* d = out + p * 3;
      d[BGR ? 2 : 0] = Elem<T>::st(r); d[1] = Elem<T>::st(g); d[BGR ? 0 : 2] = Elem<T>::st(b);
    }
  }
}

// ---- stage C (MODE 0, reflect pad, k x k) or S (MODE 1, replicate pad, 3 x 3) on an fp32 RGB scratch frame ----------------------
// tile = 16 rows x 64 pixels (+ a fixed 4-pixel apron left and right, R rows above and below) in shared memory.  Each of the 256
// threads owns 4 pixels = 12 consecutive elements of one row: per window row it pulls the 12 + 2*PAD elements it needs with
// 128-bit shared loads (conflict-free: a quarter warp's 8 x 16 B land in 32 distinct banks) and feeds 12 independent window sums,
// each strictly in avg_pool2d's row-major order (that order, not a separable sum, is what makes the result bit-identical).
// Per output element and window row: 1/12 .. 3/4 of a shared load instead of K, so the kernel is FADD-issue bound.
constexpr int ADJ_TY = 16, ADJ_TXP = 64, ADJ_TXE = ADJ_TXP * 3, ADJ_RMAX = 4, ADJ_APRON = 4;
constexpr int ADJ_SW = ADJ_TXE + 6 * ADJ_APRON;       // 216 floats per tile row
constexpr int ADJ_SH = ADJ_TY + 2 * ADJ_RMAX;         // 24 rows

template <typename T, int MODE, bool LAST, int K>
__global__ void __launch_bounds__(256)
k_adjust_box(const float* __restrict__ src, float* __restrict__ dst_scratch, T* __restrict__ out, AdjustParams A, int tiles_x, int tiles_y) {
  constexpr bool BGR = Io<T>::BGR;
  constexpr int R = K / 2;
  constexpr int PAD = ((3 * R + 3) / 4) * 4;           // elements left of the thread's first one, rounded up to a 16-byte boundary
  constexpr int NV = (12 + 2 * PAD) / 4;               // float4 loads per window row
  __shared__ __align__(16) float tile[ADJ_SH * ADJ_SW];
  const int RW = A.W * 3;
  const int tiles_per_frame = tiles_x * tiles_y;
  const int total = A.B * tiles_per_frame;
  const int ry = threadIdx.x >> 4, e0 = (threadIdx.x & 15) * 12;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int frame = t / tiles_per_frame, rem = t - frame * tiles_per_frame;
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    const int y0 = ty * ADJ_TY, x0 = tx * ADJ_TXP;
    const float* fbase = src + (int64_t)frame * A.H * RW;
    // cooperative load with padding by index mapping (reflect: -i -> i, n-1+i -> n-1-i ; replicate: clamp); apron columns
    // further out than R are never read by a window and only need an in-range address
    constexpr int rows = ADJ_TY + 2 * R, colsp = ADJ_TXP + 2 * ADJ_APRON;
    for (int i = threadIdx.x; i < rows * colsp; i += 256) {
      const int rr = i / colsp, cp = i - rr * colsp;
      int y = y0 - R + rr, x = x0 - ADJ_APRON + cp;
      if (MODE == 0) {
        y = y < 0 ? -y : (y >= A.H ? 2 * (A.H - 1) - y : y);
        x = x < 0 ? -x : (x >= A.W ? 2 * (A.W - 1) - x : x);
      }
      y = max(0, min(y, A.H - 1)); x = max(0, min(x, A.W - 1));
      const float* s = fbase + ((int64_t)y * A.W + x) * 3;
      float* d = tile + rr * ADJ_SW + cp * 3;
      d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
    }
    __syncthreads();
    const int y = y0 + ry;
    if (y < A.H) {
      const float* c0 = tile + (ry + R) * ADJ_SW + 3 * ADJ_APRON + e0;   // the thread's first element (16-byte aligned)
      float acc[12];
      float ctr[12];
#pragma unroll
      for (int dy = -R; dy <= R; ++dy) {
        float v[NV * 4];
        const float4* rowp = reinterpret_cast<const float4*>(c0 + dy * ADJ_SW - PAD);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const float4 q = rowp[k];
          v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
        if (dy == 0) {
#pragma unroll
          for (int j = 0; j < 12; ++j) ctr[j] = v[PAD + j];
        }
#pragma unroll
        for (int dx = -R; dx <= R; ++dx) {
#pragma unroll
          for (int j = 0; j < 12; ++j) {
            const float w = v[PAD + j + 3 * dx];
            acc[j] = (dy == -R && dx == -R) ? w : addx(acc[j], w);
          }
        }
      }
      float res[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const float xc = ctr[j];
        res[j] = xc;
        if constexpr (K >= 3) {
          const float blur = div_const<K * K>(acc[j]);                  // avg_pool2d: sum / (K*K), correctly rounded (div_const)
          const float detail = subx(xc, blur);
          if (MODE == 0) {
            const int p = (j / 3) * 3;                                   // this pixel's r, g, b
            const float ln = adj_luma(ctr[p], ctr[p + 1], ctr[p + 2]);
            const float mid = subx(1.0f, clamp01(mulx(fabsf(subx(ln, 0.5f)), 2.0f))   /* / 0.5 */);
            const float wgt = addx(0.35f, mulx(mid, 0.65f));
            res[j] = addx(xc, mulx(mulx(mulx(detail, A.clarity), 1.55f), wgt));   // nchw + detail * clarity * 1.55 * (0.35 + mid*0.65)
          } else {
            res[j] = addx(xc, mulx(mulx(detail, A.sharpen), 5.0f));               // nchw + fine_detail * sharpen * 5.0
          }
        }   // K < 3 (frames narrower than 3 pixels): the reference's blur returns x itself, detail == 0, result == x
      }
      const int xq = x0 + (threadIdx.x & 15) * 4;
      const int64_t o = (((int64_t)frame * A.H + y) * A.W + xq) * 3;
      // 4 whole pixels, and rows / base pointer aligned for the 4-element vector stores (16 B fp32, 8 B half, 4 B u8)
      const bool whole = xq + 3 < A.W && (A.W & 3) == 0 &&
                         (LAST ? (reinterpret_cast<uintptr_t>(out) & (4 * sizeof(T) - 1)) == 0 : (reinterpret_cast<uintptr_t>(dst_scratch) & 15) == 0);
      if (!LAST) {
        if (whole) {
          float4* d4 = reinterpret_cast<float4*>(dst_scratch + o);
          d4[0] = make_float4(res[0], res[1], res[2], res[3]);
          d4[1] = make_float4(res[4], res[5], res[6], res[7]);
          d4[2] = make_float4(res[8], res[9], res[10], res[11]);
        } else {
#pragma unroll
          for (int j = 0; j < 12; ++j)
            if (xq + j / 3 < A.W) dst_scratch[o + j] = res[j];
        }
      } else {
        union alignas(16) { T v[12]; uint4 q4[3]; uint2 q2[3]; uint32_t q1[3]; } pk;
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float m = vignette_mask(A, min(xq + px, A.W - 1), y);
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) pk.v[px * 3 + (BGR ? 2 - ch : ch)] = Elem<T>::st(adjust_stage_d(A, res[px * 3 + ch], m));
        }
        if (whole) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            if (sizeof(T) == 4) reinterpret_cast<uint4*>(out + o)[k] = pk.q4[k];
            else if (sizeof(T) == 2) reinterpret_cast<uint2*>(out + o)[k] = pk.q2[k];
            else reinterpret_cast<uint32_t*>(out + o)[k] = pk.q1[k];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 12; ++j)
            if (xq + j / 3 < A.W) out[o + j] = pk.v[j];
        }
      }
    }
    __syncthreads();
  }
}

template <typename T, int MODE, bool LAST>
cudaError_t launch_adjust_box(int K, int grid, const float* src, float* dst, T* out, const AdjustParams& A, int tiles_x, int tiles_y,
                              cudaStream_t stream) {
  switch (K) {
    case 9: k_adjust_box<T, MODE, LAST, 9><<<grid, 256, 0, stream>>>(src, dst, out, A, tiles_x, tiles_y); break;
    case 7: k_adjust_box<T, MODE, LAST, 7><<<grid, 256, 0, stream>>>(src, dst, out, A, tiles_x, tiles_y); break;
    case 5: k_adjust_box<T, MODE, LAST, 5><<<grid, 256, 0, stream>>>(src, dst, out, A, tiles_x, tiles_y); break;
    case 3: k_adjust_box<T, MODE, LAST, 3><<<grid, 256, 0, stream>>>(src, dst, out, A, tiles_x, tiles_y); break;
    default: k_adjust_box<T, MODE, LAST, 1><<<grid, 256, 0, stream>>>(src, dst, out, A, tiles_x, tiles_y); break;
  }
  count_launch();
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_adjust(const void* in, void* out, const AdjustParams& A, int enabled, float* s1, float* s2, const LaunchCtx& ctx) {
  const int64_t total = (int64_t)A.B * A.H * A.W;
  if (total == 0) return cudaSuccess;
  const int pgrid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx.sms * 16);
  const bool C = enabled && A.clarity_on, S = enabled && A.sharpen_on;
  const T* tin = reinterpret_cast<const T*>(in);
  T* tout = reinterpret_cast<T*>(out);
  if (!C && !S) {
    k_adjust_point<T, false><<<pgrid, 256, 0, ctx.stream>>>(tin, tout, nullptr, A, enabled);
    count_launch();
    return cudaGetLastError();
  }
  k_adjust_point<T, true><<<pgrid, 256, 0, ctx.stream>>>(tin, tout, s1, A, enabled);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int tiles_x = (A.W + ADJ_TXP - 1) / ADJ_TXP, tiles_y = (A.H + ADJ_TY - 1) / ADJ_TY;
  const int64_t tiles = (int64_t)A.B * tiles_x * tiles_y;
  const int tgrid = (int)std::min<int64_t>(tiles, (int64_t)ctx.sms * 8);
  const float* cur = s1;
  if (C) {
    e = S ? launch_adjust_box<T, 0, false>(A.kbox, tgrid, cur, s2, tout, A, tiles_x, tiles_y, ctx.stream)
          : launch_adjust_box<T, 0, true>(A.kbox, tgrid, cur, nullptr, tout, A, tiles_x, tiles_y, ctx.stream);
    if (e != cudaSuccess) return e;
    cur = s2;
  }
  if (S) {
    if ((e = launch_adjust_box<T, 1, true>(3, tgrid, cur, nullptr, tout, A, tiles_x, tiles_y, ctx.stream)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace vrgdg
