// vrgdg_math.cuh — per-pixel arithmetic of the post-processing hot path.
//
// Everything here is __host__ __device__ so that tests/hostcheck can compile the very same
// arithmetic with g++ and compare it with the oracle on a machine without a GPU.  The product
// only ever runs the __device__ instantiation (the library has no CPU execution path).
//
// "x" suffix = exact: one IEEE fp32 rounding per operation, no FMA contraction, so results are
// bit-identical to the reference's CPU tensor ops (each of which rounds once).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define VRGDG_HD __host__ __device__ __forceinline__
#else
#define VRGDG_HD inline
#endif

namespace vrgdg {

// ---- exact fp32 primitives ---------------------------------------------------------------
#if defined(__CUDA_ARCH__)
VRGDG_HD float addx(float a, float b) { return __fadd_rn(a, b); }
VRGDG_HD float subx(float a, float b) { return __fsub_rn(a, b); }
VRGDG_HD float mulx(float a, float b) { return __fmul_rn(a, b); }
VRGDG_HD float divx(float a, float b) { return __fdiv_rn(a, b); }
VRGDG_HD float sqrtx(float a) { return __fsqrt_rn(a); }
#else
// host build uses -ffp-contract=off; volatile keeps the optimiser from re-associating
VRGDG_HD float addx(float a, float b) { volatile float r = a + b; return r; }
VRGDG_HD float subx(float a, float b) { volatile float r = a - b; return r; }
VRGDG_HD float mulx(float a, float b) { volatile float r = a * b; return r; }
VRGDG_HD float divx(float a, float b) { volatile float r = a / b; return r; }
VRGDG_HD float sqrtx(float a) { return sqrtf(a); }
#endif

VRGDG_HD float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based ----------------------------------
struct U4 { uint32_t x, y, z, w; };

VRGDG_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

template <int ROUNDS>
VRGDG_HD U4 philox4x32(U4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint32_t h0, l0, h1, l1;
    mulhilo(M0, c.x, h0, l0);
    mulhilo(M1, c.z, h1, l1);
    U4 n;
    n.x = h1 ^ c.y ^ k0;
    n.y = l1;
    n.z = h0 ^ c.w ^ k1;
    n.w = l0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// Key/counter layout of the grain generator (documented in DESIGN.md):
//   key     = (seed_lo, seed_hi)
//   counter = (pixel index inside the frame, 0x5652 "VR", frame_lo, frame_hi)
struct GrainKey {
  uint32_t k0, k1;     // Philox key
  uint32_t f0, f1;     // frame part of the counter
};

VRGDG_HD GrainKey grain_key(uint64_t seed, int64_t frame0, int64_t frame_in_batch, int seed_mode) {
  GrainKey g;
  if (seed_mode == 1) {  // VRGDG_SEED_PER_FRAME: EnhancerNodes.py:267-268
    uint64_t s = (uint64_t)((int64_t)seed + frame0 + frame_in_batch) & 0x7FFFFFFFull;
    g.k0 = (uint32_t)s; g.k1 = 0u; g.f0 = 0u; g.f1 = 0u;
  } else {
    uint64_t f = (uint64_t)(frame0 + frame_in_batch);
    g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
    g.f0 = (uint32_t)f; g.f1 = (uint32_t)(f >> 32);
  }
  return g;
}

// Three N(0,1) per pixel from one Philox call: Box-Muller pair (x,y) -> z_r, z_g; (z,w) -> z_b.
VRGDG_HD void box_muller3(U4 r, float& zr, float& zg, float& zb) {
  const float TWO_NEG32 = 2.3283064365386963e-10f;       // 2^-32
  const float HALF_ULP = 1.1641532182693481e-10f;        // 2^-33 keeps u1 > 0
  const float TWO_PI_2NEG32 = 1.4629180792671596e-9f;    // 2*pi*2^-32
  float u1a = fmaf((float)r.x, TWO_NEG32, HALF_ULP);
  float u1b = fmaf((float)r.z, TWO_NEG32, HALF_ULP);
  float tha = (float)r.y * TWO_PI_2NEG32;
  float thb = (float)r.w * TWO_PI_2NEG32;
#if defined(__CUDA_ARCH__)
  // -2 ln u = -2 ln2 * log2 u ; MUFU.LG2, MUFU.SQRT, MUFU.SIN/COS
  float la = __log2f(u1a) * -1.3862943611198906f;
  float lb = __log2f(u1b) * -1.3862943611198906f;
  float ra, rb;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(la));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(lb));
  float sa, ca;
  __sincosf(tha, &sa, &ca);
  float cb = __cosf(thb);
#else
  float ra = sqrtf(-2.0f * logf(u1a));
  float rb = sqrtf(-2.0f * logf(u1b));
  float sa = sinf(tha), ca = cosf(tha);
  float cb = cosf(thb);
#endif
  zr = ra * ca;
  zg = ra * sa;
  zb = rb * cb;
}

template <int ROUNDS>
VRGDG_HD void grain_normals(const GrainKey& g, uint32_t pixel_in_frame, float& zr, float& zg, float& zb) {
  U4 c;
  c.x = pixel_in_frame; c.y = 0x5652u; c.z = g.f0; c.w = g.f1;
  U4 r = philox4x32<ROUNDS>(c, g.k0, g.k1);
  box_muller3(r, zr, zg, zb);
}

// ---- grain blend: nodes.py:53-60 ------------------------------------------------------------
// exact variant = the reference's op sequence, one rounding per op.
VRGDG_HD void grain_blend_exact(float& r, float& g, float& b, float zr, float zg, float zb,
                                float I, float s, float oms) {
  float zr2 = mulx(zr, 2.0f);                     // grain[...,0] *= 2.0
  float zb3 = mulx(zb, 3.0f);                     // grain[...,2] *= 3.0
  float gy = mulx(oms, zg);                       // (1.0 - s) * gray
  float gr = addx(mulx(s, zr2), gy);              // s*grain + (1-s)*gray
  float gg = addx(mulx(s, zg), gy);
  float gb = addx(mulx(s, zb3), gy);
  r = clamp01(addx(r, mulx(gr, I)));              // batch + grain*I ; clamp
  g = clamp01(addx(g, mulx(gg, I)));
  b = clamp01(addx(b, mulx(gb, I)));
}

// fused variant for in-kernel noise (the noise stream itself is ours, so contraction is free)
VRGDG_HD void grain_blend_fast(float& r, float& g, float& b, float zr, float zg, float zb,
                               float I, float s, float oms) {
  float gy = oms * zg;
  r = clamp01(fmaf(I, fmaf(2.0f * s, zr, gy), r));
  g = clamp01(fmaf(I, fmaf(s, zg, gy), g));
  b = clamp01(fmaf(I, fmaf(3.0f * s, zb, gy), b));
}

// ---- 3D LUT trilinear: VRGDG_IV_Adjustments.py:293-336 ------------------------------------------
struct LutParams {
  const float* lut;      // [S][S][S][3], [b][g][r][rgb]
  int S;
  float smax;            // float(S-1)
  float dmin[3], dspan[3];
  float blend, one_minus_blend;
};

#if defined(__CUDA_ARCH__)
#define VRGDG_LDG(p) __ldg(p)
#else
#define VRGDG_LDG(p) (*(p))
#endif

// coordinate -> (cell index, fraction); bit-exact with :296-316
VRGDG_HD void lut_coord(float v, float dmin, float dspan, float smax, int S, int& i0, int& i1, float& f) {
  float n = divx(subx(v, dmin), dspan);          // (source - domain_min) / domain_span
  n = clamp01(n);                                 // torch.clamp(normalized, 0, 1)
  float c = mulx(n, smax);                        // normalized * max_index
  float fl = floorf(c);
  i0 = (int)fl;                                   // torch.floor(r).long()
  i1 = (i0 + 1 < S - 1) ? i0 + 1 : S - 1;         // clamp(r0+1, max=max_index)
  f = subx(c, fl);                                // r - r0.float()
}

template <bool EXACT>
VRGDG_HD float lerp_ref(float a, float b, float f, float omf) {
  if (EXACT) return addx(mulx(a, omf), mulx(b, f));   // a*(1-f) + b*f, three roundings (:327-335)
  return fmaf(b, f, a * omf);
}

template <bool EXACT>
VRGDG_HD void lut3d_eval(const LutParams& P, float& r, float& g, float& b) {
  int r0, r1, g0, g1, b0, b1;
  float fr, fg, fb;
  lut_coord(r, P.dmin[0], P.dspan[0], P.smax, P.S, r0, r1, fr);
  lut_coord(g, P.dmin[1], P.dspan[1], P.smax, P.S, g0, g1, fg);
  lut_coord(b, P.dmin[2], P.dspan[2], P.smax, P.S, b0, b1, fb);
  const int S = P.S;
  const float* L = P.lut;
  // element offsets of the 8 corners: ((b*S+g)*S+r)*3
  int ob0g0 = (b0 * S + g0) * S, ob1g0 = (b1 * S + g0) * S;
  int ob0g1 = (b0 * S + g1) * S, ob1g1 = (b1 * S + g1) * S;
  const float* p000 = L + (ob0g0 + r0) * 3;   // c000 = lut[b0,g0,r0]
  const float* p001 = L + (ob1g0 + r0) * 3;   // c001 = lut[b1,g0,r0]
  const float* p010 = L + (ob0g1 + r0) * 3;   // c010 = lut[b0,g1,r0]
  const float* p011 = L + (ob1g1 + r0) * 3;   // c011 = lut[b1,g1,r0]
  const float* p100 = L + (ob0g0 + r1) * 3;   // c100 = lut[b0,g0,r1]
  const float* p101 = L + (ob1g0 + r1) * 3;
  const float* p110 = L + (ob0g1 + r1) * 3;
  const float* p111 = L + (ob1g1 + r1) * 3;
  float omb = subx(1.0f, fb), omg = subx(1.0f, fg), omr = subx(1.0f, fr);
  float o[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float c00 = lerp_ref<EXACT>(VRGDG_LDG(p000 + ch), VRGDG_LDG(p001 + ch), fb, omb);
    float c01 = lerp_ref<EXACT>(VRGDG_LDG(p010 + ch), VRGDG_LDG(p011 + ch), fb, omb);
    float c10 = lerp_ref<EXACT>(VRGDG_LDG(p100 + ch), VRGDG_LDG(p101 + ch), fb, omb);
    float c11 = lerp_ref<EXACT>(VRGDG_LDG(p110 + ch), VRGDG_LDG(p111 + ch), fb, omb);
    float c0 = lerp_ref<EXACT>(c00, c01, fg, omg);
    float c1 = lerp_ref<EXACT>(c10, c11, fg, omg);
    o[ch] = clamp01(lerp_ref<EXACT>(c0, c1, fr, omr));
  }
  r = o[0]; g = o[1]; b = o[2];
}

// strength blend of apply_lut (:355-359) on already-rounded LUT output `y` and input `x`
template <bool EXACT>
VRGDG_HD float lut_blend(float x, float y, float blend, float omb) {
  if (EXACT) return addx(mulx(x, omb), mulx(y, blend));
  return fmaf(y, blend, x * omb);
}

// ---- sRGB <-> CIE Lab (kornia.color restatement; formulas in SURVEY.md §8c) ------------------------
VRGDG_HD float srgb_to_linear(float c) {
  // where(c > 0.04045, ((c + 0.055) / 1.055) ** 2.4, c / 12.92)
  return (c > 0.04045f) ? powf(divx(addx(c, 0.055f), 1.055f), 2.4f) : divx(c, 12.92f);
}
VRGDG_HD float linear_to_srgb(float l) {
  // where(l > 0.0031308, 1.055 * clamp(l, min=thr) ** (1/2.4) - 0.055, 12.92 * l)
  return (l > 0.0031308f) ? subx(mulx(1.055f, powf(fmaxf(l, 0.0031308f), (float)(1.0 / 2.4))), 0.055f)
                          : mulx(12.92f, l);
}
VRGDG_HD float lab_f(float t) {
  // where(t > 0.008856, clamp(t, min=0.008856) ** (1/3), 7.787 t + 4/29)
  return (t > 0.008856f) ? powf(fmaxf(t, 0.008856f), (float)(1.0 / 3.0))
                         : addx(mulx(7.787f, t), (float)(4.0 / 29.0));
}
VRGDG_HD void rgb_to_lab(float r, float g, float b, float& L, float& A, float& Bv) {
  float lr = srgb_to_linear(r), lg = srgb_to_linear(g), lb = srgb_to_linear(b);
  float x = addx(addx(mulx(0.412453f, lr), mulx(0.357580f, lg)), mulx(0.180423f, lb));
  float y = addx(addx(mulx(0.212671f, lr), mulx(0.715160f, lg)), mulx(0.072169f, lb));
  float z = addx(addx(mulx(0.019334f, lr), mulx(0.119193f, lg)), mulx(0.950227f, lb));
  float fx = lab_f(divx(x, 0.95047f));
  float fy = lab_f(y);                                 // y / 1.0
  float fz = lab_f(divx(z, 1.08883f));
  L = subx(mulx(116.0f, fy), 16.0f);
  A = mulx(500.0f, subx(fx, fy));
  Bv = mulx(200.0f, subx(fy, fz));
}
VRGDG_HD float lab_finv(float f) {
  // where(f > 0.2068966, f ** 3, (f - 4/29) / 7.787)
  // torch.pow(x, 3.0) evaluates x*x*x
  return (f > 0.2068966f) ? mulx(mulx(f, f), f) : divx(subx(f, (float)(4.0 / 29.0)), 7.787f);
}
VRGDG_HD void lab_to_rgb(float L, float A, float Bv, float& r, float& g, float& b) {
  float fy = divx(addx(L, 16.0f), 116.0f);
  float fx = addx(divx(A, 500.0f), fy);
  float fz = fmaxf(subx(fy, divx(Bv, 200.0f)), 0.0f);
  float x = mulx(lab_finv(fx), 0.95047f);
  float y = lab_finv(fy);                              // * 1.0
  float z = mulx(lab_finv(fz), 1.08883f);
  float lr = addx(addx(mulx(3.2404813432005266f, x), mulx(-1.5371515162713185f, y)), mulx(-0.4985363261688878f, z));
  float lg = addx(addx(mulx(-0.9692549499965682f, x), mulx(1.8759900014898907f, y)), mulx(0.0415559265582928f, z));
  float lb = addx(addx(mulx(0.0556466391351772f, x), mulx(-0.2040413383665112f, y)), mulx(1.0573110696453443f, z));
  r = clamp01(linear_to_srgb(lr));
  g = clamp01(linear_to_srgb(lg));
  b = clamp01(linear_to_srgb(lb));
}

// nodes.py:112-115: matched = (lab - mu)/sd * sd_ref + mu_ref ; blended = t*matched + (1-t)*lab
// p = {mu_img[3], sd_img[3], mu_ref[3], sd_ref[3]}
VRGDG_HD void colormatch_pixel(float& r, float& g, float& b, const float* p, float t, float omt) {
  float lab[3];
  rgb_to_lab(r, g, b, lab[0], lab[1], lab[2]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float m = addx(mulx(divx(subx(lab[c], p[c]), p[3 + c]), p[9 + c]), p[6 + c]);
    lab[c] = addx(mulx(t, m), mulx(omt, lab[c]));
  }
  lab_to_rgb(lab[0], lab[1], lab[2], r, g, b);
}

// ---- 3x3 stencil epilogues: nodes.py:194-207, :278-287, :369-382 (numpy) and :171-174,:249-258,:345-349 (torch)
// n[0..8] = row-major 3x3 neighbourhood, n[4] = centre.
VRGDG_HD float stencil_epilogue(int op, const float* n, float s) {
  float c = n[4], v;
  switch (op) {
    case 1: {  // box unsharp: blur = sum9 / 9 ; out = c + s*(c - blur)
      float sum = ((n[0] + n[1]) + n[2]) + ((n[3] + n[4]) + n[5]) + ((n[6] + n[7]) + n[8]);
      float blur = sum / 9.0f;
      v = c + s * (c - blur);
    } break;
    case 2: {  // numpy laplacian: lap = W + N + S + E - 4c ; out = c + s*lap   (blurs; reference quirk D5)
      float lap = (((n[3] + n[1]) + n[7]) + n[5]) - 4.0f * c;
      v = c + s * lap;
    } break;
    case 3: {  // torch laplacian: edges = 4c - N - S - E - W
      float e = 4.0f * c - n[1] - n[3] - n[5] - n[7];
      v = c + s * e;
    } break;
    case 4:
    case 5: {  // sobel
      float gx = (-n[0] - 2.0f * n[3] - n[6]) + (n[2] + 2.0f * n[5] + n[8]);
      float gy = (-n[0] - 2.0f * n[1] - n[2]) + (n[6] + 2.0f * n[7] + n[8]);
      float m = gx * gx + gy * gy + ((op == 5) ? 1e-6f : 0.0f);
      v = c + s * sqrtf(m);
    } break;
    default: v = c;
  }
  return clamp01(v);
}

}  // namespace vrgdg
