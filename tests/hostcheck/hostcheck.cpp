// hostcheck.cpp — TEST INFRASTRUCTURE.  Compiles the kernels' per-pixel arithmetic header
// (comfyui-vrgamedevgirl_b200/csrc/vrgdg_math.cuh) for the host with g++ -ffp-contract=off so the CPU test
// suite can compare it with the oracle without a GPU.  Never loaded by the product.
#include "../../comfyui-vrgamedevgirl_b200/csrc/vrgdg_math.cuh"
#include <stdint.h>

using namespace vrgdg;

extern "C" {

void hc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  U4 c{ctr[0], ctr[1], ctr[2], ctr[3]};
  GrainKey K;
  for (int r = 0; r < PHILOX_ROUNDS; ++r) { K.rk[r][0] = key[0] + (uint32_t)r * PHILOX_W0; K.rk[r][1] = key[1] + (uint32_t)r * PHILOX_W1; }
  U4 r = philox4x32_rk(c, K);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// normals of a W-wide frame region: rows [y0, y0+rows), all x; out [rows][W][3]
void hc_normals(uint64_t seed, int64_t frame0, int64_t frame, int mode, int W, int y0, int rows, float* out) {
  GrainKey K;
  grain_make_key(seed, mode, K);
  GrainFrame gf = grain_frame(seed, frame0, frame, mode);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < W; ++x) {
      float* o = out + ((int64_t)y * W + x) * 3;
      grain_pixel_normals(K, gf, (uint32_t)x, (uint32_t)(y0 + y), o[0], o[1], o[2]);
    }
}

// same region through the pair interface (what the vector kernels use); W must be even
void hc_normals_pairs(uint64_t seed, int64_t frame0, int64_t frame, int mode, int W, int y0, int rows, float* out) {
  GrainKey K;
  grain_make_key(seed, mode, K);
  GrainFrame gf = grain_frame(seed, frame0, frame, mode);
  for (int y = 0; y < rows; ++y)
    for (int xp = 0; xp < W / 2; ++xp)
      grain_pair_normals(grain_pair_bits(K, gf, (uint32_t)xp, (uint32_t)(y0 + y)), out + ((int64_t)y * W + 2 * xp) * 3);
}

void hc_grain(const float* in, const float* noise, float* out, int64_t n, float I, float s, float oms, int exact) {
  for (int64_t i = 0; i < n; ++i) {
    float r = in[3 * i], g = in[3 * i + 1], b = in[3 * i + 2];
    if (exact) grain_blend_exact(r, g, b, noise[3 * i], noise[3 * i + 1], noise[3 * i + 2], I, s, oms);
    else grain_blend_fast(r, g, b, noise[3 * i], noise[3 * i + 1], noise[3 * i + 2], I, s, oms);
    out[3 * i] = r; out[3 * i + 1] = g; out[3 * i + 2] = b;
  }
}

void hc_lut3d(const float* in, float* out, int64_t n, const float* lut, int S, const float* dmin, const float* dspan,
              float blend, float omb, int exact) {
  // same packing as vrgdg_lut3d_pack
  float* packed = new float[(size_t)S * S * S * LUT_CELL_FLOATS];
  for (int bb = 0; bb < S; ++bb) for (int gg = 0; gg < S; ++gg) for (int rr = 0; rr < S; ++rr)
    lut_pack_entry(lut, S, bb, gg, rr, packed + ((size_t)(bb * S + gg) * S + rr) * LUT_CELL_FLOATS);
  LutParams P;
  P.lut = packed; P.S = S; P.smax = (float)(S - 1);
  for (int i = 0; i < 3; ++i) { P.dmin[i] = dmin[i]; P.dspan[i] = dspan[i]; }
  P.blend = blend; P.one_minus_blend = omb;
  P.unit_domain = (dmin[0] == 0.f && dmin[1] == 0.f && dmin[2] == 0.f && dspan[0] == 1.f && dspan[1] == 1.f && dspan[2] == 1.f);
  // polynomial cells (fast arithmetic of the chains): the same packing as k_lutp_pack
  float* poly = nullptr;
  if (exact == 4) {
    poly = new float[(size_t)S * S * S * LUT_CELL_FLOATS];
    for (int bb = 0; bb < S; ++bb) for (int gg = 0; gg < S; ++gg) for (int rr = 0; rr < S; ++rr)
      lutp_pack_entry(lut, S, bb, gg, rr, poly + ((size_t)(bb * S + gg) * S + rr) * LUT_CELL_FLOATS);
    P.lutp = poly;
  }
  for (int64_t i = 0; i < n; ++i) {
    float r = in[3 * i], g = in[3 * i + 1], b = in[3 * i + 2];
    float x0 = r, x1 = g, x2 = b;
    if (exact == 4) lutp_eval(P, r, g, b);
    else if (exact == 1) lut3d_eval<true>(P, r, g, b);
    else if (exact == 2) {                                           // the two-pixel form the kernels use (pixel paired with itself)
      float a[3] = {r, g, b}, c[3] = {r, g, b};
      lut3d_eval2<true>(P, a, c);
      r = c[0]; g = a[1]; b = c[2];
    } else if (exact == 3) {                                         // the element-mapped form of the tile kernels (one channel per call)
      const float o0 = lut3d_eval_channel<true>(P, r, g, b, 0), o1 = lut3d_eval_channel<true>(P, r, g, b, 1), o2 = lut3d_eval_channel<true>(P, r, g, b, 2);
      r = o0; g = o1; b = o2;
    } else lut3d_eval<false>(P, r, g, b);
    if (blend < 1.0f) {
      if (exact && exact != 4) { r = lut_blend<true>(x0, r, blend, omb); g = lut_blend<true>(x1, g, blend, omb); b = lut_blend<true>(x2, b, blend, omb); }
      else { r = lut_blend<false>(x0, r, blend, omb); g = lut_blend<false>(x1, g, blend, omb); b = lut_blend<false>(x2, b, blend, omb); }
    }
    out[3 * i] = r; out[3 * i + 1] = g; out[3 * i + 2] = b;
  }
  delete[] packed;
  delete[] poly;
}

void hc_div_const(const float* in, float* out, int64_t n, int d) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = d == 9 ? div_const<9>(in[i]) : (d == 25 ? div_const<25>(in[i]) : (d == 49 ? div_const<49>(in[i]) : (d == 81 ? div_const<81>(in[i]) : div_const<255>(in[i]))));
}

void hc_rgb_to_lab(const float* in, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) rgb_to_lab(in[3 * i], in[3 * i + 1], in[3 * i + 2], out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}

void hc_lab_to_rgb(const float* in, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) lab_to_rgb(in[3 * i], in[3 * i + 1], in[3 * i + 2], out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}

void hc_colormatch(const float* in, float* out, int64_t n, const float* params, float t, float omt) {
  for (int64_t i = 0; i < n; ++i) {
    float r = in[3 * i], g = in[3 * i + 1], b = in[3 * i + 2];
    colormatch_pixel(r, g, b, params, t, omt);
    out[3 * i] = r; out[3 * i + 1] = g; out[3 * i + 2] = b;
  }
}

// LAB raw sums {n, S_L, S_a, S_b, S_LL, S_aa, S_bb} the way the moments kernel forms them: sums over u = (fy, fx-fy, fy-fz),
// then the affine change of variables in double (cm_sums_to_lab_host)
void hc_lab_sums(const float* in, int64_t n, double* sums7) {
  double u[7] = {(double)n, 0, 0, 0, 0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    float fx, fy, fz;
    rgb_to_fxyz(in[3 * i], in[3 * i + 1], in[3 * i + 2], fx, fy, fz);
    const float u1 = fx - fy, u2 = fy - fz;
    u[1] += fy; u[2] += u1; u[3] += u2;
    u[4] += (double)fy * fy; u[5] += (double)u1 * u1; u[6] += (double)u2 * u2;
  }
  cm_sums_to_lab_host(u, sums7);
}

// frames [H][W][3], border 0 replicate / 1 zero
void hc_stencil(const float* in, float* out, int H, int W, int op, float s, int border, int exact) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < 3; ++c) {
        float n9[9];
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            int yy = y + dy, xx = x + dx;
            float v;
            if (border == 0) {
              yy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy);
              xx = xx < 0 ? 0 : (xx >= W ? W - 1 : xx);
              v = in[(yy * W + xx) * 3 + c];
            } else {
              v = (yy < 0 || yy >= H || xx < 0 || xx >= W) ? 0.0f : in[(yy * W + xx) * 3 + c];
            }
            n9[(dy + 1) * 3 + dx + 1] = v;
          }
        out[(y * W + x) * 3 + c] = exact ? stencil_epilogue_exact(op, n9, s) : stencil_epilogue(op, n9, s);
      }
}

}  // extern "C"

extern "C" void hc_pows(const float* x, float* p24, float* pinv, float* cb, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { p24[i] = pow_2p4(x[i]); pinv[i] = pow_inv2p4(x[i]); cb[i] = cbrt_pos(x[i]); }
}
