// sharp_ops.cuh -- per-column integer arithmetic of the two whole-picture passes that precede the
// block pipeline: the "sharp" RGB -> YUV 4:2:0 conversion (SJPEG_YUV_SHARP) and the riskiness
// analyser behind SJPEG_YUV_AUTO.  Written once as __host__ __device__ functions so that the same
// code runs in the sm_100a kernels (sharp.cu) and, compiled by g++, in the CPU emulation of the
// no-GPU tests (tests/emul/emul_main.cc).  Reference behaviour reproduced bit for bit:
//   fixed-point conventions, gamma tables      /root/reference/src/yuv_convert.cc:33-170
//   W/RGB import, targets, 9-3-3-1 filter      /root/reference/src/yuv_convert.cc:174-204,434-545
//   iteration and exit rule                    /root/reference/src/yuv_convert.cc:582-665
//   final Y/U/V                                /root/reference/src/yuv_convert.cc:58-71,547-577
//   7-level YUV index, riskiness score         /root/reference/src/colors_rgb.cc:1080-1122,
//                                              /root/reference/src/jpeg_tools.cc:172-236
//
// Working representation ("W/RGB"): samples carry 2 extra fractional bits (0..1023).  Per picture,
// with w, h = width, height rounded up to even and uv_w = w/2, uv_h = h/2:
//   luma state   u16 [h][w]              best estimate of W (gamma-space gray)
//   chroma state i16 [uv_h][3][uv_w]     R-W, G-W, B-W at half resolution, one row = three runs of uv_w
//   targets      same shapes              what the reconstruction must reproduce
// One iteration visits the chroma rows top to bottom; row r reads row r-1 of ITS OWN output (the
// reference updates the chroma state in place, so the row above is already refined: Gauss-Seidel
// downwards) and rows r, r+1 of the previous state.
#pragma once
#include <stdint.h>

#include "block_ops.cuh"

namespace sjb {

enum { kSharpFix = 2, kSharpMaxY = (256 << kSharpFix) - 1, kSharpGammaTab = 32, kSharpLinBits = 14,
       kSharpIterations = 4 };

struct SharpTabs {
  const uint32_t* g2l;   // [1024] gamma -> linear, 14 fractional bits
  const uint32_t* l2g;   // [34]   linear -> gamma interpolation nodes (rounding included)
};

SJB_HD int sharp_clip_y(int y) { return y < 0 ? 0 : (y > kSharpMaxY ? kSharpMaxY : y); }
SJB_HD int sharp_clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
SJB_HD uint32_t sharp_gray(uint32_t r, uint32_t g, uint32_t b) {           // yuv_convert.cc:434-437
  return (13933u * r + 46871u * g + 4732u * b + 32768u) >> 16;
}
SJB_HD uint32_t sharp_l2g(const SharpTabs& t, uint32_t value) {            // yuv_convert.cc:156-170
  const uint32_t v = value * kSharpGammaTab;
  const uint32_t pos = v >> kSharpLinBits;
  const uint32_t frac = v - (pos << kSharpLinBits);
  const uint32_t v0 = t.l2g[pos], v1 = t.l2g[pos + 1];
  return v0 + (((v1 - v0) * frac) >> kSharpLinBits);
}
SJB_HD int sharp_uplift(int a) { return (a << kSharpFix) | (1 << kSharpFix >> 1); }   // :492-494

// One 2x2 pixel cell in W/RGB precision: px[dy][dx][c].
struct SharpCell {
  int px[2][2][3];
};

// Luma candidates of the four pixels and the chroma candidate of the cell
// (UpdateW :466-474 and UpdateChroma :447-464 share the twelve gamma->linear lookups).
SJB_HD void sharp_cell_targets(const SharpTabs& t, const SharpCell& c, int (&wy)[2][2], int (&uv)[3]) {
  uint32_t lin_sum[3] = {2u, 2u, 2u};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int dy = 0; dy < 2; ++dy) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int dx = 0; dx < 2; ++dx) {
      const uint32_t r = t.g2l[c.px[dy][dx][0]], g = t.g2l[c.px[dy][dx][1]], b = t.g2l[c.px[dy][dx][2]];
      lin_sum[0] += r; lin_sum[1] += g; lin_sum[2] += b;
      wy[dy][dx] = static_cast<int>(sharp_l2g(t, sharp_gray(r, g, b)));
    }
  }
  const uint32_t r = sharp_l2g(t, lin_sum[0] >> 2), g = sharp_l2g(t, lin_sum[1] >> 2), b = sharp_l2g(t, lin_sum[2] >> 2);
  const int W = static_cast<int>(sharp_gray(r, g, b));
  uv[0] = static_cast<int16_t>(static_cast<int>(r) - W);
  uv[1] = static_cast<int16_t>(static_cast<int>(g) - W);
  uv[2] = static_cast<int16_t>(static_cast<int>(b) - W);
}

// Upsampled chroma at the two pixel columns of chroma column i, for one picture row whose
// vertical neighbour chroma row is B (the row above for the top pixel row, below for the bottom):
// A = this chroma row.  Boundary columns use the 3-1 filter (Filter2 :484-487), the interior the
// 9-3-3-1 filter (:195-204).  a_m/a_0/a_p = A[i-1], A[i], A[i+1]; same for B.
SJB_HD void sharp_upsample_pair(int i, int uv_w, int a_m, int a_0, int a_p, int b_m, int b_0, int b_p, int* even,
                                int* odd) {
  *even = (i == 0) ? ((a_0 * 3 + b_0 + 2) >> 2) : ((a_0 * 9 + a_m * 3 + b_0 * 3 + b_m + 8) >> 4);
  *odd = (i == uv_w - 1) ? ((a_0 * 3 + b_0 + 2) >> 2) : ((a_0 * 9 + a_p * 3 + b_0 * 3 + b_p + 8) >> 4);
}

// ---------------------------------------------------------------------------------------------
// Import (PreprocessARGB's first loop, :604-625): chroma column i of chroma row r of a
// width x height RGB picture -> initial luma state, luma target, chroma target (= initial state).
// Odd width / height replicate the last column / row (:506-512, :617-619).
// ---------------------------------------------------------------------------------------------
SJB_HD void sharp_import_cell(const SharpTabs& t, const uint8_t* rgb, long long stride, int width, int height, int w,
                              int uv_w, int r, int i, uint16_t* best_y, uint16_t* target_y, int16_t* best_uv,
                              int16_t* target_uv) {
  SharpCell c;
  for (int dy = 0; dy < 2; ++dy) {
    const int y = (2 * r + dy < height) ? 2 * r + dy : height - 1;
    const uint8_t* row = rgb + y * stride;
    for (int dx = 0; dx < 2; ++dx) {
      const int x = (2 * i + dx < width) ? 2 * i + dx : width - 1;
      for (int k = 0; k < 3; ++k) c.px[dy][dx][k] = sharp_uplift(row[3 * x + k]);
    }
  }
  int wy[2][2], uv[3];
  sharp_cell_targets(t, c, wy, uv);
  for (int dy = 0; dy < 2; ++dy) {
    for (int dx = 0; dx < 2; ++dx) {
      const size_t at = static_cast<size_t>(2 * r + dy) * w + 2 * i + dx;
      best_y[at] = static_cast<uint16_t>(sharp_gray(c.px[dy][dx][0], c.px[dy][dx][1], c.px[dy][dx][2]));   // StoreGray
      target_y[at] = static_cast<uint16_t>(wy[dy][dx]);
    }
  }
  for (int k = 0; k < 3; ++k) {
    const size_t at = (static_cast<size_t>(r) * 3 + k) * uv_w + i;
    target_uv[at] = static_cast<int16_t>(uv[k]);
    best_uv[at] = static_cast<int16_t>(uv[k]);
  }
}

// Inputs of one cell, split by when they become available: everything that comes from the
// previous iteration's state or from the targets can be fetched one row ahead; only `above`
// (row r-1 of the iteration in progress) has to wait for the row barrier.
struct SharpCellIn {
  int y_old[2][2];
  int cur[3][3], below[3][3];     // [channel][i-1, i, i+1] (edge columns clamped)
  int target_y[2][2], target_uv[3];
};
struct SharpCellAbove {
  int v[3][3];
};

template <class Ld>
SJB_HD void sharp_load_cell(const Ld& ld, int w, int uv_w, int i, const uint16_t* y_in, const int16_t* uv_cur,
                            const int16_t* uv_below, const uint16_t* target_y, const int16_t* target_uv,
                            SharpCellIn* in) {
  const int im = (i > 0) ? i - 1 : 0, ip = (i < uv_w - 1) ? i + 1 : i;
  for (int dy = 0; dy < 2; ++dy) {
    for (int dx = 0; dx < 2; ++dx) {
      in->y_old[dy][dx] = ld.y(y_in + dy * w + 2 * i + dx);
      in->target_y[dy][dx] = ld.y(target_y + dy * w + 2 * i + dx);
    }
  }
  for (int k = 0; k < 3; ++k) {
    const int16_t* A = uv_cur + k * uv_w;
    const int16_t* N = uv_below + k * uv_w;
    in->cur[k][0] = ld.uv(A + im); in->cur[k][1] = ld.uv(A + i); in->cur[k][2] = ld.uv(A + ip);
    in->below[k][0] = ld.uv(N + im); in->below[k][1] = ld.uv(N + i); in->below[k][2] = ld.uv(N + ip);
    in->target_uv[k] = ld.uv(target_uv + k * uv_w + i);
  }
}
template <class Ld>
SJB_HD void sharp_load_above(const Ld& ld, int uv_w, int i, const int16_t* uv_above, SharpCellAbove* a) {
  const int im = (i > 0) ? i - 1 : 0, ip = (i < uv_w - 1) ? i + 1 : i;
  for (int k = 0; k < 3; ++k) {
    const int16_t* P = uv_above + k * uv_w;
    a->v[k][0] = ld.uv(P + im); a->v[k][1] = ld.uv(P + i); a->v[k][2] = ld.uv(P + ip);
  }
}

// ---------------------------------------------------------------------------------------------
// One refinement step of chroma column i in chroma row r (the body of the iteration loop,
// :636-653): interpolate the 2x2 cell from the chroma rows above / here / below, convert it back
// to W and chroma, and move the state by the distance to the targets.  y_out / uv_out = rows 2r
// (and 2r+1 at +w) and r of THIS iteration's state.  Returns the sum of |luma correction| of the
// four pixels (SharpUpdateY's return, :174-184).
// ---------------------------------------------------------------------------------------------
struct SharpCellOut {
  int y_new[2][2];   // refined luma of the four pixels
  int uv_new[3];     // refined chroma of the cell
  uint32_t diff;     // sum of |luma correction|
};
SJB_HD void sharp_refine_compute(const SharpTabs& t, int uv_w, int i, const SharpCellIn& in,
                                 const SharpCellAbove& above, SharpCellOut* out) {
  SharpCell c;
  for (int k = 0; k < 3; ++k) {
    int e, o;
    sharp_upsample_pair(i, uv_w, in.cur[k][0], in.cur[k][1], in.cur[k][2], above.v[k][0], above.v[k][1], above.v[k][2],
                        &e, &o);
    c.px[0][0][k] = sharp_clip_y(in.y_old[0][0] + e);
    c.px[0][1][k] = sharp_clip_y(in.y_old[0][1] + o);
    sharp_upsample_pair(i, uv_w, in.cur[k][0], in.cur[k][1], in.cur[k][2], in.below[k][0], in.below[k][1],
                        in.below[k][2], &e, &o);
    c.px[1][0][k] = sharp_clip_y(in.y_old[1][0] + e);
    c.px[1][1][k] = sharp_clip_y(in.y_old[1][1] + o);
  }
  int wy[2][2], uv[3];
  sharp_cell_targets(t, c, wy, uv);
  out->diff = 0;
  for (int dy = 0; dy < 2; ++dy) {
    for (int dx = 0; dx < 2; ++dx) {
      const int d = in.target_y[dy][dx] - wy[dy][dx];
      out->y_new[dy][dx] = sharp_clip_y(in.y_old[dy][dx] + d);
      out->diff += static_cast<uint32_t>(d < 0 ? -d : d);
    }
  }
  for (int k = 0; k < 3; ++k) {   // SharpUpdateRGB :186-192 (int16 wrap-around kept)
    out->uv_new[k] = static_cast<int16_t>(in.cur[k][1] + (in.target_uv[k] - uv[k]));
  }
}
SJB_HD void sharp_store_cell(int w, int uv_w, int i, const SharpCellOut& out, uint16_t* y_out, int16_t* uv_out) {
  for (int dy = 0; dy < 2; ++dy) {
    for (int dx = 0; dx < 2; ++dx) y_out[dy * w + 2 * i + dx] = static_cast<uint16_t>(out.y_new[dy][dx]);
  }
  for (int k = 0; k < 3; ++k) uv_out[k * uv_w + i] = static_cast<int16_t>(out.uv_new[k]);
}
SJB_HD uint32_t sharp_refine_cell(const SharpTabs& t, int w, int uv_w, int i, const SharpCellIn& in,
                                  const SharpCellAbove& above, uint16_t* y_out, int16_t* uv_out) {
  SharpCellOut out;
  sharp_refine_compute(t, uv_w, i, in, above, &out);
  sharp_store_cell(w, uv_w, i, out, y_out, uv_out);
  return out.diff;
}

// Which iteration's state is the result: the loop of :628-660 always runs iterations 0 and 1 and
// stops after the first iteration >= 1 whose correction sum fell below 3*w*h or grew.
SJB_HD int sharp_final_iteration(const unsigned long long diff[kSharpIterations], int w, int h) {
  const unsigned long long threshold = static_cast<unsigned long long>(3.0 * w * h);
  for (int it = 1; it < kSharpIterations; ++it) {
    if (diff[it] < threshold || diff[it] > diff[it - 1]) return it;
  }
  return kSharpIterations - 1;
}

// Final conversion (:58-71): r, g, b in W/RGB precision, 18-bit fixed point.  Note the U and V
// rows differ from the block path's constants in the last digit (-11058, -5328).
SJB_HD int sharp_final_y(int r, int g, int b) { return sharp_clip8((19595 * r + 38469 * g + 7471 * b + (1 << 17)) >> 18); }
SJB_HD int sharp_final_u(int r, int g, int b) {
  return sharp_clip8(128 + ((-11058 * r - 21709 * g + 32768 * b + (1 << 17)) >> 18));
}
SJB_HD int sharp_final_v(int r, int g, int b) {
  return sharp_clip8(128 + ((32768 * r - 27439 * g - 5328 * b + (1 << 17)) >> 18));
}

// Pictures with a side <= 4 skip the iteration (ApplySharpYUVConversion :675-690): plain luma,
// 2x2 box chroma with the last column doubled when the width is odd.
SJB_HD int sharp_small_y(int r, int g, int b) { return (19595 * r + 38469 * g + 7471 * b + 32768) >> 16; }

// ---------------------------------------------------------------------------------------------
// riskiness: 7-level YUV index of one pixel (colors_rgb.cc:1080-1117)
// ---------------------------------------------------------------------------------------------
SJB_HD int risk_level(int v) { return static_cast<int>((static_cast<uint32_t>(v) * (0x0101u * 6u)) >> 16); }
SJB_HD int risk_index(int r, int g, int b) {
  const int y = (19595 * r + 38469 * g + 7471 * b + 32768) >> 16;
  const int u = sharp_clip8(128 + ((-11059 * r - 21709 * g + 32768 * b + 32768) >> 16));
  const int v = sharp_clip8(128 + ((32768 * r - 27439 * g - 5329 * b + 32768) >> 16));
  return risk_level(y) + 7 * risk_level(u) + 49 * risk_level(v);
}
enum { kRiskLevels3 = 343, kRiskTableBytes = 343 * 343, kRiskNoise = 4, kRiskGrayMin = 168 };

}  // namespace sjb
