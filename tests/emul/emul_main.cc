// TEST INFRASTRUCTURE: CPU emulation of the kernel pipeline, built from the product's own
// __host__ __device__ block functions (sjpeg_b200/csrc/block_ops.cuh) and host codec
// (host_codec.cc) compiled by g++.  It lets the no-GPU test tier check the per-block integer
// arithmetic, the bit packer / stuffing logic and the host-side table builders against the oracle
// before anything runs on a B200.  The thread/CTA glue of kernels.cu is NOT covered here -- that is
// what the -m gpu tests are for.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../sjpeg_b200/csrc/block_ops.cuh"
#include "../../sjpeg_b200/csrc/host_codec.h"
#include "../../sjpeg_b200/csrc/sharp_ops.cuh"

using namespace sjb;

namespace {

struct Img {
  const uint8_t* pix; long long stride; int w, h, fmt;
  void get(int x, int y, int* r, int* g, int* b) const {
    if (x > w - 1) x = w - 1;
    if (y > h - 1) y = h - 1;
    const int ps = fmt == kFmtRGB ? 3 : 4;
    const uint8_t* p = pix + y * stride + (long long)x * ps;
    *r = p[fmt == kFmtBGRA ? 2 : 0]; *g = p[1]; *b = p[fmt == kFmtBGRA ? 0 : 2];
  }
};

void LumaSamples(const Img& im, int x0, int y0, int (&v)[64]) {
  for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
    int r, g, b; im.get(x0 + x, y0 + y, &r, &g, &b); v[8 * y + x] = rgb_to_y(r, g, b);
  }
}

// mirrors f1_generic_kernel (kernels.cu) for block k of MCU (mx,my)
void BlockSamples(const Img& im, int mode, int mx, int my, int k, int (&v)[64]) {
  if (mode == kYuv420) {
    const int X = 16 * mx, Y = 16 * my;
    if (k < 4) {
      const int sub_w = im.w - X, sub_h = im.h - Y;
      int src = -1;
      if (k == 1 && sub_w <= 8) src = 0;
      if (k >= 2 && sub_h <= 8) src = (sub_w > 8) ? 1 : 0;
      else if (k == 3 && sub_w <= 8) src = 2;
      if (src < 0) { LumaSamples(im, X + 8 * (k & 1), Y + 8 * (k >> 1), v); return; }
      LumaSamples(im, X + 8 * (src & 1), Y + 8 * (src >> 1), v);
      int sum = 0;
      for (int i = 0; i < 64; ++i) sum += v[i];
      for (int i = 0; i < 64; ++i) v[i] = (sum + 32) >> 6;
    } else {
      for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
        int sr = 0, sg = 0, sb = 0;
        for (int q = 0; q < 4; ++q) {
          int r, g, b; im.get(X + 2 * x + (q & 1), Y + 2 * y + (q >> 1), &r, &g, &b);
          sr += r; sg += g; sb += b;
        }
        v[8 * y + x] = (k == 4) ? rgb4_to_u(sr, sg, sb) : rgb4_to_v(sr, sg, sb);
      }
    }
  } else {
    for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
      int r, g, b; im.get(8 * mx + x, 8 * my + y, &r, &g, &b);
      v[8 * y + x] = (k == 0) ? rgb_to_y(r, g, b) : (k == 1) ? rgb_to_u(r, g, b) : rgb_to_v(r, g, b);
    }
  }
}

const int kZZ[64] = SJB_ZIGZAG_INIT;

void QuantizeBlock(const int (&v)[64], const QuantTab& t, int16_t* zz, uint32_t* mask) {
  uint32_t m = 0;
  for (int i = 0; i < 64; ++i) {
    const int n = kZZ[i];
    const int q = quantize_coeff(v[n], t.e[i][0], t.e[i][1]);
    zz[i] = (int16_t)q;
    if (q) m |= 1u << (i >> 3);
  }
  *mask = m;
}

int DcPred(const int16_t* zz, size_t g, int k, int mb, int lb) {
  size_t prev;
  if (k < lb) { if (k > 0) prev = g - 1; else if (g == 0) return 0; else prev = g - mb + lb - 1; }
  else { if (g < (size_t)mb) return 0; prev = g - mb; }
  return zz[prev * 64];
}

struct Loader {
  const int16_t* p;
  Words4 operator()(int c) const {
    Words4 r;
    for (int j = 0; j < 4; ++j) r.w[j] = (uint16_t)p[8 * c + 2 * j] | ((uint32_t)(uint16_t)p[8 * c + 2 * j + 1] << 16);
    return r;
  }
  int value(int pos) const { return p[pos]; }
};
struct WordOut {
  std::vector<uint32_t>* w;
  void or_word(uint64_t i, uint32_t v) { (*w)[i] |= v; }
  void set_word(uint64_t i, uint32_t v) { (*w)[i] = v; }
};
struct Stats {
  uint32_t* f;
  void one(int s) { ++f[s]; }
  void many(int s, int n) { f[s] += n; }
};

}  // namespace

extern "C" {
static void AnalyseDeviceForm(const int32_t* counts, int nb_comps, uint8_t quant[2][64], const uint8_t min_quant[2][64],
                              int qdl, int qdc);

// raw (unquantised, natural order) coefficients of the whole picture
void emul_coeffs(const uint8_t* pix, int w, int h, long long stride, int mode, int fmt, int16_t* out) {
  FrameGeometry g;
  if (!MakeGeometry(mode, w, h, &g)) return;
  Img im = {pix, stride, w, h, fmt};
  for (int my = 0; my < g.mcus_y; ++my) for (int mx = 0; mx < g.mcus_x; ++mx) for (int k = 0; k < g.mcu_blocks; ++k) {
    int v[64];
    BlockSamples(im, mode, mx, my, k, v);
    fdct64(v);
    for (int i = 0; i < 64; ++i) *out++ = (int16_t)v[i];
  }
}

void emul_default_ac_syms(int chroma, uint8_t* bits, uint8_t* syms, int* n) {
  HuffSpec s; DefaultHuffSpec(true, chroma, &s);
  memcpy(bits, s.bits, 16); memcpy(syms, s.syms, s.nb_syms); *n = s.nb_syms;
}

// full pipeline, methods 0..8.  Returns size; *out malloc()ed.
size_t emul_encode(const uint8_t* pix, int w, int h, long long stride, int mode, int fmt, int method,
                   const uint8_t* quant_in /*[2][64]*/, int q_bias, int qdl, int qdc, uint8_t** out) {
  FrameGeometry g;
  if (!MakeGeometry(mode, w, h, &g)) return 0;
  method = method < 0 ? 0 : method > 8 ? 8 : method;
  const bool adaptive = method >= 3, optimize = method != 0 && method != 3, trellis = method >= 7;
  Img im = {pix, stride, w, h, fmt};
  uint8_t quant[2][64], minq[2][64];
  memcpy(quant, quant_in, 128);
  memset(minq, 1, 128);
  QuantTabs qt;
  for (int i = 0; i < 2; ++i) if (!FinalizeQuantizer(quant[i], minq[i], q_bias, &qt.m[i])) return 0;
  const size_t nb = g.nb_blocks();
  std::vector<int16_t> raw(nb * 64), zz(nb * 64);
  std::vector<uint32_t> mask(nb);
  emul_coeffs(pix, w, h, stride, mode, fmt, raw.data());
  (void)im;
  if (adaptive) {
    std::vector<int32_t> counts(2 * 64 * kHistoStride, 0);
    for (size_t b = 0; b < nb; ++b) {
      const int m = ((int)(b % g.mcu_blocks) >= g.luma_blocks) ? 1 : 0;
      for (int i = 0; i < 64; ++i) {
        const int a = abs(raw[b * 64 + i]) >> 2;
        if (a < 128) ++counts[(m * 64 + i) * kHistoStride + a];
      }
    }
    // the analysis in the form the device runs (block_ops.cuh aq_*), cross-checked against the host form
    uint8_t quant_host[2][64];
    memcpy(quant_host, quant, 128);
    AnalyseHistograms(counts.data(), g.nb_comps, quant_host, minq, qdl, qdc);
    AnalyseDeviceForm(counts.data(), g.nb_comps, quant, minq, qdl, qdc);
    if (memcmp(quant, quant_host, 128) != 0) return 0;
    for (int i = (g.nb_comps > 1 ? 1 : 0); i >= 0; --i) {
      if (!FinalizeQuantizer(quant[i], minq[i], q_bias, &qt.m[i])) return 0;
      const int zz[64] = SJB_ZIGZAG_INIT;
      for (int z = 0; z < 64; ++z) {       // the device's form of the same constants
        int32_t iq, cpos;
        if (!aq_finalize_entry(quant[i][zz[z]], zz[z] == 0, q_bias, &iq, &cpos)) return 0;
        if (iq != qt.m[i].e[z][0] || cpos != qt.m[i].e[z][1]) return 0;
      }
    }
  }
  // rate model of the trellis = code lengths of the DEFAULT AC tables (enc.cc:334)
  uint8_t default_len[2][256];
  for (int c = 0; c < 2; ++c) {
    HuffSpec hs;
    DefaultHuffSpec(true, c, &hs);
    uint32_t codes[256];
    memset(codes, 0, sizeof(codes));
    CodesFromSpec(hs, codes);
    for (int i = 0; i < 256; ++i) default_len[c][i] = (uint8_t)(codes[i] & 0xff);
  }
  for (size_t b = 0; b < nb; ++b) {
    const int c = ((int)(b % g.mcu_blocks) >= g.luma_blocks) ? 1 : 0;
    if (trellis) {
      // the kernel's dynamic programme (trellis_block_v2) on plain-array storage; the list-based
      // form it was derived from (trellis_block) must agree with it on every block
      uint32_t words[32];
      for (int i = 0; i < 32; ++i) words[i] = (uint16_t)raw[b * 64 + 2 * i] | ((uint32_t)(uint16_t)raw[b * 64 + 2 * i + 1] << 16);
      TrellisHostMem mem;
      TrellisHostTab tab = {qt.m[c].e, quant[c], default_len[c]};
      mask[b] = trellis_block_v2(words, mem, tab);
      for (int i = 0; i < 64; ++i) zz[b * 64 + i] = (int16_t)sjb_half(mem.out[i >> 1], i & 1);
      TrellisScratch scratch;
      int16_t ref_out[64];
      const uint32_t ref_mask = trellis_block(&raw[b * 64], quant[c], qt.m[c].e, default_len[c], ref_out, scratch);
      if (ref_mask != mask[b] || memcmp(ref_out, &zz[b * 64], sizeof(ref_out)) != 0) return 0;
      continue;
    }
    int v[64];
    for (int i = 0; i < 64; ++i) v[i] = raw[b * 64 + i];
    QuantizeBlock(v, qt.m[c], &zz[b * 64], &mask[b]);
  }
  HuffSpec spec[4];
  for (int i = 0; i < 4; ++i) DefaultHuffSpec(i >= 2, i & 1, &spec[i]);
  if (optimize) {
    uint32_t freq[2][272];
    memset(freq, 0, sizeof(freq));
    for (size_t b = 0; b < nb; ++b) {
      const int k = (int)(b % g.mcu_blocks), c = k >= g.luma_blocks;
      Stats st = {freq[c]};
      block_symbol_stats(Loader{&zz[b * 64]}, mask[b], zz[b * 64], DcPred(zz.data(), b, k, g.mcu_blocks, g.luma_blocks), st);
    }
    for (int c = 0; c < (g.nb_comps == 1 ? 1 : 2); ++c) {
      OptimalHuffSpec(freq[c] + 256, 12, &spec[c]);
      OptimalHuffSpec(freq[c], 256, &spec[2 + c]);
    }
  }
  CodeTabs tabs;
  memset(&tabs, 0, sizeof(tabs));
  for (int c = 0; c < 2; ++c) { CodesFromSpec(spec[c], tabs.dc[c]); CodesFromSpec(spec[2 + c], tabs.ac[c]); }
  std::vector<uint8_t> file;
  AppendHeaders(g, quant, spec, &file);
  // E1 + scan
  std::vector<uint64_t> offs(nb + 1, 0);
  for (size_t b = 0; b < nb; ++b) {
    const int k = (int)(b % g.mcu_blocks), c = k >= g.luma_blocks;
    BitCountSink s = {0};
    code_block(Loader{&zz[b * 64]}, mask[b], zz[b * 64], DcPred(zz.data(), b, k, g.mcu_blocks, g.luma_blocks), tabs.dc[c], tabs.ac[c], s);
    offs[b + 1] = offs[b] + s.total;
  }
  const uint64_t total_bits = offs[nb];
  std::vector<uint32_t> words(total_bits / 32 + 8, 0);
  WordOut wo = {&words};
  for (size_t b = nb; b-- > 0;) {   // reverse order on purpose: packing must not depend on order
    const int k = (int)(b % g.mcu_blocks), c = k >= g.luma_blocks;
    BitPackSink<WordOut> s(wo, offs[b]);
    code_block(Loader{&zz[b * 64]}, mask[b], zz[b * 64], DcPred(zz.data(), b, k, g.mcu_blocks, g.luma_blocks), tabs.dc[c], tabs.ac[c], s);
    s.finish();
  }
  const uint64_t nbytes = (total_bits + 7) >> 3;
  const unsigned pad = (unsigned)((0 - total_bits) & 7);
  for (uint64_t i = 0; i < nbytes; ++i) {
    uint32_t byte = (words[i >> 2] >> (8 * (3 - (i & 3)))) & 0xff;
    if (i == nbytes - 1 && pad) byte |= (1u << pad) - 1;
    file.push_back((uint8_t)byte);
    if (byte == 0xff) file.push_back(0);
  }
  file.push_back(0xff); file.push_back(0xd9);
  *out = (uint8_t*)malloc(file.size());
  memcpy(*out, file.data(), file.size());
  return file.size();
}

void emul_free(uint8_t* p) { free(p); }

// The analysis as the DEVICE runs it (kernels.cu: analyse_fit_kernel + analyse_pick_kernel), from the
// same block_ops.cuh functions and in the same decomposition: per position four partial sums over
// quarters of the bins, the fits in candidate order, the positions' terms added in ascending order,
// then the pick.  quant is left UNCLAMPED (the clamp to min_quant belongs to FinalizeQuantizer).
static void AnalyseDeviceForm(const int32_t* counts, int nb_comps, uint8_t quant[2][64], const uint8_t min_quant[2][64],
                              int qdl, int qdc) {
  for (int idx = (nb_comps > 1) ? 1 : 0; idx >= 0; --idx) {
    static AqFit fit[64];
    const int delta_top = (idx == 0 ? qdl : qdc) - kAqDeltaMin;
    for (int pos = 0; pos < 64; ++pos) {
      AqFit* out = &fit[pos];
      out->skip = 0;
      if (pos == 0 || pos == 1 || pos == 8) { out->skip = 1; continue; }
      const int32_t* h = counts + (static_cast<size_t>(idx) * 64 + pos) * kHistoStride;
      int total = 0, last = 0;
      for (int i = 0; i < kAqBins; ++i) {
        total += h[i];
        if (h[i]) last = i + 1;
      }
      if (total < 0.5 * last) { out->skip = 1; continue; }
      const int q0 = quant[idx][pos], qmin = min_quant[idx][pos];
      long long bits[kAqNumDelta], dist[kAqNumDelta];
      for (int d = 0; d < kAqNumDelta; ++d) {
        bits[d] = dist[d] = 0;
        if (!aq_candidate_used(q0, qmin, delta_top, d)) continue;
        for (int quarter = 0; quarter < 4; ++quarter) {
          long long bs = 0, ds = 0;
          aq_bin_sums(h, quarter * 32, std::min(last, quarter * 32 + 32), q0 + d + kAqDeltaMin, &bs, &ds);
          bits[d] += bs;
          dist[d] += ds;
        }
      }
      aq_fit_position(bits, dist, q0, qmin, delta_top, out);
    }
    double num = 0., den = 0.;
    for (int pos = 0; pos < 64; ++pos) {
      if (fit[pos].skip) continue;
      num += fit[pos].cov;
      den += fit[pos].den;
    }
    const double lambda = aq_lambda(num, den);
    for (int pos = 0; pos < 64; ++pos) {
      if (!fit[pos].skip) quant[idx][pos] = static_cast<uint8_t>(quant[idx][pos] + aq_best_delta(fit[pos], delta_top, lambda));
    }
  }
}
void emul_analyse_histo_device_form(const int32_t* counts, int nb_comps, uint8_t* quant /*[2][64]*/, const uint8_t* min_quant,
                                    int qdl, int qdc) {
  uint8_t q[2][64], mq[2][64];
  memcpy(q, quant, 128);
  memcpy(mq, min_quant, 128);
  AnalyseDeviceForm(counts, nb_comps, q, mq, qdl, qdc);
  memcpy(quant, q, 128);
}

// host_codec.cc::AnalyseHistograms on its own (the product's restatement of histogram.cc:126-315)
void emul_analyse_histo(const int32_t* counts, int nb_comps, uint8_t* quant /*[2][64]*/, const uint8_t* min_quant,
                        int qdl, int qdc) {
  uint8_t q[2][64], mq[2][64];
  memcpy(q, quant, 128);
  memcpy(mq, min_quant, 128);
  AnalyseHistograms(counts, nb_comps, q, mq, qdl, qdc);
  memcpy(quant, q, 128);
}

// the entropy kernel's walk order of busy tiles (block_ops.cuh::walk_order)
unsigned emul_walk_order(unsigned first, unsigned count, unsigned i, int mcu_blocks) {
  return walk_order(first, count, i, mcu_blocks);
}

// the sector-interleaved coefficient layout of the device arrays (block_ops.cuh): int16 offset of
// (block g, position pos) by the two routes the kernels use, and the padded block count
unsigned long long emul_coef_offset(unsigned long long g, int pos) { return coef_block_base(g) + coef_pos_offset(pos); }
unsigned long long emul_coef_chunk_offset(unsigned long long g, int c) { return coef_block_base(g) + 8ull * coef_chunk_index(c); }
unsigned long long emul_coef_padded_blocks(unsigned long long nb) { return coef_padded_blocks(nb); }

// Sharp RGB->YUV 4:2:0 as sharp.cu lays it out: import into state 0, four refinement iterations
// each writing its OWN copy of the state (the kernels run them as a pipeline), the exit rule
// applied afterwards to pick the copy to keep.  Pictures with a side <= 4 take the plain path.
struct PlainLd {
  int y(const uint16_t* p) const { return *p; }
  int uv(const int16_t* p) const { return *p; }
};
void emul_sharp_yuv(const uint8_t* rgb, int width, int height, long long stride, uint8_t* yo, uint8_t* uo,
                    uint8_t* vo) {
  const int out_uv_w = (width + 1) >> 1;
  if (width <= 4 || height <= 4) {
    const int uv_h = (height + 1) >> 1;
    for (int r = 0; r < uv_h; ++r) {
      for (int i = 0; i < out_uv_w; ++i) {
        int sum[3] = {0, 0, 0};
        for (int dy = 0; dy < 2; ++dy) {
          const int y = (2 * r + dy < height) ? 2 * r + dy : height - 1;
          for (int dx = 0; dx < 2; ++dx) {
            const int x = (2 * i + dx < width) ? 2 * i + dx : width - 1;
            const uint8_t* px = rgb + y * stride + 3 * x;
            for (int k = 0; k < 3; ++k) sum[k] += px[k];
            if (2 * r + dy < height && 2 * i + dx < width) yo[(size_t)y * width + x] = (uint8_t)sharp_small_y(px[0], px[1], px[2]);
          }
        }
        uo[(size_t)r * out_uv_w + i] = (uint8_t)sharp_final_u(sum[0], sum[1], sum[2]);
        vo[(size_t)r * out_uv_w + i] = (uint8_t)sharp_final_v(sum[0], sum[1], sum[2]);
      }
    }
    return;
  }
  uint32_t g2l[1024], l2g[34];
  MakeSharpGammaTables(g2l, l2g);
  const SharpTabs t = {g2l, l2g};
  const int w = (width + 1) & ~1, h = (height + 1) & ~1, uv_w = w >> 1, uv_h = h >> 1;
  const size_t yp = (size_t)w * h, up = (size_t)uv_w * 3 * uv_h;
  std::vector<uint16_t> ys((kSharpIterations + 1) * yp), ty(yp);
  std::vector<int16_t> uvs((kSharpIterations + 1) * up), tuv(up);
  for (int r = 0; r < uv_h; ++r) {
    for (int i = 0; i < uv_w; ++i) {
      sharp_import_cell(t, rgb, stride, width, height, w, uv_w, r, i, ys.data(), ty.data(), uvs.data(), tuv.data());
    }
  }
  unsigned long long diff[kSharpIterations] = {0, 0, 0, 0};
  const PlainLd ld;
  for (int it = 0; it < kSharpIterations; ++it) {
    const uint16_t* y_prev = ys.data() + it * yp;
    uint16_t* y_mine = ys.data() + (it + 1) * yp;
    const int16_t* uv_prev = uvs.data() + it * up;
    int16_t* uv_mine = uvs.data() + (it + 1) * up;
    for (int r = 0; r < uv_h; ++r) {
      const size_t row = (size_t)r * 3 * uv_w, y_row = (size_t)(2 * r) * w;
      const int16_t* above = (r > 0) ? uv_mine + row - 3 * uv_w : uv_prev;
      const int16_t* below = uv_prev + ((r < uv_h - 1) ? row + 3 * uv_w : row);
      for (int i = 0; i < uv_w; ++i) {
        SharpCellIn in;
        SharpCellAbove up;
        sharp_load_cell(ld, w, uv_w, i, y_prev + y_row, uv_prev + row, below, ty.data() + y_row, tuv.data() + row, &in);
        sharp_load_above(ld, uv_w, i, above, &up);
        diff[it] += sharp_refine_cell(t, w, uv_w, i, in, up, y_mine + y_row, uv_mine + row);
      }
    }
  }
  const int keep = sharp_final_iteration(diff, w, h) + 1;
  const uint16_t* yk = ys.data() + keep * yp;
  const int16_t* uvk = uvs.data() + keep * up;
  for (int j = 0; j < height; ++j) {
    const int16_t* uv = uvk + (size_t)(j >> 1) * 3 * uv_w;
    for (int i = 0; i < width; ++i) {
      const int W = yk[(size_t)j * w + i];
      yo[(size_t)j * width + i] = (uint8_t)sharp_final_y(uv[i >> 1] + W, uv[uv_w + (i >> 1)] + W, uv[2 * uv_w + (i >> 1)] + W);
    }
  }
  for (int r = 0; r < uv_h; ++r) {
    const int16_t* uv = uvk + (size_t)r * 3 * uv_w;
    for (int i = 0; i < uv_w; ++i) {
      uo[(size_t)r * out_uv_w + i] = (uint8_t)sharp_final_u(uv[i], uv[uv_w + i], uv[2 * uv_w + i]);
      vo[(size_t)r * out_uv_w + i] = (uint8_t)sharp_final_v(uv[i], uv[uv_w + i], uv[2 * uv_w + i]);
    }
  }
}

// riskiness sums + decision from the product's index function and host decision
int emul_riskiness(const uint8_t* rgb, int width, int height, long long stride, const uint8_t* table, float* risk) {
  unsigned long long sum = 0, num = 0, gray = 0;
  for (int j = 0; j + 1 < height; ++j) {
    const uint8_t* a = rgb + j * stride;
    const uint8_t* b = a + stride;
    for (int x = 0; x + 1 < width; ++x) {
      const int i0 = risk_index(a[3 * x], a[3 * x + 1], a[3 * x + 2]);
      const int i1 = risk_index(a[3 * x + 3], a[3 * x + 4], a[3 * x + 5]);
      const int i2 = risk_index(b[3 * x], b[3 * x + 1], b[3 * x + 2]);
      const int score = table[i0 + kRiskLevels3 * i1] + table[i0 + kRiskLevels3 * i2] + table[i1 + kRiskLevels3 * i2];
      if (score > kRiskNoise) { sum += score; num += 1; }
      gray += (i0 >= kRiskGrayMin && i0 < kRiskGrayMin + 7) ? 1 : 0;
    }
  }
  return RiskinessDecision(sum, num, gray, width, height, risk);
}

}  // extern "C"
