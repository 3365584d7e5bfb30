// host_codec.cc -- see host_codec.h.  Compile with -ffp-contract=off: the floating-point parts
// must evaluate in the reference's order without fused multiply-adds.
#include "host_codec.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace sjb {

namespace {

const uint8_t kZigzagToNatural[64] = SJB_ZIGZAG_INIT;

// ITU-T T.81 annex K.1 (luma, chroma)
const uint8_t kAnnexK1[2][64] = {
    {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
     14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
     18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99},
    {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
     24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};

// ITU-T T.81 annex K.3: code-length counts; the symbol lists are generated below.
const uint8_t kK3DcCounts[2][16] = {{0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},
                                    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0}};
const uint8_t kK3AcCounts[2][16] = {{0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125},
                                    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119}};
// The first 37 (luma) / 43 (chroma) AC symbols of K.3 in code order; the remaining ones are,
// in both tables, all not-yet-listed (run,size) pairs with size 1..10 in increasing numeric order.
const uint8_t kK3AcHead[2][43] = {
    {0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61,
     0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52,
     0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82},
    {0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61,
     0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33,
     0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1}};
const int kK3AcHeadLen[2] = {37, 43};

int BitLength(int v) { return v > 0 ? 32 - __builtin_clz((unsigned)v) : 0; }

// Inner sums of AnalyseHistograms: sum of the 32-bit (wrapping, then signed) products of the bin
// counts with two per-step tables.  Compiled twice -- for AVX2 (8 products per instruction) and for
// the baseline ISA -- and picked once at run time from the CPU the library finds itself on.
#define SJB_BIN_SUMS_BODY                                                                     \
  int64_t bs = 0, ds = 0;                                                                      \
  for (int i = 0; i < n; ++i) {                                                                \
    bs += static_cast<int32_t>(static_cast<uint32_t>(h[i]) * static_cast<uint32_t>(tb[i]));    \
    ds += static_cast<int32_t>(static_cast<uint32_t>(h[i]) * static_cast<uint32_t>(te[i]));    \
  }                                                                                            \
  *bits_sum = bs;                                                                              \
  *dist_sum = ds;
void BinSumsBaseline(const int32_t* h, const int32_t* tb, const int32_t* te, int n, int64_t* bits_sum,
                     int64_t* dist_sum) {
  SJB_BIN_SUMS_BODY
}
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) void BinSumsAvx2(const int32_t* h, const int32_t* tb, const int32_t* te, int n,
                                                 int64_t* bits_sum, int64_t* dist_sum) {
  SJB_BIN_SUMS_BODY
}
#endif
typedef void (*BinSumsFn)(const int32_t*, const int32_t*, const int32_t*, int, int64_t*, int64_t*);
BinSumsFn PickBinSums() {
#if defined(__x86_64__) && defined(__GNUC__)
  __builtin_cpu_init();
  const char* off = getenv("SJPEG_B200_NO_AVX2");     // lets the tests run the baseline loop too
  if (__builtin_cpu_supports("avx2") && !(off && off[0] == '1')) return BinSumsAvx2;
#endif
  return BinSumsBaseline;
}

}  // namespace

const uint8_t* DefaultMatrix(int idx) { return kAnnexK1[idx != 0]; }

void ScaleMatrix(const uint8_t in[64], float percent, uint8_t out[64]) {
  const float f = percent / 100.f;
  for (int i = 0; i < 64; ++i) {
    const int v = static_cast<int>(in[i] * f + .5f);
    out[i] = static_cast<uint8_t>(std::min(255, std::max(1, v)));
  }
}

void QualityToMatrices(float quality, uint8_t out[2][64]) {
  // libjpeg-6b style mapping, floored (quantize.cc:77-82)
  float scale;
  if (quality <= 0) scale = 5000;
  else if (quality < 50) scale = 5000 / quality;
  else if (quality < 100) scale = 2 * (100 - quality);
  else scale = 0;
  scale = floorf(scale);
  ScaleMatrix(kAnnexK1[0], scale, out[0]);
  ScaleMatrix(kAnnexK1[1], scale, out[1]);
}

void MinMatrixWithTolerance(const uint8_t in[64], int tolerance, uint8_t out[64]) {
  for (int i = 0; i < 64; ++i) {
    const int v = (in[i] * (256 - tolerance)) >> 8;
    out[i] = static_cast<uint8_t>(std::min(255, std::max(1, v)));
  }
}

bool FinalizeQuantizer(uint8_t quant[64], const uint8_t min_quant[64], int q_bias, QuantTab* out) {
  bool ok = true;
  for (int z = 0; z < 64; ++z) {
    const int i = kZigzagToNatural[z];       // table entries are stored by zig-zag position
    if (quant[i] < min_quant[i]) quant[i] = min_quant[i];
    const uint32_t q = quant[i];
    // reciprocal in 16-bit fixed point; q == 1 cannot be represented and uses 0xffff with the
    // neutral bias instead (quantize.cc:128-138)
    const uint32_t recip = (q == 1) ? 0xffffu : (((1u << 16) + q / 2) / q) & 0xffffu;
    const uint32_t bias8 = (q == 1 || i == 0) ? 0x80u : static_cast<uint32_t>(q_bias);
    const uint32_t bias = ((((bias8 * q) << 4) + 128) >> 8) & 0xffffu;
    const int thresh = static_cast<int>(((1u << 20) + recip - 1) / recip) - static_cast<int>(bias);
    const int64_t cpos = static_cast<int64_t>(bias) * recip;
    // the fused form needs: threshold expressible as the reference's uint16 (no wrap) and the
    // products inside int32
    if (thresh < 0 || thresh > 0xffff || cpos + 17000LL * recip >= (1LL << 31)) ok = false;
    out->e[z][0] = static_cast<int32_t>(recip);
    out->e[z][1] = static_cast<int32_t>(cpos);
  }
  return ok;
}

void DefaultHuffSpec(bool ac, int chroma, HuffSpec* out) {
  memset(out, 0, sizeof(*out));
  const int c = chroma ? 1 : 0;
  if (!ac) {
    memcpy(out->bits, kK3DcCounts[c], 16);
    for (int i = 0; i < 12; ++i) out->syms[i] = static_cast<uint8_t>(i);
    out->nb_syms = 12;
    return;
  }
  memcpy(out->bits, kK3AcCounts[c], 16);
  bool used[256] = {false};
  int n = 0;
  for (; n < kK3AcHeadLen[c]; ++n) {
    out->syms[n] = kK3AcHead[c][n];
    used[kK3AcHead[c][n]] = true;
  }
  for (int s = 0; s < 256; ++s) {
    const int size = s & 15;
    if (!used[s] && size >= 1 && size <= 10) out->syms[n++] = static_cast<uint8_t>(s);
  }
  out->nb_syms = n;   // 162
}

void CodesFromSpec(const HuffSpec& spec, uint32_t* table) {
  uint32_t code = 0;
  int k = 0;
  for (int len = 1; len <= 16; ++len) {
    for (int j = 0; j < spec.bits[len - 1]; ++j) {
      table[spec.syms[k++]] = (code << 16) | static_cast<uint32_t>(len);
      ++code;
    }
    code <<= 1;
  }
}

// Length-limited Huffman construction with a reserved all-ones code word, reproducing the
// reference's tie-breaking (entropy.cc:254-430): keys are (freq << 9) | symbol sorted in
// decreasing order, the pseudo symbol 'size' has key (1 << 9) | size, a merged node takes the
// key of its first member plus the frequency part of the second and is re-inserted before all
// strictly smaller keys.
void OptimalHuffSpec(const uint32_t* freq, int size, HuffSpec* out) {
  const int kMaxLen = 32;
  int depth[257], next[257], tail[257];
  uint64_t keys[257];
  int n = 0;
  for (int s = 0; s <= size; ++s) {
    depth[s] = 0;
    next[s] = -1;
    tail[s] = s;
    if (s < size && freq[s] > 0) keys[n++] = (static_cast<uint64_t>(freq[s]) << 9) | s;
  }
  out->nb_syms = n;
  std::sort(keys, keys + n, [](uint64_t a, uint64_t b) { return a > b; });   // keys are distinct
  keys[n++] = (1ull << 9) | static_cast<uint64_t>(size);

  for (int last = n - 1; last >= 1; --last) {
    const uint64_t a = keys[last - 1], b = keys[last];
    const int sa = static_cast<int>(a & 0x1ff), sb = static_cast<int>(b & 0x1ff);
    next[tail[sa]] = sb;          // concatenate the member lists
    tail[sa] = tail[sb];
    for (int s = sa; s >= 0; s = next[s]) ++depth[s];
    const uint64_t merged = a + (b & ~0x1ffull);
    int k = last - 1;
    for (; k > 0 && keys[k - 1] < merged; --k) keys[k] = keys[k - 1];
    keys[k] = merged;
  }

  uint8_t count[kMaxLen];
  memset(count, 0, sizeof(count));
  int longest = 0;
  for (int s = 0; s <= size; ++s) {
    if (depth[s] == 0) continue;
    if (depth[s] > kMaxLen) depth[s] = kMaxLen;
    ++count[depth[s] - 1];
    longest = std::max(longest, depth[s]);
  }
  // symbols by increasing code length, numeric order inside a length (pseudo symbol left out)
  int first[kMaxLen];
  for (int l = 0, pos = 0; l < longest; ++l) {
    first[l] = pos;
    pos += count[l];
  }
  for (int s = 0; s < size; ++s) {
    if (depth[s] > 0) out->syms[first[depth[s] - 1]++] = static_cast<uint8_t>(s);
  }
  // shorten codes longer than 16 bits (entropy.cc:392-416)
  for (int l = longest - 1; l >= 16; --l) {
    while (count[l] > 0) {
      int k = l - 2;
      while (count[k] == 0) --k;
      count[l] -= 2;
      count[l - 1] += 1;
      count[k] -= 1;
      count[k + 1] += 2;
    }
  }
  int l = 15;
  while (count[l] == 0) --l;
  --count[l];   // drop the pseudo symbol: it sits alone at the deepest level
  memcpy(out->bits, count, 16);
}

void AnalyseHistograms(const int32_t* counts, int nb_comps, uint8_t quant[2][64],
                       const uint8_t min_quant[2][64], int qdelta_max_luma, int qdelta_max_chroma) {
  enum { kDeltaMin = -12, kDeltaMax = 12, kNumDelta = kDeltaMax - kDeltaMin + 1, kShift = 2 };
  static const float kWeight[kNumDelta] = {0, 0, 0,  0,  0,  1,  5, 16, 43, 94, 164, 228, 255,
                                           228, 164, 94, 43, 16, 5, 1, 0,  0,  0,  0,  0};
  const double kDensity = 0.5, kCorrelation = 0.5, kFallbackLambda = 128.;
  const uint64_t kNeverTouched = 0x103ull;   // positions 0, 1 and 8
  // The int products wrap like the compiled reference's 32-bit multiplies (histogram.cc:241-244).
  // This analysis sits between two kernel launches with the GPU idle, and the reference's form of
  // the bin loop (two data-dependent branches and two serial double-precision sums per bin) cost
  // 0.24 ms (photographic) to 1.4 ms (noisy) per 4K picture -- more than the device pipeline.  The
  // loop below gives the same numbers: every term is a 32-bit integer and the sums stay far below
  // 2^53, so accumulating them in int64 and converting once equals the double-precision running
  // sum bit for bit; a level of zero leaves e = v and BitLength(0) = 0, so the reference's two
  // branches are one expression (h * v * v and h * (v * v) agree modulo 2^32), and an empty bin
  // adds zero either way; the bit length and the squared error of a bin depend on the quantiser
  // step only, so they come from a table built once (255 steps x 128 bins).
  struct BinTables {
    int32_t bits[256][kHistoBins];    // BitLength(level) of bin i quantised with step q
    int32_t err2[256][kHistoBins];    // (v - level * q)^2
    BinTables() {
      for (int q = 1; q < 256; ++q) {
        const int recip = ((1 << 16) + q - 1) / q;
        for (int i = 0; i < kHistoBins; ++i) {
          const int v = (i << kShift) + (1 << (kShift - 1));   // bin centre
          const int level = (v * recip + 32768) >> 16;
          const int e = v - level * q;
          bits[q][i] = BitLength(level);
          err2[q][i] = e * e;
        }
      }
      for (int i = 0; i < kHistoBins; ++i) bits[0][i] = err2[0][i] = 0;
    }
  };
  static const BinTables* const kBins = new BinTables();   // 256 KB, lives as long as the process
  static const BinSumsFn bin_sums = PickBinSums();

  for (int idx = (nb_comps > 1) ? 1 : 0; idx >= 0; --idx) {
    const int32_t* histo = counts + static_cast<size_t>(idx) * 64 * kHistoStride;
    const int delta_top = ((idx == 0) ? qdelta_max_luma : qdelta_max_chroma) - kDeltaMin;
    float rate[64][kNumDelta], dist[64][kNumDelta];
    double num = 0., den = 0.;
    uint64_t skip = kNeverTouched;
    for (int pos = 0; pos < 64; ++pos) {
      if ((skip >> pos) & 1) continue;
      const int32_t* h = histo + pos * kHistoStride;
      const int q0 = quant[idx][pos], qmin = min_quant[idx][pos];
      int total = 0, last = 0;
      for (int i = 0; i < kHistoBins; ++i) {
        total += h[i];
        if (h[i]) last = i + 1;
      }
      if (total < kDensity * last) {
        skip |= 1ull << pos;
        continue;
      }
      double sw = 0., sx = 0., sxx = 0., syy1 = 0., sy1 = 0., sxy1 = 0., sy2 = 0., sxy2 = 0.;
      for (int d = 0; d < kNumDelta; ++d) {
        const int q = q0 + d + kDeltaMin;
        if (q < qmin || q > 255) {
          dist[pos][d] = FLT_MAX;
          rate[pos][d] = 0;
          continue;
        }
        if (d > delta_top && kWeight[d] == 0) {   // neither fitted (weight 0) nor a candidate below
          dist[pos][d] = FLT_MAX;
          rate[pos][d] = 0;
          continue;
        }
        const int32_t* tb = kBins->bits[q];
        const int32_t* te = kBins->err2[q];
        int64_t bits_sum = 0, dist_sum = 0;
        bin_sums(h, tb, te, last, &bits_sum, &dist_sum);
        const double bsum = static_cast<double>(bits_sum), dsum = static_cast<double>(dist_sum);
        dist[pos][d] = static_cast<float>(dsum);
        rate[pos][d] = static_cast<float>(bsum);
        const double w = kWeight[d];
        if (w > 0.) {
          const double x = static_cast<double>(d + kDeltaMin);
          sw += w;
          sx += w * x;
          sxx += w * x * x;
          sy1 += w * dsum;
          syy1 += w * dsum * dsum;
          sy2 += w * bsum;
          sxy1 += w * dsum * x;
          sxy2 += w * bsum * x;
        }
      }
      const double cov = sw * sxy1 - sx * sy1;
      if (cov * cov < kCorrelation * (sw * sxx - sx * sx) * (sw * syy1 - sy1 * sy1)) {
        skip |= 1ull << pos;
        continue;
      }
      num += cov;
      den += sw * sxy2 - sx * sy2;
    }
    double lambda = kFallbackLambda;
    if (num > 1000. && den < -10.) {
      lambda = -num / den;
      if (lambda < 1.) lambda = 1.;
    }
    for (int pos = 0; pos < 64; ++pos) {
      if ((skip >> pos) & 1) continue;
      float best = FLT_MAX;
      int best_delta = 0;
      for (int d = 0; d <= delta_top; ++d) {
        if (dist[pos][d] < FLT_MAX) {
          const float score = dist[pos][d] + lambda * rate[pos][d];
          if (score < best) {
            best = score;
            best_delta = d + kDeltaMin;
          }
        }
      }
      quant[idx][pos] = static_cast<uint8_t>(quant[idx][pos] + best_delta);
    }
  }
}

bool MakeGeometry(int yuv_mode, int width, int height, FrameGeometry* g) {
  if (width <= 0 || height <= 0 || width > 65535 || height > 65535) return false;   // enc.cc:406
  g->yuv_mode = yuv_mode;
  g->width = width;
  g->height = height;
  switch (yuv_mode) {
    case kYuv420: g->nb_comps = 3; g->mcu_blocks = 6; g->luma_blocks = 4; g->mcu_size = 16; break;
    case kYuv444: g->nb_comps = 3; g->mcu_blocks = 3; g->luma_blocks = 1; g->mcu_size = 8; break;
    case kYuv400: g->nb_comps = 1; g->mcu_blocks = 1; g->luma_blocks = 1; g->mcu_size = 8; break;
    default: return false;
  }
  g->mcus_x = (width + g->mcu_size - 1) / g->mcu_size;
  g->mcus_y = (height + g->mcu_size - 1) / g->mcu_size;
  return true;
}

void AppendHeaders(const FrameGeometry& g, const uint8_t quant[2][64], const HuffSpec spec[4],
                   std::vector<uint8_t>* out) {
  auto put = [out](int v) { out->push_back(static_cast<uint8_t>(v)); };
  auto put16 = [&put](int v) { put(v >> 8); put(v); };
  // SOI + JFIF 1.01 APP0, 1:1 aspect, no thumbnail
  put16(0xffd8); put16(0xffe0); put16(16);
  for (char ch : {'J', 'F', 'I', 'F'}) put(ch);
  put(0); put16(0x0101); put(0); put16(1); put16(1); put16(0);
  // DQT: 8-bit tables in zig-zag order
  const int nb_tables = (g.nb_comps == 1) ? 1 : 2;
  put16(0xffdb); put16(2 + 65 * nb_tables);
  for (int t = 0; t < nb_tables; ++t) {
    put(t);
    for (int i = 0; i < 64; ++i) put(quant[t][kZigzagToNatural[i]]);
  }
  // SOF0
  put16(0xffc0); put16(8 + 3 * g.nb_comps); put(8); put16(g.height); put16(g.width); put(g.nb_comps);
  for (int c = 0; c < g.nb_comps; ++c) {
    put(c + 1);
    put((c == 0 && g.yuv_mode == kYuv420) ? 0x22 : 0x11);
    put(c > 0 ? 1 : 0);
  }
  // DHT, order: luma DC, luma AC, chroma DC, chroma AC.  spec[] = {dc0, dc1, ac0, ac1}
  for (int t = 0; t < nb_tables; ++t) {
    for (int ac = 0; ac <= 1; ++ac) {
      const HuffSpec& h = spec[ac * 2 + t];
      put16(0xffc4); put16(2 + 1 + 16 + h.nb_syms); put((ac << 4) | t);
      out->insert(out->end(), h.bits, h.bits + 16);
      out->insert(out->end(), h.syms, h.syms + h.nb_syms);
    }
  }
  // SOS
  put16(0xffda); put16(6 + 2 * g.nb_comps); put(g.nb_comps);
  for (int c = 0; c < g.nb_comps; ++c) { put(c + 1); put(c > 0 ? 0x11 : 0x00); }
  put(0); put(63); put(0);
}

// ---------------------------------------------------------------------------------------------
// sharp-YUV gamma tables and the riskiness decision
// ---------------------------------------------------------------------------------------------
void MakeSharpGammaTables(uint32_t g2l[1024], uint32_t l2g[34]) {   // yuv_convert.cc:102-151
  const double a = 0.099, thresh = 0.018, gamma = 1. / 0.45;
  const int max_y = 1023, lin_bits = 14, nodes = 32;
  for (int v = 0; v <= max_y; ++v) {
    const double g = (1. / max_y) * v;
    double value;
    if (g <= thresh * 4.5) {
      value = g / 4.5;
    } else {
      const double a_rec = 1. / (1. + a);
      value = pow(a_rec * (g + a), gamma);
    }
    g2l[v] = static_cast<uint32_t>(value * static_cast<double>(1 << lin_bits) + .5);
  }
  for (int v = 0; v <= nodes; ++v) {
    const double g = (1. / nodes) * v;
    double value;
    if (g <= thresh) value = 4.5 * g;
    else value = (1. + a) * pow(g, 1. / gamma) - a;
    l2g[v] = static_cast<uint32_t>(max_y * value) + (1 << lin_bits >> 1);
  }
  l2g[nodes + 1] = l2g[nodes];
}

int RiskinessDecision(unsigned long long score_sum, unsigned long long score_num, unsigned long long gray_num,
                      int width, int height, float* risk) {   // jpeg_tools.cc:212-235
  const double count = static_cast<double>(score_num);
  double gray_count = static_cast<double>(gray_num);
  double total_score = (count > 0) ? static_cast<double>(static_cast<long long>(score_sum)) / count : 0.;
  const double num_samples = (width - 1.) * (height - 1.);
  if (num_samples > 0.) gray_count /= num_samples;
  const double frac = 100. * count / (static_cast<double>(width) * height);
  if (frac < 1.) total_score = 0.;
  total_score = (total_score > 25.) ? 100. : total_score * 100. / 25.;
  if (risk != nullptr) *risk = static_cast<float>(total_score);
  return (gray_count > 0.995) ? 4 : (total_score < 40.0) ? 1 : (total_score < 70.0) ? 2 : 3;
}

}  // namespace sjb
