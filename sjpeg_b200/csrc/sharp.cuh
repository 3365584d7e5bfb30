// sharp.cuh -- launch wrappers of sharp.cu (sharp RGB->YUV 4:2:0 pre-pass, riskiness analyser).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "sharp_ops.cuh"

namespace sjb {

enum { kSharpRefineThreads = 512 };

// one picture; every pointer is a device pointer
struct SharpJob {
  const uint8_t* rgb;           // packed RGB, row 0
  long long stride;             // bytes between rows, may be negative
  int width, height;            // picture
  int w, h, uv_w, uv_h;         // even-padded size and its half
  uint16_t* y_state;            // [kSharpIterations + 1][h][w]        state 0 = import, k+1 = after iteration k
  int16_t* uv_state;            // [kSharpIterations + 1][uv_h][3][uv_w]
  uint16_t* target_y;           // [h][w]
  int16_t* target_uv;           // [uv_h][3][uv_w]
  unsigned long long* diff;     // [kSharpIterations] sum of |luma correction| per iteration
  unsigned* progress;           // [kSharpIterations] chroma rows finished per iteration
  const uint32_t* g2l;          // [1024]
  const uint32_t* l2g;          // [34]
  uint8_t* y_out;               // width x height, pitch = width
  uint8_t* u_out;               // (width+1)/2 x (height+1)/2, tight
  uint8_t* v_out;
};

struct SharpLayout {            // byte offsets inside the scratch buffer
  size_t y_state, uv_state, target_y, target_uv, counters;
};
// scratch bytes the conversion of a width x height picture needs (0-size pictures excluded)
size_t SharpScratchBytes(int width, int height, SharpLayout* lay);

// import -> pipelined refinement -> finish, all on stream s.  *launches += kernels launched.
cudaError_t LaunchSharpYuv(const uint8_t* rgb, long long stride, int width, int height, uint8_t* scratch,
                           const uint32_t* g2l, const uint32_t* l2g, uint8_t* y_out, uint8_t* u_out, uint8_t* v_out,
                           cudaStream_t s, int* launches);

// sums[0..2] = score sum, number of scores above the noise level, number of neutral-chroma pixels
// over the (width-1) x (height-1) pixel triples (jpeg_tools.cc:194-211).  table: 343*343 bytes.
cudaError_t LaunchRiskiness(const uint8_t* rgb, long long stride, int width, int height, const uint8_t* table,
                            unsigned long long* sums, int sm_count, cudaStream_t s);

}  // namespace sjb
