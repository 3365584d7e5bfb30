// sharp.cu -- sm_100a kernels of the two whole-picture passes in front of the block pipeline:
//
//   "sharp" RGB -> YUV 4:2:0 (SJPEG_YUV_SHARP; /root/reference/src/yuv_convert.cc:582-695).
//     import   : every 2x2 cell in parallel -> luma / chroma state 0 and the targets
//     refine   : the four iterations run as a PIPELINE.  Inside one iteration chroma row r needs
//                row r-1 of the same iteration (the reference refines in place, top to bottom),
//                so rows are sequential and only the columns are parallel: a thread-block cluster
//                of up to 8 CTAs splits the columns and meets at a cluster barrier per row.
//                Iteration k+1 may start row r as soon as iteration k has finished row r+1, so
//                the four iterations run concurrently, skewed by two rows, each on its own
//                cluster, chained by a progress counter in global memory (release/acquire).
//                Every iteration writes its own copy of the state; the exit rule (correction sum
//                below 3wh or growing) is evaluated afterwards and picks the copy to keep, which
//                reproduces the reference's early `break` without a host round trip.
//     finish   : state -> 8-bit Y, U, V planes (the planar 4:2:0 encoder's input)
//   riskiness analyser (SJPEG_YUV_AUTO; /root/reference/src/jpeg_tools.cc:177-236): 7-level YUV
//     index per pixel, three lookups in the caller-supplied 343x343 score table per pixel triple,
//     64-bit sums.
// Per-cell arithmetic: sharp_ops.cuh (shared with the CPU emulation of the no-GPU tests).
#include "sharp.cuh"

#include <cuda_runtime.h>
#include <stdlib.h>

namespace sjb {
namespace {

struct LdCg {   // state written by another CTA: read at L2, never from a stale L1 line
  __device__ __forceinline__ int y(const uint16_t* p) const { return __ldcg(p); }
  __device__ __forceinline__ int uv(const int16_t* p) const { return __ldcg(p); }
};

__device__ __forceinline__ void load_tabs(const SharpJob& job, uint32_t* g2l, uint32_t* l2g) {
  for (int i = threadIdx.x; i <= kSharpMaxY; i += blockDim.x) g2l[i] = job.g2l[i];
  for (int i = threadIdx.x; i < kSharpGammaTab + 2; i += blockDim.x) l2g[i] = job.l2g[i];
  __syncthreads();
}

__global__ void __launch_bounds__(256)
sharp_import_kernel(const __grid_constant__ SharpJob job) {
  __shared__ uint32_t g2l[kSharpMaxY + 1];
  __shared__ uint32_t l2g[kSharpGammaTab + 2];
  load_tabs(job, g2l, l2g);
  const SharpTabs t = {g2l, l2g};
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (i >= job.uv_w) return;
  sharp_import_cell(t, job.rgb, job.stride, job.width, job.height, job.w, job.uv_w, r, i, job.y_state, job.target_y,
                    job.uv_state, job.target_uv);
}

__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_size() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void row_barrier(unsigned nctas) {
  if (nctas == 1) {
    __syncthreads();
  } else {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid = (cluster size, kSharpIterations); cluster = (cluster size, 1, 1)
__global__ void __launch_bounds__(kSharpRefineThreads)
sharp_refine_kernel(const __grid_constant__ SharpJob job) {
  __shared__ uint32_t g2l[kSharpMaxY + 1];
  __shared__ uint32_t l2g[kSharpGammaTab + 2];
  __shared__ unsigned long long warp_sums[kSharpRefineThreads / 32];
  __shared__ unsigned seen_shared;
  load_tabs(job, g2l, l2g);
  const SharpTabs t = {g2l, l2g};
  const LdCg ld;
  const int it = blockIdx.y;
  const unsigned nctas = cluster_size(), rank = cluster_rank();
  const int w = job.w, uv_w = job.uv_w, uv_h = job.uv_h;
  const int per = (uv_w + static_cast<int>(nctas) - 1) / static_cast<int>(nctas);
  const int lo = static_cast<int>(rank) * per, hi = min(uv_w, lo + per);
  const size_t y_plane = static_cast<size_t>(w) * job.h, uv_plane = static_cast<size_t>(uv_w) * 3 * uv_h;
  const uint16_t* y_prev = job.y_state + it * y_plane;
  uint16_t* y_mine = job.y_state + (it + 1) * y_plane;
  const int16_t* uv_prev = job.uv_state + it * uv_plane;
  int16_t* uv_mine = job.uv_state + (it + 1) * uv_plane;
  const unsigned* upstream = (it > 0) ? job.progress + (it - 1) : nullptr;
  unsigned seen = 0;                 // rows the previous iteration is known to have finished
  unsigned long long diff = 0;

  // Row r of this iteration reads rows r, r+1 of the previous state.  The first cell of every
  // thread is fetched one row ahead (everything but the row above, which is the critical path),
  // so the wait on the upstream iteration covers row r+2.
  auto wait_upstream = [&](int rows_needed) {
    if (upstream == nullptr) return;
    const unsigned need = static_cast<unsigned>(min(rows_needed, uv_h));
    if (seen >= need) return;          // uniform across the CTA
    if (threadIdx.x == 0) {
      unsigned v;
      do { v = ld_acquire(upstream); } while (v < need);
      seen_shared = v;
    }
    __syncthreads();
    seen = seen_shared;
    __syncthreads();
  };
  auto load_row_cell = [&](int r, int i, SharpCellIn* in) {
    const size_t uv_row = static_cast<size_t>(r) * 3 * uv_w, y_row = static_cast<size_t>(2 * r) * w;
    const int16_t* below = uv_prev + ((r < uv_h - 1) ? uv_row + 3 * uv_w : uv_row);
    sharp_load_cell(ld, w, uv_w, i, y_prev + y_row, uv_prev + uv_row, below, job.target_y + y_row,
                    job.target_uv + uv_row, in);
  };
  const int first = lo + static_cast<int>(threadIdx.x);
  SharpCellIn ahead;
  wait_upstream(2);
  if (first < hi) load_row_cell(0, first, &ahead);
  for (int r = 0; r < uv_h; ++r) {
    const size_t uv_row = static_cast<size_t>(r) * 3 * uv_w, y_row = static_cast<size_t>(2 * r) * w;
    const int16_t* above = (r > 0) ? uv_mine + uv_row - 3 * uv_w : uv_prev;
    SharpCellIn in = ahead;
    wait_upstream(r + 3);
    for (int i = first; i < hi; i += blockDim.x) {
      SharpCellAbove up;
      sharp_load_above(ld, uv_w, i, above, &up);
      if (i != first) load_row_cell(r, i, &in);
      else if (r + 1 < uv_h) load_row_cell(r + 1, first, &ahead);     // in flight during the compute below
      diff += sharp_refine_cell(t, w, uv_w, i, in, up, y_mine + y_row, uv_mine + uv_row);
    }
    // The row barrier releases this row to the other CTAs of the cluster; the release store of
    // rank 0 then extends it (cumulatively) to the next iteration's cluster.
    row_barrier(nctas);
    if (rank == 0 && threadIdx.x == 0) st_release(job.progress + it, static_cast<unsigned>(r + 1));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = diff;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) s += warp_sums[k];
    atomicAdd(job.diff + it, s);
  }
}

// Fast form for pictures up to 8192 pixels wide (one cell per thread).  Two things are taken off
// the per-row critical path of the kernel above:
//  * the row above travels through shared memory -- every thread leaves its new chroma in a
//    double-buffered row in its CTA's shared memory and the two edge threads also drop theirs
//    into the neighbouring CTAs' halo slots over distributed shared memory -- so the next row
//    starts from a shared-memory read instead of an L2 round trip;
//  * the global stores of a row (needed by the NEXT iteration's cluster and by the finish kernel,
//    not by this one) are issued one row late, so the release at the row barrier finds them
//    drained instead of waiting for them.  The progress counter accordingly lags one row.
__device__ __forceinline__ uint32_t dsmem_addr(const void* local, unsigned target_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
               : "=r"(r) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(local))), "r"(target_rank));
  return r;
}
__device__ __forceinline__ void dsmem_store_s16(uint32_t addr, int v) {
  asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<short>(v)) : "memory");
}

__global__ void __launch_bounds__(kSharpRefineThreads)
sharp_refine_fast_kernel(const __grid_constant__ SharpJob job) {
  __shared__ uint32_t g2l[kSharpMaxY + 1];
  __shared__ uint32_t l2g[kSharpGammaTab + 2];
  __shared__ unsigned long long warp_sums[kSharpRefineThreads / 32];
  __shared__ unsigned seen_shared;
  __shared__ int16_t rowbuf[2][3][kSharpRefineThreads + 2];   // [row parity][channel][halo | cells | halo]
  load_tabs(job, g2l, l2g);
  const SharpTabs t = {g2l, l2g};
  const LdCg ld;
  const int it = blockIdx.y, tid = threadIdx.x;
  const unsigned nctas = cluster_size(), rank = cluster_rank();
  const int w = job.w, uv_w = job.uv_w, uv_h = job.uv_h;
  const int per = (uv_w + static_cast<int>(nctas) - 1) / static_cast<int>(nctas);   // <= blockDim.x
  const int lo = static_cast<int>(rank) * per, hi = min(uv_w, lo + per);
  const int i = lo + tid;
  const bool valid = i < hi;
  const bool feeds_left = valid && tid == 0 && rank > 0;                        // my cell is the left CTA's right halo
  const bool feeds_right = valid && i == hi - 1 && rank + 1 < nctas && hi < uv_w;   // ... the right CTA's left halo
  const size_t y_plane = static_cast<size_t>(w) * job.h, uv_plane = static_cast<size_t>(uv_w) * 3 * uv_h;
  const uint16_t* y_prev = job.y_state + it * y_plane;
  uint16_t* y_mine = job.y_state + (it + 1) * y_plane;
  const int16_t* uv_prev = job.uv_state + it * uv_plane;
  int16_t* uv_mine = job.uv_state + (it + 1) * uv_plane;
  const unsigned* upstream = (it > 0) ? job.progress + (it - 1) : nullptr;
  unsigned seen = 0;
  unsigned long long diff = 0;

  auto wait_upstream = [&](int rows_needed) {
    if (upstream == nullptr) return;
    const unsigned need = static_cast<unsigned>(min(rows_needed, uv_h));
    if (seen >= need) return;          // uniform across the CTA
    if (tid == 0) {
      unsigned v;
      do { v = ld_acquire(upstream); } while (v < need);
      seen_shared = v;
    }
    __syncthreads();
    seen = seen_shared;
    __syncthreads();
  };
  auto load_row_cell = [&](int r, SharpCellIn* in) {
    const size_t uv_row = static_cast<size_t>(r) * 3 * uv_w, y_row = static_cast<size_t>(2 * r) * w;
    const int16_t* below = uv_prev + ((r < uv_h - 1) ? uv_row + 3 * uv_w : uv_row);
    sharp_load_cell(ld, w, uv_w, i, y_prev + y_row, uv_prev + uv_row, below, job.target_y + y_row,
                    job.target_uv + uv_row, in);
  };
  SharpCellIn ahead;
  SharpCellOut pending;
  row_barrier(nctas);                  // every CTA of the cluster is running before anyone writes into its halo
  wait_upstream(2);
  if (valid) load_row_cell(0, &ahead);
  for (int r = 0; r < uv_h; ++r) {
    const SharpCellIn in = ahead;
    if (valid && r > 0) {              // last row's results: out to global memory, one row late
      sharp_store_cell(w, uv_w, i, pending, y_mine + static_cast<size_t>(2 * (r - 1)) * w,
                       uv_mine + static_cast<size_t>(r - 1) * 3 * uv_w);
    }
    wait_upstream(r + 3);
    if (valid) {
      SharpCellAbove up;
      if (r == 0) {
        sharp_load_above(ld, uv_w, i, uv_prev, &up);
      } else {
        const int16_t(*prev_row)[kSharpRefineThreads + 2] = rowbuf[(r - 1) & 1];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          up.v[k][1] = prev_row[k][tid + 1];
          up.v[k][0] = (i > 0) ? prev_row[k][tid] : up.v[k][1];
          up.v[k][2] = (i < uv_w - 1) ? prev_row[k][tid + 2] : up.v[k][1];
        }
      }
      if (r + 1 < uv_h) load_row_cell(r + 1, &ahead);      // in flight during the compute below
      sharp_refine_compute(t, uv_w, i, in, up, &pending);
      diff += pending.diff;
      int16_t(*row)[kSharpRefineThreads + 2] = rowbuf[r & 1];
#pragma unroll
      for (int k = 0; k < 3; ++k) row[k][tid + 1] = static_cast<int16_t>(pending.uv_new[k]);
      if (feeds_left) {
#pragma unroll
        for (int k = 0; k < 3; ++k) dsmem_store_s16(dsmem_addr(&row[k][per + 1], rank - 1), pending.uv_new[k]);
      }
      if (feeds_right) {
#pragma unroll
        for (int k = 0; k < 3; ++k) dsmem_store_s16(dsmem_addr(&row[k][0], rank + 1), pending.uv_new[k]);
      }
    }
    row_barrier(nctas);
    // rows < r are in global memory and released (cumulatively) by the barrier just passed
    if (rank == 0 && tid == 0 && r > 0) st_release(job.progress + it, static_cast<unsigned>(r));
  }
  if (valid) {
    sharp_store_cell(w, uv_w, i, pending, y_mine + static_cast<size_t>(2 * (uv_h - 1)) * w,
                     uv_mine + static_cast<size_t>(uv_h - 1) * 3 * uv_w);
  }
  row_barrier(nctas);                  // also: nobody leaves while a neighbour may still write its halo
  if (rank == 0 && tid == 0) st_release(job.progress + it, static_cast<unsigned>(uv_h));
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, d);
  if ((tid & 31) == 0) warp_sums[tid >> 5] = diff;
  __syncthreads();
  if (tid == 0) {
    unsigned long long s = 0;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) s += warp_sums[k];
    atomicAdd(job.diff + it, s);
  }
}

__global__ void __launch_bounds__(256)
sharp_finish_kernel(const __grid_constant__ SharpJob job) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (i >= job.uv_w) return;
  unsigned long long d[kSharpIterations];
#pragma unroll
  for (int k = 0; k < kSharpIterations; ++k) d[k] = job.diff[k];
  const int keep = sharp_final_iteration(d, job.w, job.h) + 1;   // state index (0 = import)
  const uint16_t* ys = job.y_state + static_cast<size_t>(keep) * job.w * job.h;
  const int16_t* uv = job.uv_state + (static_cast<size_t>(keep) * job.uv_h + r) * 3 * job.uv_w;
  const int cr = uv[i], cg = uv[job.uv_w + i], cb = uv[2 * job.uv_w + i];
  const int out_uv_w = (job.width + 1) >> 1;
  if (i < out_uv_w) {
    job.u_out[static_cast<size_t>(r) * out_uv_w + i] = static_cast<uint8_t>(sharp_final_u(cr, cg, cb));
    job.v_out[static_cast<size_t>(r) * out_uv_w + i] = static_cast<uint8_t>(sharp_final_v(cr, cg, cb));
  }
  for (int dy = 0; dy < 2; ++dy) {
    const int y = 2 * r + dy;
    if (y >= job.height) break;
    for (int dx = 0; dx < 2; ++dx) {
      const int x = 2 * i + dx;
      if (x >= job.width) break;
      const int W = ys[static_cast<size_t>(y) * job.w + x];
      job.y_out[static_cast<size_t>(y) * job.width + x] = static_cast<uint8_t>(sharp_final_y(cr + W, cg + W, cb + W));
    }
  }
}

// width <= 4 or height <= 4: no iteration (yuv_convert.cc:675-690)
__global__ void __launch_bounds__(256)
sharp_small_kernel(const __grid_constant__ SharpJob job) {
  const int W = job.width, H = job.height, uv_w = (W + 1) >> 1, uv_h = (H + 1) >> 1;
  const long long cells = static_cast<long long>(uv_w) * uv_h;
  for (long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; c < cells;
       c += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(c % uv_w), r = static_cast<int>(c / uv_w);
    int s[3] = {0, 0, 0};
    for (int dy = 0; dy < 2; ++dy) {
      const int y = min(2 * r + dy, H - 1);
      const uint8_t* row = job.rgb + y * job.stride;
      for (int dx = 0; dx < 2; ++dx) {
        const int x = min(2 * i + dx, W - 1);
        const int rr = row[3 * x], gg = row[3 * x + 1], bb = row[3 * x + 2];
        s[0] += rr; s[1] += gg; s[2] += bb;
        if (2 * r + dy < H && 2 * i + dx < W) {
          job.y_out[static_cast<size_t>(y) * W + x] = static_cast<uint8_t>(sharp_small_y(rr, gg, bb));
        }
      }
    }
    job.u_out[static_cast<size_t>(r) * uv_w + i] = static_cast<uint8_t>(sharp_final_u(s[0], s[1], s[2]));
    job.v_out[static_cast<size_t>(r) * uv_w + i] = static_cast<uint8_t>(sharp_final_v(s[0], s[1], s[2]));
  }
}

// ---------------------------------------------------------------------------------------------
// riskiness
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
riskiness_kernel(const uint8_t* __restrict__ rgb, long long stride, int width, int height,
                 const uint8_t* __restrict__ table, unsigned long long* __restrict__ sums /* [3] */) {
  __shared__ unsigned long long red[3][8];
  unsigned long long score_sum = 0, score_num = 0, gray_num = 0;
  const int cols = width - 1, rows = height - 1;
  // one thread walks a short horizontal run so that the right-hand neighbour's index is reused
  enum { kRun = 8 };
  const int runs_x = (cols + kRun - 1) / kRun;
  const long long total = static_cast<long long>(runs_x) * rows;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(q / runs_x), x0 = static_cast<int>(q % runs_x) * kRun;
    const uint8_t* a = rgb + j * stride;
    const uint8_t* b = a + stride;
    int idx0 = risk_index(a[3 * x0], a[3 * x0 + 1], a[3 * x0 + 2]);
    const int x1 = min(x0 + kRun, cols);
    for (int x = x0; x < x1; ++x) {
      const int idx1 = risk_index(a[3 * x + 3], a[3 * x + 4], a[3 * x + 5]);
      const int idx2 = risk_index(b[3 * x], b[3 * x + 1], b[3 * x + 2]);
      const int score = __ldg(table + idx0 + kRiskLevels3 * idx1) + __ldg(table + idx0 + kRiskLevels3 * idx2) +
                        __ldg(table + idx1 + kRiskLevels3 * idx2);
      if (score > kRiskNoise) { score_sum += score; score_num += 1; }
      gray_num += (idx0 >= kRiskGrayMin && idx0 < kRiskGrayMin + 7) ? 1 : 0;
      idx0 = idx1;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    score_sum += __shfl_xor_sync(0xffffffffu, score_sum, d);
    score_num += __shfl_xor_sync(0xffffffffu, score_num, d);
    gray_num += __shfl_xor_sync(0xffffffffu, gray_num, d);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = score_sum; red[1][threadIdx.x >> 5] = score_num; red[2][threadIdx.x >> 5] = gray_num;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    unsigned long long s = 0;
    for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
    if (s) atomicAdd(&sums[threadIdx.x], s);
  }
}

}  // namespace

size_t SharpScratchBytes(int width, int height, SharpLayout* lay) {
  const int w = (width + 1) & ~1, h = (height + 1) & ~1;
  const size_t y_plane = static_cast<size_t>(w) * h * sizeof(uint16_t);
  const size_t uv_plane = static_cast<size_t>(w >> 1) * 3 * (h >> 1) * sizeof(int16_t);
  auto align = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  size_t off = 0;
  lay->y_state = off;   off += align((kSharpIterations + 1) * y_plane);
  lay->uv_state = off;  off += align((kSharpIterations + 1) * uv_plane);
  lay->target_y = off;  off += align(y_plane);
  lay->target_uv = off; off += align(uv_plane);
  lay->counters = off;  off += 256;   // diff[4] (u64) then progress[4] (u32)
  return off;
}

cudaError_t LaunchSharpYuv(const uint8_t* rgb, long long stride, int width, int height, uint8_t* scratch,
                           const uint32_t* g2l, const uint32_t* l2g, uint8_t* y_out, uint8_t* u_out, uint8_t* v_out,
                           cudaStream_t s, int* launches) {
  SharpJob job = {};
  job.rgb = rgb; job.stride = stride; job.width = width; job.height = height;
  job.w = (width + 1) & ~1; job.h = (height + 1) & ~1; job.uv_w = job.w >> 1; job.uv_h = job.h >> 1;
  job.g2l = g2l; job.l2g = l2g;
  job.y_out = y_out; job.u_out = u_out; job.v_out = v_out;
  if (width <= 4 || height <= 4) {
    sharp_small_kernel<<<1, 256, 0, s>>>(job);
    if (launches) *launches += 1;
    return cudaGetLastError();
  }
  SharpLayout lay;
  SharpScratchBytes(width, height, &lay);
  job.y_state = reinterpret_cast<uint16_t*>(scratch + lay.y_state);
  job.uv_state = reinterpret_cast<int16_t*>(scratch + lay.uv_state);
  job.target_y = reinterpret_cast<uint16_t*>(scratch + lay.target_y);
  job.target_uv = reinterpret_cast<int16_t*>(scratch + lay.target_uv);
  job.diff = reinterpret_cast<unsigned long long*>(scratch + lay.counters);
  job.progress = reinterpret_cast<unsigned*>(scratch + lay.counters + kSharpIterations * sizeof(unsigned long long));
  cudaError_t e = cudaMemsetAsync(scratch + lay.counters, 0, 256, s);
  if (e != cudaSuccess) return e;
  const dim3 cells((job.uv_w + 255) / 256, job.uv_h);
  sharp_import_kernel<<<cells, 256, 0, s>>>(job);

  // columns per cluster: enough CTAs that one row is about one cell per thread, at most 8
  int nctas = 1;
  while (nctas < 8 && job.uv_w > nctas * kSharpRefineThreads) nctas *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nctas, kSharpIterations);
  cfg.blockDim = dim3(kSharpRefineThreads);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // one cell per thread (pictures up to 8 x 512 x 2 = 8192 pixels wide): shared-memory row exchange
  const bool fast = job.uv_w <= nctas * kSharpRefineThreads && getenv("SJB_SHARP_GENERIC") == nullptr;
  e = fast ? cudaLaunchKernelEx(&cfg, sharp_refine_fast_kernel, job) : cudaLaunchKernelEx(&cfg, sharp_refine_kernel, job);
  if (e != cudaSuccess) return e;
  sharp_finish_kernel<<<cells, 256, 0, s>>>(job);
  if (launches) *launches += 3;
  return cudaGetLastError();
}

cudaError_t LaunchRiskiness(const uint8_t* rgb, long long stride, int width, int height, const uint8_t* table,
                            unsigned long long* sums, int sm_count, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(sums, 0, 3 * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  if (width > 1 && height > 1) {
    riskiness_kernel<<<sm_count * 8, 256, 0, s>>>(rgb, stride, width, height, table, sums);
  }
  return cudaGetLastError();
}

}  // namespace sjb
