// engine_stripes.inl -- row stripes of pictures split across the GPUs of one box, exchanged with
// NCCL from inside the library (included at the end of engine.cu; SURVEY.md 8e, BASELINE.json
// config 5).  One process per GPU; every rank holds the same MCU-row range [y0, y1) of each of the
// n pictures of a batch and calls sjb_stripes_encode() collectively.
//
// The reference writes no restart markers (/root/reference/src/headers.cc:242-258) and predicts DC
// across the whole picture (enc.cc:277, entropy.cc:133-136), so the stripes of a picture are ONE
// bit string.  Everything between the collectives stays on the device and stream-ordered; per
// batch (not per picture) the ranks exchange
//   [adaptive]  ncclAllReduce  coefficient histograms  int32[n][2][64][129]   (histogram.cc:317-339)
//               ncclAllGather  last quantised DC of each component            int[n][3]
//   [optimise]  ncclAllReduce  symbol counts           uint32[n][2][272]      (enc.cc:344-374)
//               ncclAllGather  bit counts of the stripes                      u64[n]
//               ncclAllGather  size / shared-byte record of the stuffed stripes  u64[n][2]
//               ncclSend/Recv  (grouped) the compressed stripes, exact sizes, to rank 0
// and the host waits once per phase that needs host arithmetic (matrices, Huffman tables, sizes).
// Every rank derives the same matrices and tables from the same reduced counters, deterministically.
// Rank 0 prepends the header, concatenates the stripes and OR-merges the byte two (or more)
// neighbours share (sjb_stripes_assemble, host only, also exported for the CPU tests).

namespace {

void StripeRows(int height, int yuv_mode, int world, int rank, int* y0, int* y1) {
  const int mcu = (yuv_mode == SJB_YUV_420) ? 16 : 8;
  const int rows = (height + mcu - 1) / mcu;
  const int base = rows / world, extra = rows % world;
  const int a = rank * base + std::min(rank, extra);
  const int b = a + base + (rank < extra ? 1 : 0);
  *y0 = std::min(a * mcu, height);
  *y1 = std::min(b * mcu, height);
}

}  // namespace

struct sjb_comm {
  sjb_context* ctx = nullptr;
  ncclComm_t nccl = nullptr;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;        // every kernel and every collective, in the same order on all ranks
  cudaStream_t upload = nullptr;        // every host-to-device copy of the pixels, in order (see phase1)
  std::vector<Lane*> sets;              // one buffer set per group of stripes (their own streams stay idle)
  DeviceBuffer dc_local, dc_all, bits_local, bits_all, offsets, meta_local, meta_all, send, recv, zeros;
  unsigned long long* h_meta = nullptr; // pinned [world][n][2]
  size_t h_meta_cap = 0;
  uint8_t* h_recv = nullptr;            // pinned, rank 0
  size_t h_recv_cap = 0;
  cudaEvent_t joined = nullptr;
};

#define NC(expr)                                                                         \
  do {                                                                                   \
    const ncclResult_t r_ = (expr);                                                      \
    if (r_ != ncclSuccess) {                                                             \
      ctx->err = std::string(#expr) + ": " + api->GetErrorString(r_);                    \
      return SJB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

extern "C" {

int sjb_comm_unique_id(uint8_t id[128]) try {
  const NcclApi* api = Nccl();
  if (api == nullptr || id == nullptr) return SJB_ERR_CUDA;
  ncclUniqueId u;
  if (api->GetUniqueId(&u) != ncclSuccess) return SJB_ERR_CUDA;
  static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id, &u, 128);
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_comm_create(sjb_context* ctx, const uint8_t id[128], int rank, int world, sjb_comm** out) try {
  if (ctx == nullptr || id == nullptr || out == nullptr || world < 1 || world > 64 || rank < 0 || rank >= world) return SJB_ERR_ARG;
  *out = nullptr;
  ctx->err.clear();
  const NcclApi* api = Nccl();
  if (api == nullptr) {
    ctx->err = "libnccl.so.2 could not be loaded";
    return SJB_ERR_CUDA;
  }
  CU(cudaSetDevice(ctx->device));
  std::unique_ptr<sjb_comm> c(new (std::nothrow) sjb_comm());
  if (!c) return SJB_ERR_NOMEM;
  c->ctx = ctx;
  c->rank = rank;
  c->world = world;
  ncclUniqueId u;
  memcpy(&u, id, 128);
  NC(api->CommInitRank(&c->nccl, world, u, rank));
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->upload, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->joined, cudaEventDisableTiming));
  *out = c.release();
  return SJB_OK;
} SJB_NOTHROW_END

void sjb_comm_destroy(sjb_comm* c) {
  if (c == nullptr) return;
  cudaSetDevice(c->ctx->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (Lane* L : c->sets) {
    DestroyLane(L);
    delete L;
  }
  for (DeviceBuffer* b : {&c->dc_local, &c->dc_all, &c->bits_local, &c->bits_all, &c->offsets, &c->meta_local, &c->meta_all,
                          &c->send, &c->recv, &c->zeros}) {
    b->Release();
  }
  if (c->h_meta) cudaFreeHost(c->h_meta);
  if (c->h_recv) cudaFreeHost(c->h_recv);
  if (c->joined) cudaEventDestroy(c->joined);
  const NcclApi* api = Nccl();
  if (api != nullptr && c->nccl != nullptr) api->CommDestroy(c->nccl);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->upload) cudaStreamDestroy(c->upload);
  delete c;
}

int sjb_stripe_rows(int height, int yuv_mode, int world, int rank, int* y0, int* y1) {
  if (height <= 0 || world < 1 || rank < 0 || rank >= world || y0 == nullptr || y1 == nullptr) return SJB_ERR_ARG;
  if (yuv_mode != SJB_YUV_420 && yuv_mode != SJB_YUV_444 && yuv_mode != SJB_YUV_400) return SJB_ERR_ARG;
  StripeRows(height, yuv_mode, world, rank, y0, y1);
  return SJB_OK;
}

// Host only.  part[r] = bytes stripe r emitted (size[r]); flags[r] = head_byte | tail_byte << 8 |
// tail_bits << 16 | head_open << 24 as the stuffing kernel reports them (kernels.cuh StreamInfo):
// head_byte = the stripe's share of the byte it begins in (when that byte is shared with the
// previous stripe), tail_byte / tail_bits = its share of the byte it ends in, head_open = the stripe
// ends INSIDE its head byte (it owns no byte boundary at all).
int sjb_stripes_assemble(const uint8_t* header, size_t header_len, int stripes, const uint8_t* const* part,
                         const size_t* size, const unsigned* flags, uint8_t* out, size_t out_capacity,
                         size_t* out_size) try {
  if (header == nullptr || stripes < 0 || part == nullptr || size == nullptr || flags == nullptr || out_size == nullptr)
    return SJB_ERR_ARG;
  size_t need = header_len;
  for (int r = 0; r < stripes; ++r) need += size[r] + 2;
  *out_size = 0;
  if (out == nullptr || need > out_capacity) {
    *out_size = need;                  // upper bound
    return SJB_ERR_CAPACITY;
  }
  size_t pos = 0;
  memcpy(out, header, header_len);
  pos += header_len;
  unsigned carry = 0, carry_bits = 0;
  for (int r = 0; r < stripes; ++r) {
    const unsigned head = flags[r] & 0xff, tail = (flags[r] >> 8) & 0xff, tbits = (flags[r] >> 16) & 0xff;
    const bool head_open = ((flags[r] >> 24) & 1) != 0;
    if (carry_bits != 0) {
      const unsigned merged = carry | head;
      if (head_open) {                 // still the same byte: keep collecting
        carry = merged;
        carry_bits = tbits;
        continue;
      }
      out[pos++] = static_cast<uint8_t>(merged);
      if (merged == 0xff) out[pos++] = 0x00;          // bit_writer.h:172-196
    }
    if (size[r] != 0) {
      if (part[r] == nullptr) return SJB_ERR_ARG;
      memcpy(out + pos, part[r], size[r]);
      pos += size[r];
    }
    carry = tail;
    carry_bits = tbits;
  }
  *out_size = pos;
  return SJB_OK;
} SJB_NOTHROW_END

int sjb_stripes_encode(sjb_comm* comm, int n, const uint8_t* const* pix, int pix_on_device, int width, int height,
                       long long stride, const sjb_params* params, uint8_t* const* out, size_t out_capacity,
                       size_t* sizes) try {
  if (comm == nullptr || n <= 0 || pix == nullptr || params == nullptr) return SJB_ERR_ARG;
  sjb_context* ctx = comm->ctx;
  ctx->err.clear();
  const NcclApi* api = Nccl();
  if (api == nullptr) return SJB_ERR_CUDA;
  const int rank = comm->rank, world = comm->world;
  if (rank == 0 && (out == nullptr || sizes == nullptr)) return SJB_ERR_ARG;
  const int pstep = (params->pix_fmt != SJB_PIX_RGB) ? 4 : 3;
  Plan full;                                               // whole picture: header geometry, argument checks
  RC(MakePlan(width, height, stride, params, &full));
  int y0, y1;
  StripeRows(height, full.g.yuv_mode, world, rank, &y0, &y1);
  const int hs = y1 - y0;
  const bool active = hs > 0;
  int first_holder = 0, last_holder = 0, prev_holder = -1;
  for (int r = 0; r < world; ++r) {
    int a, b;
    StripeRows(height, full.g.yuv_mode, world, r, &a, &b);
    if (b > a) {
      last_holder = r;
      if (r < rank) prev_holder = r;
    }
  }
  const bool is_first = active && rank == first_holder, is_last = active && rank == last_holder;
  Plan plan = full;                                        // this rank's stripe
  if (active) RC(MakePlan(width, hs, stride, params, &plan));
  (void)pstep;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = comm->stream;
  const int groups = (n + kMaxGroup - 1) / kMaxGroup;
  while (static_cast<int>(comm->sets.size()) < groups) {
    Lane* L = new (std::nothrow) Lane();
    if (L == nullptr) return SJB_ERR_NOMEM;
    comm->sets.push_back(L);
    RC(InitLane(ctx, L));
  }
  const size_t un = static_cast<size_t>(n);
  CU(comm->dc_local.Reserve(un * 3 * sizeof(int)));
  CU(comm->dc_all.Reserve(un * 3 * sizeof(int) * world));
  CU(comm->bits_local.Reserve(un * 8));
  CU(comm->bits_all.Reserve(un * 8 * world));
  CU(comm->offsets.Reserve(un * 8));
  CU(comm->meta_local.Reserve(un * 16));
  CU(comm->meta_all.Reserve(un * 16 * world));
  if (comm->h_meta_cap < un * 16 * world) {
    if (comm->h_meta) cudaFreeHost(comm->h_meta);
    comm->h_meta = nullptr;
    CU(cudaMallocHost(reinterpret_cast<void**>(&comm->h_meta), un * 16 * world));
    comm->h_meta_cap = un * 16 * world;
  }

  uint8_t quant0[2][64], min_quant[2][64];
  QuantTabs qt;
  if (!MakeQuantTabs(plan, quant0, min_quant, &qt)) return SJB_ERR_ARG;
  HuffSpec def_spec[4];
  CodeTabs def_tabs;
  memset(&def_tabs, 0, sizeof(def_tabs));
  for (int i = 0; i < 4; ++i) DefaultHuffSpec(i >= 2, i & 1, &def_spec[i]);
  for (int c = 0; c < 2; ++c) {
    CodesFromSpec(def_spec[c], def_tabs.dc[c]);
    CodesFromSpec(def_spec[2 + c], def_tabs.ac[c]);
  }
  std::vector<uint8_t> quant(un * 128);
  for (int i = 0; i < n; ++i) memcpy(&quant[i * 128], quant0, 128);
  std::vector<HuffSpec> spec(un * 4);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) spec[i * 4 + k] = def_spec[k];
  std::vector<CodeTabs> tabs(un, def_tabs);
  std::vector<FrameSet> fsets(groups);

  // The batch is cut into chunks of kChunkGroups groups.  A chunk's uploads and first kernels run on
  // its sets' own streams (phase 1); its collectives, host phases and the gather to rank 0 run on
  // the communicator's stream (the rest).  Phase 1 of chunk c+1 is enqueued BEFORE the rest of
  // chunk c, so the host-to-device copies of the next chunk -- what bounds a batch that comes from
  // host memory -- proceed while this chunk's exchange waits for its small messages.
  // One group (16 pictures) per chunk: what follows the LAST chunk's upload is exposed, the exchanges of
  // the earlier ones hide under the next upload (two groups per chunk: 64 x 1080p on 4 GPUs 3.14 ms, of
  // which 0.45 ms behind the last copy).
  enum { kChunkGroups = 1 };
  const int chunks = (groups + kChunkGroups - 1) / kChunkGroups;

  // ---- phase 1, every set on its own stream: upload, F1 (+ H1) ---------------------------------
  // Phase 1 enqueues COPIES only, all on one stream.  (Copies issued on several streams share the
  // link concurrently, so every chunk's pixels would arrive at the end of the whole batch's upload;
  // in order, chunk c is complete while c+1 is still being copied.  And no kernel of chunk c+1 is
  // enqueued before the exchange of chunk c: a kernel waiting for pixels that are still in flight
  // held up the collectives queued after it -- measured: every chunk's sizes arrived only when the
  // next chunk's upload had finished.)
  auto phase1 = [&](int c) -> int {
    for (int k = c * kChunkGroups; k < std::min(groups, (c + 1) * kChunkGroups); ++k) {
      Lane* L = comm->sets[k];
      const int frames = std::min<int>(kMaxGroup, n - k * kMaxGroup);
      RC(ReserveLane(ctx, L, plan, frames));
      FrameSet& fs = fsets[k];
      FillFrameSet(plan, stride, &fs);
      fs.frames = frames;
      CU(cudaStreamWaitEvent(comm->upload, comm->joined, 0));   // previous batch's last use of these buffers
      if (active) {
        if (!pix_on_device) RC(ReservePix(ctx, L, plan, stride, frames));
        for (int f = 0; f < frames; ++f) {
          const uint8_t* p = pix[k * kMaxGroup + f];
          if (p == nullptr) return SJB_ERR_ARG;
          fs.pix[f] = p;
          if (!pix_on_device) {
            long long ds = stride;
            RC(UploadPicture(ctx, L, p, plan, stride, f, &fs.pix[f], &ds, comm->upload));
            fs.stride = ds;
          }
        }
      }
      CU(cudaEventRecord(L->ev[2], comm->upload));
      L->header_valid = 0;
      L->tabs_valid = 0;
    }
    return SJB_OK;
  };

  // ---- the rest of a chunk: on the communicator's stream, in the same order on every rank -------
  static const bool trace = getenv("SJB_STRIPES_TRACE") != nullptr;
  static const bool trace_sync = trace && atoi(getenv("SJB_STRIPES_TRACE")) >= 2;
  auto now_us = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
  auto rest = [&](int c) -> int {
    double t_mark = now_us();
    auto mark = [&](const char* what) {
      if (!trace) return;
      if (trace_sync) cudaStreamSynchronize(st);      // SJB_STRIPES_TRACE=2: device time of each step
      const double t = now_us();
      fprintf(stderr, "[stripes r%d c%d] %-22s %8.1f us\n", rank, c, what, t - t_mark);
      t_mark = t;
    };
    const int g0 = c * kChunkGroups, g1 = std::min(groups, (c + 1) * kChunkGroups);
    const int i0 = g0 * kMaxGroup, i1 = std::min(n, g1 * kMaxGroup);
    const int nc = i1 - i0;                                  // pictures of this chunk
    const size_t unc = static_cast<size_t>(nc);
    // colour conversion + fDCT (+ quantise | histogram) once the chunk's pixels have landed
    for (int k = g0; k < g1; ++k) {
      Lane* L = comm->sets[k];
      CU(cudaStreamWaitEvent(st, L->ev[2], 0));
      if (L->words_dirty) {
        CU(cudaMemsetAsync(L->words.ptr, 0, L->words.bytes, st));
        L->words_dirty = false;
      }
      if (active) {
        Lane view;                       // LaunchF1 reads gb / stream / launches only
        view.gb = L->gb;
        view.stream = st;
        LaunchF1(&view, fsets[k], plan.g, /*raw=*/plan.adaptive, qt);
        view.stream = nullptr;
      }
      if (plan.adaptive) {
        CU(cudaMemsetAsync(L->d_small()->hist, 0, fsets[k].frames * sizeof(L->d_small()->hist[0]), st));
        if (active) LaunchHistogram(fsets[k], L->gb, st);
      }
    }
    CU(cudaGetLastError());
    mark("F1");

    // adaptive quantisation: all-reduce the histograms, every rank derives the same matrices
    if (plan.adaptive) {
      NC(api->GroupStart());
      for (int k = g0; k < g1; ++k) {
        SmallLayout* D = comm->sets[k]->d_small();
        NC(api->AllReduce(D->hist, D->hist, static_cast<size_t>(fsets[k].frames) * 2 * 64 * kHistoStride, ncclInt32, ncclSum,
                          comm->nccl, st));
      }
      NC(api->GroupEnd());
      // every rank derives the same matrices from the same reduced counters, on the device (A1); the
      // copies for the DQT segments are picked up behind a later wait (rank 0 only needs them)
      for (int k = g0; k < g1; ++k) {
        RC(EnqueueAdaptiveQuantise(ctx, comm->sets[k], fsets[k], plan, quant0, min_quant, def_tabs, active, st));
      }
    }

    // DC predictors: last quantised DC of every component, handed to the next stripe
    int* d_dc_local = comm->dc_local.as<int>();
    if (active) {
      for (int k = g0; k < g1; ++k) LaunchLastDc(fsets[k], comm->sets[k]->gb, d_dc_local + (k - g0) * kMaxGroup * 3, st);
    } else {
      CU(cudaMemsetAsync(d_dc_local, 0, unc * 3 * sizeof(int), st));
    }
    NC(api->AllGather(d_dc_local, comm->dc_all.ptr, unc * 3, ncclInt32, comm->nccl, st));
    mark("enqueue to AG(dc)");
    auto dc_init_of = [&](int k) -> const int* {
      return (prev_holder < 0) ? nullptr
                               : comm->dc_all.as<int>() + (static_cast<size_t>(prev_holder) * nc + (k - g0) * kMaxGroup) * 3;
    };

    // optimised Huffman tables: all-reduce the symbol counts
    if (plan.optimize) {
      for (int k = g0; k < g1; ++k) {
        Lane* L = comm->sets[k];
        SmallLayout* D = L->d_small();
        CU(cudaMemsetAsync(D->freq, 0, fsets[k].frames * sizeof(D->freq[0]), st));
        if (!active) continue;
        GroupBuffers gb = L->gb;
        gb.dc_init = dc_init_of(k);
        LaunchSymbolStats(fsets[k], gb, st);
      }
      CU(cudaGetLastError());
      NC(api->GroupStart());
      for (int k = g0; k < g1; ++k) {
        SmallLayout* D = comm->sets[k]->d_small();
        NC(api->AllReduce(D->freq, D->freq, static_cast<size_t>(fsets[k].frames) * 2 * 272, ncclUint32, ncclSum, comm->nccl, st));
      }
      NC(api->GroupEnd());
      for (int k = g0; k < g1; ++k) {
        Lane* L = comm->sets[k];
        CU(cudaMemcpyAsync(L->host->freq, L->d_small()->freq, fsets[k].frames * sizeof(L->host->freq[0]),
                           cudaMemcpyDeviceToHost, st));
      }
      CU(cudaStreamSynchronize(st));
      const int nb_tables = (full.g.nb_comps == 1) ? 1 : 2;
      ctx->pool.ParallelFor(nc, [&](int j) {
        const int i = i0 + j;
        const uint32_t* freq = comm->sets[i / kMaxGroup]->host->freq[i % kMaxGroup];
        for (int cc = 0; cc < nb_tables; ++cc) {
          OptimalHuffSpec(freq + 272 * cc + 256, 12, &spec[i * 4 + cc]);
          OptimalHuffSpec(freq + 272 * cc, 256, &spec[i * 4 + 2 + cc]);
          CodesFromSpec(spec[i * 4 + cc], tabs[i].dc[cc]);
          CodesFromSpec(spec[i * 4 + 2 + cc], tabs[i].ac[cc]);
        }
      });
    }

    // entropy coding, bit counts, global bit offsets, byte stuffing
    unsigned long long* d_bits = comm->bits_local.as<unsigned long long>();
    CU(cudaMemsetAsync(d_bits, 0, unc * 8, st));
    for (int k = g0; k < g1 && active; ++k) {
      Lane* L = comm->sets[k];
      const int frames = fsets[k].frames;
      for (int f = 0; f < frames; ++f) L->host->tabs[f] = tabs[k * kMaxGroup + f];
      LaunchCopySmall(L->d_small()->tabs, L->host->tabs, frames * sizeof(CodeTabs), st);
      CU(cudaMemsetAsync(L->state.ptr, 0, L->state.bytes, st));
      GroupBuffers gb = L->gb;
      gb.dc_init = dc_init_of(k);
      L->words_dirty = true;
      LaunchEntropyPack(fsets[k], gb, st);
      LaunchStripeBits(gb, frames, d_bits + (k - g0) * kMaxGroup, st);
    }
    CU(cudaGetLastError());
    mark("E");
    NC(api->AllGather(d_bits, comm->bits_all.ptr, unc, ncclUint64, comm->nccl, st));
    mark("AG(bits)");
    unsigned long long* d_off = comm->offsets.as<unsigned long long>();
    LaunchStripeOffsets(comm->bits_all.as<unsigned long long>(), nc, rank, d_off, st);
    unsigned long long* d_meta = comm->meta_local.as<unsigned long long>();
    CU(cudaMemsetAsync(d_meta, 0, unc * 16, st));
    for (int k = g0; k < g1 && active; ++k) {
      Lane* L = comm->sets[k];
      StuffArgs sa;
      memset(&sa, 0, sizeof(sa));
      for (int f = 0; f < fsets[k].frames; ++f) sa.flags[f] = (is_first ? kStuffFirst : 0) | (is_last ? kStuffLast : 0) | kStuffKeepWords;
      GroupBuffers gb = L->gb;
      gb.bit_offsets = d_off + (k - g0) * kMaxGroup;
      LaunchStuff(fsets[k], gb, sa, st);
      LaunchStripeMeta(gb, fsets[k].frames, d_meta + static_cast<size_t>(k - g0) * kMaxGroup * 2, st);
    }
    CU(cudaGetLastError());
    NC(api->AllGather(d_meta, comm->meta_all.ptr, unc * 2, ncclUint64, comm->nccl, st));
    LaunchCopySmall(comm->h_meta, comm->meta_all.ptr, unc * 16 * world, st);
    mark("enqueue to AG(meta)");
    CU(cudaStreamSynchronize(st));
    mark("sync after meta");
    if (plan.adaptive) {
      for (int j = 0; j < nc; ++j) {
        const int i = i0 + j;
        Lane* L = comm->sets[i / kMaxGroup];
        if (L->host->aq_fail[i % kMaxGroup]) return SJB_ERR_ARG;
        memcpy(&quant[static_cast<size_t>(i) * 128], L->host->quant[i % kMaxGroup], 128);
      }
    }

    // gather the compressed stripes on rank 0: exact sizes, one grouped send/recv
    std::vector<size_t> rank_bytes(world, 0);
    for (int r = 0; r < world; ++r) {
      for (int j = 0; j < nc; ++j) rank_bytes[r] += static_cast<size_t>(comm->h_meta[(static_cast<size_t>(r) * nc + j) * 2]);
    }
    size_t total = 0;
    for (int r = 0; r < world; ++r) total += rank_bytes[r];
    if (active && rank_bytes[rank] > 0) {
      CU(comm->send.Reserve(rank_bytes[rank]));
      for (int k = g0; k < g1; ++k) {
        Lane* L = comm->sets[k];
        LaunchStripeCompact(L->gb.out, L->gb.out_pitch, (k - g0) * kMaxGroup, fsets[k].frames, d_meta, comm->send.as<uint8_t>(), st);
      }
      CU(cudaGetLastError());
    }
    if (rank == 0) {
      CU(comm->recv.Reserve(std::max<size_t>(total, 1)));
      if (comm->h_recv_cap < total) {
        if (comm->h_recv) cudaFreeHost(comm->h_recv);
        comm->h_recv = nullptr;
        comm->h_recv_cap = 0;
        CU(cudaMallocHost(reinterpret_cast<void**>(&comm->h_recv), total + (total >> 2) + 4096));
        comm->h_recv_cap = total + (total >> 2) + 4096;
      }
    }
    std::vector<size_t> rank_base(world, 0);
    for (int r = 1; r < world; ++r) rank_base[r] = rank_base[r - 1] + rank_bytes[r - 1];
    if (world > 1) {
      NC(api->GroupStart());
      if (rank != 0 && rank_bytes[rank] > 0) NC(api->Send(comm->send.ptr, rank_bytes[rank], ncclUint8, 0, comm->nccl, st));
      if (rank == 0) {
        for (int r = 1; r < world; ++r) {
          if (rank_bytes[r] > 0) NC(api->Recv(comm->recv.as<uint8_t>() + rank_base[r], rank_bytes[r], ncclUint8, r, comm->nccl, st));
        }
      }
      NC(api->GroupEnd());
    }
    for (int k = g0; k < g1; ++k) comm->sets[k]->words_dirty = true;   // shifted reads cannot self-clean
    mark("compact + send/recv enq");
    if (rank != 0) {
      CU(cudaStreamSynchronize(st));       // the send buffer is reused by the next chunk
      return SJB_OK;
    }
    if (rank_bytes[0] > 0) CU(cudaMemcpyAsync(comm->h_recv, comm->send.ptr, rank_bytes[0], cudaMemcpyDeviceToHost, st));
    if (total > rank_bytes[0]) {
      CU(cudaMemcpyAsync(comm->h_recv + rank_bytes[0], comm->recv.as<uint8_t>() + rank_bytes[0], total - rank_bytes[0],
                         cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    mark("sync after gather");

    // rank 0: header + stripes, boundary bytes merged; the pictures are independent, so they are put
    // together on the context's worker threads (one thread took 125-190 us per 32 pictures, all of it
    // exposed behind the last chunk)
    int rc = SJB_OK;
    std::vector<size_t> start(static_cast<size_t>(world) * nc);      // where stripe (r, j) begins in h_recv
    for (int r = 0; r < world; ++r) {
      size_t at = rank_base[r];
      for (int j = 0; j < nc; ++j) {
        start[static_cast<size_t>(r) * nc + j] = at;
        at += static_cast<size_t>(comm->h_meta[(static_cast<size_t>(r) * nc + j) * 2]);
      }
    }
    std::vector<int> arcs(nc, SJB_OK);
    ctx->pool.ParallelFor(nc, [&](int j) {
      const int i = i0 + j;
      std::vector<uint8_t> header;
      header.reserve(1024);
      AppendHeaders(full.g, reinterpret_cast<const uint8_t(*)[64]>(&quant[static_cast<size_t>(i) * 128]), &spec[i * 4], &header);
      const uint8_t* part[64];
      size_t psize[64];
      unsigned pflags[64];
      int holders = 0;
      for (int r = 0; r <= last_holder && holders < 64; ++r) {
        const unsigned long long* m = comm->h_meta + (static_cast<size_t>(r) * nc + j) * 2;
        part[holders] = comm->h_recv + start[static_cast<size_t>(r) * nc + j];
        psize[holders] = static_cast<size_t>(m[0]);
        pflags[holders] = static_cast<unsigned>(m[1]);
        ++holders;
      }
      size_t size = 0;
      arcs[j] = sjb_stripes_assemble(header.data(), header.size(), holders, part, psize, pflags, out[i], out[i] ? out_capacity : 0,
                                     &size);
      sizes[i] = size;
    });
    for (int j = 0; j < nc; ++j) if (arcs[j] != SJB_OK) rc = arcs[j];
    mark("assemble");
    return rc;
  };

  // How far the uploads run ahead of the exchange: pinned (or device) pixels are copied
  // asynchronously, so every chunk's copies are queued at once and the link never idles; pageable
  // pixels go through the context's staging threads, which block the caller, one chunk ahead.
  int ahead = 1;
  if (pix_on_device) {
    ahead = chunks;
  } else if (active) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, pix[0]) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered) ahead = chunks;
    cudaGetLastError();
  }
  int result = SJB_OK;
  const double t_begin = now_us();
  int queued = 0;
  for (; queued < std::min(chunks, ahead); ++queued) RC(phase1(queued));
  for (int c = 0; c < chunks; ++c) {
    if (queued < chunks && queued <= c + ahead) RC(phase1(queued++));
    if (trace) fprintf(stderr, "[stripes r%d c%d] phase1 enqueued at     %8.1f us since entry\n", rank, c, now_us() - t_begin);
    const int rc = rest(c);
    if (rc != SJB_OK && rc != SJB_ERR_CAPACITY) return rc;
    if (rc != SJB_OK) result = rc;
  }
  CU(cudaEventRecord(comm->joined, st));
  return result;
} SJB_NOTHROW_END

// Frames sharded across the ranks (each rank encodes its own pictures; SURVEY.md 8e "frames"): the
// finished JPEGs, left in DEVICE memory by sjb_encode_batch(out_on_device = 1), are collected on
// rank 0 -- an all-gather of the counts, one of the sizes, then one grouped ncclSend/ncclRecv of the
// exact bytes.  Rank 0 gets them in rank order, back to back in `blob` (host memory), their sizes in
// out_sizes[0 .. *n_total).
int sjb_gather_frames(sjb_comm* comm, int n_local, const uint8_t* const* dev_jpegs, const size_t* sizes_local,
                      uint8_t* blob, size_t blob_capacity, size_t* out_sizes, int out_sizes_capacity, int* n_total) try {
  if (comm == nullptr || n_local < 0 || (n_local > 0 && (dev_jpegs == nullptr || sizes_local == nullptr))) return SJB_ERR_ARG;
  sjb_context* ctx = comm->ctx;
  ctx->err.clear();
  const NcclApi* api = Nccl();
  if (api == nullptr) return SJB_ERR_CUDA;
  const int rank = comm->rank, world = comm->world;
  if (rank == 0 && (out_sizes == nullptr || n_total == nullptr)) return SJB_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = comm->stream;
  // counts
  CU(comm->bits_local.Reserve(8));
  CU(comm->bits_all.Reserve(8 * static_cast<size_t>(world)));
  if (comm->h_meta_cap < 8 * static_cast<size_t>(world)) {
    if (comm->h_meta) cudaFreeHost(comm->h_meta);
    comm->h_meta = nullptr;
    comm->h_meta_cap = 0;
    CU(cudaMallocHost(reinterpret_cast<void**>(&comm->h_meta), 4096));
    comm->h_meta_cap = 4096;
  }
  comm->h_meta[0] = static_cast<unsigned long long>(n_local);
  LaunchCopySmall(comm->bits_local.ptr, comm->h_meta, 8, st);
  NC(api->AllGather(comm->bits_local.ptr, comm->bits_all.ptr, 1, ncclUint64, comm->nccl, st));
  std::vector<unsigned long long> counts(world);
  CU(cudaMemcpyAsync(counts.data(), comm->bits_all.ptr, 8 * static_cast<size_t>(world), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  size_t maxn = 1, total_n = 0;
  for (int r = 0; r < world; ++r) {
    maxn = std::max<size_t>(maxn, counts[r]);
    total_n += counts[r];
  }
  // sizes, padded to the largest count
  CU(comm->meta_local.Reserve(8 * maxn));
  CU(comm->meta_all.Reserve(8 * maxn * world));
  std::vector<unsigned long long> mine(maxn, 0), all(maxn * world);
  size_t my_bytes = 0;
  for (int i = 0; i < n_local; ++i) {
    mine[i] = sizes_local[i];
    my_bytes += sizes_local[i];
  }
  CU(cudaMemcpyAsync(comm->meta_local.ptr, mine.data(), 8 * maxn, cudaMemcpyHostToDevice, st));
  NC(api->AllGather(comm->meta_local.ptr, comm->meta_all.ptr, maxn, ncclUint64, comm->nccl, st));
  CU(cudaMemcpyAsync(all.data(), comm->meta_all.ptr, 8 * maxn * world, cudaMemcpyDeviceToHost, st));
  // this rank's JPEGs back to back
  CU(comm->send.Reserve(std::max<size_t>(my_bytes, 1)));
  size_t off = 0;
  for (int i = 0; i < n_local; ++i) {
    if (dev_jpegs[i] == nullptr) return SJB_ERR_ARG;
    CU(cudaMemcpyAsync(comm->send.as<uint8_t>() + off, dev_jpegs[i], sizes_local[i], cudaMemcpyDeviceToDevice, st));
    off += sizes_local[i];
  }
  CU(cudaStreamSynchronize(st));
  std::vector<size_t> rank_bytes(world, 0), rank_base(world, 0);
  for (int r = 0; r < world; ++r) for (size_t i = 0; i < counts[r]; ++i) rank_bytes[r] += static_cast<size_t>(all[r * maxn + i]);
  for (int r = 1; r < world; ++r) rank_base[r] = rank_base[r - 1] + rank_bytes[r - 1];
  const size_t total = rank_base[world - 1] + rank_bytes[world - 1];
  if (rank == 0) CU(comm->recv.Reserve(std::max<size_t>(total, 1)));
  if (world > 1) {
    NC(api->GroupStart());
    if (rank != 0 && rank_bytes[rank] > 0) NC(api->Send(comm->send.ptr, rank_bytes[rank], ncclUint8, 0, comm->nccl, st));
    if (rank == 0) {
      for (int r = 1; r < world; ++r) {
        if (rank_bytes[r] > 0) NC(api->Recv(comm->recv.as<uint8_t>() + rank_base[r], rank_bytes[r], ncclUint8, r, comm->nccl, st));
      }
    }
    NC(api->GroupEnd());
  }
  if (rank != 0) {
    CU(cudaStreamSynchronize(st));
    return SJB_OK;
  }
  *n_total = static_cast<int>(total_n);
  int rc = SJB_OK;
  if (static_cast<size_t>(out_sizes_capacity) < total_n || blob == nullptr || blob_capacity < total) rc = SJB_ERR_CAPACITY;
  if (rc == SJB_OK) {
    size_t k = 0;
    for (int r = 0; r < world; ++r) for (size_t i = 0; i < counts[r]; ++i) out_sizes[k++] = static_cast<size_t>(all[r * maxn + i]);
    if (rank_bytes[0] > 0) CU(cudaMemcpyAsync(blob, comm->send.ptr, rank_bytes[0], cudaMemcpyDeviceToHost, st));
    if (total > rank_bytes[0]) {
      CU(cudaMemcpyAsync(blob + rank_bytes[0], comm->recv.as<uint8_t>() + rank_bytes[0], total - rank_bytes[0],
                         cudaMemcpyDeviceToHost, st));
    }
  }
  CU(cudaStreamSynchronize(st));
  return rc;
} SJB_NOTHROW_END

}  // extern "C"
#undef NC
