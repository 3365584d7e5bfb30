// TEST INFRASTRUCTURE: drives csrc/host_stager.cc (compiled against tests/emul/fake_cuda) through
// many uploads of awkward sizes from several owner threads (one stager each, as one context per
// thread in the product) and checks every byte.  Built with -fsanitize=thread by the test.
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>
#include "../../sjpeg_b200/csrc/host_stager.h"

static int run_owner(int seed, int uploads) {
  sjb::HostStager stager;
  uint32_t s = 12345u + seed;
  auto rnd = [&]() { s = s * 1103515245u + 12345u; return (s >> 8) & 0xffffff; };
  int bad = 0;
  for (int u = 0; u < uploads; ++u) {
    // sizes around the 128 KB piece and 2 MB chunk boundaries, and a few large ones
    size_t n;
    switch (rnd() % 5) {
      case 0: n = 1 + rnd() % 300000; break;
      case 1: n = (2u << 20) * (1 + rnd() % 5) + (rnd() % 3) - 1; break;
      case 2: n = (128u << 10) * (1 + rnd() % 40) + (rnd() % 3) - 1; break;
      case 3: n = 9000000 + rnd() % 9000000; break;
      default: n = 1 + rnd() % 6000000; break;
    }
    std::vector<uint8_t> src(n), dst(n, 0);
    for (size_t i = 0; i < n; i += 97) src[i] = static_cast<uint8_t>(rnd());
    src[n - 1] = 0x5a;
    if (stager.Upload(dst.data(), src.data(), n, nullptr) != cudaSuccess) return 1000;
    if (memcmp(src.data(), dst.data(), n) != 0) ++bad;
  }
  return bad;
}

int main() {
  int results[4] = {0, 0, 0, 0};
  std::thread owners[4];
  for (int t = 0; t < 4; ++t) owners[t] = std::thread([&, t] { results[t] = run_owner(t, 40); });
  int bad = 0;
  for (int t = 0; t < 4; ++t) { owners[t].join(); bad += results[t]; }
  printf("stager uploads with wrong bytes: %d\n", bad);
  return bad ? 1 : 0;
}
