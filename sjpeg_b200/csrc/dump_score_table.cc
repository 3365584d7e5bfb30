// Build-time helper (csrc/Makefile, target `table`): linked with the reference's score_7.cc where
// the reference tree is present, writes sjpeg::kSharpnessScore (343 x 343 bytes,
// /root/reference/src/sjpegi.h:89, jpeg_tools.cc:204-206) to the file named on the command line.
#include <stdint.h>
#include <stdio.h>

namespace sjpeg {
extern const uint8_t kSharpnessScore[];
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "wb");
  if (f == nullptr) return 1;
  const size_t n = 343 * 343;
  const bool ok = fwrite(sjpeg::kSharpnessScore, 1, n, f) == n;
  return (fclose(f) == 0 && ok) ? 0 : 1;
}
