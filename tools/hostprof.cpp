// tools/hostprof.cpp — DEVELOPMENT ONLY: times the host phases of ppd_block_decode (parse, shape,
// dump) on a machine without a GPU.  Linked against a -DPPD_HOSTPROF build of csrc/ppd_host.cu in
// which node refs are not computed (see csrc/hostprof_stub.h); its output is only good for diffing
// two host implementations against each other and for timing.  Not part of libppd_b200.so.
//   build/hostprof <flat block file> [repeats] [output file]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/ppd_b200.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> flat(n);
  if (fread(flat.data(), 1, n, f) != (size_t)n) return 2;
  fclose(f);
  int reps = argc > 2 ? atoi(argv[2]) : 3;
  ppd_ctx* c = nullptr;
  if (ppd_ctx_create(0, &c) != 0) return 3;
  for (int r = 0; r < reps; r++) {
    uint8_t* out = nullptr;
    size_t out_len = 0;
    auto t0 = std::chrono::steady_clock::now();
    int rc = ppd_block_decode(c, flat.data(), flat.size(), &out, &out_len);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "decode rc=%d out=%zu bytes  %.2f ms\n", rc, out_len, ms);
    if (rc != 0) fprintf(stderr, "  error: %s\n", ppd_last_error(c));
    if (r == reps - 1 && argc > 3 && out) {
      FILE* o = fopen(argv[3], "wb");
      fwrite(out, 1, out_len, o);
      fclose(o);
    }
    ppd_free(out);
  }
  ppd_ctx_destroy(c);
  return 0;
}
