// Test harness for gemma_b200/host/line_pipeline.h: token_to_double must equal atof bit for bit, and LinePipeline must
// hand blocks back in file order with every line intact (plain and gzip input).
#include <cstdio>
#include <cstring>
#include <random>
#include <string>

#include "../../gemma_b200/host/line_pipeline.h"

static bool same_bits(double a, double b) { return std::memcmp(&a, &b, sizeof(double)) == 0; }

int main(int argc, char **argv) {
  std::mt19937_64 rng(12345);
  const char *fixed[] = {"0", "1", "2", "0.5", "1.999", "-0", "+1.25", "0.000", ".5", "5.", "1e-3", "2E2", "NA", "nan", "inf", "-inf",
                         "0x10", "1.0abc", "", "-", ".", "123456789012345", "1234567890123456", "0.1234567890123456789", "00001.50",
                         "9007199254740993", "0.30000000000000004", "4.35", "0.1", "0.7", "2.675", "1e400", "-1e-400"};
  long bad = 0, n = 0;
  for (const char *t : fixed) { ++n; if (!same_bits(token_to_double(t), std::atof(t))) { std::printf("MISMATCH '%s'\n", t); ++bad; } }
  char buf[64];
  for (long k = 0; k < 3000000; ++k) {
    const int kind = (int)(rng() % 4);
    if (kind == 0) std::snprintf(buf, sizeof buf, "%.*f", (int)(rng() % 7), (double)(rng() % 2000001) / 1e6);
    else if (kind == 1) std::snprintf(buf, sizeof buf, "%.*g", 1 + (int)(rng() % 17), std::ldexp((double)(rng() >> 11), -52) * 2.0);
    else if (kind == 2) std::snprintf(buf, sizeof buf, "%s%llu.%0*llu", (rng() & 1) ? "-" : "", (unsigned long long)(rng() % 1000), 1 + (int)(rng() % 18),
                                      (unsigned long long)(rng() % 1000000000ULL));
    else std::snprintf(buf, sizeof buf, "%.*e", (int)(rng() % 12), std::ldexp((double)(rng() >> 11), -40));
    ++n;
    if (!same_bits(token_to_double(buf), std::atof(buf))) { if (bad < 10) std::printf("MISMATCH '%s'\n", buf); ++bad; }
  }
  std::printf("tokens %ld mismatches %ld\n", n, bad);
  if (argc > 1) {          // file order / integrity check: every line is "<index> <payload...>"
    long expect = 0, lines_bad = 0;
    struct Out { std::vector<long> idx; size_t first = 0; };
    LinePipeline<Out> pipe(argv[1], [](LineBlock &blk, Out &o) {
      o.first = blk.first_line;
      for (char *ln : blk.lines) { char *cur = ln; char *t = next_token(cur); o.idx.push_back(t ? std::atol(t) : -1); }
    }, 5);
    if (!pipe.ok()) { std::printf("cannot open %s\n", argv[1]); return 2; }
    Out o;
    while (pipe.next(o)) {
      if ((long)o.first != expect) ++lines_bad;
      for (long v : o.idx) { if (v != expect) ++lines_bad; ++expect; }
    }
    std::printf("lines %ld out_of_order %ld\n", expect, lines_bad);
    bad += lines_bad;
  }
  return bad ? 1 : 0;
}
