#include "../../oracle/libm_restate.h"   // g++ -O2 -ffp-contract=off -o libm_check.bin libm_check.cc -lm (about one minute)
#include <cstdio>
#include <cstdlib>
using namespace libm_restate;
int main() {
  long bad = 0, n = 0; int shown = 0;
  // tanf over all floats in [0, 3pi/4) and negatives
  for (uint32_t u = 0; u < 0x4016cbe4u; ++u) {
    float x = wf((int32_t)u);
    float a = tanf(x), b = tanf_r(x);
    if (fw(a) != fw(b)) { ++bad; if (shown++ < 10) printf("tanf %a: libm %a restate %a\n", x, a, b); }
    ++n;
    if ((u & 0xff) == 0) { float xn = -x; a = tanf(xn); b = tanf_r(xn); if (fw(a) != fw(b)) { ++bad; if (shown++ < 10) printf("tanf %a: libm %a restate %a\n", xn, a, b); } }
  }
  printf("tanf: %ld of %ld differ\n", bad, n);
  bad = 0; n = 0; shown = 0;
  for (uint64_t u = 0; u < 0x7f800000ull; ++u) {
    float x = wf((int32_t)u);
    float a = atanf(x), b = atanf_r(x);
    if (fw(a) != fw(b)) { ++bad; if (shown++ < 10) printf("atanf %a: libm %a restate %a\n", x, a, b); }
    ++n;
  }
  printf("atanf: %ld of %ld differ\n", bad, n);
  bad = 0; n = 0; shown = 0;
  srand(1);
  for (long i = 0; i < 200000000; ++i) {
    uint32_t r1 = ((uint32_t)rand() << 16) ^ (uint32_t)rand(), r2 = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    float y = wf((int32_t)r1), x = wf((int32_t)r2);
    if (!(fabsf(x) < 3e38f) || !(fabsf(y) < 3e38f)) continue;
    if (i & 1) { y = (float)((int)(r1 % 2001) - 1000) * 0.37f; x = (float)((int)(r2 % 2001) - 1000) * 0.11f; }
    float a = atan2f(y, x), b = atan2f_r(y, x);
    if (fw(a) != fw(b)) { ++bad; if (shown++ < 10) printf("atan2f %a %a: libm %a restate %a\n", y, x, a, b); }
    ++n;
  }
  printf("atan2f: %ld of %ld differ\n", bad, n);
  return 0;
}
