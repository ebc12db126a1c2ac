/* Harness for tests/test_host_cli.py: compares the product's fast "%.12E"/"%d" formatters
 * (mpilattice-boltzmann_b200/host/fast_format.h) with glibc printf on many floats.
 * usage: fmt_check <count> <seed>; prints the number of mismatches. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../mpilattice-boltzmann_b200/host/fast_format.h"

static uint64_t state;
static uint32_t next_u32(void)
{
  state ^= state << 13; state ^= state >> 7; state ^= state << 17;
  return (uint32_t)(state >> 16);
}

static int check_one(float f)
{
  char a[64], b[64];
  char* end = fmt_e12(a, f);
  *end = 0;
  snprintf(b, sizeof b, "%.12E", (double)f);
  if (strcmp(a, b) != 0) { fprintf(stderr, "mismatch: fast '%s' printf '%s'\n", a, b); return 1; }
  return 0;
}

int main(int argc, char** argv)
{
  long count = argc > 1 ? atol(argv[1]) : 1000000;
  state = argc > 2 ? (uint64_t)atoll(argv[2]) * 2654435761u + 88172645463325252ull : 88172645463325252ull;
  long bad = 0;
  /* raw bit patterns: every exponent, denormals, infinities, NaNs */
  for (long i = 0; i < count; i++) {
    uint32_t u = next_u32();
    float f;
    memcpy(&f, &u, 4);
    if (f != f) continue;                    /* NaN payload/sign printing is not part of the contract */
    bad += check_one(f);
  }
  /* the ranges the solver actually writes: velocities ~1e-9..1e-1, pressures ~0.03 */
  for (long i = 0; i < count; i++) {
    float f = (float)((next_u32() / 4294967296.0) * 0.2 - 0.1);
    bad += check_one(f);
    bad += check_one(f * 1e-6f);
    bad += check_one(0.0333333f + f * 1e-3f);
  }
  const float special[] = {0.0f, -0.0f, 1.0f, -1.0f, 9.9999995e-1f, 9.99999999e8f, 1e-45f, 3.4028235e38f,
                           1.17549435e-38f, 0.1f, 0.5f, 1e10f, 1e-10f, 123456.789f, 9.5f, 99.5f, 0.95f};
  for (size_t i = 0; i < sizeof special / sizeof special[0]; i++) bad += check_one(special[i]);
  /* integers */
  for (long i = 0; i < 100000; i++) {
    unsigned v = i < 70000 ? (unsigned)i : next_u32();
    char a[32], b[32];
    *fmt_uint(a, v) = 0;
    snprintf(b, sizeof b, "%u", v);
    if (strcmp(a, b) != 0) { fprintf(stderr, "uint mismatch %s %s\n", a, b); bad++; }
  }
  printf("%ld\n", bad);
  return bad != 0;
}
