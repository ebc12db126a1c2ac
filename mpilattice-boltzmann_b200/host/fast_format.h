/*
 * fast_format.h -- exact, fast replacements for the two printf conversions that dominate the
 * writing of final_state.dat / av_vels.dat ("%d" and "%.12E", reference d2q9-bgk.c:1115, 1136).
 * Checked digit for digit against glibc printf by tests/test_host_cli.py (tests/c/fmt_check.c).
 */
#ifndef LBM_FAST_FORMAT_H
#define LBM_FAST_FORMAT_H

#include <math.h>
#include <stdio.h>
#include <string.h>

/*
 * final_state.dat has four %.12E columns per cell; glibc printf spends most of the output
 * time there.  fmt_e12 writes exactly what printf("%.12E", (double)v) writes for every
 * finite float: a float widened to double has at most 24 significant bits, so the decimal
 * expansion is formed exactly with 128-bit integer arithmetic and rounded half-to-even like
 * glibc.  Anything unusual (NaN, infinity) falls back to snprintf. */
static inline char* fmt_e12_slow(char* out, double v) { return out + sprintf(out, "%.12E", v); }

static inline char* fmt_e12(char* out, float f)
{
  if (f == 0.0f) {
    if (signbit(f)) *out++ = '-';
    memcpy(out, "0.000000000000E+00", 18);
    return out + 18;
  }
  if (!isfinite(f)) return fmt_e12_slow(out, (double)f);
  int e2;
  double m = frexp(fabs((double)f), &e2);               /* |f| = m * 2^e2, m in [0.5, 1) */
  unsigned long long mant = (unsigned long long)ldexp(m, 24);   /* exact: <= 24 significant bits */
  e2 -= 24;                                             /* |f| = mant * 2^e2 */
  if (e2 > 40 || e2 < -200) return fmt_e12_slow(out, (double)f);
  /* decimal exponent estimate, then exact digits = floor(|f| * 10^(12-e10)) with remainder */
  int e10 = (int)floor(log10(fabs((double)f)));
  for (int attempt = 0; attempt < 3; attempt++) {
    /* value * 10^s with s = 12 - e10, as num/den in 128-bit integers */
    int s = 12 - e10;
    unsigned __int128 num = mant, den = 1;
    if (e2 >= 0) num <<= e2; else {
      if (-e2 >= 120) return fmt_e12_slow(out, (double)f);
      den <<= -e2;
    }
    int ok = 1;
    if (s >= 0) { for (int i = 0; i < s; i++) { if (num >> 123) { ok = 0; break; } num *= 10; } }
    else { for (int i = 0; i < -s; i++) { if (den >> 123) { ok = 0; break; } den *= 10; } }
    if (!ok) return fmt_e12_slow(out, (double)f);
    unsigned __int128 q = num / den, r = num - q * den;
    /* round half to even on the exact remainder */
    unsigned __int128 twice = r * 2;
    if (twice > den || (twice == den && (q & 1))) q++;
    const unsigned __int128 lo = (unsigned __int128)1000000000000ULL;          /* 10^12 */
    const unsigned __int128 hi = (unsigned __int128)10000000000000ULL;         /* 10^13 */
    if (q < lo) { e10--; continue; }
    if (q >= hi) {
      /* either the estimate was one too low, or rounding carried 9.99..->10.0 */
      e10++; continue;
    }
    unsigned long long digits = (unsigned long long)q;  /* 13 digits */
    char buf[13];
    for (int i = 12; i >= 0; i--) { buf[i] = (char)('0' + digits % 10); digits /= 10; }
    if (signbit(f)) *out++ = '-';
    *out++ = buf[0];
    *out++ = '.';
    memcpy(out, buf + 1, 12);
    out += 12;
    *out++ = 'E';
    int ae = e10 < 0 ? -e10 : e10;
    *out++ = e10 < 0 ? '-' : '+';
    if (ae >= 100) { *out++ = (char)('0' + ae / 100); ae %= 100; }
    *out++ = (char)('0' + ae / 10);
    *out++ = (char)('0' + ae % 10);
    return out;
  }
  return fmt_e12_slow(out, (double)f);
}

static inline char* fmt_uint(char* out, unsigned v)
{
  char tmp[12];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *out++ = tmp[--n];
  return out;
}

#endif /* LBM_FAST_FORMAT_H */
