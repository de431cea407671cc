/* Exhaustive proof for the packed-bf16 AsymQuantizer path of K2 (bf16 tensors).
 *
 * The reference (utils_quant.py:144-147, every op rounded to bf16 after an fp32
 * computation) needs two quotients per element:
 *     n = fl_bf16( fl_f32( d / a ) )     d = fl_bf16(x - beta) in [0, alpha],  a = fl_bf16(alpha + 1e-8)
 *     u = fl_bf16( fl_f32( c / S ) )     c an integer code in [0, S],  S = 2^bits - 1
 * K2 computes them as ONE fp32 multiply by a per-row reciprocal and one rounding to bf16:
 *     n' = fl_bf16( fl_f32( d * ra ) ),  ra = RN_f32(1/a)        (__frcp_rn, once per row)
 *     u' = fl_bf16( fl_f32( c * rS ) ),  rS = RN_f32(1/S)
 * and SymQuantizer's dequantization (utils_quant.py:72) one more of the same kind:
 *     y = fl_bf16( fl_f32( q / e ) )     q = rint(fl_bf16(x*s)): an integer that is a bf16 value,
 *                                        e = fl_bf16(s + 1e-6)
 *     y' = fl_bf16( fl_f32( q * re ) ),  re = RN_f32(1/e)
 * (checked for every bf16 e in the window and every integer-valued bf16 q in [0, 512]).
 * Claim: n' == n and u' == u (and y' == y), bit for bit, for EVERY bf16 a in the guarded window
 * [2^-100, 2^100] with every bf16 d in [0, a], and for every bits in 2..8 with every c.
 * Why it holds: d and a carry 8 significant bits, so d/a is either exactly
 * representable in bf16 or at least 2^-17 (relative) away from every bf16
 * rounding midpoint (a midpoint has an odd 9-bit significand M; M*A = D*2^k has no
 * solution with 8-bit D), while the reciprocal route is off by < 2^-22.  This
 * program checks all (a, d) pairs instead of trusting the argument
 * (~4.2e8 quotients, seconds).
 *
 * Second claim (argv[1] == "addsub", ~4.3e9 pairs): for bf16 operands the packed
 * single-rounding instructions add/sub.rn.bf16x2 equal the reference's
 * fl_bf16(fl_f32(x +- y)): the fp32 sum of two 8-bit significands is exact when the
 * exponents differ by <= 16 and otherwise the small operand cannot move the bf16
 * rounding either way.  Checked for every pair of finite bf16 values against the
 * sum taken in double.  (bf16 x bf16 products are exact in fp32: 16 bits.)  Test infrastructure; build:
 *   gcc -O2 -ffp-contract=off -fopenmp bf16_quotient_by_reciprocal.c -o bf16q && ./bf16q
 * Exit status 0 and "0 mismatches" == proven.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
/* fp32 -> bf16 bits, round to nearest even (finite inputs only) */
static inline uint16_t bf16_rn(float f) {
  uint32_t u = f2bits(f);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf16_to_f(uint16_t h) { return bits2f((uint32_t)h << 16); }

/* double -> bf16 bits, round to nearest even, via the exact fp32 candidates around it */
static inline uint16_t bf16_rn_d(double v) {
  float f = (float)v;                    /* RN to fp32 */
  double fd = (double)f;
  uint32_t u = f2bits(f);
  /* sticky: if the double was not exactly f, nudge the fp32 pattern's LSB-below bits so that a
   * tie at the bf16 midpoint is broken the way the exact value demands */
  if (fd != v && (u & 0xffffu) == 0x8000u) {
    int up = (v > fd) == (f >= 0);       /* exact magnitude above f's magnitude? */
    return (uint16_t)((u >> 16) + (up ? 1 : 0));
  }
  return bf16_rn(f);
}

static int check_addsub(void) {
  long long bad = 0, total = 0;
#pragma omp parallel for reduction(+ : bad, total) schedule(dynamic, 16)
  for (int xb = 0; xb < 65536; ++xb) {
    if (((xb >> 7) & 0xff) == 0xff) continue;              /* inf / NaN */
    const float x = bf16_to_f((uint16_t)xb);
    for (int yb = 0; yb < 65536; ++yb) {
      if (((yb >> 7) & 0xff) == 0xff) continue;
      const float y = bf16_to_f((uint16_t)yb);
      const float s32 = x + y;
      if ((f2bits(s32) & 0x7f800000u) == 0x7f800000u) continue;   /* overflow: both routes give inf */
      const uint16_t twice = bf16_rn(s32);                 /* the reference: fp32 op, then bf16 */
      const uint16_t once = bf16_rn_d((double)x + (double)y);   /* what add.rn.bf16x2 returns */
      bad += twice != once;
      ++total;
    }
  }
  printf("add/sub: %lld pairs, %lld mismatches\n", total, bad);
  return bad != 0;
}

/* Third claim (argv[1] == "anyratio", ~8.4e8 quotients): the same single-multiply quotient for
 * ANY finite bf16 numerator, not only d <= a — the W1/W2 weight path divides each weight by its
 * row's mean |w| (utils_quant.py:211,229), and a weight can be far above or below the mean. */
static int check_anyratio(void) {
  long long bad = 0, total = 0;
#pragma omp parallel for reduction(+ : bad, total) schedule(dynamic, 64)
  for (int ab = (27 << 7); ab < (228 << 7); ++ab) {
    const float a = bf16_to_f((uint16_t)ab);
    const float ra = 1.0f / a;
    for (int db = 0; db < (255 << 7); ++db) {            /* every finite non-negative bf16 numerator */
      const float d = bf16_to_f((uint16_t)db);
      const float quo = d / a;
      if ((f2bits(quo) & 0x7f800000u) == 0x7f800000u) continue;   /* overflow: inf both ways */
      bad += bf16_rn(quo) != bf16_rn(d * ra);
      ++total;
    }
  }
  printf("any ratio: %lld quotients, %lld mismatches\n", total, bad);
  return bad != 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && strcmp(argv[1], "addsub") == 0) return check_addsub();
  if (argc > 1 && strcmp(argv[1], "anyratio") == 0) return check_anyratio();
  long long bad = 0, total = 0;
  /* a: every positive bf16 with exponent in [-100, 100]  (biased 27 .. 227) */
#pragma omp parallel for reduction(+ : bad, total) schedule(dynamic, 64)
  for (int ab = (27 << 7); ab < (228 << 7); ++ab) {
    const float a = bf16_to_f((uint16_t)ab);
    const float ra = 1.0f / a;                       /* correctly rounded, like __frcp_rn */
    for (int db = 0; db <= ab; ++db) {               /* every bf16 d with 0 <= d <= a (incl. subnormals) */
      const float d = bf16_to_f((uint16_t)db);
      const uint16_t ref = bf16_rn(d / a);
      const uint16_t got = bf16_rn(d * ra);
      bad += ref != got;
      ++total;
    }
  }
  /* Sym dequantization: integer-valued bf16 codes 0..512 over every bf16 divisor in the window */
#pragma omp parallel for reduction(+ : bad, total) schedule(dynamic, 64)
  for (int eb = (27 << 7); eb < (228 << 7); ++eb) {
    const float e = bf16_to_f((uint16_t)eb);
    const float re = 1.0f / e;
    for (int q = 0; q <= 512; ++q) {
      const float qf = (float)q;
      if (bf16_to_f(bf16_rn(qf)) != qf) continue;      /* not a bf16 value (odd numbers above 256) */
      const float quo = qf / e;
      if ((f2bits(quo) & 0x7f800000u) == 0x7f800000u) continue;
      bad += bf16_rn(quo) != bf16_rn(qf * re);
      ++total;
    }
  }
  for (int bits = 2; bits <= 8; ++bits) {
    const float S = (float)((1 << bits) - 1);
    const float rS = 1.0f / S;
    for (int c = 0; c <= (1 << bits) - 1; ++c) {
      bad += bf16_rn((float)c / S) != bf16_rn((float)c * rS);
      ++total;
    }
  }
  printf("%lld quotients, %lld mismatches\n", total, bad);
  return bad != 0;
}
