/* Exhaustive proof that the per-row-reciprocal quotient used by K1/K2 is
 * bit-identical to IEEE division for every operand the fake-quant path sees.
 *
 *   r  = RN(1/e)                   (once per row:  __frcp_rn)
 *   q0 = RN(q*r)
 *   rem= fma(-e, q0, q)            (exact)
 *   y  = fma(r, rem, q0)           == RN(q/e) ?
 *
 * Scaling e by a power of two scales every intermediate exactly (no
 * over/underflow in the path's range: e in [1e-6, 1.3e8], |q| <= 256), so it
 * suffices to sweep all 2^23 mantissas of e in [1,2) for every integer code
 * q in [1, QMAX].  Also sweeps the constant-divisor case u = q/S of the Asym
 * path (S = 2^bits - 1).   Test infrastructure; build:
 *   gcc -O2 -march=native -ffp-contract=off -fopenmp div_by_reciprocal.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
  int qmax = argc > 1 ? atoi(argv[1]) : 256;
  long long bad = 0, total = 0;
#pragma omp parallel for reduction(+ : bad, total) schedule(static)
  for (uint32_t m = 0; m < (1u << 23); ++m) {
    const float e = bits2f(0x3f800000u | m);
    const float r = 1.0f / e;
    for (int qi = 1; qi <= qmax; ++qi) {
      const float q = (float)qi;
      const float q0 = q * r;
      const float rem = fmaf(-e, q0, q);
      const float y = fmaf(r, rem, q0);
      const float ref = q / e;
      bad += f2bits(y) != f2bits(ref);
      ++total;
    }
  }
  printf("per-row divisor: %lld quotients, %lld mismatches (q in [1,%d], all 2^23 mantissas)\n", total, bad, qmax);
  long long bad2 = 0, total2 = 0;
  for (int bits = 1; bits <= 15; ++bits) {
    const float S = (float)((1 << bits) - 1);
    const float r = 1.0f / S;
    for (int qi = 0; qi <= (1 << bits) - 1; ++qi) {
      const float q = (float)qi;
      const float q0 = q * r;
      const float y = fmaf(r, fmaf(-S, q0, q), q0);
      bad2 += f2bits(y) != f2bits(q / S);
      ++total2;
    }
  }
  printf("constant divisor S=2^b-1: %lld quotients, %lld mismatches (b in [1,15])\n", total2, bad2);
  return (bad || bad2) ? 1 : 0;
}
