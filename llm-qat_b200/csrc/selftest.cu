// selftest.cu — device self-test of the hoisted-reciprocal divisions used by
// K1/K2 (common.cuh: FastRecip, div_code_by_recip) against div.rn, over the
// operand domain the kernels see:
//   general numerators (Asym's n = fl(x - beta) / a):  divisor a = alpha + 1e-8
//     in [2^-27, 2^100] (the row guard), numerator in [0, a], quotient >= 2^-41.
//     Quotients below 2^-17 give code 0 for every bits <= 15 whatever their low
//     bits are, so the domain is 2^24 wider than what can influence a result.
//   integer numerators (the codes): |q| <= 32767 over divisors in [2^-100, 2^100].
// counters[4..5] additionally sweep numerators down to the denormal boundary,
// where div.rn leaves its fast path and bit-equality is NOT expected (and not
// needed): reported for information only.
#define QAT_PDL_FAMILY 10   // bit of QAT_B200_PDL_MASK (common.cuh)
#include "common.cuh"

namespace qat {
namespace {

__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d2049bb133111bull;
  return z ^ (z >> 31);
}

// counters[0] = mismatches (general numerators), [1] = tested,
// counters[2] = mismatches (integer numerators),  [3] = tested,
// counters[4] = mismatches (wide domain, informational), [5] = tested
__global__ void __launch_bounds__(256) fastdiv_selftest_kernel(uint64_t seed, int per_row, int bf16_operands,
                                                               unsigned long long* counters) {
  uint64_t s = seed ^ (0xd1342543de82ef95ull * (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x + 1));
  unsigned long long bad = 0, n = 0, bad_i = 0, n_i = 0, bad_w = 0, n_w = 0;
  const uint32_t mant_mask = bf16_operands ? 0x007f0000u : 0x007fffffu;
  // one "row" per thread: a divisor, then per_row numerators
  const uint64_t r = splitmix(s);
  const int eb = (int)(r % 128) - 27;  // divisor exponent in [-27, 100]
  const float b = __uint_as_float(((uint32_t)(eb + 127) << 23) | ((uint32_t)(r >> 20) & mant_mask));
  const int ebw = (int)((r >> 8) % 201) - 100;  // wide sweep: [-100, 100]
  const float bw = __uint_as_float(((uint32_t)(ebw + 127) << 23) | ((uint32_t)(r >> 20) & mant_mask));
  if (!recip_range_ok(b) || !recip_range_ok(bw)) return;
  FastRecip fr, frw;
  fr.set(b);
  frw.set(bw);
  const float rb = __frcp_rn(bw);
  for (int k = 0; k < per_row; ++k) {
    const uint64_t t = splitmix(s);
    // numerator: random mantissa, exponent 0..40 below the divisor's, clamped to <= b
    const int ea = eb - (int)((t >> 56) % 41);
    float a = __uint_as_float(((uint32_t)(ea + 127) << 23) | ((uint32_t)t & mant_mask));
    if (a > b) a = b;
    if ((t >> 52 & 0xf) == 0) a = 0.0f;
    {
      const float ref = __fdiv_rn(a, b);
      const float got = or_sign(fr.div(a, b), a);
      ++n;
      bad += __float_as_uint(ref) != __float_as_uint(got);
    }
    {  // informational: numerators down to the smallest normal
      const int eaw = ebw - (int)((t >> 56) % 61);
      float aw = __uint_as_float(((uint32_t)max(eaw + 127, 1) << 23) | ((uint32_t)t & mant_mask));
      if (aw > bw) aw = bw;
      const float ref = __fdiv_rn(aw, bw);
      const float got = or_sign(frw.div(aw, bw), aw);
      ++n_w;
      bad_w += __float_as_uint(ref) != __float_as_uint(got);
    }
    // integer numerator (a code) over the same divisor
    const float q = (float)((int)((t >> 20) % 65535) - 32767);
    const float ref_i = __fdiv_rn(q, bw);
    const float got_i = or_sign(div_code_by_recip(q, bw, rb), q);
    if (fabsf(ref_i) < 0x1p120f) {
      ++n_i;
      bad_i += __float_as_uint(ref_i) != __float_as_uint(got_i);
    }
  }
  atomicAdd(counters + 0, bad);
  atomicAdd(counters + 1, n);
  atomicAdd(counters + 2, bad_i);
  atomicAdd(counters + 3, n_i);
  atomicAdd(counters + 4, bad_w);
  atomicAdd(counters + 5, n_w);
}

}  // namespace
}  // namespace qat

extern "C" int qat_selftest_fastdiv(uint64_t seed, int64_t rows, int per_row, int bf16_operands,
                                    uint64_t* dev_counters, void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(rows > 0 && per_row > 0 && dev_counters != nullptr, "bad self-test arguments");
  const int64_t blocks = (rows + 255) / 256;
  QAT_CHECK_ARG(blocks < (1ll << 31), "too many rows");
  fastdiv_selftest_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      seed, per_row, bf16_operands, reinterpret_cast<unsigned long long*>(dev_counters));
  QAT_CHECK_LAUNCH("fastdiv_selftest_kernel");
  return QAT_OK;
}
