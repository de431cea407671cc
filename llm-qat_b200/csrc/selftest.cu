// selftest.cu — device self-test of the hoisted-reciprocal divisions used by
// K1/K2 (common.cuh: FastRecip, div_code_by_recip) against div.rn, over the
// operand domain the kernels guard: divisor b in [2^-100, 2^100], numerator a in
// [0, b] (Asym's fl(x - beta) <= alpha <= a) or an integer code |q| <= 32767.
// A mismatch is counted only where the exact quotient is a normal number — below
// that the quantizer's code is 0 whatever the low bits are.
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
// counters[2] = mismatches (integer numerators),  [3] = tested
__global__ void __launch_bounds__(256) fastdiv_selftest_kernel(uint64_t seed, int per_row, int bf16_operands,
                                                               unsigned long long* counters) {
  uint64_t s = seed ^ (0xd1342543de82ef95ull * (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x + 1));
  unsigned long long bad = 0, n = 0, bad_i = 0, n_i = 0;
  const uint32_t mant_mask = bf16_operands ? 0x007f0000u : 0x007fffffu;
  // one "row" per thread: a divisor, then per_row numerators
  const uint64_t r = splitmix(s);
  const int eb = (int)(r % 201) - 100;  // exponent in [-100, 100]
  const float b = __uint_as_float(((uint32_t)(eb + 127) << 23) | ((uint32_t)(r >> 20) & mant_mask));
  if (!recip_range_ok(b)) return;
  FastRecip fr;
  fr.set(b);
  const float rb = __frcp_rn(b);
  for (int k = 0; k < per_row; ++k) {
    const uint64_t t = splitmix(s);
    // numerator: random mantissa, exponent 0..60 below the divisor's, clamped to <= b
    int ea = eb - (int)((t >> 56) % 61);
    float a = __uint_as_float(((uint32_t)max(ea + 127, 1) << 23) | ((uint32_t)t & mant_mask));
    if (a > b) a = b;
    if ((t >> 52 & 0xf) == 0) a = 0.0f;
    const float ref = __fdiv_rn(a, b);
    const float got = or_sign(fr.div(a, b), a);
    if (fabsf(ref) >= 0x1p-120f || ref == 0.0f) {
      ++n;
      bad += __float_as_uint(ref) != __float_as_uint(got);
    }
    // integer numerator (a code) over the same divisor
    const float q = (float)((int)((t >> 20) % 65535) - 32767);
    const float ref_i = __fdiv_rn(q, b);
    const float got_i = or_sign(div_code_by_recip(q, b, rb), q);
    if (fabsf(ref_i) < 0x1p120f) {
      ++n_i;
      bad_i += __float_as_uint(ref_i) != __float_as_uint(got_i);
    }
  }
  atomicAdd(counters + 0, bad);
  atomicAdd(counters + 1, n);
  atomicAdd(counters + 2, bad_i);
  atomicAdd(counters + 3, n_i);
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
