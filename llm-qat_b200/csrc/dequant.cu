// dequant.cu — rebuild the fake-quantized tensor from K1's int8 codes:
//     out[r, c] = fl_dtype(codes[r, c] / e[r])
// which is bit-identical to SymQuantizer's forward output
// (/root/reference/models/utils_quant.py:72) wherever the int8 feed did not
// saturate.  QuantizeLinear's backward uses it for the dgrad/wgrad operands, so
// the forward can save 1 B/elem of codes instead of the reference's two
// dequantized tensors.  HBM-bound: 1 + sizeof(T) bytes per element.
#include "common.cuh"

namespace qat {
namespace {

constexpr int kThreads = 256;

constexpr int kUnroll = 2;  // 16-byte code vectors per thread per tile, both loaded before use

template <int DT>
__device__ __forceinline__ void dequant_vec(const uint4& c, float e, void* out, int64_t j) {
  const bool fast = recip_range_ok(e);
  const float r = __frcp_rn(e);
  // bf16 output and a bf16-valued divisor (what K1 emits for bf16 tensors): one multiply is
  // exact — see SymScale::mulq in common.cuh
  const bool mulq = DT == QAT_BF16 && fast && Num<QAT_BF16>::fl(e) == e;
  const uint32_t w[4] = {c.x, c.y, c.z, c.w};
  float y[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float q = (float)(int)(int8_t)((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
    y[k] = mulq ? __fmul_rn(q, r) : fast ? div_code_by_recip(q, e, r) : __fdiv_rn(q, e);
  }
  if (DT == QAT_BF16) {
    uint4* o = reinterpret_cast<uint4*>(out) + 2 * j;
    stg_stream(o, make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                             pack_bf16x2(y[6], y[7])));
    stg_stream(o + 1, make_uint4(pack_bf16x2(y[8], y[9]), pack_bf16x2(y[10], y[11]),
                                 pack_bf16x2(y[12], y[13]), pack_bf16x2(y[14], y[15])));
  } else {
    uint4* o = reinterpret_cast<uint4*>(out) + 4 * j;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      stg_stream(o + k, make_uint4(__float_as_uint(y[4 * k]), __float_as_uint(y[4 * k + 1]),
                                   __float_as_uint(y[4 * k + 2]), __float_as_uint(y[4 * k + 3])));
  }
}

// one 16-byte vector of codes (16 elements of one row) per thread and unroll
// step; IDX32: the vector index fits 32 bits, so the row lookup is a 32-bit
// division instead of a 64-bit one (~60 instructions per vector).
template <int DT, bool IDX32>
__global__ void __launch_bounds__(kThreads) dequant_codes_kernel(const int8_t* __restrict__ codes,
                                                                 const float* __restrict__ row_e,
                                                                 void* __restrict__ out, int64_t nvec,
                                                                 int vec_per_row) {
  constexpr int64_t kTile = (int64_t)kThreads * kUnroll;
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t base = (int64_t)blockIdx.x * kTile; base < nvec; base += (int64_t)gridDim.x * kTile) {
    uint4 c[kUnroll];
    float e[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t j = base + (int64_t)u * kThreads + threadIdx.x;
      if (j < nvec) {
        c[u] = ldg_stream(reinterpret_cast<const uint4*>(codes) + j);
        e[u] = row_e[IDX32 ? (int64_t)((uint32_t)j / (uint32_t)vec_per_row) : j / vec_per_row];
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t j = base + (int64_t)u * kThreads + threadIdx.x;
      if (j < nvec) dequant_vec<DT>(c[u], e[u], out, j);
    }
  }
}

}  // namespace
}  // namespace qat

extern "C" int qat_dequant_codes(const int8_t* codes, const float* row_e, void* out, int64_t rows,
                                 int64_t cols, int dtype, void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(dtype == QAT_F32 || dtype == QAT_BF16, "dtype must be QAT_F32 or QAT_BF16 (got %d)", dtype);
  QAT_CHECK_ARG(rows >= 0 && cols >= 0, "negative shape");
  if (rows == 0 || cols == 0) return QAT_OK;
  QAT_CHECK_ARG(codes && row_e && out, "NULL operand");
  QAT_CHECK_ARG(cols % 16 == 0, "cols must be a multiple of 16 (got %lld)", (long long)cols);
  QAT_CHECK_ARG(((uintptr_t)codes & 15) == 0 && ((uintptr_t)out & 15) == 0, "pointers must be 16-byte aligned");
  const int64_t nvec = rows * cols / 16;
  QAT_CHECK_ARG(cols / 16 < (1ll << 31), "row too long");
  int64_t grid = (nvec + kThreads * kUnroll - 1) / (kThreads * kUnroll);
  const int64_t cap = (int64_t)num_sms() * 8;   // 8 resident CTAs per SM
  if (grid > cap) grid = cap;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool idx32 = nvec < (1ll << 32);
  const int vpr = (int)(cols / 16);
  const dim3 g((unsigned)grid), b(kThreads);
  if (dtype == QAT_BF16) {
    if (idx32) (void)launch_pdl(dequant_codes_kernel<QAT_BF16, true>, g, b, 0, st, codes, row_e, out, nvec, vpr);
    else (void)launch_pdl(dequant_codes_kernel<QAT_BF16, false>, g, b, 0, st, codes, row_e, out, nvec, vpr);
  } else {
    if (idx32) (void)launch_pdl(dequant_codes_kernel<QAT_F32, true>, g, b, 0, st, codes, row_e, out, nvec, vpr);
    else (void)launch_pdl(dequant_codes_kernel<QAT_F32, false>, g, b, 0, st, codes, row_e, out, nvec, vpr);
  }
  QAT_CHECK_LAUNCH("dequant_codes_kernel");
  return QAT_OK;
}
