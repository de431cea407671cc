// lowbit.cu — QuantizeLinear's w_bits in {1, 2} weight path (secondary; no
// BASELINE config exercises it).  Replaces the eager chain at
// /root/reference/models/utils_quant.py:202-242:
//   1-bit : sf = mean|w|      ; q = sf * sign(w / sf)
//   2-bit : sf = 2 * mean|w|  ; q = sf * (round(clamp(w/sf, -0.99, 0.99) * 2 - 0.5) + 0.5) / 2
//   forward value = (q - w) + w           (gradient is the identity)
// The mean is per output row, or over the whole tensor when layerwise.
//
// Two launches: a double-precision sum of |w| merged with atomics (the sum
// order of torch's vectorised CPU reduction is not a portable contract; fp64
// makes ours order-independent to within fp32 rounding), then one streaming
// pass that re-reads w (L2-resident for LLaMA-sized weights) and applies.
#include "common.cuh"

namespace qat {
namespace {

constexpr int kThreads = 256;

struct LowbitParams {
  const void* w;
  void* out;
  double* sums;  // [rows] or [1]
  int64_t rows, cols;
  int64_t chunk;  // elements of a row per CTA
  int w_bits;
  int layerwise;
};

template <int DT>
__device__ __forceinline__ float load_elem(const void* p, int64_t i) {
  if (DT == QAT_F32) return reinterpret_cast<const float*>(p)[i];
  return bf16lo(reinterpret_cast<const uint16_t*>(p)[i]);
}

template <int DT>
__global__ void __launch_bounds__(kThreads) lowbit_sum_kernel(const LowbitParams p) {
  __shared__ double sm[kThreads / 32];
  const int64_t row = blockIdx.y;
  const int64_t j0 = (int64_t)blockIdx.x * p.chunk;
  const int64_t j1 = min(j0 + p.chunk, p.cols);
  double acc = 0.0;
  for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads)
    acc += (double)fabsf(load_elem<DT>(p.w, row * p.cols + j));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) t += sm[i];
    atomicAdd(p.sums + (p.layerwise ? 0 : row), t);
  }
}

template <int DT>
__global__ void __launch_bounds__(kThreads) lowbit_apply_kernel(const LowbitParams p) {
  using N = Num<DT>;
  const int64_t row = blockIdx.y;
  const int64_t j0 = (int64_t)blockIdx.x * p.chunk;
  const int64_t j1 = min(j0 + p.chunk, p.cols);
  const double count = p.layerwise ? (double)p.rows * (double)p.cols : (double)p.cols;
  const float mean_abs = N::fl((float)(p.sums[p.layerwise ? 0 : row] / count));  // :205-210 / :219-224
  const float sf = (p.w_bits == 1) ? mean_abs : N::fl(__fmul_rn(2.0f, mean_abs));
  const float clip = N::fl(0.99f);  // 1 - 1e-2, cast to the tensor dtype by clamp (:218)
  for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
    const int64_t i = row * p.cols + j;
    const float w = load_elem<DT>(p.w, i);
    const float r = N::fl(__fdiv_rn(w, sf));
    float q;
    if (p.w_bits == 1) {
      const float sgn = (r > 0.f) ? 1.f : (r < 0.f) ? -1.f : 0.f;  // torch.sign: NaN -> 0
      q = N::fl(__fmul_rn(sf, sgn));                               // :211-213
    } else {
      float t = (r != r) ? r : fminf(fmaxf(r, -clip), clip);       // clamp keeps NaN (:229-231)
      t = N::fl(__fsub_rn(N::fl(__fmul_rn(t, 2.0f)), 0.5f));       // :232-233
      t = N::fl(__fadd_rn(rintf(t), 0.5f));                        // :228,235
      q = N::fl(__fdiv_rn(N::fl(__fmul_rn(sf, t)), 2.0f));         // :226-237
    }
    const float eff = N::fl(__fadd_rn(N::fl(__fsub_rn(q, w)), w));  // :240-242
    if (DT == QAT_F32)
      reinterpret_cast<float*>(p.out)[i] = eff;
    else
      reinterpret_cast<__nv_bfloat16*>(p.out)[i] = __float2bfloat16_rn(eff);
  }
}

}  // namespace
}  // namespace qat

extern "C" {

size_t qat_lowbit_workspace_bytes(int64_t rows, int layerwise) {
  if (rows <= 0) return 0;
  return (size_t)(layerwise ? 1 : rows) * sizeof(double);
}

int qat_lowbit_weight_fwd(const void* w, void* w_eff, int64_t rows, int64_t cols, int dtype,
                          int w_bits, int layerwise, void* workspace, size_t workspace_bytes,
                          void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(dtype == QAT_F32 || dtype == QAT_BF16, "dtype must be QAT_F32 or QAT_BF16 (got %d)", dtype);
  QAT_CHECK_ARG(w_bits == 1 || w_bits == 2, "low-bit path handles w_bits 1 or 2 (got %d)", w_bits);
  QAT_CHECK_ARG(rows >= 0 && cols >= 0, "negative shape");
  if (rows == 0 || cols == 0) return QAT_OK;
  QAT_CHECK_ARG(w != nullptr && w_eff != nullptr && w != w_eff, "w / w_eff NULL or aliased");
  const size_t need = qat_lowbit_workspace_bytes(rows, layerwise);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("low-bit weight path needs %zu bytes of workspace (got %zu)", need, workspace_bytes);
    return QAT_ERR_WORKSPACE;
  }
  if (rows > 65535) {
    set_error("low-bit weight path supports at most 65535 rows (got %lld)", (long long)rows);
    return QAT_ERR_UNSUPPORTED;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
  LowbitParams p{};
  p.w = w;
  p.out = w_eff;
  p.sums = reinterpret_cast<double*>(workspace);
  p.rows = rows;
  p.cols = cols;
  p.w_bits = w_bits;
  p.layerwise = layerwise ? 1 : 0;
  p.chunk = 8192;
  dim3 grid((unsigned)((cols + p.chunk - 1) / p.chunk), (unsigned)rows);
  if (dtype == QAT_F32) {
    lowbit_sum_kernel<QAT_F32><<<grid, kThreads, 0, st>>>(p);
    QAT_CHECK_LAUNCH("lowbit_sum_kernel");
    lowbit_apply_kernel<QAT_F32><<<grid, kThreads, 0, st>>>(p);
    QAT_CHECK_LAUNCH("lowbit_apply_kernel");
  } else {
    lowbit_sum_kernel<QAT_BF16><<<grid, kThreads, 0, st>>>(p);
    QAT_CHECK_LAUNCH("lowbit_sum_kernel");
    lowbit_apply_kernel<QAT_BF16><<<grid, kThreads, 0, st>>>(p);
    QAT_CHECK_LAUNCH("lowbit_apply_kernel");
  }
  return QAT_OK;
}

}  // extern "C"
