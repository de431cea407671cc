// lowbit.cu — QuantizeLinear's w_bits in {1, 2} weight path (secondary; no
// BASELINE config exercises it).  Replaces the eager chain at
// /root/reference/models/utils_quant.py:202-242:
//   1-bit : sf = mean|w|      ; q = sf * sign(w / sf)
//   2-bit : sf = 2 * mean|w|  ; q = sf * (round(clamp(w/sf, -0.99, 0.99) * 2 - 0.5) + 0.5) / 2
//   forward value = (q - w) + w           (gradient is the identity)
// The mean is per output row, or over the whole tensor when layerwise.
//
// Row-wise mode with rows that fit registers (every LLaMA weight): ONE pass, the
// K1 structure — a thread group owns a row, loads it with 128-bit streaming
// loads, sums |w| in double precision (the sum order of torch's vectorised CPU
// reduction is not a portable contract; fp64 makes ours order-independent to
// within fp32 rounding), derives the scale and applies from registers: 2e B/elem.
// Layerwise mode, very long or unaligned rows: two launches, a double-precision
// sum merged with atomics, then a pass that re-reads w and applies (3e B/elem).
#define QAT_PDL_FAMILY 3   // bit of QAT_B200_PDL_MASK (common.cuh)
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

// (q - w) + w for one element, every op rounded to the tensor dtype — :205-242
// `rsf`: RN(1/sf) when the quotient may be taken as one multiply (bf16 tensors, sf in the proven
// window: oracle/proofs/bf16_quotient_by_reciprocal.c "anyratio", every bf16 numerator x every
// bf16 divisor, 0 mismatches), else 0 -> exact IEEE division.
template <int DT>
__device__ __forceinline__ float lowbit_eff(float w, float sf, float rsf, float clip, int w_bits) {
  using N = Num<DT>;
  const float r = N::fl((DT == QAT_BF16 && rsf != 0.f) ? __fmul_rn(w, rsf) : __fdiv_rn(w, sf));
  float q;
  if (w_bits == 1) {
    const float sgn = (r > 0.f) ? 1.f : (r < 0.f) ? -1.f : 0.f;  // torch.sign: NaN -> 0
    q = N::fl(__fmul_rn(sf, sgn));                               // :211-213
  } else {
    float t = (r != r) ? r : fminf(fmaxf(r, -clip), clip);       // clamp keeps NaN (:229-231)
    t = N::fl(__fsub_rn(N::fl(__fmul_rn(t, 2.0f)), 0.5f));       // :232-233
    t = N::fl(__fadd_rn(rintf(t), 0.5f));                        // :228,235
    q = N::fl(__fmul_rn(N::fl(__fmul_rn(sf, t)), 0.5f));         // :226-237  (/ 2 == * 0.5 exactly)
  }
  return N::fl(__fadd_rn(N::fl(__fsub_rn(q, w)), w));            // :240-242
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
  const float rsf = (DT == QAT_BF16 && recip_range_ok(sf)) ? __frcp_rn(sf) : 0.f;
  for (int64_t j = j0 + threadIdx.x; j < j1; j += kThreads) {
    const int64_t i = row * p.cols + j;
    const float eff = lowbit_eff<DT>(load_elem<DT>(p.w, i), sf, rsf, clip, p.w_bits);
    if (DT == QAT_F32)
      reinterpret_cast<float*>(p.out)[i] = eff;
    else
      reinterpret_cast<__nv_bfloat16*>(p.out)[i] = __float2bfloat16_rn(eff);
  }
}

// bf16, two elements per instruction wherever the op exists packed: mul/add/sub.rn.bf16x2 round
// once where the reference rounds twice, which is identical for bf16 operands (products are
// exact in fp32; sums: oracle/proofs "addsub"); rint(t) for t in [-2.5, 1.5] is (t + 192) - 192
// in bf16 arithmetic (spacing 1 in [128, 256), ties to even; the lost sign of -0 dies in + 0.5).
__device__ __forceinline__ uint32_t lowbit_pair_bf16(uint32_t w2, float sf, float rsf, uint32_t sf2, uint32_t clip2,
                                                     int w_bits) {
  const float r0 = __fmul_rn(bf16lo(w2), rsf), r1 = __fmul_rn(bf16hi(w2), rsf);   // w / sf  (one multiply: "anyratio")
  uint32_t q2;
  if (w_bits == 1) {
    const uint32_t rb = pack_bf16x2(r0, r1);   // fl_bf16 first: a quotient may round to zero
    const float a = bf16lo(rb), b = bf16hi(rb);
    const float s0 = (a > 0.f) ? sf : (a < 0.f) ? -sf : __fmul_rn(sf, 0.f);        // sf * sign(r); sign(NaN) = 0
    const float s1 = (b > 0.f) ? sf : (b < 0.f) ? -sf : __fmul_rn(sf, 0.f);
    q2 = pack_bf16x2(s0, s1);
  } else {
    const uint32_t k2 = 0x40004000u, kh = 0x3f003f00u, k192 = 0x43404340u;        // 2.0, 0.5, 192.0 as bf16x2
    uint32_t t2 = clamp_nan_bf16x2(pack_bf16x2(r0, r1), clip2 ^ 0x80008000u, clip2);  // :229-231
    t2 = sub_bf16x2(mul_bf16x2(t2, k2), kh);                                      // :232-233  * 2 - 0.5
    t2 = sub_bf16x2(add_bf16x2(t2, k192), k192);                                  // :228      round
    t2 = add_bf16x2(t2, kh);                                                      // :235      + 0.5
    q2 = mul_bf16x2(mul_bf16x2(sf2, t2), kh);                                     // :226-237  sf * t / 2
  }
  return add_bf16x2(sub_bf16x2(q2, w2), w2);                                      // :240-242  (q - w) + w
}

// one thread group (a power of two >= 32 threads) per row, ITERS 16-byte vectors per thread
template <int DT, int ITERS>
__global__ void __launch_bounds__(1024) lowbit_row_kernel(const void* __restrict__ w, void* __restrict__ out,
                                                          int64_t rows, int64_t nvec, int64_t cols, int log2_group,
                                                          int w_bits) {
  using N = Num<DT>;
  __shared__ double sm[32];
  const uint32_t group = 1u << log2_group;
  const uint32_t t = threadIdx.x & (group - 1u);
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> log2_group) + (threadIdx.x >> log2_group);
  const bool row_ok = row < rows;
  const uint32_t nv = row_ok ? (uint32_t)nvec : 0u;
  const uint4* wrow = reinterpret_cast<const uint4*>(w) + row * nvec;
  pdl_wait();
  pdl_launch_dependents();
  uint4 v[ITERS];
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const uint32_t j = t + (uint32_t)i * group;
    v[i] = (j < nv) ? ldg_stream(wrow + j) : make_uint4(0u, 0u, 0u, 0u);
  }
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    if (DT == QAT_BF16) {
      // eight bf16 magnitudes summed in fp32 (exact unless their exponents spread over more than
      // 2^13 — and then off by < 2^-24, below what torch's own fp32 accumulation loses), fp64 across
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < N::kPerVec; ++k) part += fabsf(vec_get<DT>(v[i], k));
      acc += (double)part;
    } else {
#pragma unroll
      for (int k = 0; k < N::kPerVec; ++k) acc += (double)fabsf(vec_get<DT>(v[i], k));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (group > 32) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    const int nw = group >> 5;
    const int base = (warp / nw) * nw;
    double tsum = (lane < nw) ? sm[base + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
    acc = tsum;
  }
  const float mean_abs = N::fl((float)(acc / (double)cols));                 // :205-210 / :219-224
  const float sf = (w_bits == 1) ? mean_abs : N::fl(__fmul_rn(2.0f, mean_abs));
  const float clip = N::fl(0.99f);  // 1 - 1e-2, cast to the tensor dtype by clamp (:218)
  const float rsf = (DT == QAT_BF16 && recip_range_ok(sf)) ? __frcp_rn(sf) : 0.f;
  uint4* orow = reinterpret_cast<uint4*>(out) + row * nvec;
  if constexpr (DT == QAT_BF16) {
    if (rsf != 0.f) {   // row-uniform: the packed chain, from registers
      const uint32_t sf2 = pack_bf16x2(sf, sf), clip2 = pack_bf16x2(clip, clip);
#pragma unroll
      for (int i = 0; i < ITERS; ++i) {
        const uint32_t j = t + (uint32_t)i * group;
        const uint4 o = make_uint4(lowbit_pair_bf16(v[i].x, sf, rsf, sf2, clip2, w_bits),
                                   lowbit_pair_bf16(v[i].y, sf, rsf, sf2, clip2, w_bits),
                                   lowbit_pair_bf16(v[i].z, sf, rsf, sf2, clip2, w_bits),
                                   lowbit_pair_bf16(v[i].w, sf, rsf, sf2, clip2, w_bits));
        if (j < nv) stg_stream(orow + j, o);
      }
    } else if (row_ok) {
      // scale 0 / inf / NaN / outside the proven window (all-zero or overflowed rows): exact IEEE
      // division, element by element straight from global memory — rare, and kept away from the
      // register-resident path (the division's slow path is a subroutine call)
      const uint16_t* wr = reinterpret_cast<const uint16_t*>(w) + row * cols;
      __nv_bfloat16* outr = reinterpret_cast<__nv_bfloat16*>(out) + row * cols;
#pragma unroll 1
      for (int64_t c = t; c < cols; c += group)
        outr[c] = __float2bfloat16_rn(lowbit_eff<DT>(bf16lo(wr[c]), sf, 0.f, clip, w_bits));
    }
  } else {
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const uint32_t j = t + (uint32_t)i * group;
      float y[N::kPerVec];
#pragma unroll
      for (int k = 0; k < N::kPerVec; ++k) y[k] = lowbit_eff<DT>(vec_get<DT>(v[i], k), sf, rsf, clip, w_bits);
      const uint4 o = make_uint4(__float_as_uint(y[0]), __float_as_uint(y[1]), __float_as_uint(y[2 % N::kPerVec]),
                                 __float_as_uint(y[3 % N::kPerVec]));
      if (j < nv) stg_stream(orow + j, o);
    }
  }
}

template <int DT>
bool launch_row_kernel(const void* w, void* out, int64_t rows, int64_t cols, int w_bits, cudaStream_t st) {
  const int per = 16 / Num<DT>::kBytes;
  if (cols % per != 0 || ((uintptr_t)w & 15) != 0 || ((uintptr_t)out & 15) != 0) return false;
  const int64_t nvec = cols / per;
  int group = 32, lg = 5;
  while (group < 1024 && (nvec + group - 1) / group > 8) {
    group <<= 1;
    ++lg;
  }
  const int64_t iters = (nvec + group - 1) / group;
  if (iters > 8) return false;
  const int block = group < 256 ? 256 : group;
  const int64_t grid64 = (rows + block / group - 1) / (block / group);
  if (grid64 > 0x7fffffffLL) return false;
  const dim3 g((unsigned)grid64), b((unsigned)block);
#define QAT_LB(IT) (void)launch_pdl(lowbit_row_kernel<DT, IT>, g, b, 0, st, w, out, rows, nvec, cols, lg, w_bits)
  switch ((int)iters) {
    case 1: QAT_LB(1); break;
    case 2: QAT_LB(2); break;
    case 3: QAT_LB(3); break;
    case 4: QAT_LB(4); break;
    case 5: QAT_LB(5); break;
    case 6: QAT_LB(6); break;
    case 7: QAT_LB(7); break;
    default: QAT_LB(8); break;
  }
#undef QAT_LB
  return true;
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
  if (!layerwise) {   // one pass with the row in registers whenever the layout allows
    const bool done = dtype == QAT_F32 ? launch_row_kernel<QAT_F32>(w, w_eff, rows, cols, w_bits, st)
                                       : launch_row_kernel<QAT_BF16>(w, w_eff, rows, cols, w_bits, st);
    if (done) {
      QAT_CHECK_LAUNCH("lowbit_row_kernel");
      return QAT_OK;
    }
  }
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
