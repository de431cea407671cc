// kd_loss.cu — the logit-distillation loss of /root/reference/utils/kd_trainer.py:42-48
//     loss = KLDivLoss(batchmean)( log_softmax(student, dim=2), softmax(teacher, dim=2) )
//          = (1 / B) * sum_{b,t,v} p_t (log p_t - log q_s)
// which eager PyTorch runs as log_softmax + softmax + kl_div (+ their three backward kernels) over
// fp32 [B, S, V] tensors (V = 32000: 262 MB each at S = 2048).  Here: one pass per direction.
//   forward : one CTA per (b, t) row reads the student and teacher logits once (online soft-max:
//             running max / sum / weighted sum per thread, merged across the CTA), writes the row's
//             KL and four row statistics; a second single-CTA kernel adds the rows in a fixed order
//             (deterministic) and scales by 1 / B.
//   backward: d loss / d student = (softmax(student) - softmax(teacher)) * grad / B, one read of both
//             logits + one write, from the saved row statistics; `grad` is read from device memory.
// HBM-bound: forward 2e B/elem, backward 3e B/elem (e = bytes per logit).
#define QAT_PDL_FAMILY 8   // bit of QAT_B200_PDL_MASK (common.cuh)
#include "common.cuh"

namespace qat {
namespace {

constexpr int kThreads = 256;

struct Online {   // soft-max state of a set of logits: max m, z = sum e^(x - m), a = sum e^(x - m) * w
  float m, z, a;
};
__device__ __forceinline__ Online merge(const Online& x, const Online& y) {
  Online r;
  r.m = fmaxf(x.m, y.m);
  const float fx = (x.m == -INFINITY) ? 0.f : __expf(x.m - r.m);
  const float fy = (y.m == -INFINITY) ? 0.f : __expf(y.m - r.m);
  r.z = x.z * fx + y.z * fy;
  r.a = x.a * fx + y.a * fy;
  return r;
}
__device__ __forceinline__ void push(Online& st, float x, float w) {
  if (x > st.m) {
    const float f = (st.m == -INFINITY) ? 0.f : __expf(st.m - x);
    st.z *= f;
    st.a *= f;
    st.m = x;
  }
  const float e = __expf(x - st.m);
  st.z += e;
  st.a += e * w;
}
__device__ __forceinline__ Online block_merge(Online v, Online* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Online y;
    y.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    y.z = __shfl_xor_sync(0xffffffffu, v.z, o);
    y.a = __shfl_xor_sync(0xffffffffu, v.a, o);
    v = merge(v, y);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  Online r = sm[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) r = merge(r, sm[w]);
  __syncthreads();
  return r;
}

template <int DT>
__device__ __forceinline__ void load8(const void* base, int64_t vec, float (&x)[8]) {
  // DT == QAT_BF16: one 16-byte vector = 8 logits; QAT_F32: two vectors
  if (DT == QAT_BF16) {
    const uint4 v = ldg_stream(reinterpret_cast<const uint4*>(base) + vec);
    x[0] = bf16lo(v.x); x[1] = bf16hi(v.x); x[2] = bf16lo(v.y); x[3] = bf16hi(v.y);
    x[4] = bf16lo(v.z); x[5] = bf16hi(v.z); x[6] = bf16lo(v.w); x[7] = bf16hi(v.w);
  } else {
    const uint4 a = ldg_stream(reinterpret_cast<const uint4*>(base) + 2 * vec);
    const uint4 b = ldg_stream(reinterpret_cast<const uint4*>(base) + 2 * vec + 1);
    x[0] = __uint_as_float(a.x); x[1] = __uint_as_float(a.y); x[2] = __uint_as_float(a.z); x[3] = __uint_as_float(a.w);
    x[4] = __uint_as_float(b.x); x[5] = __uint_as_float(b.y); x[6] = __uint_as_float(b.z); x[7] = __uint_as_float(b.w);
  }
}
template <int DT>
__device__ __forceinline__ float load1(const void* base, int64_t i) {
  if (DT == QAT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
  return reinterpret_cast<const float*>(base)[i];
}

// row statistics: {max_s, 1/Z_s, max_t, 1/Z_t}
template <int DT>
__global__ void __launch_bounds__(kThreads) kd_fwd_kernel(const void* __restrict__ student,
                                                          const void* __restrict__ teacher, float* __restrict__ row_kl,
                                                          float4* __restrict__ row_stat, int64_t V, int vec_ok) {
  __shared__ Online sm[kThreads / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int64_t row = blockIdx.x;
  constexpr int kB = DT == QAT_BF16 ? 2 : 4;
  const char* s_row = reinterpret_cast<const char*>(student) + row * V * kB;
  const char* t_row = reinterpret_cast<const char*>(teacher) + row * V * kB;
  Online ss{-INFINITY, 0.f, 0.f}, tt{-INFINITY, 0.f, 0.f};
  const int64_t nvec = vec_ok ? V / 8 : 0;
  for (int64_t v = threadIdx.x; v < nvec; v += kThreads) {
    float xs[8], xt[8];
    load8<DT>(s_row, v, xs);
    load8<DT>(t_row, v, xt);
    // one rescale per vector, not per element
    float ms = xs[0], mt = xt[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      ms = fmaxf(ms, xs[i]);
      mt = fmaxf(mt, xt[i]);
    }
    if (ms > ss.m) {
      const float f = (ss.m == -INFINITY) ? 0.f : __expf(ss.m - ms);
      ss.z *= f;
      ss.m = ms;
    }
    if (mt > tt.m) {
      const float f = (tt.m == -INFINITY) ? 0.f : __expf(tt.m - mt);
      tt.z *= f;
      tt.a *= f;
      tt.m = mt;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ss.z += __expf(xs[i] - ss.m);
      const float e = __expf(xt[i] - tt.m);
      tt.z += e;
      tt.a += e * (xt[i] - xs[i]);
    }
  }
  for (int64_t i = nvec * 8 + threadIdx.x; i < V; i += kThreads) {
    const float xs = load1<DT>(s_row, i), xt = load1<DT>(t_row, i);
    push(ss, xs, 0.f);
    push(tt, xt, xt - xs);
  }
  const Online S = block_merge(ss, sm);
  const Online T = block_merge(tt, sm);
  if (threadIdx.x == 0) {
    // sum_v p_t [(t - m_t - log Z_t) - (s - m_s - log Z_s)] = a_t / Z_t - m_t - log Z_t + m_s + log Z_s
    row_kl[row] = T.a / T.z - T.m - __logf(T.z) + S.m + __logf(S.z);
    row_stat[row] = make_float4(S.m, 1.0f / S.z, T.m, 1.0f / T.z);
  }
}

// fixed-order sum of the row terms in fp64, x scale
__global__ void __launch_bounds__(1024) kd_reduce_kernel(const float* __restrict__ row_kl, int64_t rows, float scale,
                                                         float* __restrict__ loss) {
  __shared__ double sm[32];
  pdl_wait();
  pdl_launch_dependents();
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < rows; i += 1024) acc += (double)row_kl[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += sm[w];
    *loss = (float)(t * (double)scale);
  }
}

template <int DT>
__global__ void __launch_bounds__(kThreads) kd_bwd_kernel(const void* __restrict__ student,
                                                          const void* __restrict__ teacher,
                                                          const float4* __restrict__ row_stat,
                                                          const float* __restrict__ grad_loss, float scale,
                                                          void* __restrict__ grad_student, int64_t V, int vec_ok) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t row = blockIdx.x;
  constexpr int kB = DT == QAT_BF16 ? 2 : 4;
  const char* s_row = reinterpret_cast<const char*>(student) + row * V * kB;
  const char* t_row = reinterpret_cast<const char*>(teacher) + row * V * kB;
  char* g_row = reinterpret_cast<char*>(grad_student) + row * V * kB;
  const float4 st = row_stat[row];
  const float g = (grad_loss != nullptr ? *grad_loss : 1.0f) * scale;
  const float qs = st.y * g, pt = st.w * g;
  const int64_t nvec = vec_ok ? V / 8 : 0;
  for (int64_t v = threadIdx.x; v < nvec; v += kThreads) {
    float xs[8], xt[8], o[8];
    load8<DT>(s_row, v, xs);
    load8<DT>(t_row, v, xt);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __expf(xs[i] - st.x) * qs - __expf(xt[i] - st.z) * pt;
    if (DT == QAT_BF16) {
      stg_stream(reinterpret_cast<uint4*>(g_row) + v,
                 make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                            pack_bf16x2(o[6], o[7])));
    } else {
      stg_stream(reinterpret_cast<uint4*>(g_row) + 2 * v,
                 make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3])));
      stg_stream(reinterpret_cast<uint4*>(g_row) + 2 * v + 1,
                 make_uint4(__float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7])));
    }
  }
  for (int64_t i = nvec * 8 + threadIdx.x; i < V; i += kThreads) {
    const float xs = load1<DT>(s_row, i), xt = load1<DT>(t_row, i);
    const float o = __expf(xs - st.x) * qs - __expf(xt - st.z) * pt;
    if (DT == QAT_BF16) reinterpret_cast<__nv_bfloat16*>(g_row)[i] = __float2bfloat16_rn(o);
    else reinterpret_cast<float*>(g_row)[i] = o;
  }
}

bool rows_vec_ok(const void* a, const void* b, const void* c, int64_t V, int dtype) {
  const int64_t row_bytes = V * (dtype == QAT_BF16 ? 2 : 4);
  return V % 8 == 0 && row_bytes % 16 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

}  // namespace
}  // namespace qat

extern "C" int qat_kd_loss_fwd(const void* student, const void* teacher, float* loss, float* row_kl,
                               float* row_stat, int64_t rows, int64_t V, int64_t batch, int dtype, void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(dtype == QAT_F32 || dtype == QAT_BF16, "dtype must be QAT_F32 or QAT_BF16 (got %d)", dtype);
  QAT_CHECK_ARG(rows > 0 && V > 0 && batch > 0, "bad shape rows=%lld V=%lld batch=%lld", (long long)rows, (long long)V,
                (long long)batch);
  QAT_CHECK_ARG(student && teacher && loss && row_kl && row_stat, "NULL operand");
  QAT_CHECK_ARG(((uintptr_t)row_stat & 15) == 0, "row_stat must be 16-byte aligned");
  QAT_CHECK_ARG(rows < (1ll << 31), "too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int vec_ok = rows_vec_ok(student, teacher, student, V, dtype) ? 1 : 0;
  cudaError_t e;
  if (dtype == QAT_BF16)
    e = launch_pdl(kd_fwd_kernel<QAT_BF16>, dim3((unsigned)rows), dim3(kThreads), 0, st, student, teacher, row_kl,
                   reinterpret_cast<float4*>(row_stat), V, vec_ok);
  else
    e = launch_pdl(kd_fwd_kernel<QAT_F32>, dim3((unsigned)rows), dim3(kThreads), 0, st, student, teacher, row_kl,
                   reinterpret_cast<float4*>(row_stat), V, vec_ok);
  if (e != cudaSuccess) return cuda_fail(e, "kd_fwd_kernel launch");
  QAT_CHECK_LAUNCH("kd_fwd_kernel");
  e = launch_pdl(kd_reduce_kernel, dim3(1), dim3(1024), 0, st, (const float*)row_kl, rows, 1.0f / (float)batch, loss);
  if (e != cudaSuccess) return cuda_fail(e, "kd_reduce_kernel launch");
  QAT_CHECK_LAUNCH("kd_reduce_kernel");
  return QAT_OK;
}

extern "C" int qat_kd_loss_bwd(const void* student, const void* teacher, const float* row_stat,
                               const float* grad_loss, void* grad_student, int64_t rows, int64_t V, int64_t batch,
                               int dtype, void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(dtype == QAT_F32 || dtype == QAT_BF16, "dtype must be QAT_F32 or QAT_BF16 (got %d)", dtype);
  QAT_CHECK_ARG(rows > 0 && V > 0 && batch > 0, "bad shape");
  QAT_CHECK_ARG(student && teacher && row_stat && grad_loss && grad_student, "NULL operand");
  QAT_CHECK_ARG(rows < (1ll << 31), "too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int vec_ok = rows_vec_ok(student, teacher, grad_student, V, dtype) ? 1 : 0;
  cudaError_t e;
  if (dtype == QAT_BF16)
    e = launch_pdl(kd_bwd_kernel<QAT_BF16>, dim3((unsigned)rows), dim3(kThreads), 0, st, student, teacher,
                   reinterpret_cast<const float4*>(row_stat), grad_loss, 1.0f / (float)batch, grad_student, V, vec_ok);
  else
    e = launch_pdl(kd_bwd_kernel<QAT_F32>, dim3((unsigned)rows), dim3(kThreads), 0, st, student, teacher,
                   reinterpret_cast<const float4*>(row_stat), grad_loss, 1.0f / (float)batch, grad_student, V, vec_ok);
  if (e != cudaSuccess) return cuda_fail(e, "kd_bwd_kernel launch");
  QAT_CHECK_LAUNCH("kd_bwd_kernel");
  return QAT_OK;
}
