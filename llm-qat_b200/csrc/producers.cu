// producers.cu — the producers of QuantizeLinear's inputs, each fused with the per-token
// fake-quantization of what it produces, and the K/V fake-quant call site
// fused with RoPE:
//   rmsnorm_feed     modeling_llama_quant.py:112-129 (LlamaRMSNorm)  -> y + int8 codes / divisors / STE mask of y
//   swiglu_feed      modeling_llama_quant.py:235  act_fn(gate) * up  -> act + its codes / divisors / mask
//   qkv_prep         modeling_llama_quant.py:320-341: SymQuantizer.apply on K and V (per token over all
//                    heads' channels, utils_quant.py:53-72) then rotary embedding of Q and K (:174-196)
// plus their backward kernels.  The codes are exactly what qat_sym_fwd's GEMM-feed mode would emit for
// the same tensor (same SymScale chain from common.cuh, same packing), so QuantizeLinear consumes them
// as if it had quantized its input itself.  All HBM-bound, one CTA (256 threads) per token row, the row
// held in registers between the reductions and the store.
#define QAT_PDL_FAMILY 7   // bit of QAT_B200_PDL_MASK (common.cuh)
#include "common.cuh"

namespace qat {
namespace {

constexpr int kThreads = 256;
constexpr uint32_t kFull = 0xffffffffu;

__device__ __forceinline__ float block_sum(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) r += sm[w];
  __syncthreads();
  return r;
}
__device__ __forceinline__ uint32_t block_max_u32(uint32_t v, uint32_t* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFull, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  uint32_t r = 0u;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) r = max(r, sm[w]);
  __syncthreads();
  return r;
}
// max |x| over the 8 bf16 values of a vector, as a 15-bit pattern (NaN sorts above inf => propagates)
__device__ __forceinline__ uint32_t amax_bits8(const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t m = 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    m = max(m, w[k] & 0x7fffu);
    m = max(m, (w[k] >> 16) & 0x7fffu);
  }
  return m;
}
__device__ __forceinline__ float elem(const uint4& v, int i) { return vec_get<QAT_BF16>(v, i); }
__device__ __forceinline__ uint4 pack8(const float (&y)[8]) {
  return make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                    pack_bf16x2(y[6], y[7]));
}

// int8 codes + pass-mask byte of one vector of 8 bf16 values — the arithmetic of quant_vec_feed
// (fakequant.cu) for the symmetric quantizer in dtype DT (QAT_BF16 or QAT_BF16_AMP)
template <int DT>
__device__ __forceinline__ void feed8(const SymScale<DT>& sc, const uint4& v, float lo, float hi, uint2* codes,
                                      uint8_t* mask) {
  float qv[8];
  if (DT == QAT_BF16) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t pw = mul_bf16x2(w[k], sc.s2);
      qv[2 * k] = bf16lo(pw);
      qv[2 * k + 1] = bf16hi(pw);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = __fmul_rn(elem(v, i), sc.s);
  }
  *codes = make_uint2(pack_codes4<true>(qv[0], qv[1], qv[2], qv[3]), pack_codes4<true>(qv[4], qv[5], qv[6], qv[7]));
  uint32_t pass = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float xf = elem(v, i);
    pass |= ((xf >= hi || xf <= lo) ? 0u : 1u) << i;   // utils_quant.py:85-86
  }
  *mask = (uint8_t)pass;
}

struct FeedOut {
  int8_t* codes;    // [rows, cols] (NULL: no feed)
  float* row_e;     // [rows]
  uint8_t* mask;    // packed, bit i%8 of byte i/8 over the flattened tensor (may be NULL)
  float lo, hi;     // clip bounds, rounded to bf16 by the host
  float qmax;       // 2^(bits-1) - 1
};

// ================================================================================================
// RMSNorm (+ feed)
// ================================================================================================
constexpr int kRmsMaxIters = 4;   // hidden <= 256 * 4 * 8 = 8192; kernels are instantiated for the exact count (registers)

template <int DT, int kRmsIters>
__global__ void __launch_bounds__(kThreads) rmsnorm_feed_kernel(const uint4* __restrict__ x,
                                                                const uint4* __restrict__ w, uint4* __restrict__ y,
                                                                float* __restrict__ rstd_out, int nvec, float eps,
                                                                float inv_cols, const FeedOut f) {
  __shared__ float sm_f[kThreads / 32];
  __shared__ uint32_t sm_u[kThreads / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int64_t row = blockIdx.x;
  const uint4* xr = x + row * nvec;
  uint4 v[kRmsIters];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kRmsIters; ++i) {
    const int j = threadIdx.x + i * kThreads;
    v[i] = (j < nvec) ? ldg_stream(xr + j) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a = elem(v[i], e);
      ss = fmaf(a, a, ss);
    }
  }
  const float var = block_sum(ss, sm_f) * inv_cols;          // :122  x.to(fp32).pow(2).mean(-1)
  const float rstd = rsqrtf(var + eps);                      // :123
  if (threadIdx.x == 0 && rstd_out != nullptr) rstd_out[row] = rstd;
  uint32_t amax = 0u;
#pragma unroll
  for (int i = 0; i < kRmsIters; ++i) {
    const int j = threadIdx.x + i * kThreads;
    if (j < nvec) {
      const uint4 wv = __ldg(w + j);
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = Num<QAT_BF16>::fl(elem(v[i], e) * rstd);   // :123 fp32 product, :127 .to(bf16)
        o[e] = elem(wv, e) * xh;                                     // :129 weight * hidden (bf16 x bf16)
      }
      v[i] = pack8(o);
      stg_stream(y + row * nvec + j, v[i]);
      amax = max(amax, amax_bits8(v[i]));
    }
  }
  if (f.codes == nullptr) return;
  amax = block_max_u32(amax, sm_u);
  SymScale<DT> sc;
  sc.derive(__uint_as_float(amax << 16), f.qmax);
  if (threadIdx.x == 0) f.row_e[row] = (amax == 0x7f80u) ? __int_as_float(0x7fc00000) : sc.e;   // see sym_fwd_feed
#pragma unroll
  for (int i = 0; i < kRmsIters; ++i) {
    const int j = threadIdx.x + i * kThreads;
    if (j < nvec) {
      uint2 c;
      uint8_t m;
      feed8<DT>(sc, v[i], f.lo, f.hi, &c, &m);
      reinterpret_cast<uint2*>(f.codes)[row * nvec + j] = c;
      if (f.mask != nullptr) f.mask[row * nvec + j] = m;
    }
  }
}

// backward: gx = rstd * (g*w - xh * mean(g*w*xh)),  gw partial sums over this CTA's rows
constexpr int kRmsBwdRows = 8;
template <int kRmsIters>
__global__ void __launch_bounds__(kThreads) rmsnorm_bwd_kernel(const uint4* __restrict__ g, const uint4* __restrict__ x,
                                                               const uint4* __restrict__ w,
                                                               const float* __restrict__ rstd_in, uint4* __restrict__ gx,
                                                               float* __restrict__ gw_partial, int64_t rows, int nvec,
                                                               float inv_cols) {
  __shared__ float sm_f[kThreads / 32];
  pdl_wait();
  pdl_launch_dependents();
  float gw_acc[kRmsIters][8];
#pragma unroll
  for (int i = 0; i < kRmsIters; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) gw_acc[i][e] = 0.f;
  uint4 wv[kRmsIters];
#pragma unroll
  for (int i = 0; i < kRmsIters; ++i) {
    const int j = threadIdx.x + i * kThreads;
    wv[i] = (j < nvec) ? __ldg(w + j) : make_uint4(0u, 0u, 0u, 0u);
  }
  const int64_t r0 = (int64_t)blockIdx.x * kRmsBwdRows;
  for (int rr = 0; rr < kRmsBwdRows; ++rr) {
    const int64_t row = r0 + rr;
    if (row >= rows) break;   // block-uniform
    const float rstd = rstd_in[row];
    uint4 xv[kRmsIters], gv[kRmsIters];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kRmsIters; ++i) {
      const int j = threadIdx.x + i * kThreads;
      const bool ok = j < nvec;
      xv[i] = ok ? ldg_stream(x + row * nvec + j) : make_uint4(0u, 0u, 0u, 0u);
      gv[i] = ok ? ldg_stream(g + row * nvec + j) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = Num<QAT_BF16>::fl(elem(xv[i], e) * rstd);
        const float ge = elem(gv[i], e);
        gw_acc[i][e] = fmaf(ge, xh, gw_acc[i][e]);
        dot = fmaf(ge * elem(wv[i], e), xh, dot);
      }
    }
    const float mean = block_sum(dot, sm_f) * inv_cols;
#pragma unroll
    for (int i = 0; i < kRmsIters; ++i) {
      const int j = threadIdx.x + i * kThreads;
      if (j < nvec) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = elem(xv[i], e) * rstd;
          o[e] = rstd * (elem(gv[i], e) * elem(wv[i], e) - xh * mean);
        }
        stg_stream(gx + row * nvec + j, pack8(o));
      }
    }
  }
  float* part = gw_partial + (int64_t)blockIdx.x * nvec * 8;
#pragma unroll
  for (int i = 0; i < kRmsIters; ++i) {
    const int j = threadIdx.x + i * kThreads;
    if (j < nvec) {
      *reinterpret_cast<float4*>(part + (int64_t)j * 8) = make_float4(gw_acc[i][0], gw_acc[i][1], gw_acc[i][2], gw_acc[i][3]);
      *reinterpret_cast<float4*>(part + (int64_t)j * 8 + 4) = make_float4(gw_acc[i][4], gw_acc[i][5], gw_acc[i][6], gw_acc[i][7]);
    }
  }
}
// gw[c] = sum over partial blocks (fixed order), stored as bf16
__global__ void __launch_bounds__(kThreads) colsum_kernel(const float* __restrict__ partial, int nblk, int cols,
                                                          __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int c = blockIdx.x * kThreads + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int b = 0; b < nblk; ++b) acc += partial[(int64_t)b * cols + c];
  out[c] = __float2bfloat16_rn(acc);
}

// ================================================================================================
// SiLU(gate) * up (+ feed)
// ================================================================================================
constexpr int kActMaxIters = 8;   // intermediate <= 256 * 8 * 8 = 16384

__device__ __forceinline__ float silu_f(float xv) { return __fdividef(xv, 1.0f + __expf(-xv)); }

template <int DT, int kActIters>
__global__ void __launch_bounds__(kThreads) swiglu_feed_kernel(const uint4* __restrict__ gate,
                                                               const uint4* __restrict__ up, uint4* __restrict__ act,
                                                               int nvec, const FeedOut f) {
  __shared__ uint32_t sm_u[kThreads / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int64_t row = blockIdx.x;
  uint4 v[kActIters];
  uint32_t amax = 0u;
#pragma unroll
  for (int i = 0; i < kActIters; ++i) {
    const int j = threadIdx.x + i * kThreads;
    if (j < nvec) {
      const uint4 gv = ldg_stream(gate + row * nvec + j);
      const uint4 uv = ldg_stream(up + row * nvec + j);
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        o[e] = Num<QAT_BF16>::fl(silu_f(elem(gv, e))) * elem(uv, e);   // :235 act_fn(gate) (bf16) * up
      v[i] = pack8(o);
      stg_stream(act + row * nvec + j, v[i]);
      amax = max(amax, amax_bits8(v[i]));
    } else {
      v[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (f.codes == nullptr) return;
  amax = block_max_u32(amax, sm_u);
  SymScale<DT> sc;
  sc.derive(__uint_as_float(amax << 16), f.qmax);
  if (threadIdx.x == 0) f.row_e[row] = (amax == 0x7f80u) ? __int_as_float(0x7fc00000) : sc.e;
#pragma unroll
  for (int i = 0; i < kActIters; ++i) {
    const int j = threadIdx.x + i * kThreads;
    if (j < nvec) {
      uint2 c;
      uint8_t m;
      feed8<DT>(sc, v[i], f.lo, f.hi, &c, &m);
      reinterpret_cast<uint2*>(f.codes)[row * nvec + j] = c;
      if (f.mask != nullptr) f.mask[row * nvec + j] = m;
    }
  }
}

// d gate = g * up * silu'(gate),  d up = g * silu(gate)
__global__ void __launch_bounds__(kThreads) swiglu_bwd_kernel(const uint4* __restrict__ g, const uint4* __restrict__ gate,
                                                              const uint4* __restrict__ up, uint4* __restrict__ d_gate,
                                                              uint4* __restrict__ d_up, int64_t nvec_total) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x; j < nvec_total; j += (int64_t)gridDim.x * kThreads) {
    const uint4 gv = ldg_stream(g + j), av = ldg_stream(gate + j), uv = ldg_stream(up + j);
    float dg[8], du[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a = elem(av, e), ge = elem(gv, e);
      const float sig = __fdividef(1.0f, 1.0f + __expf(-a));
      const float sl = Num<QAT_BF16>::fl(a * sig);
      du[e] = ge * sl;
      const float g_sl = Num<QAT_BF16>::fl(ge * elem(uv, e));    // grad wrt the bf16 silu output
      dg[e] = g_sl * (sig * (1.0f + a * (1.0f - sig)));
    }
    stg_stream(d_gate + j, pack8(dg));
    stg_stream(d_up + j, pack8(du));
  }
}

// ================================================================================================
// K/V fake-quant + rotary embedding of Q and K, one launch (modeling_llama_quant.py:320-341)
// ================================================================================================
// One CTA per token: hidden = H * 128 <= 256 * 2 * 8 * ... each thread owns `kQkvIters` vectors of each
// of q, k, v.  Vector j covers elements [8j, 8j+8) of the token's row; inside a head (128 = 16
// vectors) the rotation partner of vector t is vector t ^ 8, i.e. thread (tid ^ 8) of the same warp.
constexpr int kQkvMaxIters = 4;   // hidden <= 8192

template <int DT>
__device__ __forceinline__ float rope_round(float v) { return DT == QAT_BF16 ? Num<QAT_BF16>::fl(v) : v; }

// y = x * cos + rotate_half(x) * sin for the 8 elements of vector `j` (head offset d0 = (j % 16) * 8);
// xp = the partner vector's values (d0 ^ 64).  Plain bf16 tensors round every op to bf16 like eager
// PyTorch; under autocast (DT == QAT_BF16_AMP) q * cos promotes to fp32 and only the matmul's cast
// rounds, i.e. one rounding at the end.
template <int DT>
__device__ __forceinline__ void rope8(const float (&xv)[8], const float (&xp)[8], const float* cs, const float* sn,
                                      int d0, float (&o)[8]) {
  const float sign = (d0 < 64) ? -1.0f : 1.0f;    // rotate_half: (-x2, x1)
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float c = rope_round<DT>(cs[d0 + e]), s = rope_round<DT>(sn[d0 + e]);
    const float a = rope_round<DT>(xv[e] * c);
    const float b = rope_round<DT>((sign * xp[e]) * s);
    o[e] = a + b;   // rounded to bf16 when packed
  }
}

struct QkvParams {
  const uint4* q;
  const uint4* k;
  const uint4* v;
  uint4* q_out;
  uint4* k_out;
  uint4* v_out;
  uint8_t* k_mask;   // packed STE pass-masks of the unquantized K / V (NULL when kv_bits >= 32)
  uint8_t* v_mask;
  const float* cos;  // [max_pos, 128] fp32 tables (LlamaRotaryEmbedding.cos_cached / sin_cached)
  const float* sin;
  const int64_t* pos;  // [tokens] position ids
  int64_t max_pos;     // rows of the tables
  int nvec;            // vectors per token row = H * 16
  int kv_bits;         // >= 32: no K/V fake-quant
  float lo, hi, qmax;
};

// table row of a position id: negative ids count from the end like the reference's `cos[position_ids]`
// (torch indexing); ids past the table — an IndexError there — are clamped so that no launch reads out of bounds
__device__ __forceinline__ int64_t rope_row(int64_t pos, int64_t max_pos) {
  if (pos < 0) pos += max_pos;
  return pos < 0 ? 0 : (pos >= max_pos ? max_pos - 1 : pos);
}

template <int DT, int kQkvIters>
__global__ void __launch_bounds__(kThreads) qkv_prep_kernel(const QkvParams p) {
  __shared__ uint32_t sm_u[kThreads / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int64_t row = blockIdx.x;
  const int64_t trow = rope_row(p.pos[row], p.max_pos);
  const float* cs = p.cos + trow * 128;
  const float* sn = p.sin + trow * 128;
  uint4 kv[kQkvIters], vv[kQkvIters];
  uint32_t kmax = 0u, vmax = 0u;
#pragma unroll
  for (int i = 0; i < kQkvIters; ++i) {
    const int j = threadIdx.x + i * kThreads;
    const bool ok = j < p.nvec;
    kv[i] = ok ? ldg_stream(p.k + row * p.nvec + j) : make_uint4(0u, 0u, 0u, 0u);
    vv[i] = ok ? ldg_stream(p.v + row * p.nvec + j) : make_uint4(0u, 0u, 0u, 0u);
    kmax = max(kmax, amax_bits8(kv[i]));
    vmax = max(vmax, amax_bits8(vv[i]));
  }
  const bool quant = p.kv_bits < 32;
  SymScale<DT> sk, sv;
  if (quant) {
    kmax = block_max_u32(kmax, sm_u);
    vmax = block_max_u32(vmax, sm_u);
    sk.derive(__uint_as_float(kmax << 16), p.qmax);
    sv.derive(__uint_as_float(vmax << 16), p.qmax);
  }
#pragma unroll
  for (int i = 0; i < kQkvIters; ++i) {
    if (i * kThreads >= p.nvec) break;           // block-uniform
    const int j = threadIdx.x + i * kThreads;    // nvec is a multiple of 16, so a warp's 32 vectors are 2 whole heads
    const bool ok = j < p.nvec;
    const int d0 = (j & 15) * 8;
    // ---- V: fake-quant only
    float vq[8], kq[8], qf[8];
    uint32_t vpass = 0, kpass = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xv = elem(vv[i], e), xk = elem(kv[i], e);
      float code;
      vq[e] = quant ? (sv.fast ? sv.template apply<true>(xv, &code) : sv.template apply<false>(xv, &code)) : xv;
      kq[e] = quant ? (sk.fast ? sk.template apply<true>(xk, &code) : sk.template apply<false>(xk, &code)) : xk;
      if (DT == QAT_BF16) {   // the quantizer's output tensor is bf16
        vq[e] = Num<QAT_BF16>::fl(vq[e]);
        kq[e] = Num<QAT_BF16>::fl(kq[e]);
      }
      vpass |= ((xv >= p.hi || xv <= p.lo) ? 0u : 1u) << e;
      kpass |= ((xk >= p.hi || xk <= p.lo) ? 0u : 1u) << e;
    }
    const uint4 qv = ok ? ldg_stream(p.q + row * p.nvec + j) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int e = 0; e < 8; ++e) qf[e] = elem(qv, e);
    // ---- rotation partners live in lane ^ 8
    float kp[8], qp[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      kp[e] = __shfl_xor_sync(kFull, kq[e], 8);
      qp[e] = __shfl_xor_sync(kFull, qf[e], 8);
    }
    if (ok) {
      float ko[8], qo[8];
      rope8<DT>(kq, kp, cs, sn, d0, ko);
      rope8<DT>(qf, qp, cs, sn, d0, qo);
      stg_stream(p.q_out + row * p.nvec + j, pack8(qo));
      stg_stream(p.k_out + row * p.nvec + j, pack8(ko));
      stg_stream(p.v_out + row * p.nvec + j, pack8(vq));
      if (quant && p.k_mask != nullptr) {
        p.k_mask[row * p.nvec + j] = (uint8_t)kpass;
        p.v_mask[row * p.nvec + j] = (uint8_t)vpass;
      }
    }
  }
}

// backward: dq = rope^T(dq_rot); dk = mask_k .* rope^T(dk_rot); dv = mask_v .* dv_q
//   rope^T(dy)[d] = dy[d] cos[d] + dy[d+64] sin[d+64]   (d <  64)
//                 = dy[d] cos[d] - dy[d-64] sin[d-64]   (d >= 64)
template <int kQkvIters>
__global__ void __launch_bounds__(kThreads) qkv_prep_bwd_kernel(const uint4* __restrict__ dq_rot,
                                                                const uint4* __restrict__ dk_rot,
                                                                const uint4* __restrict__ dv_q,
                                                                const uint8_t* __restrict__ k_mask,
                                                                const uint8_t* __restrict__ v_mask,
                                                                const float* __restrict__ cos_t,
                                                                const float* __restrict__ sin_t,
                                                                const int64_t* __restrict__ pos, int64_t max_pos,
                                                                uint4* __restrict__ dq, uint4* __restrict__ dk,
                                                                uint4* __restrict__ dv, int nvec) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t row = blockIdx.x;
  const int64_t trow = rope_row(pos[row], max_pos);
  const float* cs = cos_t + trow * 128;
  const float* sn = sin_t + trow * 128;
#pragma unroll 1
  for (int i = 0; i < kQkvIters; ++i) {
    const int j = threadIdx.x + i * kThreads;
    if (i * kThreads >= nvec) break;   // block-uniform
    const bool ok = j < nvec;
    const int d0 = (j & 15) * 8;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const uint4 a = ok ? ldg_stream(dq_rot + row * nvec + j) : z;
    const uint4 b = ok ? ldg_stream(dk_rot + row * nvec + j) : z;
    const uint4 c = ok ? ldg_stream(dv_q + row * nvec + j) : z;
    const uint32_t km = (ok && k_mask != nullptr) ? k_mask[row * nvec + j] : 0xffu;
    const uint32_t vm = (ok && v_mask != nullptr) ? v_mask[row * nvec + j] : 0xffu;
    float oq[8], ok_[8], ov[8];
    const float sign = (d0 < 64) ? 1.0f : -1.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sp = sn[(d0 ^ 64) + e];                 // sin of the partner position (== sin[d0 + e])
      const float aq = elem(a, e), ak = elem(b, e);
      const float pq = __shfl_xor_sync(kFull, aq, 8), pk = __shfl_xor_sync(kFull, ak, 8);
      oq[e] = aq * cs[d0 + e] + sign * pq * sp;
      const float kk = ak * cs[d0 + e] + sign * pk * sp;
      ok_[e] = ((km >> e) & 1u) ? kk : 0.f;
      ov[e] = ((vm >> e) & 1u) ? elem(c, e) : 0.f;
    }
    if (ok) {
      stg_stream(dq + row * nvec + j, pack8(oq));
      stg_stream(dk + row * nvec + j, pack8(ok_));
      stg_stream(dv + row * nvec + j, pack8(ov));
    }
  }
}

float bf16_round_host(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// vectors per thread for a row of `nvec` 16-byte vectors: the kernels keep the row in registers, so they are
// instantiated for the exact count (LLaMA-7B: 2 for hidden 4096, 6 for 11008) instead of the maximum
inline int iters_for(int nvec) { return (nvec + kThreads - 1) / kThreads; }

#define QAT_DISPATCH_ITERS(IT, MAXIT, CALL)                 \
  switch (IT) {                                             \
    case 1: { constexpr int I = 1; CALL; } break;            \
    case 2: { constexpr int I = 2; CALL; } break;            \
    case 3: { constexpr int I = 3; CALL; } break;            \
    case 4: { constexpr int I = 4; CALL; } break;            \
    case 5: { constexpr int I = (MAXIT) >= 5 ? 5 : 4; CALL; } break; \
    case 6: { constexpr int I = (MAXIT) >= 6 ? 6 : 4; CALL; } break; \
    case 7: { constexpr int I = (MAXIT) >= 7 ? 7 : 4; CALL; } break; \
    default: { constexpr int I = (MAXIT) >= 8 ? 8 : 4; CALL; } break; \
  }

int check_feed(int dtype, int bits, int64_t cols, int max_cols) {
  QAT_CHECK_ARG(dtype == QAT_BF16 || dtype == QAT_BF16_AMP, "dtype must be QAT_BF16 or QAT_BF16_AMP (got %d)", dtype);
  QAT_CHECK_ARG(bits >= 2 && bits <= 8, "int8 feed needs 2 <= bits <= 8 (got %d)", bits);
  QAT_CHECK_ARG(cols > 0 && cols % 8 == 0 && cols <= max_cols, "cols must be a multiple of 8 and <= %d (got %lld)",
                max_cols, (long long)cols);
  return QAT_OK;
}

FeedOut make_feed(int8_t* codes, float* row_e, uint8_t* mask, float lo, float hi, int bits) {
  FeedOut f{};
  f.codes = codes;
  f.row_e = row_e;
  f.mask = mask;
  f.lo = bf16_round_host(lo);
  f.hi = bf16_round_host(hi);
  f.qmax = (float)((1 << (bits - 1)) - 1);
  return f;
}

}  // namespace
}  // namespace qat

extern "C" int qat_rmsnorm_feed_fwd(const void* x, const void* weight, void* y, float* rstd, int8_t* codes,
                                    float* row_e, uint8_t* mask, float clip_lo, float clip_hi, int64_t rows,
                                    int64_t cols, float eps, int dtype, int bits, void* stream) {
  using namespace qat;
  int rc = check_feed(dtype, codes ? bits : 8, cols, kThreads * kRmsMaxIters * 8);
  if (rc != QAT_OK) return rc;
  QAT_CHECK_ARG(rows >= 0, "negative rows");
  if (rows == 0) return QAT_OK;
  QAT_CHECK_ARG(x && weight && y, "NULL operand");
  QAT_CHECK_ARG(codes == nullptr || row_e != nullptr, "row_e is required with codes");
  QAT_CHECK_ARG((((uintptr_t)x | (uintptr_t)weight | (uintptr_t)y | (uintptr_t)codes) & 15) == 0,
                "operands must be 16-byte aligned");
  QAT_CHECK_ARG(rows < (1ll << 31), "too many rows");
  const FeedOut f = make_feed(codes, row_e, mask, clip_lo, clip_hi, bits);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nvec = (int)(cols / 8);
  cudaError_t e = cudaSuccess;
  if (dtype == QAT_BF16) {
    QAT_DISPATCH_ITERS(iters_for(nvec), kRmsMaxIters,
                       e = launch_pdl(rmsnorm_feed_kernel<QAT_BF16, I>, dim3((unsigned)rows), dim3(kThreads), 0, st,
                                      (const uint4*)x, (const uint4*)weight, (uint4*)y, rstd, nvec, eps,
                                      1.0f / (float)cols, f));
  } else {
    QAT_DISPATCH_ITERS(iters_for(nvec), kRmsMaxIters,
                       e = launch_pdl(rmsnorm_feed_kernel<QAT_BF16_AMP, I>, dim3((unsigned)rows), dim3(kThreads), 0, st,
                                      (const uint4*)x, (const uint4*)weight, (uint4*)y, rstd, nvec, eps,
                                      1.0f / (float)cols, f));
  }
  if (e != cudaSuccess) return cuda_fail(e, "rmsnorm_feed_kernel launch");
  QAT_CHECK_LAUNCH("rmsnorm_feed_kernel");
  return QAT_OK;
}

extern "C" size_t qat_rmsnorm_bwd_workspace_bytes(int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  return (size_t)((rows + qat::kRmsBwdRows - 1) / qat::kRmsBwdRows) * (size_t)cols * 4;
}

extern "C" int qat_rmsnorm_bwd(const void* grad_y, const void* x, const void* weight, const float* rstd, void* grad_x,
                               void* grad_weight, void* workspace, size_t workspace_bytes, int64_t rows, int64_t cols,
                               void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(cols > 0 && cols % 8 == 0 && cols <= kThreads * kRmsMaxIters * 8, "unsupported cols %lld", (long long)cols);
  QAT_CHECK_ARG(rows >= 0, "negative rows");
  if (rows == 0) return QAT_OK;
  QAT_CHECK_ARG(grad_y && x && weight && rstd && grad_x && grad_weight && workspace, "NULL operand");
  QAT_CHECK_ARG(workspace_bytes >= qat_rmsnorm_bwd_workspace_bytes(rows, cols), "workspace too small");
  QAT_CHECK_ARG((((uintptr_t)grad_y | (uintptr_t)x | (uintptr_t)weight | (uintptr_t)grad_x | (uintptr_t)workspace) & 15) == 0,
                "operands must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nblk = (int)((rows + kRmsBwdRows - 1) / kRmsBwdRows);
  cudaError_t e = cudaSuccess;
  QAT_DISPATCH_ITERS(iters_for((int)(cols / 8)), kRmsMaxIters,
                     e = launch_pdl(rmsnorm_bwd_kernel<I>, dim3((unsigned)nblk), dim3(kThreads), 0, st,
                                    (const uint4*)grad_y, (const uint4*)x, (const uint4*)weight, rstd, (uint4*)grad_x,
                                    (float*)workspace, rows, (int)(cols / 8), 1.0f / (float)cols));
  if (e != cudaSuccess) return cuda_fail(e, "rmsnorm_bwd_kernel launch");
  QAT_CHECK_LAUNCH("rmsnorm_bwd_kernel");
  e = launch_pdl(colsum_kernel, dim3((unsigned)((cols + kThreads - 1) / kThreads)), dim3(kThreads), 0, st,
                 (const float*)workspace, nblk, (int)cols, (__nv_bfloat16*)grad_weight);
  if (e != cudaSuccess) return cuda_fail(e, "colsum_kernel launch");
  QAT_CHECK_LAUNCH("colsum_kernel");
  return QAT_OK;
}

extern "C" int qat_swiglu_feed_fwd(const void* gate, const void* up, void* act, int8_t* codes, float* row_e,
                                   uint8_t* mask, float clip_lo, float clip_hi, int64_t rows, int64_t cols, int dtype,
                                   int bits, void* stream) {
  using namespace qat;
  int rc = check_feed(dtype, codes ? bits : 8, cols, kThreads * kActMaxIters * 8);
  if (rc != QAT_OK) return rc;
  QAT_CHECK_ARG(rows >= 0, "negative rows");
  if (rows == 0) return QAT_OK;
  QAT_CHECK_ARG(gate && up && act, "NULL operand");
  QAT_CHECK_ARG(codes == nullptr || row_e != nullptr, "row_e is required with codes");
  QAT_CHECK_ARG((((uintptr_t)gate | (uintptr_t)up | (uintptr_t)act | (uintptr_t)codes) & 15) == 0,
                "operands must be 16-byte aligned");
  QAT_CHECK_ARG(rows < (1ll << 31), "too many rows");
  const FeedOut f = make_feed(codes, row_e, mask, clip_lo, clip_hi, bits);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nvec = (int)(cols / 8);
  cudaError_t e = cudaSuccess;
  if (dtype == QAT_BF16) {
    QAT_DISPATCH_ITERS(iters_for(nvec), kActMaxIters,
                       e = launch_pdl(swiglu_feed_kernel<QAT_BF16, I>, dim3((unsigned)rows), dim3(kThreads), 0, st,
                                      (const uint4*)gate, (const uint4*)up, (uint4*)act, nvec, f));
  } else {
    QAT_DISPATCH_ITERS(iters_for(nvec), kActMaxIters,
                       e = launch_pdl(swiglu_feed_kernel<QAT_BF16_AMP, I>, dim3((unsigned)rows), dim3(kThreads), 0, st,
                                      (const uint4*)gate, (const uint4*)up, (uint4*)act, nvec, f));
  }
  if (e != cudaSuccess) return cuda_fail(e, "swiglu_feed_kernel launch");
  QAT_CHECK_LAUNCH("swiglu_feed_kernel");
  return QAT_OK;
}

extern "C" int qat_swiglu_bwd(const void* grad_act, const void* gate, const void* up, void* grad_gate, void* grad_up,
                              int64_t n, void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(n >= 0 && n % 8 == 0, "element count must be a multiple of 8 (got %lld)", (long long)n);
  if (n == 0) return QAT_OK;
  QAT_CHECK_ARG(grad_act && gate && up && grad_gate && grad_up, "NULL operand");
  QAT_CHECK_ARG((((uintptr_t)grad_act | (uintptr_t)gate | (uintptr_t)up | (uintptr_t)grad_gate | (uintptr_t)grad_up) & 15) == 0,
                "operands must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t nvec = n / 8;
  int64_t grid = (nvec + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (grid > cap) grid = cap;
  cudaError_t e = launch_pdl(swiglu_bwd_kernel, dim3((unsigned)grid), dim3(kThreads), 0, st, (const uint4*)grad_act,
                             (const uint4*)gate, (const uint4*)up, (uint4*)grad_gate, (uint4*)grad_up, nvec);
  if (e != cudaSuccess) return cuda_fail(e, "swiglu_bwd_kernel launch");
  QAT_CHECK_LAUNCH("swiglu_bwd_kernel");
  return QAT_OK;
}

extern "C" int qat_qkv_prep_fwd(const void* q, const void* k, const void* v, void* q_out, void* k_out, void* v_out,
                                uint8_t* k_mask, uint8_t* v_mask, const float* cos_table, const float* sin_table,
                                const int64_t* position_ids, int64_t max_pos, int64_t tokens, int heads, int head_dim,
                                int kv_bits, float clip_lo, float clip_hi, int dtype, void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(dtype == QAT_BF16 || dtype == QAT_BF16_AMP, "dtype must be QAT_BF16 or QAT_BF16_AMP (got %d)", dtype);
  QAT_CHECK_ARG(head_dim == 128, "head_dim must be 128 (got %d)", head_dim);
  QAT_CHECK_ARG(heads > 0 && heads * 16 <= kThreads * kQkvMaxIters, "unsupported head count %d", heads);
  QAT_CHECK_ARG(kv_bits >= 2, "kv_bits must be >= 2 (got %d)", kv_bits);
  QAT_CHECK_ARG(tokens >= 0, "negative token count");
  if (tokens == 0) return QAT_OK;
  QAT_CHECK_ARG(max_pos > 0, "max_pos (rows of the cos / sin tables) must be positive (got %lld)", (long long)max_pos);
  QAT_CHECK_ARG(q && k && v && q_out && k_out && v_out && cos_table && sin_table && position_ids, "NULL operand");
  QAT_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)q_out | (uintptr_t)k_out | (uintptr_t)v_out) & 15) == 0,
                "operands must be 16-byte aligned");
  QAT_CHECK_ARG(tokens < (1ll << 31), "too many tokens");
  QkvParams p{};
  p.q = (const uint4*)q;
  p.k = (const uint4*)k;
  p.v = (const uint4*)v;
  p.q_out = (uint4*)q_out;
  p.k_out = (uint4*)k_out;
  p.v_out = (uint4*)v_out;
  p.k_mask = k_mask;
  p.v_mask = v_mask;
  p.cos = cos_table;
  p.sin = sin_table;
  p.pos = position_ids;
  p.max_pos = max_pos;
  p.nvec = heads * 16;
  p.kv_bits = kv_bits;
  p.lo = bf16_round_host(clip_lo);
  p.hi = bf16_round_host(clip_hi);
  p.qmax = kv_bits < 32 ? (float)((1ll << (kv_bits - 1)) - 1) : 0.f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
  if (dtype == QAT_BF16) {
    QAT_DISPATCH_ITERS(iters_for(p.nvec), kQkvMaxIters,
                       e = launch_pdl(qkv_prep_kernel<QAT_BF16, I>, dim3((unsigned)tokens), dim3(kThreads), 0, st, p));
  } else {
    QAT_DISPATCH_ITERS(iters_for(p.nvec), kQkvMaxIters,
                       e = launch_pdl(qkv_prep_kernel<QAT_BF16_AMP, I>, dim3((unsigned)tokens), dim3(kThreads), 0, st, p));
  }
  if (e != cudaSuccess) return cuda_fail(e, "qkv_prep_kernel launch");
  QAT_CHECK_LAUNCH("qkv_prep_kernel");
  return QAT_OK;
}

extern "C" int qat_qkv_prep_bwd(const void* dq_rot, const void* dk_rot, const void* dv_q, const uint8_t* k_mask,
                                const uint8_t* v_mask, const float* cos_table, const float* sin_table,
                                const int64_t* position_ids, int64_t max_pos, void* dq, void* dk, void* dv,
                                int64_t tokens, int heads, int head_dim, void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(head_dim == 128, "head_dim must be 128 (got %d)", head_dim);
  QAT_CHECK_ARG(heads > 0 && heads * 16 <= kThreads * kQkvMaxIters, "unsupported head count %d", heads);
  if (tokens <= 0) return QAT_OK;
  QAT_CHECK_ARG(max_pos > 0, "max_pos (rows of the cos / sin tables) must be positive (got %lld)", (long long)max_pos);
  QAT_CHECK_ARG(dq_rot && dk_rot && dv_q && dq && dk && dv && cos_table && sin_table && position_ids, "NULL operand");
  QAT_CHECK_ARG((((uintptr_t)dq_rot | (uintptr_t)dk_rot | (uintptr_t)dv_q | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) == 0,
                "operands must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
  QAT_DISPATCH_ITERS(iters_for(heads * 16), kQkvMaxIters,
                     e = launch_pdl(qkv_prep_bwd_kernel<I>, dim3((unsigned)tokens), dim3(kThreads), 0, st,
                                    (const uint4*)dq_rot, (const uint4*)dk_rot, (const uint4*)dv_q, k_mask, v_mask,
                                    cos_table, sin_table, position_ids, max_pos, (uint4*)dq, (uint4*)dk, (uint4*)dv,
                                    heads * 16));
  if (e != cudaSuccess) return cuda_fail(e, "qkv_prep_bwd_kernel launch");
  QAT_CHECK_LAUNCH("qkv_prep_bwd_kernel");
  return QAT_OK;
}
