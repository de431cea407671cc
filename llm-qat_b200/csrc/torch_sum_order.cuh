// torch_sum_order.cuh — the summation ORDER of torch's CPU `sum(dim=-1)` over a contiguous fp32 row,
// restated so that `w.abs().mean(dim=-1, keepdim=True)` of the reference's W1 / W2 weight path
// (/root/reference/models/utils_quant.py:205-210, 219-224) can be reproduced bit for bit on the GPU.
//
// What ATen does (aten/src/ATen/native/cpu/SumKernel.cpp, the AVX2 build every x86 server runs — the
// kernel is not compiled for AVX-512, and ATEN_CPU_CAPABILITY=avx2 / avx512 give identical bits here):
// a row of K floats is read as vectors of 8 lanes, four vectors per step ("ILP"), i.e. as
// 32 independent CHAINS — chain c = (vector k of the step, lane l) = k*8 + l owns elements
// step*32 + c.  Each chain is summed by a 4-level cascade: 16 serial adds into level 0, level 0 is
// folded into level 1 after every 16 steps, level 1 into level 2 after every 256, level 2 into 3
// after every 4096; the steps past the last full 16 go into level 0; then level 0 += 1, 2, 3.
// Up to three left-over vectors are added to the chains of k = 0, the four k's are folded
// (k0 + k1) + k2) + k3 per lane, and one scalar accumulator takes the < 8 tail elements followed by
// the 8 lanes in order.  Rows shorter than one vector (K < 8) use the same scheme on scalars
// (4 chains, no lanes).  bf16 rows: torch casts to fp32, sums as above, divides in fp32 and rounds
// the mean to bf16 once (ReduceOps.cpp mean_out) — same order, elements widened on load.
// Every function takes the element loader as a functor (index -> |w| as float) and compiles for the
// host too: tests/test_torch_sum_order.py builds them with g++ and checks them against torch.sum
// itself on the CPU, so the order the kernel uses is pinned without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QAT_HD __host__ __device__ __forceinline__
#else
#define QAT_HD inline
#endif

namespace qat {
namespace tso {

constexpr int kLanes = 8;                 // Vectorized<float>::size() of the AVX2 build
constexpr int kIlp = 4;                   // row_sum's ilp_factor
constexpr int kChains = kLanes * kIlp;    // 32: one warp lane per chain
constexpr int kGroup = 16;                // level_step (level_power 4 for every row below 2^24 steps)

QAT_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}

// steps of one chain: floor(floor(K / 8) / 4)
QAT_HD int64_t chain_steps(int64_t cols) { return (cols / kLanes) / kIlp; }

// level 0 of chain c over the full group `g` (steps 16 g ... 16 g + 15)
template <class Load>
QAT_HD float group_sum(const Load& ld, int64_t g, int c) {
  float a = 0.f;
#pragma unroll
  for (int j = 0; j < kGroup; ++j) a = add_rn(a, ld((g * kGroup + j) * kChains + c));
  return a;
}

// the whole cascade of chain c, given its level-0 group sums (gsum(g)) for the n / 16 full groups
template <class Load, class GroupSum>
QAT_HD float chain_sum(const Load& ld, const GroupSum& gsum, int64_t n, int c) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int64_t groups = n / kGroup;
  for (int64_t g = 0; g < groups; ++g) {
    const int64_t i = (g + 1) * kGroup;
    a1 = add_rn(a1, gsum(g));
    if ((i & 0xf0) == 0) {
      a2 = add_rn(a2, a1);
      a1 = 0.f;
      if ((i & 0xf00) == 0) {
        a3 = add_rn(a3, a2);
        a2 = 0.f;
      }
    }
  }
  for (int64_t i = groups * kGroup; i < n; ++i) a0 = add_rn(a0, ld(i * kChains + c));
  a0 = add_rn(a0, a1);
  a0 = add_rn(a0, a2);
  a0 = add_rn(a0, a3);
  return a0;
}

// from the 32 chain sums to the row sum (one thread): left-over vectors, the fold over k, tail, lanes
template <class Load, class Chain>
QAT_HD float finalize(const Load& ld, const Chain& chain, int64_t cols) {
  const int64_t nvec = cols / kLanes, n = nvec / kIlp;
  float p[kLanes];
#pragma unroll
  for (int l = 0; l < kLanes; ++l) p[l] = chain(l);
  for (int64_t v = n * kIlp; v < nvec; ++v) {
#pragma unroll
    for (int l = 0; l < kLanes; ++l) p[l] = add_rn(p[l], ld(v * kLanes + l));
  }
#pragma unroll
  for (int k = 1; k < kIlp; ++k) {
#pragma unroll
    for (int l = 0; l < kLanes; ++l) p[l] = add_rn(p[l], chain(k * kLanes + l));
  }
  float fin = 0.f;
  for (int64_t e = nvec * kLanes; e < cols; ++e) fin = add_rn(fin, ld(e));
#pragma unroll
  for (int l = 0; l < kLanes; ++l) fin = add_rn(fin, p[l]);
  return fin;
}

// rows shorter than one vector (cols < 8): scalar_inner_sum -> row_sum on scalars, 4 chains
template <class Load>
QAT_HD float short_row_sum(const Load& ld, int64_t cols) {
  float p[kIlp] = {0.f, 0.f, 0.f, 0.f};
  const int64_t n = cols / kIlp;   // 0 or 1
  for (int64_t i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < kIlp; ++k) p[k] = add_rn(p[k], ld(i * kIlp + k));
  }
  for (int64_t e = n * kIlp; e < cols; ++e) p[0] = add_rn(p[0], ld(e));
#pragma unroll
  for (int k = 1; k < kIlp; ++k) p[0] = add_rn(p[0], p[k]);
  return add_rn(0.f, p[0]);
}

// reference composition of the pieces above for one row (what the kernel distributes over a CTA)
template <class Load>
QAT_HD float row_sum_serial(const Load& ld, int64_t cols, float* scratch /* >= 32 + 32 * (steps / 16) floats */) {
  if (cols < kLanes) return short_row_sum(ld, cols);
  const int64_t n = chain_steps(cols), groups = n / kGroup;
  float* chain = scratch;
  float* gs = scratch + kChains;
  for (int64_t g = 0; g < groups; ++g)
    for (int c = 0; c < kChains; ++c) gs[g * kChains + c] = group_sum(ld, g, c);
  for (int c = 0; c < kChains; ++c)
    chain[c] = chain_sum(ld, [&](int64_t g) { return gs[g * kChains + c]; }, n, c);
  return finalize(ld, [&](int c) { return chain[c]; }, cols);
}

}  // namespace tso
}  // namespace qat
