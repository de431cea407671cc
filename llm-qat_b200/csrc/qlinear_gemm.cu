// qlinear_gemm.cu — K4: the QuantizeLinear contraction on the integer grid.
//
// Replaces /root/reference/models/utils_quant.py:250  F.linear(x_q, W_q)  (a
// cuBLAS bf16 GEMM over two *dequantized* tensors) with
//     out[t, n] = (sum_k qx[t,k] * qw[n,k]) * (1/ex[t]) * (1/ew[n])
// where qx / qw are the int8 codes and ex / ew the dequant divisors that K1
// emits in its codes-only mode.  The integer dot product is exact (s32
// accumulators), so the result is at least as accurate as the reference's.
//
// sm_100a design
//   * persistent grid, one CTA per SM, static round-robin over 128x256 tiles;
//   * warp 0   : TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle, 4 stages)
//   * warp 1   : MMA issuer    (one thread, tcgen05.mma.cta_group::1.kind::i8,
//                               M128 x N256 x K32, s32 accumulators in TMEM)
//   * warp 2   : TMEM allocator (512 columns = 2 accumulator stages)
//   * warps 4-7: epilogue      (tcgen05.ld 32x32b.x32 -> I2F -> x row/col
//                               scales -> bf16/fp32 -> 16-byte global stores),
//                               overlapped with the next tile's MMAs.
//   * smem: 4 x (16 KB A + 32 KB B) + column-scale staging = ~195 KB.
// Tensor-bound: 2*T*N*K integer ops; see DESIGN.md "K4".
#include <cuda.h>

#include <cstdio>

#include <mutex>

#include "common.cuh"

namespace qat {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 128;  // bytes == int8 elements: one 128B swizzle row
constexpr int UMMA_K = 32;    // elements per tcgen05.mma.kind::i8
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * BLOCK_N;  // 512
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;  // warps 4..7; (warp % 4) selects the TMEM lane quadrant
constexpr int kEpiThreads = 128;

constexpr uint32_t kABytes = BLOCK_M * BLOCK_K;  // 16 KB
constexpr uint32_t kBBytes = BLOCK_N * BLOCK_K;  // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;

struct SmemLayout {
  // offsets from the 1024-byte aligned base
  static __host__ __device__ constexpr uint32_t a(int s) { return (uint32_t)s * kStageBytes; }
  static __host__ __device__ constexpr uint32_t b(int s) { return (uint32_t)s * kStageBytes + kABytes; }
  static constexpr uint32_t colscale = kStages * kStageBytes;             // 2 x 256 floats
  static constexpr uint32_t bars = colscale + kAccStages * BLOCK_N * 4;   // 8-byte aligned
  static __host__ __device__ constexpr uint32_t full(int s) { return bars + 8u * s; }
  static __host__ __device__ constexpr uint32_t empty(int s) { return bars + 8u * (kStages + s); }
  static __host__ __device__ constexpr uint32_t tfull(int a) { return bars + 8u * (2 * kStages + a); }
  static __host__ __device__ constexpr uint32_t tempty(int a) { return bars + 8u * (2 * kStages + kAccStages + a); }
  static constexpr uint32_t tmem_ptr = bars + 8u * (2 * kStages + 2 * kAccStages);
  static constexpr uint32_t total = tmem_ptr + 16;
};
constexpr uint32_t kSmemBytes = SmemLayout::total + 1024;  // slack for manual 1024B alignment

// ---- PTX wrappers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch failure the host reports)
// instead of hanging the GPU.  ~4 s at 2 GHz; never reached in a correct run.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) {
      printf("qat_qlinear: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> s32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile whose rows are 128 bytes (one swizzle
// atom wide): 8-row groups are 1024 bytes apart (SBO); LBO is unused for
// swizzled K-major layouts (encoded 1, as CUTLASS does); version = 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);   // [0,14)  start address >> 4
  d |= (uint64_t)1 << 16;                      // [16,30) leading byte offset >> 4
  d |= (uint64_t)(1024u >> 4) << 32;           // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                      // [46,48) descriptor version
  d |= (uint64_t)2 << 61;                      // [61,64) SWIZZLE_128B
  return d;
}
// kind::i8 instruction descriptor: s32 accumulate, signed int8 A and B, both K-major
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n) {
  return (2u << 4)                  // c_format  = S32
         | (1u << 7) | (1u << 10)   // a_format = b_format = INT8 (signed)
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct GemmParams {
  const float* ex;  // [T] dequant divisors of the activation rows
  const float* ew;  // [N] dequant divisors of the weight rows
  void* out;
  int64_t T, N, K;
  int m_blocks, n_blocks, k_blocks;
};

template <int OUT_DT>
__device__ __forceinline__ void store_chunk(const GemmParams& p, int64_t row, int64_t col0,
                                            const uint32_t (&acc)[32], float row_scale,
                                            const float* colscale /* smem, 32 entries */, bool vec_ok) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j)
    v[j] = __fmul_rn(__fmul_rn((float)(int32_t)acc[j], row_scale), colscale[j]);
  if (OUT_DT == QAT_BF16) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.N + col0;
    if (vec_ok && col0 + 32 <= p.N) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 o;
        o.x = pack_bf16x2(v[j], v[j + 1]);
        o.y = pack_bf16x2(v[j + 2], v[j + 3]);
        o.z = pack_bf16x2(v[j + 4], v[j + 5]);
        o.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(dst + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) dst[j] = __float2bfloat16_rn(v[j]);
    }
  } else {
    float* dst = reinterpret_cast<float*>(p.out) + row * p.N + col0;
    if (vec_ok && col0 + 32 <= p.N) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) dst[j] = v[j];
    }
  }
}

template <int OUT_DT>
__global__ void __launch_bounds__(kThreads, 1)
qlinear_i8_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.m_blocks * p.n_blocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(base + SmemLayout::full(s), 1);
      mbar_init(base + SmemLayout::empty(s), 1);
    }
#pragma unroll
    for (int a = 0; a < kAccStages; ++a) {
      mbar_init(base + SmemLayout::tfull(a), 1);
      mbar_init(base + SmemLayout::tempty(a), kEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     base + SmemLayout::tmem_ptr),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + SmemLayout::tmem_ptr);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % p.m_blocks, n_blk = tile / p.m_blocks;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(base + SmemLayout::empty(stage), phase ^ 1u);
          const uint32_t full = base + SmemLayout::full(stage);
          mbar_expect_tx(full, kStageBytes);
          tma_load_2d(base + SmemLayout::a(stage), &map_a, full, kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(base + SmemLayout::b(stage), &map_b, full, kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_i8(BLOCK_M, BLOCK_N);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(base + SmemLayout::tempty(acc), acc_phase ^ 1u);  // epilogue drained this stage
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(base + SmemLayout::full(stage), phase);
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc(base + SmemLayout::a(stage));
          const uint64_t bdesc = make_smem_desc(base + SmemLayout::b(stage));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes inside the 128B swizzle atom: +2 in the (>>4) address field
            umma_i8(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(base + SmemLayout::empty(stage));  // frees the smem stage when the MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(base + SmemLayout::tfull(acc));  // accumulator complete -> epilogue
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue (4 warps, 128 threads) =====================
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..127
    const bool vec_ok = (OUT_DT == QAT_BF16) ? (p.N % 8 == 0) : (p.N % 4 == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % p.m_blocks, n_blk = tile / p.m_blocks;
      const int64_t row = (int64_t)m_blk * BLOCK_M + quad * 32 + lane;
      const int64_t col_base = (int64_t)n_blk * BLOCK_N;
      float* colscale = reinterpret_cast<float*>(base_ptr + SmemLayout::colscale) + acc * BLOCK_N;
      // stage 1/ew for this tile's 256 columns (2 per thread) while the MMAs run
#pragma unroll
      for (int j = et; j < BLOCK_N; j += kEpiThreads) {
        const int64_t c = col_base + j;
        colscale[j] = (c < p.N) ? __frcp_rn(p.ew[c]) : 0.f;
      }
      const float row_scale = (row < p.T) ? __frcp_rn(p.ex[row]) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // colscale visible to the 4 warps
      mbar_wait(base + SmemLayout::tfull(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (row < p.T && col_base + c * 32 < p.N)
          store_chunk<OUT_DT>(p, row, col_base + c * 32, r, row_scale, colscale + c * 32, vec_ok);
      }
      tcgen05_fence_before();
      mbar_arrive(base + SmemLayout::tempty(acc));
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols)
                 : "memory");
  }
}

// ---- host: tensor maps --------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                              CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

// [rows, K] int8 row-major -> 2-D map, box = {128 bytes of K, box_rows}, 128B swizzle
int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int box_rows) {
  EncodeFn enc = get_encode();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return QAT_ERR_UNSUPPORTED;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K};  // bytes between rows (int8)
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for [%lld, %lld] int8", (int)r, (long long)rows,
              (long long)K);
    return QAT_ERR_BAD_ARG;
  }
  return QAT_OK;
}

template <int OUT_DT>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
  static bool attr_set = false;  // per instantiation; benign race (idempotent)
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(qlinear_i8_kernel<OUT_DT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemBytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(qlinear_i8_kernel)");
    attr_set = true;
  }
  const int tiles = p.m_blocks * p.n_blocks;
  int grid = num_sms();
  if (grid > tiles) grid = tiles;
  qlinear_i8_kernel<OUT_DT><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, p);
  QAT_CHECK_LAUNCH("qlinear_i8_kernel");
  return QAT_OK;
}

}  // namespace
}  // namespace qat

extern "C" int qat_qlinear_i8_fwd(const int8_t* qx, const int8_t* qw, const float* ex, const float* ew,
                                  void* out, int64_t T, int64_t N, int64_t K, int out_dtype, void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(out_dtype == QAT_F32 || out_dtype == QAT_BF16, "out_dtype must be QAT_F32 or QAT_BF16");
  QAT_CHECK_ARG(T >= 0 && N >= 0 && K > 0, "bad GEMM shape [%lld, %lld, %lld]", (long long)T, (long long)N,
                (long long)K);
  if (T == 0 || N == 0) return QAT_OK;
  QAT_CHECK_ARG(qx && qw && ex && ew && out, "NULL operand");
  QAT_CHECK_ARG(K % 16 == 0, "K must be a multiple of 16 (TMA row pitch), got %lld", (long long)K);
  QAT_CHECK_ARG(((uintptr_t)qx & 15) == 0 && ((uintptr_t)qw & 15) == 0, "code pointers must be 16-byte aligned");
  QAT_CHECK_ARG(((uintptr_t)out & 15) == 0, "out must be 16-byte aligned");
  QAT_CHECK_ARG(T < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "dimension too large");
  CUtensorMap ma, mb;
  int rc = make_map(&ma, qx, T, K, BLOCK_M);
  if (rc != QAT_OK) return rc;
  rc = make_map(&mb, qw, N, K, BLOCK_N);
  if (rc != QAT_OK) return rc;
  GemmParams p{};
  p.ex = ex;
  p.ew = ew;
  p.out = out;
  p.T = T;
  p.N = N;
  p.K = K;
  p.m_blocks = (int)((T + BLOCK_M - 1) / BLOCK_M);
  p.n_blocks = (int)((N + BLOCK_N - 1) / BLOCK_N);
  p.k_blocks = (int)((K + BLOCK_K - 1) / BLOCK_K);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_dtype == QAT_BF16) return launch<QAT_BF16>(ma, mb, p, st);
  return launch<QAT_F32>(ma, mb, p, st);
}

// One call for QuantizeLinear.forward's main path (utils_quant.py:197-201,244-250):
// K1 (codes-only) on the activations, K1 on the weights, then the tcgen05 GEMM.
// `reuse_x` / `reuse_w` skip a quantization whose outputs the caller still holds
// (q/k/v share one input; weights only change at optimizer steps).
extern "C" int qat_qlinear_fused_fwd(const void* x, const void* w, void* out, int8_t* qx, float* ex,
                                     uint8_t* mx, int8_t* qw, float* ew, uint8_t* mw, int64_t T, int64_t N,
                                     int64_t K, int dtype, int a_bits, int w_bits, float clip_lo,
                                     float clip_hi, int reuse_x, int reuse_w, void* stream) {
  using namespace qat;
  QAT_CHECK_ARG(qx && ex && qw && ew, "code / scale buffers must be provided");
  QAT_CHECK_ARG(a_bits >= 2 && a_bits <= 8 && w_bits >= 2 && w_bits <= 8, "int8 grid needs 2 <= bits <= 8");
  if (!reuse_x) {
    int rc = qat_sym_fwd(x, nullptr, qx, QAT_CODES_I8, nullptr, ex, mx, clip_lo, clip_hi, T, K, dtype, a_bits,
                         nullptr, 0, stream);
    if (rc != QAT_OK) return rc;
  }
  if (!reuse_w) {
    int rc = qat_sym_fwd(w, nullptr, qw, QAT_CODES_I8, nullptr, ew, mw, clip_lo, clip_hi, N, K, dtype, w_bits,
                         nullptr, 0, stream);
    if (rc != QAT_OK) return rc;
  }
  return qat_qlinear_i8_fwd(qx, qw, ex, ew, out, T, N, K, dtype, stream);
}
