// host.cu — host-buffer entry points: the same fake-quant forward + STE
// backward, called with HOST pointers (pinned for full speed).  This is the
// end-to-end shape of the reference's call when tensors live on the CPU
// (SymQuantizer.apply / .backward, utils_quant.py:37-87): copy in, run, copy out.
//
// Rows are independent, so the tensor is cut into row chunks that flow through
// a three-stage pipeline on three streams — H2D(x, g) | K1/K2 + K3 | D2H(y, gx)
// — overlapping both PCIe directions with the kernels.  Everything is ordered
// after prior work on the caller's stream and the caller's stream waits for the
// last D2H, so stream semantics are those of a single asynchronous call.
#include <vector>

#include "common.cuh"

namespace qat {
namespace {

struct Pipe {
  cudaStream_t in = nullptr, out = nullptr;
  std::vector<cudaEvent_t> ev;
  int device = -1;
  cudaEvent_t event(size_t i) {
    while (ev.size() <= i) {
      cudaEvent_t e;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      ev.push_back(e);
    }
    return ev[i];
  }
};

// one pipe per host thread and device; streams/events are the only persistent
// resources the library creates (include/qat_b200.h "Ownership").
Pipe* get_pipe() {
  static thread_local std::vector<Pipe> pipes;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  for (auto& p : pipes)
    if (p.device == dev) return &p;
  Pipe p;
  p.device = dev;
  if (cudaStreamCreateWithFlags(&p.in, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  if (cudaStreamCreateWithFlags(&p.out, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  pipes.push_back(p);
  return &pipes.back();
}

constexpr int64_t kTargetChunkBytes = 8ll << 20;  // per tensor per chunk

template <bool SYM>
int fwd_bwd_host(const void* x_host, const void* g_host, void* y_host, void* gx_host, float lo,
                 float hi, int64_t rows, int64_t cols, int dtype, int bits, void* scratch,
                 size_t scratch_bytes, void* stream) {
  QAT_CHECK_ARG(dtype == QAT_F32 || dtype == QAT_BF16, "dtype must be QAT_F32 or QAT_BF16 (got %d)", dtype);
  QAT_CHECK_ARG(rows >= 0 && cols >= 0, "negative shape");
  if (rows == 0 || cols == 0) return QAT_OK;
  QAT_CHECK_ARG(x_host != nullptr && y_host != nullptr, "x_host / y_host is NULL");
  const bool bwd = g_host != nullptr;
  QAT_CHECK_ARG(!bwd || gx_host != nullptr, "gx_host is NULL but g_host is given");
  const size_t need = qat_host_scratch_bytes(rows, cols, dtype, bwd ? 1 : 0);
  if (scratch == nullptr || scratch_bytes < need) {
    set_error("host entry point needs %zu bytes of device scratch (got %zu)", need, scratch_bytes);
    return QAT_ERR_WORKSPACE;
  }
  if (qat_fwd_workspace_bytes(rows, cols, dtype) != 0) {
    set_error("host entry points handle register-resident rows only (cols=%lld too long)", (long long)cols);
    return QAT_ERR_UNSUPPORTED;
  }
  Pipe* pipe = get_pipe();
  if (pipe == nullptr) return cuda_fail(cudaGetLastError(), "creating pipeline streams");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  const int64_t esz = dtype == QAT_F32 ? 4 : 2;
  const int64_t row_bytes = cols * esz;
  // device rows start 16-byte aligned so the vector path is taken whenever the
  // row pitch allows; tensors are laid out back to back, each 256-byte aligned.
  const int64_t tensor_bytes = ((rows * row_bytes + 255) / 256) * 256;
  char* dx = reinterpret_cast<char*>(scratch);
  char* dy = dx + tensor_bytes;
  char* dg = dy + tensor_bytes;
  char* dgx = dg + tensor_bytes;

  int64_t chunk_rows = kTargetChunkBytes / (row_bytes > 0 ? row_bytes : 1);
  if (chunk_rows < 1) chunk_rows = 1;
  // keep chunk starts 16-byte aligned for any row pitch
  while ((chunk_rows * row_bytes) % 16 != 0) ++chunk_rows;
  const int64_t nchunks = (rows + chunk_rows - 1) / chunk_rows;

  cudaError_t e;
#define QAT_TRY(call)                                  \
  do {                                                 \
    e = (call);                                        \
    if (e != cudaSuccess) return cuda_fail(e, #call);  \
  } while (0)

  cudaEvent_t ev_start = pipe->event(0);
  if (ev_start == nullptr) return cuda_fail(cudaGetLastError(), "cudaEventCreate");
  QAT_TRY(cudaEventRecord(ev_start, st));
  QAT_TRY(cudaStreamWaitEvent(pipe->in, ev_start, 0));
  QAT_TRY(cudaStreamWaitEvent(pipe->out, ev_start, 0));

  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t r0 = c * chunk_rows;
    const int64_t nr = (rows - r0 < chunk_rows) ? rows - r0 : chunk_rows;
    const int64_t off = r0 * row_bytes, bytes = nr * row_bytes;
    cudaEvent_t ev_in = pipe->event(1 + 2 * c), ev_k = pipe->event(2 + 2 * c);
    if (ev_in == nullptr || ev_k == nullptr) return cuda_fail(cudaGetLastError(), "cudaEventCreate");
    QAT_TRY(cudaMemcpyAsync(dx + off, reinterpret_cast<const char*>(x_host) + off, bytes,
                            cudaMemcpyHostToDevice, pipe->in));
    if (bwd)
      QAT_TRY(cudaMemcpyAsync(dg + off, reinterpret_cast<const char*>(g_host) + off, bytes,
                              cudaMemcpyHostToDevice, pipe->in));
    QAT_TRY(cudaEventRecord(ev_in, pipe->in));
    QAT_TRY(cudaStreamWaitEvent(st, ev_in, 0));
    int rc = SYM ? qat_sym_fwd(dx + off, dy + off, nullptr, QAT_CODES_NONE, nullptr, nullptr, nullptr,
                               lo, hi, nr, cols, dtype, bits, nullptr, 0, st)
                 : qat_asym_fwd(dx + off, dy + off, nullptr, QAT_CODES_NONE, nullptr, nullptr, nullptr,
                                lo, hi, nr, cols, dtype, bits, nullptr, 0, st);
    if (rc != QAT_OK) return rc;
    if (bwd) {
      rc = qat_ste_bwd(dg + off, dx + off, dgx + off, nullptr, lo, hi, nr * cols, dtype, st);
      if (rc != QAT_OK) return rc;
    }
    QAT_TRY(cudaEventRecord(ev_k, st));
    QAT_TRY(cudaStreamWaitEvent(pipe->out, ev_k, 0));
    QAT_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(y_host) + off, dy + off, bytes,
                            cudaMemcpyDeviceToHost, pipe->out));
    if (bwd)
      QAT_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(gx_host) + off, dgx + off, bytes,
                              cudaMemcpyDeviceToHost, pipe->out));
  }
  cudaEvent_t ev_done = pipe->event(1 + 2 * nchunks);
  if (ev_done == nullptr) return cuda_fail(cudaGetLastError(), "cudaEventCreate");
  QAT_TRY(cudaEventRecord(ev_done, pipe->out));
  QAT_TRY(cudaStreamWaitEvent(st, ev_done, 0));
#undef QAT_TRY
  return QAT_OK;
}

}  // namespace
}  // namespace qat

extern "C" {

size_t qat_host_scratch_bytes(int64_t rows, int64_t cols, int dtype, int with_backward) {
  if (rows <= 0 || cols <= 0) return 0;
  const int64_t esz = dtype == QAT_F32 ? 4 : 2;
  const int64_t tensor_bytes = ((rows * cols * esz + 255) / 256) * 256;
  return (size_t)tensor_bytes * (with_backward ? 4 : 2);
}

int qat_sym_fwd_bwd_host(const void* x_host, const void* g_host, void* y_host, void* gx_host,
                         float clip_lo, float clip_hi, int64_t rows, int64_t cols, int dtype,
                         int bits, void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
  return qat::fwd_bwd_host<true>(x_host, g_host, y_host, gx_host, clip_lo, clip_hi, rows, cols,
                                 dtype, bits, dev_scratch, dev_scratch_bytes, stream);
}

int qat_asym_fwd_bwd_host(const void* x_host, const void* g_host, void* y_host, void* gx_host,
                          float clip_lo, float clip_hi, int64_t rows, int64_t cols, int dtype,
                          int bits, void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
  return qat::fwd_bwd_host<false>(x_host, g_host, y_host, gx_host, clip_lo, clip_hi, rows, cols,
                                  dtype, bits, dev_scratch, dev_scratch_bytes, stream);
}

}  // extern "C"
