"""The W1 / W2 weight path needs `w.abs().mean(dim=-1)` bit for bit (/root/reference/models/utils_quant.py:205-210,
219-224), i.e. the summation ORDER of torch's CPU reduction.  `llm-qat_b200/csrc/torch_sum_order.cuh` restates that
order once, for the kernel and for the host; this file compiles the very same header with g++ and compares it —
and the numpy statement in oracle/quant_oracle.py — with `torch.sum` / `torch.mean` themselves on the CPU.
No GPU, no /root/reference needed: the arithmetic lives in torch ("third-party arithmetic")."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import quant_oracle as qo  # noqa: E402

WIDTHS = (list(range(1, 70)) + [127, 128, 129, 172, 255, 256, 257, 300, 511, 512, 513, 1000, 1023, 1024, 2048, 4096,
                               4097, 5120, 8191, 8192, 8200, 11008, 13824, 16384, 16385, 20000, 32768, 40000, 70001,
                               131072 + 37])

HOST_SRC = r'''
#include "torch_sum_order.cuh"
#include <vector>
extern "C" void tso_row_sums(const float* a, long long rows, long long cols, float* out) {
  std::vector<float> scratch(32 + 32 * (qat::tso::chain_steps(cols) / 16 + 1));
  for (long long r = 0; r < rows; ++r) {
    const float* row = a + r * cols;
    out[r] = qat::tso::row_sum_serial([row](long long e) { return row[e]; }, cols, scratch.data());
  }
}
'''


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _cases(seed):
    rng = np.random.default_rng(seed)
    for k in WIDTHS:
        for r in (1, 3, 37):
            if r == 1 and k >= 32768:
                continue   # a single output element of >= 32768 inputs: torch splits the row over threads
            a = np.abs(rng.standard_normal((r, k)).astype(np.float32)) * np.float32(0.02)
            if k > 3 and r > 1:
                a[1, rng.integers(0, k)] = 3e4          # outlier: rounding depends on the order for real
            yield a


def test_numpy_statement_of_the_row_sum_equals_torch_cpu_sum():
    for a in _cases(0):
        want = torch.from_numpy(a).sum(dim=-1).numpy()
        got = qo.torch_cpu_row_sum(a)
        assert np.array_equal(_bits(want), _bits(got)), a.shape


def test_row_mean_fp32_and_bf16_equal_torch_cpu_mean():
    rng = np.random.default_rng(5)
    for k in (1, 5, 7, 8, 31, 172, 300, 4096, 4097, 11008, 13824):
        w = (rng.standard_normal((19, k)) * 0.02).astype(np.float32)
        t = torch.from_numpy(w)
        want = t.abs().mean(dim=-1, keepdim=True).numpy()
        assert np.array_equal(_bits(want), _bits(qo._mean_rows(np.abs(w), False, "fp32"))), k
        tb = t.to(torch.bfloat16)
        wantb = tb.abs().mean(dim=-1, keepdim=True).float().numpy()
        gotb = qo._mean_rows(np.abs(tb.float().numpy()), False, "bf16")
        assert np.array_equal(_bits(wantb), _bits(gotb)), k


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    d = tmp_path_factory.mktemp("tso")
    src = d / "tso_host.cpp"
    src.write_text(HOST_SRC)
    so = d / "libtso_host.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                    "-I", os.path.join(ROOT, "llm-qat_b200", "csrc"), str(src), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.tso_row_sums.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p]
    lib.tso_row_sums.restype = None
    return lib


def test_kernel_header_compiled_for_the_host_equals_torch_cpu_sum(host_lib):
    """The functions the CUDA kernel calls (group_sum / chain_sum / finalize / short_row_sum), built by g++."""
    for a in _cases(1):
        a = np.ascontiguousarray(a)
        out = np.empty(a.shape[0], dtype=np.float32)
        host_lib.tso_row_sums(a.ctypes.data, a.shape[0], a.shape[1], out.ctypes.data)
        want = torch.from_numpy(a).sum(dim=-1).numpy()
        assert np.array_equal(_bits(want), _bits(out)), a.shape


def test_layerwise_mean_of_the_reference_depends_on_its_thread_count():
    """Why the layerwise W1 / W2 scale of a large tensor has a tolerance and not a bit contract: a full reduction of
    >= 32768 elements is split over torch's intra-op threads (TensorIterator two-pass reduction), so the reference's
    own bits change with the machine it runs on.  Below 32768 elements the reduction is one serial cascade over the
    flattened tensor (pinned exactly), and per-row means never split a row (pinned exactly, any thread count)."""
    rng = np.random.default_rng(3)
    n0 = torch.get_num_threads()
    varied = 0
    try:
        for shape in ((192, 4096), (300, 4096), (33, 1000), (40, 192), (7, 300)):
            a = np.abs((rng.standard_normal(shape) * 0.02).astype(np.float32))
            t = torch.from_numpy(a)
            seen = set()
            for n in (1, 2, 3, 5, 8):
                torch.set_num_threads(n)
                if torch.get_num_threads() != n:
                    continue
                seen.add(float(t.sum().item()).hex())
                assert np.array_equal(_bits(t.sum(dim=-1).numpy()), _bits(qo.torch_cpu_row_sum(a))), (shape, n)
            if a.size < 32768:
                assert seen == {float(qo.torch_cpu_row_sum(a.reshape(1, -1))[0]).hex()}, shape
            varied += len(seen) > 1
    finally:
        torch.set_num_threads(n0)
    if n0 == 1:
        pytest.skip("single-threaded host: the split cannot be shown")
    assert varied >= 1
