"""CPU: the C-ABI library loads, exports every symbol include/qat_b200.h
declares, and the Python boundary mirrors the reference's interface.  No
kernel is launched here."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qat_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qat_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = header_functions()
    for must in ("qat_sym_fwd", "qat_asym_fwd", "qat_ste_bwd", "qat_ste_bwd_from_mask", "qat_qlinear_i8_fwd",
                 "qat_lowbit_weight_fwd", "qat_version", "qat_last_error", "qat_sym_fwd_bwd_host"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import llm_qat_b200

    path = llm_qat_b200._lib.LIB_PATH
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in qat_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (qat_[a-z0-9_]+)", out))
    assert exported == set(header_functions()), exported ^ set(header_functions())


def test_binding_covers_the_header():
    import llm_qat_b200

    assert sorted(llm_qat_b200._lib.exported_symbols()) == header_functions()
    assert llm_qat_b200._lib.lib().qat_version() == 100


def test_library_is_sm100a_native():
    """SASS carries the Blackwell-only mnemonics: tcgen05 MMA, TMEM load, TMA."""
    import llm_qat_b200

    sass = subprocess.run(["cuobjdump", "-sass", llm_qat_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCIMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic


def test_bad_arguments_return_errors_not_crashes():
    import llm_qat_b200

    L = llm_qat_b200._lib.lib()
    # dtype 7 is invalid; pointers are never dereferenced on this path
    rc = L.qat_sym_fwd(16, 32, 0, 0, 0, 0, 0, -2.0, 2.0, 4, 4, 7, 4, 0, 0, 0)
    assert rc == 1001 and "dtype" in llm_qat_b200._lib.last_error()
    rc = L.qat_sym_fwd(16, 32, 0, 0, 0, 0, 0, -2.0, 2.0, 4, 4, 0, 1, 0, 0, 0)
    assert rc == 1001 and "num_bits" in llm_qat_b200._lib.last_error()
    rc = L.qat_qlinear_i8_fwd(16, 16, 16, 16, 16, 8, 8, 12, 1, 0)
    assert rc == 1001 and "multiple of 16" in llm_qat_b200._lib.last_error()
    assert L.qat_fwd_workspace_bytes(8192, 4096, 0) == 0
    assert L.qat_fwd_workspace_bytes(1, 1 << 25, 0) == 16
    assert L.qat_host_scratch_bytes(8, 16, 1, 1) == 4 * 256


def test_no_cpu_fallback():
    from llm_qat_b200 import AsymQuantizer, QuantizeLinear, SymQuantizer

    x = torch.randn(4, 8)
    clip = torch.tensor([-2.0, 2.0])
    for q in (SymQuantizer, AsymQuantizer):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            q.apply(x, clip, 4, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        QuantizeLinear(8, 4, w_bits=4, a_bits=8)(x)


def test_quantize_linear_mirrors_reference_interface():
    from llm_qat_b200 import AsymQuantizer, QuantizeLinear, SymQuantizer

    lin = QuantizeLinear(16, 8, bias=True, w_bits=4, a_bits=8, act_layerwise=False, weight_layerwise=False)
    assert isinstance(lin, torch.nn.Linear)
    assert lin.bias is None                       # reference ignores bias= (utils_quant.py:176)
    assert list(lin.state_dict().keys()) == ["weight"] and lin.weight.shape == (8, 16)
    assert (lin.w_bits, lin.a_bits, lin.act_layerwise, lin.weight_layerwise) == (4, 8, False, False)
    assert lin.act_quantizer is SymQuantizer
    assert QuantizeLinear(16, 8, symmetric=False, a_bits=8).act_quantizer is AsymQuantizer
    assert not hasattr(QuantizeLinear(16, 8, a_bits=32), "act_quantizer")   # :184
    assert not hasattr(QuantizeLinear(16, 8, a_bits=2), "act_quantizer")
    assert len(list(lin.buffers())) == 0 and len(list(lin.parameters())) == 1


def test_install_aliases_reference_module_name():
    import sys

    import llm_qat_b200

    llm_qat_b200.install("qat_test_models.utils_quant")
    mod = sys.modules["qat_test_models.utils_quant"]
    assert mod.QuantizeLinear is llm_qat_b200.QuantizeLinear and mod.SymQuantizer is llm_qat_b200.SymQuantizer
    del sys.modules["qat_test_models.utils_quant"]


def test_reduction_view_matches_reference_rules():
    from llm_qat_b200.utils_quant import _reduction_view

    assert _reduction_view(torch.zeros(6, 8), False) == (6, 8)
    assert _reduction_view(torch.zeros(2, 3, 8), False) == (6, 8)
    assert _reduction_view(torch.zeros(8), False) == (1, 8)
    assert _reduction_view(torch.zeros(()), False) == (1, 1)
    assert _reduction_view(torch.zeros(2, 3, 4, 8), False) == (6, 32)
    assert _reduction_view(torch.zeros(2, 3, 4, 8), True) == (1, 192)
    with pytest.raises(ValueError):
        _reduction_view(torch.zeros(1, 1, 1, 1, 1), False)      # utils_quant.py:69-70
    with pytest.raises(RuntimeError):                           # non-viewable 4-D raises like .view()
        _reduction_view(torch.zeros(2, 3, 4, 8).transpose(1, 2), False)


def test_host_side_policy_knobs(monkeypatch):
    """Host logic that needs no GPU: cache-mode parsing, the blob layout the forward and the
    backward must agree on, the gates of the fused path, and the new ABI knobs."""
    import llm_qat_b200
    from llm_qat_b200 import QuantizeLinear
    from llm_qat_b200 import utils_quant as uq

    for env, want in (("0", 0), ("1", 1), ("2", 2), ("", 1), ("yes", 1)):
        monkeypatch.setenv("QAT_B200_CACHE", env)
        assert uq._cache_mode() == want
    monkeypatch.delenv("QAT_B200_CACHE")
    assert uq._cache_mode() == 1
    assert not uq._in_backward_pass()                       # not inside autograd's engine here

    off_e, off_m, total = uq._feed_layout(2048, 4096)       # codes | divisors | packed mask, 256 B aligned
    assert off_e == 2048 * 4096 and off_m == off_e + 2048 * 4 and total == off_m + 2048 * 4096 // 8
    off_e, off_m, total = uq._feed_layout(3, 40)
    assert off_e == 256 and off_m == 512 and total == 512 + 15

    lin = QuantizeLinear(32, 8, w_bits=4, a_bits=8)
    assert not lin._can_fuse(torch.zeros(4, 32))            # CPU tensor: the unfused path raises "no CPU fallback"
    assert not QuantizeLinear(32, 8, w_bits=16, a_bits=8)._can_fuse(torch.zeros(4, 32))
    assert not uq._sym_amp(torch.zeros(4, 32, dtype=torch.bfloat16))   # CPU tensors never take the autocast variant

    L = llm_qat_b200._lib.lib()
    assert L.qat_set_gemm_cta_group(3) == 1001 and "cta_group" in llm_qat_b200._lib.last_error()
    assert L.qat_set_gemm_cta_group(0) == 0 and L.qat_set_pdl(1) == 0
    # the autocast variant exists for SymQuantizer only
    assert L.qat_asym_fwd(16, 32, 0, 0, 0, 0, 0, -2.0, 2.0, 4, 4, 2, 4, 0, 0, 0) == 1001
    assert L.qat_ste_bwd_devclip(16, 16, 32, 0, 0, 8, 0, 0) == 1001


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/qat_b200.h must be consumable by a C compiler (no C++ or torch types across the
    ABI), and a C program must link against the shared library and reach the no-GPU entry points."""
    import shutil
    import subprocess

    import llm_qat_b200

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include "qat_b200.h"
int main(void) {
  if (qat_version() != QAT_B200_VERSION) return 1;
  /* bad dtype: rejected before any pointer is touched, with a message */
  if (qat_sym_fwd((const void*)16, (void*)32, 0, QAT_CODES_NONE, 0, 0, 0, -2.0f, 2.0f, 4, 4, 9, 4, 0, 0, 0) != QAT_ERR_BAD_ARG) return 2;
  if (qat_last_error()[0] == 0) return 3;
  if (qat_fwd_workspace_bytes(8192, 4096, QAT_BF16) != 0) return 4;
  if (qat_host_scratch_bytes(8, 16, QAT_BF16, 1) != 4 * 256) return 5;
  printf("abi ok %d\n", qat_version());
  return 0;
}
''')
    lib_dir = os.path.dirname(llm_qat_b200._lib.LIB_PATH)
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src),
                    "-L", lib_dir, "-l:libqat_b200.so", "-Wl,-rpath," + lib_dir, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "abi ok 100" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_pdl_is_switched_off_while_a_torch_profiler_is_active():
    """Programmatic dependent launch x CUPTI: the binding turns PDL off for the duration of a torch.profiler
    session (the stall it avoids is reproduced by tests/gpu_pdl_profiler_soak.py)."""
    import torch
    from torch.profiler import ProfilerActivity, profile

    from llm_qat_b200 import _lib

    if not _lib._pdl_auto:
        import pytest

        pytest.skip("QAT_B200_PDL is forced or off in this environment")
    _lib.lib()
    assert _lib._pdl_suppressed is False or _lib._INJECTED
    with profile(activities=[ProfilerActivity.CPU]):
        _lib.lib()
        assert _lib._pdl_suppressed is True
    _lib.lib()
    assert _lib._pdl_suppressed is _lib._INJECTED
