"""CPU checks of the drop-in boundary: both shared libraries load without a GPU and export every symbol the headers
in include/ declare; argument errors are reported through the return code + db1_last_error (no compute is launched)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(db1_[a-z0-9_]+)\s*\(", src)))


def test_cuda_library_exports_every_declared_symbol():
    from db1_sm100 import _lib
    lib = _lib.lib()
    names = _declared("db1_sm100.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libdb1_sm100.so does not export %s" % n
    assert lib.db1_abi_version() >= 1


def test_host_library_exports_every_declared_symbol():
    from db1_sm100 import _lib
    lib = _lib.hostlib()
    names = _declared("db1_host.h")
    assert names == ["db1_build_rl_sample_idx", "db1_build_sample_idx", "db1_collate_plan", "db1_concat_rows", "db1_decode",
                     "db1_discretize", "db1_rl_assemble", "db1_rl_layout"]
    for n in names:
        assert hasattr(lib, n), "libdb1_host.so does not export %s" % n


def test_argument_errors_do_not_launch_and_set_last_error():
    from db1_sm100 import _lib
    lib = _lib.lib()
    rc = lib.db1_gemm_f16(None, None)
    assert rc < 0 and b"null descriptor" in lib.db1_last_error()
    d = _lib.GemmDesc()
    d.M, d.N, d.K = 128, 128, 0
    rc = lib.db1_gemm_f16(C.byref(d), None)
    assert rc < 0 and b"bad shape" in lib.db1_last_error()
    rc = lib.db1_layernorm_fwd(None, None, None, None, None, 4, 64, C.c_float(1e-5), None)
    assert rc < 0 and b"null pointer" in lib.db1_last_error()


def test_product_path_refuses_cpu_tensors():
    """No CPU / PyTorch fallback: the module surface raises instead of computing off the kernels."""
    import pytest
    import torch
    from db1_sm100._lib import Db1Error
    from db1_sm100 import ops
    with pytest.raises(Db1Error):
        ops.gemm(torch.zeros(8, 8, dtype=torch.half), torch.zeros(8, 8, dtype=torch.half),
                 torch.zeros(8, 8, dtype=torch.half), 8, 8, 8, lda=8, ldb=8, ldc=8)
