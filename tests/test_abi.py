"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/vldd_b200.h declares."""
import ctypes
import os
import subprocess

import pytest


def test_library_builds_and_exports_header_symbols():
    from multimodal_dataset_distillation_b200 import _lib, build
    path = build.build_library()
    assert os.path.exists(path)
    handle = _lib.lib()
    declared = _lib.header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/vldd_b200.h but not exported"
    assert set(declared) == set(_lib._SIGS), "ctypes signatures out of sync with the header"
    assert handle.vldd_version() == 100


def test_library_is_sm100a_and_has_no_torch_dependency():
    from multimodal_dataset_distillation_b200 import build
    path = build.build_library()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    ldd = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd


def test_argument_errors_do_not_need_a_gpu():
    from multimodal_dataset_distillation_b200 import _lib
    h = _lib.lib()
    rc = h.vldd_flat_sgd_step(None, None, None, None, 8, None)
    assert rc == -1 and b"null" in h.vldd_last_error()
    assert h.vldd_unrolled_match_workspace_bytes(100, 200, 8, 768, 2304) == 0      # B > N is rejected
    assert b"B <= N" in h.vldd_last_error()
    n = h.vldd_unrolled_match_workspace_bytes(100, 100, 8, 768, 2304)
    assert 300e6 < n < 2e9


def test_product_path_refuses_cpu_tensors():
    import torch
    from multimodal_dataset_distillation_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.flat_sgd_step(torch.zeros(4), torch.zeros(4), 0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.proj_head_forward(torch.zeros(ops.head_numel(4, 8)), torch.zeros(2, 4), 8)


def test_product_package_never_imports_the_oracle():
    import pathlib
    root = pathlib.Path(__file__).resolve().parents[1] / "multimodal_dataset_distillation_b200"
    for f in root.rglob("*.py"):
        text = f.read_text()
        assert "import oracle" not in text and "from oracle" not in text, f
