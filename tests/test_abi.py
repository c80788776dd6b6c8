"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/sailgpu.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sailgpu.h")).read()
    return sorted(set(re.findall(r"SAILGPU_API [\w\s\*]+?(sailgpu_\w+)\(", text)))


def test_header_declares_the_documented_surface():
    syms = declared_symbols()
    for s in ["sailgpu_ctx_create", "sailgpu_op_create", "sailgpu_op_push", "sailgpu_op_push_device",
              "sailgpu_op_finish_input", "sailgpu_op_pull", "sailgpu_op_pull_device", "sailgpu_op_metrics",
              "sailgpu_last_error", "sailgpu_op_destroy", "sailgpu_exchange", "sailgpu_ctx_comm_init"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from sail_b200 import build
    lib = ctypes.CDLL(build.build())
    for s in declared_symbols():
        assert hasattr(lib, s), f"libsailgpu.so does not export {s}"
    assert lib.sailgpu_version() >= 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sail_b200 import engine
    with pytest.raises(engine.GpuUnavailable):
        engine.Context(0)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sail_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".hpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f"{f} imports the oracle"
