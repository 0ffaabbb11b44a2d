"""CPU-side checks of the drop-in boundary: libvsom_b200.so builds for sm_100a, loads, exports every symbol
include/vsom_b200.h declares, and refuses to run without a GPU instead of falling back to the CPU."""
import os
import subprocess

import pytest

from conftest import REPO


def test_library_exports_every_declared_symbol(vsom):
    from importlib import import_module

    binding = import_module(vsom.__name__ + ".binding")
    declared = binding.header_symbols()
    assert len(declared) >= 20
    assert vsom.exported_symbols() == declared


def test_library_holds_sm100a_code_only(vsom):
    out = subprocess.run(["cuobjdump", "--list-elf", vsom.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = {line.split(".")[-2] for line in out.stdout.split() if line.endswith(".cubin")}
    assert archs == {"sm_100a"}, out.stdout


def test_model_length(vsom):
    assert vsom.model_length(784, vsom.STANDARD) == 784
    assert vsom.model_length(128, vsom.MEDIAN) == 128
    assert vsom.model_length(32, vsom.CLR) == 992


def test_no_cpu_fallback(vsom):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(vsom.VsomError) as e:
        vsom.VsomContext(4, 3, 3)
    assert e.value.code == -5 and "no CPU path" in str(e.value)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(REPO, "variational-self-organizing-maps_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text and "libvsom_ref" not in text and "vsom_oracle" not in text, f
