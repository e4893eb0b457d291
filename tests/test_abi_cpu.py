"""No-GPU checks of the boundary: the C-ABI library builds for sm_100a, loads, exports every symbol that
include/qvz_gpu.h declares, and refuses to run without a device (there is no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from qvz_b200 import lib
    return lib.build()


def test_exports_match_header(libpath):
    import ctypes
    from qvz_b200 import lib
    header = open(os.path.join(ROOT, "include", "qvz_gpu.h")).read()
    declared = set(re.findall(r"\b(qvz_gpu_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lib.EXPORTS)
    L = ctypes.CDLL(libpath)
    for name in declared:
        assert hasattr(L, name), name


def test_header_compiles_as_c(tmp_path):
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "qvz_gpu.h"\nint main(void){struct qvz_flat_tables t; (void)t; return sizeof(struct qvz_gpu_timings) != 44;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o",
                    str(tmp_path / "t")], check=True)
    subprocess.run([str(tmp_path / "t")], check=True)


def test_no_cpu_fallback(libpath):
    import torch
    from qvz_b200 import lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.QvzError):
        lib.Handle(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qvz_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".cc", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/_ref", ""), os.path.join(dirpath, f)
