import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["small_f05_M_c2", "small_r2_L_c1", "small_f10_A_c3"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    from oracle.bindings import FlatTables
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["columns"], g["clusters"] = int(g["columns"]), int(g["clusters"])
    g["tables"] = FlatTables(g["clusters"], g["columns"], g["t_nctx"], g["t_ctx_of"], g["t_q_off"],
                             g["t_qratio"], g["t_qmap"], g["t_smap"], g["t_distortion"])
    return g


@pytest.fixture(scope="session")
def oracle():
    from oracle.bindings import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.bindings import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libqvzref.so not built (reference sources absent)")
    return Ref()


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
