import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.build()
    return oracle


def _build_host_sim(brick_steps: int):
    import ctypes as C
    d = os.path.join(ROOT, "tests", "host_sim")
    so = os.path.join(d, "libhost_sim_brick.so" if brick_steps else "libhost_sim.so")
    src = os.path.join(d, "host_sim.cpp")
    hdrs = [os.path.join(ROOT, "nerfacc_b200", "csrc", h) for h in ("lattice.cuh", "march.cuh", "march_generic.cuh", "expand.cuh", "occ_pack.cuh", "nfa_math.cuh",
                                                                     "pdf.cuh")]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-Wno-unknown-pragmas", f"-DNFA_BRICK_STEPS={brick_steps}", "-o", so, src])
    lib = C.CDLL(so)
    lib.sim_chain.restype = C.c_float
    lib.sim_chain.argtypes = [C.c_float, C.c_float, C.c_uint32]
    lib.sim_seek.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    lib.sim_expand_run.argtypes = [C.c_float, C.c_float, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.sim_occ_words.restype = C.c_int64
    lib.sim_occ_coarse_words.restype = C.c_int64
    lib.sim_march.restype = C.c_int64
    lib.sim_philox_uniform.restype = C.c_float
    lib.sim_philox_uniform.argtypes = [C.c_uint64] * 3
    return lib


@pytest.fixture(scope="session")
def host_sim():
    """CPU build of the product's host+device traversal headers (tests/host_sim), as shipped."""
    return _build_host_sim(0)


@pytest.fixture(scope="session")
def host_sim_brick():
    """The same headers with the whole-brick loop of the walk compiled in (NFA_BRICK_STEPS=1, march.cuh): exact, but
    slower on the GPU under SIMT divergence, so not shipped; kept tested."""
    return _build_host_sim(1)


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_bins(z):
    import numpy as np
    shp = tuple(int(v) for v in z["binaries_shape"])
    return np.unpackbits(z["binaries_bits"])[: int(np.prod(shp))].reshape(shp).astype(bool)
