import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/libvisma_oracle.so), built on demand with gcc."""
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference behind oracle/ref_shim.cpp; built in the dev container only."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/libvisma_ref.so not built (needs /root/reference: `make -C oracle ref`)")
    pyref.lib()
    return pyref


@pytest.fixture(scope="session")
def kat():
    d = np.load(os.path.join(GOLDEN, "icp_kat.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def pair12():
    """cloud_bin_2 -> cloud_bin_1 from the reference's examples/TestData/ICP/init.log entry "1 2"; expected
    outputs = the unmodified reference's (scripts/make_golden.py)."""
    d = np.load(os.path.join(GOLDEN, "icp_pair12.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def unit_rand():
    d = np.load(os.path.join(GOLDEN, "unit_rand.npz"))
    return d["rand"].astype(np.float64), float(d["rand_max"])


@pytest.fixture(scope="session")
def vb():
    """The product library through its Python mirror; fails loudly if the CUDA extension is missing."""
    import visma_b200
    from visma_b200 import _lib, registration, renderer, synth
    _lib.lib()

    class NS:
        pass
    ns = NS()
    ns.lib, ns.reg, ns.ren, ns.synth, ns.pkg = _lib, registration, renderer, synth, visma_b200
    return ns


def small_scene(n_scene=60000, n_objects=3, m=4000, seed=5):
    from visma_b200 import synth
    return synth.make_room_scene(n_scene, n_objects, m, seed=seed)
