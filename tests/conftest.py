import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from graspa_b200.types import Box, ForceField, System  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_config(name):
    """-> (box, ff, system, raw npz dict) from the committed golden fixture tests/golden/config_<name>.npz"""
    z = dict(np.load(os.path.join(GOLDEN, f"config_{name}.npz")))
    box = Box(z["cell"], alpha=float(z["alpha"]), kmax=tuple(int(k) for k in z["kmax"]), recip_cutoff=float(z["recip_cutoff"]),
              prefactor=float(z["prefactor"]))
    ff = ForceField(z["eps"], z["sigma"], z["shift"], float(z["cutoff_vdw"]), float(z["cutoff_coul"]), overlap=float(z["overlap"]),
                    no_charges=bool(int(z["no_charges"])), use_tail=z["use_tail"], tail_energy=z["tail_energy"])
    system = System(int(z["nhost"]), z["natoms"], z["molsize"], z["pos"], z["charge"], z["type"], z["molid"], alloc=z["alloc"])
    return box, ff, system, z


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(b), floor), 1e-300)))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def gpu_engine_factory():
    """GPU tests go through the C ABI; the library must exist and a device must be present."""
    from graspa_b200 import engine
    engine.load_library()

    def make(box, ff, system, beta=None, ntrials=10, norient=10):
        return engine.Engine().setup(box, ff, system, beta, ntrials, norient)
    return make
