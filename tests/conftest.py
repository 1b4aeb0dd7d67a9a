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


def _cuda_device_present():
    """True when the CUDA driver reports a device (checked through libcuda directly: no torch import at collection time)"""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        return cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests` on a host without a GPU skips the gpu-marked tests instead of failing them (there is no CPU path)"""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the engine has no CPU path")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_config(name):
    """-> (box, ff, system, raw npz dict) from the committed golden fixture tests/golden/config_<name>.npz"""
    z = dict(np.load(os.path.join(GOLDEN, f"config_{name}.npz")))
    box = Box(z["cell"], alpha=float(z["alpha"]), kmax=tuple(int(k) for k in z["kmax"]), recip_cutoff=float(z["recip_cutoff"]),
              prefactor=float(z["prefactor"]))
    ff = ForceField(z["eps"], z["sigma"], z["shift"], float(z["cutoff_vdw"]), float(z["cutoff_coul"]), overlap=float(z["overlap"]),
                    no_charges=bool(int(z["no_charges"])), use_tail=z["use_tail"], tail_energy=z["tail_energy"])
    system = System(int(z["nhost"]), z["natoms"], z["molsize"], z["pos"], z["charge"], z["type"], z["molid"], alloc=z["alloc"])
    return box, ff, system, z


ENERGY_TO_KELVIN = 1.20272430057          # the reference's internal energy unit -> K (read_data.cpp:1200)


def load_nist(b):
    """-> (box, ff, system, printed) of Box-<b> of the reference's NIST SPC/E known-answer example
    (tests/golden/nist_spce.npz, written by tests/golden/make_nist.py).  printed = the reference's own output.txt values
    {vdw_gg, real_gg, ewald_gg (Fourier - self - intra), tail, total, fourier_gg}, internal units."""
    z = dict(np.load(os.path.join(GOLDEN, "nist_spce.npz")))
    ff = ForceField(z["eps"], z["sigma"], z["shift"], float(z["cutoff_vdw"]), float(z["cutoff_coul"]), overlap=float(z["overlap"]),
                    use_tail=z["use_tail"], tail_energy=z["tail_energy"])
    box = Box(z[f"b{b}_cell"], alpha=float(z[f"b{b}_alpha"]), kmax=tuple(int(k) for k in z[f"b{b}_kmax"]),
              recip_cutoff=float(z[f"b{b}_recip_cutoff"]), use_lammps_ewald=True)
    pos = z[f"b{b}_pos"]; n = len(pos); ms = len(z["mol_type"]); nm = n // ms
    system = System(1, np.array([0, n]), np.array([0, ms]), pos, np.tile(z["mol_charge"], nm), np.tile(z["mol_type"], nm),
                    np.repeat(np.arange(nm), ms), alloc=np.array([0, n]))
    keys = ("vdw_gg", "real_gg", "ewald_gg", "tail", "total", "fourier_gg")
    return box, ff, system, dict(zip(keys, (float(x) for x in z[f"b{b}_printed"])))


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
