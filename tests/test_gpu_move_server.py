"""The resident move server (graspa_b200/csrc/move_server.cuh) against the launch-per-move path: the same chain of Monte Carlo moves --
every kind, accepted and rejected, on a charged system (config B: CO2 in MFI, Ewald) and on the Xe/Kr mixture (config D: identity
swaps, tail corrections) -- must give the same selections and decisions, energies equal to rounding (1e-11; bitwise on the uncharged system) and leave bitwise the same atoms; the server must survive being
stopped by other calls in between, leaving on its idle limit, and a refill of the random pool."""
import gc
import os
import time

import numpy as np
import pytest

from graspa_b200.types import TRANSLATION, ROTATION
from tests.conftest import load_config
from tests.test_gpu_moves import _grow

pytestmark = pytest.mark.gpu


def _engine(gpu_engine_factory, name, server):
    box, ff, s, z = load_config(name)
    if name == "B":
        s = _grow(s, 1, 600)
    else:
        s = _grow(_grow(s, 1, 300), 2, 300)
    eng = gpu_engine_factory(box, ff, s, float(z["beta"]), int(z["ntrials"]), int(z["norient"]))
    if "sf_ads" in z:
        eng.upload_structure_factors(z["sf_ads"], z["sf_fw"])
        eng.set_exclusion_constants(int(z["comp"]), float(z["excl"][0]), float(z["excl"][1]))
    eng.move_server(server)
    return s, z, eng


def _flat(m):
    out = []
    for k in ("first_bead", "chain", "old_first_bead", "old_chain"):
        d = m[k]
        out += [float(d["rosenbluth"]), float(d["stored_r"]), float(d["selected"]), float(d["success"]), float(d["n_survivors"])]
        out += [float(v) for v in np.ravel(d["energy"])] + [float(v) for v in np.ravel(d["selected_pos"])]
    out += [m["ewald"][0], m["ewald"][1], m["tail"], float(m["overlap"]), float(m["uniforms_used"]), float(m["pool_used"]), float(m["success"])]
    out += [float(m["delta"][k]) for k in ("HHVDW", "HHReal", "HGVDW", "HGReal", "GGVDW", "GGReal")]
    return np.array(out)


def _chain(eng, s, z, comps, nsteps, seed, interrupt_every=0, pool_refill_at=-1, sleep_at=-1):
    """a deterministic chain of moves; decisions depend only on the rng and on the results (equal results -> equal chains)"""
    rng = np.random.default_rng(seed)
    pool = rng.random((4096, 3)); eng.upload_random_pool(pool)
    off = 0; log = []
    for step in range(nsteps):
        if off > 4000:
            off = 0
        if step == pool_refill_at:
            pool = rng.random((4096, 3)); eng.upload_random_pool(pool); off = 0
        if step == sleep_at:
            time.sleep(0.35)                                   # longer than the server's idle limit: it leaves and is started again
        if interrupt_every and step % interrupt_every == interrupt_every - 1:
            log.append(eng.download_atoms(comps[0])["pos"].ravel().copy())     # not a server call: stops it, commits are flushed
        c = int(rng.choice(comps)); ms = int(s.molsize[c])
        nmol = eng.number_of_molecules(c)
        kind = rng.choice(["ins", "del", "tr", "rot", "rei", "swap"] if len(comps) > 1 else ["ins", "del", "tr", "rot", "rei"])
        u = rng.random(2); acc = rng.random()
        if kind != "ins" and nmol == 0:
            continue
        mol = int(rng.integers(0, max(nmol, 1)))
        if kind == "ins":
            m = eng.move_insertion(c, off, u); off += m["pool_used"]
            if m["success"] and acc < 0.6:
                eng.accept_insertion(c)
        elif kind == "del":
            m = eng.move_deletion(c, mol, off); off += m["pool_used"]
            if m["success"] and acc < 0.4:
                eng.accept_deletion(c, mol)
        elif kind in ("tr", "rot"):
            if kind == "rot" and ms == 1:
                continue
            m = eng.move_single_body(TRANSLATION if kind == "tr" else ROTATION, c, mol, (0.4, 0.5, 0.3), off); off += m["pool_used"]
            if m["success"] and acc < 0.5:
                eng.accept_translation(c)
        elif kind == "rei":
            m = eng.move_reinsertion(c, mol, off, u); off += m["pool_used"]
            if m["success"] and acc < 0.5:
                eng.accept_reinsertion(c, mol)
        else:
            newc = int(rng.choice(comps))
            m = eng.move_identity_swap(c, mol, newc, off, u[0]); off += m["pool_used"]
            if m["success"] and acc < 0.6:
                eng.accept_identity_swap(c, mol, newc)
        log.append(_flat(m))
    atoms = [eng.download_atoms(c) for c in comps]
    return log, atoms


def _compare(a, b):
    la, aa = a; lb, ab = b
    assert len(la) == len(lb)
    worst = 0.0
    for k, (x, y) in enumerate(zip(la, lb)):
        assert x.shape == y.shape
        if not np.array_equal(x, y, equal_nan=True):
            d = np.abs(x - y) / np.maximum(np.abs(y), 1e-6)
            worst = max(worst, float(np.nanmax(d)))
            assert np.nanmax(d) < 1e-11, (k, int(np.nanargmax(d)), x[int(np.nanargmax(d))], y[int(np.nanargmax(d))])
    print("largest relative difference between the two paths:", worst)
    for x, y in zip(aa, ab):
        n = x["n_live"]
        assert n == y["n_live"]
        for k in ("pos", "scale", "charge", "scale_coul", "type"):
            assert np.array_equal(x[k][:n], y[k][:n]), k


@pytest.mark.parametrize("name,comps,nsteps", [("B", [1], 260), ("D", [1, 2], 400)])
def test_server_gives_the_launch_per_move_results(gpu_engine_factory, name, comps, nsteps):
    gc.collect()
    s, z, eng = _engine(gpu_engine_factory, name, True)
    a = _chain(eng, s, z, comps, nsteps, 77)
    starts, cmds = eng.move_server()
    assert starts >= 1 and cmds >= nsteps // 2, "the resident server did not run (another engine alive in this process?)"
    eng.close(); del eng; gc.collect()
    s, z, eng = _engine(gpu_engine_factory, name, False)
    b = _chain(eng, s, z, comps, nsteps, 77)
    assert eng.move_server() == (0, 0)
    eng.close()
    _compare(a, b)


def test_server_is_stopped_and_restarted_by_other_calls_idle_limit_and_pool_refills(gpu_engine_factory, monkeypatch):
    gc.collect()
    monkeypatch.setenv("GB_MOVE_SERVER_IDLE_MS", "100")
    s, z, eng = _engine(gpu_engine_factory, "D", True)
    a = _chain(eng, s, z, [1, 2], 240, 5, interrupt_every=25, pool_refill_at=120, sleep_at=60)
    starts, cmds = eng.move_server()
    assert starts >= 240 // 25 and cmds > 100
    eng.close(); del eng; gc.collect()
    s, z, eng = _engine(gpu_engine_factory, "D", False)
    b = _chain(eng, s, z, [1, 2], 240, 5, interrupt_every=25, pool_refill_at=120)
    eng.close()
    _compare(a, b)
