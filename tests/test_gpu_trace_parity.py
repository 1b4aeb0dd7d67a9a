"""Accept/reject-sequence parity against the REFERENCE CUDA PROGRAM itself, run in the same test on the same GPU
(BASELINE.json: "an identical accept/reject sequence for the first 10^4 cycles", "Widom <W> and Henry coefficients within 1e-9
relative").

oracle/_ref/graspa_ref_cuda_trace.x is the reference's own program built for sm_100 with ONE added line (oracle/build_ref.sh
trace: RunMoves, axpy.cu:297, appends "component movetype deltaE" per move); oracle/_ref/graspa_ref_cuda.x is the unmodified
program.  Both are built in the build container from the sources under /root/reference and travel to the GPU box as prebuilt
files; nothing here reads /root/reference.  graspa_b200_mc --trace writes the same line per move from the host driver above
the C ABI.  A rejected move carries a zeroed MoveEnergy in the reference, so "accepted" = non-zero energy change."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TRACE = os.path.join(ROOT, "oracle", "_ref", "graspa_ref_cuda_trace.x")
REF_PLAIN = os.path.join(ROOT, "oracle", "_ref", "graspa_ref_cuda.x")
DRIVER = os.path.join(ROOT, "graspa_b200", "host", "graspa_b200_mc")


def _deck_copy(name, tmp_path, n_init, n_prod):
    src = os.path.join(ROOT, "oracle", "_ref", "examples", name)
    for need in (src, DRIVER):
        if not os.path.exists(need):
            pytest.skip(f"{need} not built (python -c 'import __graft_entry__ as g; g.build()' in the build container)")
    dst = str(tmp_path / name)
    shutil.copytree(src, dst)
    p = os.path.join(dst, "simulation.input")
    os.chmod(p, 0o644)
    out = []
    for ln in open(p).read().splitlines():
        if ln.startswith("NumberOfInitializationCycles"): ln = f"NumberOfInitializationCycles {n_init}"
        elif ln.startswith("NumberOfEquilibrationCycles"): ln = "NumberOfEquilibrationCycles 0"
        elif ln.startswith("NumberOfProductionCycles"): ln = f"NumberOfProductionCycles {n_prod}"
        out.append(ln)
    open(p, "w").write("\n".join(out) + "\n")
    return dst


def _compare_traces(ref_path, our_path):
    ref = [l.split() for l in open(ref_path)]
    our = [l.split() for l in open(our_path)]
    assert len(ref) == len(our) and len(ref) > 0, (len(ref), len(our))
    differ = 0; first = None; accepted = 0; worst = 0.0; zero_swaps = 0
    for k in range(len(ref)):
        rc, rd = int(ref[k][0]), float(ref[k][2])
        kind, oc, oa, od = our[k][1], int(our[k][2]), int(our[k][4]), float(our[k][5])
        ra = 1 if rd != 0.0 else 0
        accepted += ra
        if kind == "identity_swap":
            oc = rc                     # the reference's TempVal.component is the NEW species until the retrace starts, ours prints the OLD one
            if oa == 1 and od == 0.0 and ra == 0:
                zero_swaps += 1         # a monatomic molecule regrown in place as its own species: accepted with exactly zero energy change
                continue
        if ra != oa or rc != oc:
            differ += 1
            first = k if first is None else first
        if ra and oa:
            worst = max(worst, abs(rd - od) / max(abs(rd), 1e-300))
    return dict(moves=len(ref), accepted=accepted, differ=differ, first=first, worst=worst, zero_swaps=zero_swaps)


@pytest.mark.parametrize("deck,n_init,n_prod,min_moves", [
    ("CO2-MFI", 10000, 0, 200000),              # config B: 10^4 cycles of max(20, N) moves, CBMC insertion / deletion / reinsertion, Ewald
    ("XeKr-Mixture", 10000, 0, 10000),          # config D: identity swaps, tail corrections, no charges
    ("CO2_NaX_Zeolite", 5000, 5000, 10000),     # config C: movable Na+ framework component, block pockets, cubic cell
])
def test_accept_reject_sequence_equals_the_reference_program(deck, n_init, n_prod, min_moves, tmp_path):
    if not os.path.exists(REF_TRACE):
        pytest.skip("oracle/_ref/graspa_ref_cuda_trace.x not built (oracle/build_ref.sh trace)")
    d = _deck_copy(deck, tmp_path, n_init, n_prod)
    ref_trace = str(tmp_path / "ref_trace.txt"); our_trace = str(tmp_path / "our_trace.txt")
    env = dict(os.environ, GRASPA_TRACE=ref_trace)
    with open(str(tmp_path / "ref_out.txt"), "w") as fo:
        r = subprocess.run([REF_TRACE], cwd=d, env=env, stdout=fo, stderr=subprocess.STDOUT, timeout=1200)
    assert r.returncode == 0, open(str(tmp_path / "ref_out.txt")).read()[-2000:]
    o = subprocess.run([DRIVER, d, "--init", str(n_init), "--equil", "0", "--prod", str(n_prod), "--trace", our_trace],
                       capture_output=True, text=True, timeout=1200)
    assert o.returncode == 0, o.stderr[-2000:]
    c = _compare_traces(ref_trace, our_trace)
    print(deck, c)
    assert c["moves"] >= min_moves
    assert c["differ"] == 0, c
    assert c["accepted"] > c["moves"] // 50
    assert c["worst"] < 1e-8, c          # energy change of every accepted move, relative (7e-9 on the longest CO2-MFI chain: sums differ in order)


def _widom_lines(text):
    w = [float(ln.split(":")[1]) for ln in text.splitlines() if ln.startswith("(Total) Averaged Rosenbluth Weight:")]
    avg = [ln for ln in text.splitlines() if ln.startswith("Averaged Rosenbluth Weight:")]
    kh = [ln for ln in text.splitlines() if ln.startswith("Averaged Henry Coefficient")]
    return w, (float(avg[0].split(":")[1].split("+/-")[0]) if avg else None), (float(kh[0].split(":")[1].split("+/-")[0]) if kh else None)


def test_henry_coefficient_deck_equals_the_reference_program(tmp_path):
    """Examples/Henrys_coefficient (config A), 20 000 Widom insertions, seed 0: the per-block <W>, the average and the Henry
    coefficient as the UNMODIFIED reference program prints them, against the batched RNG-exact path of the host driver and
    against its one-insertion-at-a-time path."""
    if not os.path.exists(REF_PLAIN):
        pytest.skip("oracle/_ref/graspa_ref_cuda.x not built (oracle/build_ref.sh cuda)")
    d = _deck_copy("Henrys_coefficient", tmp_path, 0, 20000)
    r = subprocess.run([REF_PLAIN], cwd=d, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    ref_w, ref_avg, ref_kh = _widom_lines(r.stdout)
    assert len(ref_w) == 5 and ref_avg is not None and ref_kh is not None
    for flags in ((), ("--sequential-widom",)):
        o = subprocess.run([DRIVER, d, "--init", "0", "--equil", "0", "--prod", "20000", *flags], capture_output=True, text=True, timeout=1200)
        assert o.returncode == 0, o.stderr[-2000:]
        w, avg, kh = _widom_lines(o.stdout)
        assert len(w) == 5
        # 10 printed decimals of a number ~1e2: 1e-9 relative is the tolerance BASELINE.json states and about what the print resolves
        for a, b in zip(w, ref_w):
            assert abs(a - b) <= 1e-9 * abs(b) + 2e-10, (flags, w, ref_w)
        assert abs(avg - ref_avg) <= 1e-9 * abs(ref_avg) + 2e-10
        assert abs(kh - ref_kh) <= 1e-9 * abs(ref_kh) + 1e-14


def test_pipelined_widom_replay_gives_the_single_lane_averages(tmp_path):
    """a long RNG-exact replay alternates pools between two lanes of engines (evaluation of pool k in a background thread beside the
    upload, classification and walk of pool k + 1) and books the averages in cycle order: every printed digit of the per-block <W>,
    of the average and of the Henry coefficient equals the single-lane run's (300 000 insertions = 18 pools)"""
    d = _deck_copy("Henrys_coefficient", tmp_path, 0, 300000)
    runs = []
    for flags in ((), ("--no-pipeline",)):
        o = subprocess.run([DRIVER, d, "--init", "0", "--equil", "0", "--prod", "300000", *flags], capture_output=True, text=True, timeout=1200)
        assert o.returncode == 0, o.stderr[-2000:]
        assert '"widom_path": "batched-exact"' in o.stdout
        runs.append([ln for ln in o.stdout.splitlines() if "Averaged" in ln or ln.startswith("AVG WIDOM")])
    assert len(runs[0]) >= 12 and runs[0] == runs[1]


REF_OVERLAY = os.path.join(ROOT, "oracle", "_ref", "graspa_ref_overlay.x")


def _stock_and_overlay(d, tmp_path):
    """runs the stock reference (trace build) and the reference bound to libgraspa_b200.so on deck d; -> their outputs, both traces,
    the number of moves whose component / move type / accepted-or-not differ, accepted moves, worst relative energy difference"""
    traces = {}
    outs = {}
    for tag, exe in (("stock", REF_TRACE), ("overlay", REF_OVERLAY)):
        tr = str(tmp_path / f"{tag}_trace.txt")
        r = subprocess.run([exe], cwd=d, env=dict(os.environ, GRASPA_TRACE=tr), capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0, (tag, (r.stdout + r.stderr)[-3000:])
        traces[tag] = [l.split() for l in open(tr)]
        outs[tag] = r
    a, b = traces["stock"], traces["overlay"]
    differ = 0; worst = 0.0; accepted = 0
    for x, y in zip(a, b):
        dx, dy = float(x[2]), float(y[2])
        if x[0] != y[0] or x[1] != y[1] or (dx != 0.0) != (dy != 0.0):
            differ += 1
        elif dx != 0.0:
            accepted += 1
            worst = max(worst, abs(dx - dy) / max(abs(dx), 1e-4))     # floor: a move of a molecule at lambda = 0 changes the energy by rounding noise only
    print("overlay vs stock:", len(a), "moves,", accepted, "accepted,", differ, "differ, worst relative energy difference", worst)
    return outs, a, b, differ, accepted, worst


@pytest.mark.parametrize("deck,n_init,n_prod,min_moves,min_accepted,min_launches", [
    ("CO2-MFI", 10000, 0, 200000, 50000, 400000),          # translation / rotation / CBMC insertion / deletion / reinsertion, Ewald
    ("XeKr-Mixture", 10000, 0, 9999, 2000, 20000),         # + IdentitySwapMove (mc_swap_moves.h:199-431), two species, tail corrections
    ("CO2_NaX_Zeolite", 5000, 5000, 9999, 500, 30000),     # + moves of a separated framework component (Na+), block pockets, cubic cell
    ("NPTMC", 300, 0, 30000, 10000, 100000),               # + VolumeMove (mc_box.h:196-320): 100 CO2 created in an empty box, ~900 volume moves
])
def test_reference_drivers_bound_to_the_c_abi_reproduce_the_stock_reference(deck, n_init, n_prod, min_moves, min_accepted, min_launches, tmp_path):
    """The drop-in, demonstrated: the reference's OWN program with its hot-path call sites bound to libgraspa_b200.so
    (oracle/overlay/: the adapter header a maintainer would add + the call-site patch, applied to a scratch copy by
    oracle/build_ref.sh overlay; RunMoves, Insertion_Body, Deletion_Body, ReinsertionMove, SingleBodyMove, the acceptance tests and
    every random-number draw remain the reference's code) against the stock reference program, CO2-MFI, 10^4 cycles, seed 0:
    the same accept/reject sequence move by move, and the reference's own FINAL energy check (its CPU and GPU total-energy
    routines run on the state the engine leaves) reports no drift."""
    for need in (REF_TRACE, REF_OVERLAY):
        if not os.path.exists(need):
            pytest.skip(f"{need} not built (oracle/build_ref.sh trace / overlay)")
    d = _deck_copy(deck, tmp_path, n_init, n_prod)
    outs, a, b, differ, accepted, worst = _stock_and_overlay(d, tmp_path)
    assert len(a) == len(b) and len(a) > min_moves, (len(a), len(b))
    assert differ == 0 and accepted > min_accepted and worst < 1e-8, (differ, accepted, worst)
    if deck == "NPTMC":
        vol = lambda text: [ln for ln in text.splitlines() if ln.startswith("VOLUME MOVE A") or ln.startswith("CYCLE:")]
        assert len(vol(outs["stock"].stdout)) >= 2 and vol(outs["stock"].stdout) == vol(outs["overlay"].stdout)
        assert int(vol(outs["stock"].stdout)[-1].split(":")[1]) > 100       # accepted volume moves
    # the engine really served the run, and the reference's own end-of-run check is content with the state it left
    assert "engine kernel launches served the reference's drivers" in outs["overlay"].stderr
    launches = int(outs["overlay"].stderr.split("graspa_b200 overlay:")[1].split()[0])
    assert launches > min_launches
    def final_total(text):
        lines = text.splitlines()
        k = max(i for i, ln in enumerate(lines) if "*** FINAL STAGE ***" in ln)
        return float([ln for ln in lines[k:k + 25] if ln.startswith("Total Energy:")][0].split(":")[1].split("(")[0])
    assert abs(final_total(outs["overlay"].stdout) - final_total(outs["stock"].stdout)) < 2e-5
    def block_total(text, header):
        lines = text.splitlines()
        k = max(i for i, ln in enumerate(lines) if header in ln)
        return float([ln for ln in lines[k:k + 25] if ln.startswith("Total Energy:")][0].split(":")[1].split("(")[0])
    # the reference's own criterion (test_examples.py:59-61), against the energy scale of the run: the NaX deck starts from overlapping
    # cations (initial total 9.1e13), so the running sum of the move deltas carries rounding of that size in BOTH programs (stock: 18.2)
    scale = abs(block_total(outs["overlay"].stdout, "*** INITIAL STAGE ***"))
    assert abs(block_total(outs["overlay"].stdout, "ENERGY DRIFT (CPU FINAL - RUNNING FINAL)")) < max(1e-3, 1e-11 * scale)


def test_widom_deck_through_the_bound_reference_prints_the_stock_averages(tmp_path):
    """Examples/Henrys_coefficient through the drop-in: the reference's Widom move (axpy.cu:163-186) is Insertion_Body, whose stages the
    overlay binds; per-block <W>, the average and the Henry coefficient of the bound program against the UNMODIFIED reference program."""
    for need in (REF_PLAIN, REF_OVERLAY):
        if not os.path.exists(need):
            pytest.skip(f"{need} not built (oracle/build_ref.sh cuda / overlay)")
    d = _deck_copy("Henrys_coefficient", tmp_path, 0, 20000)
    res = {}
    for tag, exe in (("stock", REF_PLAIN), ("overlay", REF_OVERLAY)):
        r = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=1200)
        assert r.returncode == 0, (tag, (r.stdout + r.stderr)[-2000:])
        res[tag] = _widom_lines(r.stdout)
        if tag == "overlay":
            assert "engine kernel launches served the reference's drivers" in r.stderr
    (w0, a0, k0), (w1, a1, k1) = res["stock"], res["overlay"]
    assert len(w0) == 5 and len(w1) == 5 and a0 is not None and k0 is not None
    for x, y in zip(w1, w0):
        assert abs(x - y) <= 1e-9 * abs(y) + 2e-10, (w1, w0)
    assert abs(a1 - a0) <= 1e-9 * abs(a0) + 2e-10 and abs(k1 - k0) <= 1e-9 * abs(k0) + 1e-14


def test_cbcf_moves_through_the_bound_reference_reproduce_the_stock_reference(tmp_path):
    """CB/CFC through the drop-in: the CO2-MFI deck with `CBCFProbability` added (no example deck of the reference uses CB/CFC), 8
    molecules created, 3000 cycles, stock reference against the reference bound to libgraspa_b200.so.  CBCFMove, its lambda bins and
    its Wang-Landau bookkeeping stay the reference's code (mc_cbcfc.h:224-495); the kernels under it go through gb_lambda_change_delta,
    gb_ewald_delta_lambda_change, gb_cbcf_set_scale, gb_cbcf_deletion_stage and gb_accept_lambda_change.  Same decisions move by move,
    the same CBCF statistics, no drift in the reference's own final energy check.  (The stock program accepts lambda changes only: its
    fractional insertion / deletion branches test a local SuccessConstruction that nothing sets, mc_cbcfc.h:307-318, :372-393 -- their
    first steps and reversals still run and are compared; the accepted compositions are pinned in
    tests/test_gpu_moves.py::test_cbcf_insertion_and_deletion_compositions.)"""
    for need in (REF_TRACE, REF_OVERLAY):
        if not os.path.exists(need):
            pytest.skip(f"{need} not built (oracle/build_ref.sh trace / overlay)")
    d = _deck_copy("CO2-MFI", tmp_path, 3000, 0)
    p = os.path.join(d, "simulation.input")
    out = []
    for ln in open(p).read().splitlines():
        if "CreateNumberOfMolecules" in ln:
            ln = ln.replace(" 0", " 8")
        out.append(ln)
        if "SwapProbability" in ln:
            pad = ln[:len(ln) - len(ln.lstrip())]
            out += [pad + "CBCFProbability          1.0", pad + "LambdaType               ShiMaginn"]
    open(p, "w").write("\n".join(out) + "\n")
    outs, a, b, differ, accepted, worst = _stock_and_overlay(d, tmp_path)
    assert len(a) == len(b) and len(a) > 50000, (len(a), len(b))
    assert differ == 0 and accepted > 5000 and worst < 1e-8, (differ, accepted, worst)
    def stats(text):
        return [ln for ln in text.splitlines() if ln.startswith("CBCF ")]
    st = stats(outs["stock"].stdout)
    assert len(st) >= 8 and st == stats(outs["overlay"].stdout), (st, stats(outs["overlay"].stdout))
    val = lambda key: int([ln for ln in st if ln.startswith(key)][0].split(":")[1])
    assert val("CBCF Lambda Accepted") > 1000 and val("CBCF Insertion Performed") > 1000 and val("CBCF Deletion Performed") > 10
    assert val("CBCF Insertion Accepted") == 0 and val("CBCF Deletion Accepted") == 0          # the reference's dead branches, see above
    def block_total(text, header):
        lines = text.splitlines()
        k = max(i for i, ln in enumerate(lines) if header in ln)
        return float([ln for ln in lines[k:k + 25] if ln.startswith("Total Energy:")][0].split(":")[1].split("(")[0])
    assert abs(block_total(outs["overlay"].stdout, "*** FINAL STAGE ***") - block_total(outs["stock"].stdout, "*** FINAL STAGE ***")) < 2e-5
    assert abs(block_total(outs["overlay"].stdout, "ENERGY DRIFT (CPU FINAL - RUNNING FINAL)")) < 1e-3
