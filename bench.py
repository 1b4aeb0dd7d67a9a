#!/usr/bin/env python
"""bench.py -- Widom insertions/s of the B200-native gRASPA energy engine (BASELINE.json north_star).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU arm: oracle port on the host cores)

Workload (config.workload): BASELINE.json configs[4] -- synthetic 4x4x4 Mg-MOF-74 supercell (3456 framework atoms,
triclinic, Ewald 5800 k-vectors), CO2 Widom insertions with 10 trial positions + 10 trial orientations, FP64.
One "step" = one batch of --batch ghost insertions per GPU through Insertion_Body's whole path (first bead, chain,
Rosenbluth selection, Ewald Fourier delta, tail, block sums).  Insertions are independent, so ranks take disjoint
index ranges (weak scaling: fixed per-GPU batch) and only the block sums are all-reduced over NCCL.

value : inputs (random pool, uniforms) already resident in HBM; sums read back.
e2e   : the same step through the C ABI with HOST (pinned) buffers: H2D of the randoms and D2H of the sums inside
        the timed region.
Timing: CUDA events recorded on the engine's own stream (made torch's current stream), max over ranks.
Extra keys (N=1 only, skipped with --no-secondary): "gcmc" -- cycles/s of the sequential Markov chain on the reference's
CO2-MFI example through the host driver (one kernel per move), with the reference's own CUDA build (oracle/_ref) timed
beside it on the same GPU when the binary is present.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.conftest import load_config  # noqa: E402  (fixture loader only: committed .npz, no oracle)

# dram bytes per insertion of k_widom_pair measured by ncu (profiles/r1_pair_kernel.md); bench.py cannot run under ncu itself
NCU_DRAM_BYTES_PER_INSERTION = 501.4

METRIC = "widom_insertions_per_s"
UNIT = "insertions/s"
WORKLOAD = "synthetic 4x4x4 Mg-MOF-74 supercell (3456 atoms, 5800 k), CO2 Widom, 10 positions + 10 orientations, FP64"


def flops_per_insertion(counts, n, triclinic=True, natoms_mol=3):
    """SURVEY section 8(d) nominal flop count: F_pair = Npairs*(44|20) + Nvdw*17 + Ncoul*9 + 2*Nin;
    Fourier: (17*n_atoms + 55) per active k."""
    pairs, nvdw, ncoul, nin, act_atoms = [float(c) / n for c in counts]
    f_pair = pairs * (44.0 if triclinic else 20.0) + nvdw * 17.0 + ncoul * 9.0 + 2.0 * nin
    nact = act_atoms / natoms_mol
    f_k = nact * (17.0 * natoms_mol + 55.0)
    return f_pair, f_k


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.stop_flag = False; self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def host_cores():
    """threads the CPU arm uses: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which would
    otherwise silently make the baseline single-threaded)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def oracle_setup(box, ff, s, z, comp):
    from oracle import oracle as orc
    orc.build()
    return orc, orc.WidomSetup(box, ff, s, comp, float(z["beta"]), 10, 10, z["sf_ads"], z["sf_fw"])


def cpu_sample(box, ff, s, z, comp, target_s=12.0, seed=7):
    """oracle Widom batch on the host cores, bounded to ~target_s seconds; also yields the per-insertion pair counts"""
    orc, ws = oracle_setup(box, ff, s, z, comp)
    rng = np.random.default_rng(seed)
    nt = host_cores()
    n0 = 64 * max(1, nt // 8)
    t0 = time.perf_counter(); orc.widom_batch(ws, rng.random((n0, 20, 3)), rng.random((n0, 2)), nthreads=nt); dt = time.perf_counter() - t0
    n = int(max(n0, min(200000, n0 * target_s / max(dt, 1e-3))))
    rnd = rng.random((n, 20, 3)); uni = rng.random((n, 2))
    t0 = time.perf_counter(); out, stage, counts = orc.widom_batch(ws, rnd, uni, nthreads=nt); dt = time.perf_counter() - t0
    return dict(value=n / dt, n=n, seconds=dt, cores=nt, counts=[int(c) for c in counts], mean_W=float(out[:, 0].mean()))


def reference_host_sample(box, ff, s, z, comp, target_s=6.0, seed=11):
    """The reference's OWN host routines (PBC / VDW / CoulombReal of maths.cuh:427-500, compiled in place into
    oracle/_ref/libgraspa_ref_host.so by oracle/build_ref.sh) on the trial-energy part of the same workload: first-bead style
    trial atoms against the framework, one host thread (the reference's host code is serial).  Reported as pairs/s and as the
    insertions/s it would sustain on the pair part alone (103 680 pairs per insertion in config E)."""
    from oracle import oracle as orc
    from graspa_b200.types import CBMC_INSERTION
    if not orc.ref_available():
        return {"unavailable": "oracle/_ref/libgraspa_ref_host.so not built (needs /root/reference at build time)"}
    rng = np.random.default_rng(seed)
    new_molid = int(s.natoms[comp]) // int(s.molsize[comp])
    nsys = int(s.natoms[:s.nhost].sum())
    n = 64; dt = 0.0; t = None
    while True:
        t = orc.trial_positions(box, s, CBMC_INSERTION, comp, 0, n, rng.random((n, 3)))
        t0 = time.perf_counter(); orc.ref_trial_energies(box, ff, s, n, 1, t, comp, new_molid); dt = time.perf_counter() - t0
        if dt > 0.5 * target_s or n >= 1 << 20:
            break
        n = int(min(1 << 20, max(2 * n, n * 0.8 * target_s / max(dt, 1e-4))))
    pairs = float(n) * nsys
    per_ins = float(nsys) * (10 + 10 * (int(s.molsize[comp]) - 1))
    return {"kind": "reference", "cores": 1, "pairs_per_s": pairs / dt, "insertions_per_s_pair_part": pairs / dt / per_ins,
            "sample": f"{n} trial atoms x {nsys} framework atoms in {dt:.1f} s through ref_trial_energies (the reference's PBC/VDW/CoulombReal)"}


def _tail_json(text):
    for ln in reversed(text.strip().splitlines()):
        ln = ln.strip()
        if ln.startswith("{") and ln.endswith("}"):
            try:
                return json.loads(ln)
            except Exception:
                pass
    return None


def _deck_pair(name, init, prod, unit):
    """One example deck of the reference through the host driver (graspa_b200/host/graspa_b200_mc) and, when present, through
    the reference's own CUDA program built from /root/reference by oracle/build_ref.sh, same seed, same cycle counts.
    unit "cycles/s": Monte Carlo cycles of the sequential Markov chain; "insertions/s": Widom moves of a Widom-only deck."""
    import shutil
    import tempfile
    deck = os.path.join(ROOT, "oracle", "_ref", "examples", name)
    drv = os.path.join(ROOT, "graspa_b200", "host", "graspa_b200_mc")
    n = init + prod
    out = {"workload": f"{name} example deck, {init} initialisation + {prod} production cycles, seed 0", "unit": unit}
    if not (os.path.isdir(deck) and os.path.exists(drv)):
        out["unavailable"] = "host driver or example deck not built (oracle/build_ref.sh, make -C graspa_b200/host)"
        return out
    try:
        r = subprocess.run([drv, deck, "--init", str(init), "--equil", "0", "--prod", str(prod)], capture_output=True, text=True, timeout=300)
        j = _tail_json(r.stdout.split("host time inside")[0]) or {}
        out["value"] = j.get("cycles_per_s"); out["moves_per_s"] = j.get("moves_per_s"); out["kernel_launches"] = j.get("kernel_launches")
        out["seconds"] = j.get("seconds"); out["path"] = j.get("widom_path") if unit == "insertions/s" else j.get("move_calls")
        for ln in r.stdout.splitlines():
            if ln.startswith("ENERGY DRIFT"):
                out["energy_drift"] = float(ln.split(":")[-1])
            if ln.startswith("Averaged Rosenbluth Weight:"):
                out["mean_W"] = float(ln.split(":")[1].split("+/-")[0])
            if ln.startswith("FINAL"):
                out["final_total_energy"] = float(ln.split("Total:")[-1])
    except Exception as ex:  # noqa: BLE001
        out["error"] = str(ex)
    ref = os.path.join(ROOT, "oracle", "_ref", "graspa_ref_cuda.x")
    if os.path.exists(ref):
        d = tempfile.mkdtemp(prefix="deck_ref_")
        try:
            for f in os.listdir(deck):
                shutil.copy(os.path.join(deck, f), d)
            inp = os.path.join(d, "simulation.input")
            os.chmod(inp, 0o644)
            txt = open(inp).read().splitlines()
            sub = {"NumberOfInitializationCycles": init, "NumberOfEquilibrationCycles": 0, "NumberOfProductionCycles": prod}
            txt = [next((f"{k} {v}" for k, v in sub.items() if t.startswith(k)), t) for t in txt]
            open(inp, "w").write("\n".join(txt) + "\n")
            t0 = time.perf_counter()
            r = subprocess.run([ref], cwd=d, capture_output=True, text=True, timeout=600)
            wall = time.perf_counter() - t0
            took = [ln for ln in r.stdout.splitlines() if ln.startswith("Work took")]
            if not took and os.path.exists(os.path.join(d, "output.txt")):
                took = [ln for ln in open(os.path.join(d, "output.txt")).read().splitlines() if ln.startswith("Work took")]
            secs = float(took[-1].split()[2]) if took else wall
            out["reference_cuda"] = {"value": n / secs, "seconds": secs,
                                     "what": "the reference's own CUDA program (sm_100 build of /root/reference) on the same GPU, same deck, same seed"}
            if out.get("value"):
                out["speedup_vs_reference_cuda"] = out["value"] / (n / secs)
        except Exception as ex:  # noqa: BLE001
            out["reference_cuda"] = {"error": str(ex)}
        finally:
            shutil.rmtree(d, ignore_errors=True)
    return out


def gcmc_secondary(cycles=5000):
    """GCMC cycles/s of the sequential Markov chain on the reference's CO2-MFI example (one k_move launch per move)"""
    return _deck_pair("CO2-MFI", cycles, 0, "cycles/s")


def more_secondaries():
    """the other example decks of BASELINE.json's configs, each beside the reference's own CUDA build:
    Xe/Kr mixture with identity swaps, NaX with movable cations and block pockets, and the Henry-coefficient Widom deck
    (the host driver replays the reference's random stream there: RNG-exact batched insertions)"""
    return {"gcmc_xekr": _deck_pair("XeKr-Mixture", 20000, 0, "cycles/s"),
            "gcmc_nax": _deck_pair("CO2_NaX_Zeolite", 5000, 0, "cycles/s"),
            "widom_henry": _deck_pair("Henrys_coefficient", 0, 20000, "insertions/s"),
            # the total-energy path in a loop: NPT volume moves (one box) and the two-box Gibbs ensemble
            "npt_co2": _deck_pair("NPTMC", 600, 0, "cycles/s"),
            "gibbs_co2": _deck_pair("NVT-Gibbs", 20, 0, "cycles/s")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    box, ff, s, z = load_config("E")
    comp = int(z["comp"])
    orc, ws = oracle_setup(box, ff, s, z, comp)
    rng = np.random.default_rng(17)
    # bounded sample per step: sized from a probe so that warmup+steps stay within a few minutes
    nt = host_cores()
    n0 = 64 * max(1, nt // 8)
    t0 = time.perf_counter(); orc.widom_batch(ws, rng.random((n0, 20, 3)), rng.random((n0, 2)), nthreads=nt); rate = n0 / (time.perf_counter() - t0)
    per_step = int(max(64, min(100000, rate * 60.0 / max(1, args.steps + args.warmup))))
    rnd = rng.random((per_step, 20, 3)); uni = rng.random((per_step, 2))
    for _ in range(args.warmup):
        orc.widom_batch(ws, rnd, uni, nthreads=nt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.widom_batch(ws, rnd, uni, nthreads=nt)
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "insertions_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": nt, "kind": "port",
                             "sample": f"{per_step} insertions/step x {args.steps} steps of the same workload, OpenMP over insertions"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=400000, help="Widom insertions per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the GCMC cycles/s section")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from graspa_b200 import engine
    from graspa_b200.shard import reduce_block_sums

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: graspa_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    box, ff, s, z = load_config("E")
    comp = int(z["comp"]); B = args.batch
    eng = engine.Engine(local).setup(box, ff, s, float(z["beta"]), 10, 10)
    # the engine launches on its own stream: make it torch's current stream so that events, copies and NCCL share it
    ext = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
    torch.cuda.set_stream(ext)
    eng.total_ewald(store=True)                                  # structure factors built on the GPU
    eng.set_exclusion_constants(comp, float(z["excl"][0]), float(z["excl"][1]))
    fp64_peak = eng.measure_fp64_peak()
    hbm_peak, hbm_src = 6458.7, "fallback of B200_PROFILING.md"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:  # noqa: BLE001
        pass

    # ---- synthetic inputs: this rank's contiguous index range of the job
    gen = torch.Generator(device="cuda"); gen.manual_seed(1234 + rank)
    d_pool = torch.rand((B * 20, 3), dtype=torch.float64, device="cuda", generator=gen)
    d_uni = torch.rand((B, 2), dtype=torch.float64, device="cuda", generator=gen)
    h_pool = torch.empty((B * 20, 3), dtype=torch.float64).pin_memory(); h_pool.copy_(d_pool)
    h_uni = torch.empty((B, 2), dtype=torch.float64).pin_memory(); h_uni.copy_(d_uni)
    torch.cuda.synchronize()

    # weak scaling: every rank owns B insertions of a job of B*world; bins are assigned on the global index
    my_shard = (rank * B, B * world)
    dev = torch.device("cuda", local)

    def step_device():
        sums = eng.widom_batch_device(comp, B, d_pool.data_ptr(), B * 20, d_uni.data_ptr(), shard=my_shard)
        return reduce_block_sums(sums, device=dev)          # NCCL all-reduce of the 5 x 12 block sums (identity at N=1)

    def step_e2e():
        _, _, sums = eng.widom_batch(comp, h_pool.numpy(), h_uni.numpy(), want_outputs=False, shard=my_shard)
        return reduce_block_sums(sums, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        barrier(); ev0.record(ext)
        for _ in range(steps):
            out = fn()
        ev1.record(ext); barrier()
        dt = ev0.elapsed_time(ev1) * 1e-3
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        return dt, out

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local); sampler.start()
    eng.timing_enable(True); eng.timing_read(2, reset=True); eng.launch_count(reset=True)
    dt, sums = timed(step_device, args.steps)
    launches = eng.launch_count()
    ms_pair, n_pair = eng.timing_read(0); ms_ew, n_ew = eng.timing_read(1)
    eng.timing_enable(False)
    for _ in range(2):
        step_e2e()
    dt_e2e, sums_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True; sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = B * world * args.steps / dt
    e2e_value = B * world * args.steps / dt_e2e
    total_count = float(sums[:, 2].sum())
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "insertions_per_gpu_per_step": B, "global_insertions_per_step": B * world,
                       "parallelism": f"widom-shard x{world}", "cache": "inputs per step (random pool %.0f MB) exceed the 126 MB L2" % (B * 20 * 24 / 1e6),
                       "mean_W": float(sums[:, 0].sum() / max(total_count, 1.0)), "failed_fraction": float(sums[:, 10].sum() / max(total_count, 1.0))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * 20 * 24 + B * 16), "d2h_bytes_per_step": int(5 * 12 * 8)},
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "kernels": {"k_widom_pair_ms": ms_pair / max(n_pair, 1), "k_widom_ewald_ms": ms_ew / max(n_ew / 2, 1),
                        "how": "CUDA events around each launch on the engine's stream, averaged over the timed steps"}}
    # ---- CPU baseline + algorithmic flop count on a bounded sample of the same workload
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_sample(box, ff, s, z, comp)
        line["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                                "sample": f"{cpu['n']} insertions of the same workload in {cpu['seconds']:.1f} s, OpenMP over insertions"}
        try:
            line["cpu_baseline"]["reference_host_routines"] = reference_host_sample(box, ff, s, z, comp)
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"]["reference_host_routines"] = {"error": str(ex)}
        f_pair, f_k = flops_per_insertion(cpu["counts"], cpu["n"])
        t_pair = ms_pair / max(n_pair, 1) * 1e-3
        achieved = f_pair * B / t_pair / 1e12
        line["roofline"] = {"bound": "fp64", "kernel": "k_widom_pair", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": achieved / fp64_peak, "traffic": None,
                            "peak_source": "DFMA microbenchmark run in this process (gb_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                            "flop_per_insertion": f_pair, "launch_ms": t_pair * 1e3, "share_of_step": ms_pair / max(ms_pair + ms_ew, 1e-9),
                            "ewald_kernel": {"achieved": f_k * B / (ms_ew / max(n_ew / 2, 1) * 1e-3) / 1e12, "unit": "TFLOP/s", "launch_ms": ms_ew / max(n_ew / 2, 1)},
                            "hbm": {"algorithmic_bytes_per_insertion": 20 * 24 + 16 + 8 * 8,
                                    "achieved_gbs": (20 * 24 + 16 + 8 * 8) * B / t_pair / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                                    "note": "framework atoms, erfc and LJ tables are shared-memory resident: the pair kernel reads 496 B and writes 64 B per insertion, "
                                            "so HBM (frac %.4f) is not the binding roof; the FP64 pipe is" % ((20 * 24 + 16 + 8 * 8) * B / t_pair / 1e9 / hbm_peak)}}
        line["roofline"]["traffic"] = NCU_DRAM_BYTES_PER_INSERTION * B
        line["roofline"]["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one k_widom_pair launch from profiles/ (ncu --set full), "
                                              "scaled per insertion: %.0f B" % NCU_DRAM_BYTES_PER_INSERTION)
    if world == 1 and not args.no_secondary:
        torch.cuda.set_stream(torch.cuda.default_stream())
        eng.close()
        line["gcmc"] = gcmc_secondary()
        line.update(more_secondaries())
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
