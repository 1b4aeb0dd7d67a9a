"""Independent simulation boxes across the GPUs of one node (BASELINE.json configs[3]).

A GCMC box is one sequential Markov chain and does not shard; what does partition is the SET of boxes -- isotherm
points (one pressure each) or mixture boxes.  The reference runs them one after the other on one GPU
(`Run_Simulation_MultipleBoxes`, axpy.cu:593-625); here every box is its own host-driver process bound to one GPU
(`graspa_b200_mc --device g --pressure P`), boxes are dealt round-robin over the GPUs, each GPU works through its queue,
and nothing is exchanged until the per-box results (loadings, energies, move statistics) are gathered.  No collective.

    python -m graspa_b200.boxes <deck dir> --pressures 1e4,3e4,1e5,3e5 --gpus 8 --init 20000 --prod 20000
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "graspa_b200", "host", "graspa_b200_mc")


def assign_boxes(n_boxes: int, n_gpus: int):
    """round-robin deal: box b runs on GPU b % n_gpus; -> list (one entry per GPU) of box indices in execution order"""
    if n_gpus < 1 or n_boxes < 0:
        raise ValueError("bad box / GPU count")
    return [list(range(g, n_boxes, n_gpus)) for g in range(n_gpus)]


def _json_lines(text):
    out = []
    for ln in text.splitlines():
        ln = ln.strip()
        if ln.startswith("{") and ln.endswith("}"):
            try:
                out.append(json.loads(ln))
            except ValueError:
                pass
    return out


def run_boxes(deck, points, gpus=1, init=None, prod=None, equil=0, driver=DRIVER, extra=(), timeout=3600, devices=None):
    """points: list of dicts with optional keys pressure (Pa), temperature (K), seed.  -> list of per-box results in the
    order of `points`: {"box", "gpu", "point", "seconds", "returncode", "loading": [...], "run": {...}, "final_total_energy"}.
    One worker thread per GPU feeds that GPU's queue; the driver processes do the work.  devices: the device of worker g
    (default: entry g of CUDA_VISIBLE_DEVICES, else g) -- a rank of a multi-process launch passes its own device."""
    if not os.path.exists(driver):
        raise FileNotFoundError(f"{driver} is missing: make -C graspa_b200/csrc && make -C graspa_b200/host (there is no CPU path)")
    queues = assign_boxes(len(points), gpus)
    results = [None] * len(points)

    # each driver process sees only its own GPU: CUDA start-up enumerates every visible device, which costs seconds per
    # process on an 8-GPU node and is paid once per box
    visible = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]

    def worker(g):
        env = dict(os.environ)
        dev = devices[g] if devices is not None else g
        env["CUDA_VISIBLE_DEVICES"] = visible[dev] if dev < len(visible) else str(dev)
        for b in queues[g]:
            pt = points[b]
            cmd = [driver, deck, "--device", "0"]
            if init is not None:
                cmd += ["--init", str(init)]
            cmd += ["--equil", str(equil)]
            if prod is not None:
                cmd += ["--prod", str(prod)]
            if "pressure" in pt:
                cmd += ["--pressure", repr(float(pt["pressure"]))]
            if "temperature" in pt:
                cmd += ["--temperature", repr(float(pt["temperature"]))]
            if "seed" in pt:
                cmd += ["--seed", str(int(pt["seed"]))]
            cmd += list(extra)
            t0 = time.perf_counter()
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
                rc, text, err = r.returncode, r.stdout, r.stderr
            except subprocess.TimeoutExpired as ex:
                rc, text, err = -9, ex.stdout or "", "timeout"
            js = _json_lines(text)
            res = {"box": b, "gpu": g, "point": pt, "seconds": time.perf_counter() - t0, "returncode": rc}
            for j in js:
                if "loading" in j:
                    res["loading"] = j["loading"]; res["pressure_pa"] = j.get("pressure_pa"); res["temperature"] = j.get("temperature")
                if "cycles_per_s" in j:
                    res["run"] = j
            for ln in text.splitlines():
                if ln.startswith("FINAL"):
                    res["final_total_energy"] = float(ln.split("Total:")[-1])
                if ln.startswith("ENERGY DRIFT"):
                    res["energy_drift"] = float(ln.split(":")[-1])
            if rc != 0:
                res["stderr"] = err[-500:]
            results[b] = res

    threads = [threading.Thread(target=worker, args=(g,), daemon=True) for g in range(gpus)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    wall = time.perf_counter() - t0
    return results, wall


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("deck")
    ap.add_argument("--pressures", default="", help="comma-separated pressures in Pa: one box per value")
    ap.add_argument("--replicas", type=int, default=0, help="instead of pressures: N replicas of the deck with seeds 0..N-1")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--init", type=int, default=None)
    ap.add_argument("--prod", type=int, default=None)
    args = ap.parse_args(argv)
    if args.pressures:
        points = [{"pressure": float(p)} for p in args.pressures.split(",") if p]
    else:
        points = [{"seed": k} for k in range(max(1, args.replicas))]
    results, wall = run_boxes(args.deck, points, gpus=args.gpus, init=args.init, prod=args.prod)
    cycles = sum((r.get("run") or {}).get("cycles", 0) for r in results)
    print(json.dumps({"boxes": len(points), "gpus": args.gpus, "wall_seconds": wall, "total_cycles": cycles,
                      "aggregate_cycles_per_s": cycles / wall if wall > 0 else None, "results": results}))
    return 0 if all(r and r["returncode"] == 0 for r in results) else 1


if __name__ == "__main__":
    sys.exit(main())
