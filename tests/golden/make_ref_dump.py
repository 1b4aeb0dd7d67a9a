"""Turns the text dumps of the instrumented reference program (oracle/build_ref.sh dump, run on the GPU box by
scripts/make_ref_dump.sh -> gpurun_out/ref_dump_<deck>.txt) into the committed fixtures

  tests/golden/ref_dump_widom_A.npz     first 150 Insertion_Body calls of Examples/Henrys_coefficient (seed 0): pool randoms, trial
                                        positions, per-trial energies of the reference's CUDA kernel, log Boltzmann factors,
                                        uniforms, selections, Rosenbluth sums, chain positions, final weight and energies
  tests/golden/ref_dump_select.npz      the Boltzmann-selection records (log factors, uniform, selected, sum) of every dumped
                                        stage of CO2-MFI, CO2_NaX_Zeolite and BlockPocket (few survivors, pockets, adsorbates)
  tests/golden/ref_dump_pockets.npz     BlockedPocket() verdicts of CO2_NaX_Zeolite and BlockPocket with the boxes and pocket lists

    python -m tests.golden.make_ref_dump          (build container; needs the dumps under gpurun_out/)
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")


def parse(path):
    ins = []; bp = []; cur = None
    with open(path) as f:
        lines = f.read().splitlines()
    k = 0
    while k < len(lines):
        t = lines[k].split()
        if not t:
            k += 1; continue
        tag = t[0]
        if tag == "BEGIN":
            cur = dict(); ins.append(cur)
        elif tag == "BP":
            bp.append([float(t[1]), float(t[2]), float(t[3]), float(t[4]), float(t[5])])
        elif cur is None:
            pass
        elif tag == "FB":
            n = int(t[2]); cur["fb_off"] = int(t[3])
            rows = [lines[k + 1 + i].split() for i in range(n)]
            cur["fb_rnd"] = np.array([[float(x) for x in r[1:4]] for r in rows]); cur["fb_pos"] = np.array([[float(x) for x in r[5:8]] for r in rows])
            k += n
        elif tag in ("FBE", "CHE"):
            n = int(t[1])
            rows = np.array([[float(x) for x in lines[k + 1 + i].split()[1:]] for i in range(n)]).reshape(n, 6)
            cur[tag.lower()] = rows            # trial index, log Boltzmann factor, HGVDW, HGReal, GGVDW, GGReal
            k += n
        elif tag in ("FBS", "CHS"):
            cur[tag.lower()] = np.array([float(x) for x in t[1:5]])      # good, selected (among survivors), Rosenbluth sum, uniform
        elif tag == "CH":
            no, cs = int(t[1]), int(t[2]); cur["ch_off"] = int(t[3]); cur["ch_fb"] = np.array([float(x) for x in t[4:7]])
            cur["ch_rnd"] = np.array([[float(x) for x in lines[k + 1 + i].split()[1:4]] for i in range(no)])
            cur["ch_pos"] = np.array([[float(x) for x in lines[k + 1 + no + i].split()[1:4]] for i in range(no * cs)])
            k += no + no * cs
        elif tag == "INS":
            cur["ins"] = np.array([float(x) for x in t[1:9]])           # W, HGVDW, HGReal, GGVDW, GGReal, GGEwaldE, HGEwaldE, TailE
        k += 1
    return ins, np.array(bp).reshape(-1, 5)


def widom_fixture(ins, n):
    ins = ins[:n]
    ntr = len(ins[0]["fb_rnd"])
    out = dict(fb_rnd=np.array([i["fb_rnd"] for i in ins]), fb_pos=np.array([i["fb_pos"] for i in ins]), fb_off=np.array([i["fb_off"] for i in ins]))
    def surv(key, width):
        a = np.full((len(ins), width, 6), np.nan); cnt = np.zeros(len(ins), dtype=np.int64)
        for k, i in enumerate(ins):
            if key in i:
                m = len(i[key]); a[k, :m] = i[key]; cnt[k] = m
        return a, cnt
    out["fbe"], out["fbe_n"] = surv("fbe", ntr)
    out["fbs"] = np.array([i.get("fbs", np.full(4, np.nan)) for i in ins])
    has_ch = np.array(["ch_rnd" in i for i in ins])
    nor = max((len(i["ch_rnd"]) for i in ins if "ch_rnd" in i), default=0)
    ncp = max((len(i["ch_pos"]) for i in ins if "ch_pos" in i), default=0)
    out["has_chain"] = has_ch
    out["ch_rnd"] = np.array([i["ch_rnd"] if "ch_rnd" in i else np.full((nor, 3), np.nan) for i in ins])
    out["ch_pos"] = np.array([i["ch_pos"] if "ch_pos" in i else np.full((ncp, 3), np.nan) for i in ins])
    out["ch_fb"] = np.array([i.get("ch_fb", np.full(3, np.nan)) for i in ins])
    out["ch_off"] = np.array([i.get("ch_off", -1) for i in ins])
    out["che"], out["che_n"] = surv("che", max(nor, 1))
    out["chs"] = np.array([i.get("chs", np.full(4, np.nan)) for i in ins])
    out["has_ins"] = np.array(["ins" in i for i in ins])
    out["ins"] = np.array([i.get("ins", np.full(8, np.nan)) for i in ins])
    return out


def select_records(ins):
    """every dumped selection: (log Boltzmann factors of the survivors, uniform, good, selected, Rosenbluth sum)"""
    logs, meta = [], []
    for i in ins:
        for e, s in (("fbe", "fbs"), ("che", "chs")):
            if e in i and s in i and len(i[e]) > 0:
                lb = np.full(16, np.nan); lb[:len(i[e])] = i[e][:, 1]
                logs.append(lb); meta.append([len(i[e]), i[s][3], i[s][0], i[s][1], i[s][2]])
    return np.array(logs), np.array(meta)


def deck_json(name):
    exe = os.path.join(ROOT, "graspa_b200", "host", "graspa_b200_mc")
    return json.loads(subprocess.check_output([exe, "--dump-deck", f"/root/reference/Examples/{name}"]))


def main():
    src = os.path.join(ROOT, "gpurun_out")
    ins, _ = parse(os.path.join(src, "ref_dump_Henrys_coefficient.txt"))
    np.savez_compressed(os.path.join(HERE, "ref_dump_widom_A.npz"), **widom_fixture(ins, 150))
    logs, meta, origin = [], [], []
    pockets = {}
    for k, deck in enumerate(("CO2-MFI", "CO2_NaX_Zeolite", "BlockPocket")):
        ins, bp = parse(os.path.join(src, f"ref_dump_{deck}.txt"))
        lg, mt = select_records(ins)
        logs.append(lg); meta.append(mt); origin.append(np.full(len(lg), k))
        if len(bp):
            d = deck_json(deck)
            comp = int(bp[0, 0]); a = d["adsorbates"][comp - len(d["framework"])]
            tag = deck.replace("-", "_")
            # distinct verdict-relevant calls only: keep every blocked one and an equal number of free ones, in order
            blocked = bp[:, 4] > 0.5
            keep = np.zeros(len(bp), dtype=bool); keep[blocked] = True
            free = np.flatnonzero(~blocked)[:max(400, int(blocked.sum()))]; keep[free] = True
            pockets.update({f"{tag}_cell": np.array(d["cell"]), f"{tag}_centers": np.array(a["pocket_centers"]).reshape(-1, 3),
                            f"{tag}_radii": np.array(a["pocket_radii"]), f"{tag}_invert": int(a["invert_pockets"]), f"{tag}_calls": bp[keep][:, 1:5]})
    np.savez_compressed(os.path.join(HERE, "ref_dump_select.npz"), logs=np.concatenate(logs), meta=np.concatenate(meta), origin=np.concatenate(origin))
    np.savez_compressed(os.path.join(HERE, "ref_dump_pockets.npz"), **pockets)
    for f in ("ref_dump_widom_A.npz", "ref_dump_select.npz", "ref_dump_pockets.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
