#!/bin/bash
# GPU box: move-by-move comparison of the accept/reject sequence.  The reference's CUDA program with ONE added line
# (oracle/build_ref.sh trace: RunMoves appends "component movetype deltaE" per move) against graspa_b200_mc --trace,
# same deck, same seed.  Usage: scripts/compare_trace.sh <deck> <init cycles> [production cycles]
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; NI="$2"; NP="${3:-0}"
OUT="$ROOT/gpurun_out/trace_$NAME"; rm -rf "$OUT"; mkdir -p "$OUT/ref"
cp -r "$ROOT/oracle/_ref/examples/$NAME/"* "$OUT/ref/"; chmod -R u+w "$OUT/ref/"*
sed -i "s/^NumberOfInitializationCycles.*/NumberOfInitializationCycles $NI/; s/^NumberOfEquilibrationCycles.*/NumberOfEquilibrationCycles 0/; s/^NumberOfProductionCycles.*/NumberOfProductionCycles $NP/" "$OUT/ref/simulation.input"
( cd "$OUT/ref" && GRASPA_TRACE="$OUT/ref_trace.txt" timeout 1500 "$ROOT/oracle/_ref/graspa_ref_cuda_trace.x" > output.txt 2> stderr.txt; echo "reference exit $?" )
timeout 1500 "$ROOT/graspa_b200/host/graspa_b200_mc" "$OUT/ref" --init "$NI" --equil 0 --prod "$NP" --trace "$OUT/our_trace.txt" > "$OUT/ours.txt" 2>&1; echo "ours exit $?"
python - "$OUT/ref_trace.txt" "$OUT/our_trace.txt" <<'PY'
import sys
ref = [l.split() for l in open(sys.argv[1])]
our = [l.split() for l in open(sys.argv[2])]
n = min(len(ref), len(our))
mism = None; acc = 0; worst = 0.0; comp_mism = 0; nmism = 0; zero_swaps = 0; box_moves = 0; box_acc = 0
for k in range(n):
    rc, rd = int(ref[k][0]), float(ref[k][2])
    oc, oa, od = int(our[k][2]), int(our[k][4]), float(our[k][5])
    ra = 1 if rd != 0.0 else 0          # the reference returns a zeroed MoveEnergy for a rejected move
    acc += ra
    if our[k][1] == "identity_swap": oc = rc      # the reference's TempVal.component is the NEW species until the retrace starts, ours prints the OLD one
    if rc != oc: comp_mism += 1
    if our[k][1] in ("volume", "gibbs_volume", "gibbs_transfer"):
        # these moves add to the running energy on their own (mc_box.h:288, move_struct.h:488-489): RunMoves, and with it the
        # reference's trace line, carries a zero whether they were accepted or not.  Their decisions show in every later
        # move (the state differs otherwise) and in the attempt / acceptance counters of the two outputs.
        box_moves += 1; box_acc += oa
        continue
    if oa == 1 and od == 0.0 and our[k][1] == "identity_swap" and ra == 0:
        zero_swaps += 1                 # an accepted swap of a monatomic molecule into its own species changes no energy at all
        continue
    if ra != oa or rc != oc:
        nmism += 1
        if mism is None: mism = k
    if ra and oa: worst = max(worst, abs(rd - od) / max(abs(rd), 1e-300))
print(f"moves: reference {len(ref)}, ours {len(our)}; compared {n}; accepted (non-zero energy change) in the reference {acc}")
print(f"moves whose component or accept/reject decision differs: {nmism} (first: {mism}); component mismatches: {comp_mism}")
print(f"accepted same-species identity swaps with exactly zero energy change (indistinguishable from a rejection in the reference's trace): {zero_swaps}")
print(f"volume / Gibbs moves (decision not in the reference's trace line, see the counters of the outputs): {box_moves}, accepted here {box_acc}")
print(f"largest relative difference of an accepted move's energy change: {worst:.3e}")
PY
