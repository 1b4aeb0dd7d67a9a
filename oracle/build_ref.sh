#!/bin/bash
# TEST INFRASTRUCTURE ONLY.
#
# Builds, from the reference sources WHERE THEY LIE under /root/reference
# (nothing is copied into this repository), two artefacts into oracle/_ref/
# (git-ignored, but shipped to the GPU box by gpurun):
#
#   oracle/_ref/libgraspa_ref_host.so   host harness around the reference's own
#                                       PBC / VDW / CoulombReal / Ewald_Total /
#                                       tail-correction routines (ref_harness.cu
#                                       #includes the reference headers)
#   oracle/_ref/graspa_ref_cuda.x       the reference's own CUDA program, built
#                                       for sm_100 (the "reference CUDA build on
#                                       the same B200" baseline of BASELINE.md 2a)
#   oracle/_ref/examples/<name>/        the example input decks the baseline
#                                       binary is run on (input data, not code)
#
# The reference's own build system (nvc++ scripts) is NOT used: only nvcc on
# its five translation units, as BASELINE.md 2a records.  The full-binary
# build needs one narrowing cast (mc_swap_moves.h:268), applied to a scratch
# copy under /tmp, never to /root/reference.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${GRASPA_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
WHAT="${1:-all}"
if [ ! -d "$REF/src_clean" ]; then
  echo "build_ref.sh: $REF/src_clean not present; keeping prebuilt oracle/_ref as is"
  exit 0
fi
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"

if [ "$WHAT" = "all" ] || [ "$WHAT" = "host" ]; then
  echo "[build_ref] host harness"
  "$NVCC" -O2 -std=c++20 -arch=sm_100 --expt-relaxed-constexpr -w \
      -Xcompiler -fopenmp -Xcompiler -fPIC -shared -x cu \
      -I"$REF/src_clean" "$HERE/ref_harness.cu" -o "$OUT/libgraspa_ref_host.so"
fi

if [ "$WHAT" = "all" ] || [ "$WHAT" = "examples" ]; then
  echo "[build_ref] example decks"
  for ex in Henrys_coefficient CO2-MFI CO2_NaX_Zeolite XeKr-Mixture Ar_MgMOF74_UFF Bae-Mixture BlockPocket CO2_MgMOF74_UFF Tail-Correction Ionic-MOF-mixtures TIP4PEW-MgMOF-LJ1264; do
    mkdir -p "$OUT/examples/$ex"
    # inputs only: no committed outputs, no restart dumps
    find "$REF/Examples/$ex" -maxdepth 1 -type f \
        \( -name '*.def' -o -name '*.cif' -o -name '*.block' -o -name 'simulation.input' \) \
        -exec cp {} "$OUT/examples/$ex/" \;
  done
  # the NPT example predates keywords the current reader insists on (Check_Inputs_In_read_data_cpp, read_data.cpp:115-131):
  # its Widom_Trials / Widom_Orientation are spelled the current way and UseChargesFromCIFFile is stated
  mkdir -p "$OUT/examples/NPTMC"
  find "$REF/Examples/NPTMC" -maxdepth 1 -type f \( -name '*.def' -o -name '*.cif' -o -name 'simulation.input' \) -exec cp {} "$OUT/examples/NPTMC/" \;
  chmod u+w "$OUT/examples/NPTMC/simulation.input"
  sed -i 's/^Widom_Trials .*/NumberOfTrialPositions 10/; s/^Widom_Orientation .*/NumberOfTrialOrientations 10\nUseChargesFromCIFFile yes/' "$OUT/examples/NPTMC/simulation.input"
  # the Gibbs-ensemble example (two boxes run together) lacks UseChargesFromCIFFile, which the current reader requires
  for ex in NVT-Gibbs NPT-Gibbs; do
    mkdir -p "$OUT/examples/$ex"
    find "$REF/Examples/$ex" -maxdepth 1 -type f \( -name '*.def' -o -name '*.cif' -o -name 'simulation.input' \) -exec cp {} "$OUT/examples/$ex/" \;
    chmod u+w "$OUT/examples/$ex/simulation.input"
    grep -q UseChargesFromCIFFile "$OUT/examples/$ex/simulation.input" || sed -i 's/^ChargeMethod .*/&\nUseChargesFromCIFFile yes/' "$OUT/examples/$ex/simulation.input"
    sed -i '/^SaveOutputToFile/d' "$OUT/examples/$ex/simulation.input"      # output goes to stdout like everywhere else
  done
  # GCMC of TIP4P water continuing from a RASPA-2 restart file
  d="$OUT/examples/Restart-Examples"
  mkdir -p "$d/RestartInitial/System_0"
  find "$REF/Examples/Restart-Examples" -maxdepth 1 -type f \( -name '*.def' -o -name '*.cif' -o -name 'simulation.input' \) -exec cp {} "$d/" \;
  cp "$REF/Examples/Restart-Examples/RestartInitial/System_0/restartfile" "$d/RestartInitial/System_0/"
  # ... and from a LAMMPS data file (RestartInputFileType LAMMPS)
  d="$OUT/examples/Restart-LAMMPS"
  mkdir -p "$d/LMPDataInitial/System_0"
  find "$REF/Examples/Restart-Examples/Read-LAMMPS" -maxdepth 1 -type f \( -name '*.def' -o -name '*.cif' -o -name 'simulation.input' \) -exec cp {} "$d/" \;
  cp "$REF/Examples/Restart-Examples/Read-LAMMPS/LMPDataInitial/System_0/init.data" "$d/LMPDataInitial/System_0/"
  # the NIST SPC/E known-answer decks: inputs + the RASPA-2 restart file they start from
  for b in 1 2 3 4; do
    d="$OUT/examples/Reference_NIST_SPCE/Box-$b"
    mkdir -p "$d/RestartInitial/System_0"
    find "$REF/Examples/Reference_NIST_SPCE/Box-$b" -maxdepth 1 -type f \
        \( -name '*.def' -o -name "Box-$b.cif" -o -name 'simulation.input' \) -exec cp {} "$d/" \;
    cp "$REF/Examples/Reference_NIST_SPCE/Box-$b/RestartInitial/System_0/restartfile" "$d/RestartInitial/System_0/"
  done
fi

if [ "$WHAT" = "all" ] || [ "$WHAT" = "cuda" ]; then
  echo "[build_ref] reference CUDA program (sm_100)"
  SCR="$(mktemp -d /tmp/graspa_ref_build.XXXXXX)"
  cp -r "$REF/src_clean/." "$SCR/"
  chmod -R u+w "$SCR"
  # the one accommodation nvcc needs: size_t -> int narrowing in a braced init
  sed -i '268s/{OLDComponent, OLDMolInComponent}/{(int) OLDComponent, (int) OLDMolInComponent}/' "$SCR/mc_swap_moves.h"
  FLAGS="-O3 -std=c++20 -arch=sm_100 --expt-relaxed-constexpr -w -Xcompiler -fopenmp -rdc=true -x cu"
  ( cd "$SCR"
    for f in axpy.cu main.cpp read_data.cpp data_struct.cpp VDW_Coulomb.cu; do
      "$NVCC" $FLAGS -c "$f" -o "${f%.*}.o" &
    done
    wait
    "$NVCC" -arch=sm_100 -rdc=true -Xcompiler -fopenmp main.o read_data.o axpy.o data_struct.o VDW_Coulomb.o -o "$OUT/graspa_ref_cuda.x"
  )
  rm -rf "$SCR"
fi
if [ "$WHAT" = "trace" ]; then
  # the same program with ONE added line: RunMoves (axpy.cu:297) appends "component movetype deltaE" of every move to the
  # file named by $GRASPA_TRACE -- the accept/reject sequence the parity check compares move by move
  # (scripts/compare_trace.sh).  The line is inserted into the scratch copy only.
  echo "[build_ref] reference CUDA program with a move trace (sm_100)"
  SCR="$(mktemp -d /tmp/graspa_ref_trace.XXXXXX)"
  cp -r "$REF/src_clean/." "$SCR/"
  chmod -R u+w "$SCR"
  sed -i '268s/{OLDComponent, OLDMolInComponent}/{(int) OLDComponent, (int) OLDMolInComponent}/' "$SCR/mc_swap_moves.h"
  grep -n "SystemComponents.deltaE += DeltaE;" "$SCR/axpy.cu" | head -1 | grep -q "^297:" || { echo "axpy.cu:297 is not the deltaE accumulation any more"; exit 1; }
  sed -i '297i\  { static FILE* gtf = getenv("GRASPA_TRACE") ? fopen(getenv("GRASPA_TRACE"), "w") : nullptr; if(gtf) fprintf(gtf, "%zu %d %.12e\\n", comp, MoveType, DeltaE.total()); }' "$SCR/axpy.cu"
  FLAGS="-O3 -std=c++20 -arch=sm_100 --expt-relaxed-constexpr -w -Xcompiler -fopenmp -rdc=true -x cu"
  ( cd "$SCR"
    for f in axpy.cu main.cpp read_data.cpp data_struct.cpp VDW_Coulomb.cu; do
      "$NVCC" $FLAGS -c "$f" -o "${f%.*}.o" &
    done
    wait
    "$NVCC" -arch=sm_100 -rdc=true -Xcompiler -fopenmp main.o read_data.o axpy.o data_struct.o VDW_Coulomb.o -o "$OUT/graspa_ref_cuda_trace.x"
  )
  sed -n '295,299p' "$SCR/axpy.cu"
  rm -rf "$SCR"
fi
if [ "$WHAT" = "dump" ]; then
  # the same program instrumented by oracle/ref_dump_patch.py (scratch copy only): Insertion_Body writes its trial positions,
  # per-trial energies, Boltzmann selection and Rosenbluth weights, BlockedPocket its verdicts, to $GRASPA_DUMP.  Run once on the
  # GPU box (scripts/make_ref_dump.sh); tests/golden/make_ref_dump.py turns the text into the committed fixtures.
  echo "[build_ref] reference CUDA program with the CBMC dump (sm_100)"
  SCR="$(mktemp -d /tmp/graspa_ref_dump.XXXXXX)"
  cp -r "$REF/src_clean/." "$SCR/"
  chmod -R u+w "$SCR"
  sed -i '268s/{OLDComponent, OLDMolInComponent}/{(int) OLDComponent, (int) OLDMolInComponent}/' "$SCR/mc_swap_moves.h"
  python "$HERE/ref_dump_patch.py" "$SCR"
  FLAGS="-O3 -std=c++20 -arch=sm_100 --expt-relaxed-constexpr -w -Xcompiler -fopenmp -rdc=true -x cu"
  ( cd "$SCR"
    for f in axpy.cu main.cpp read_data.cpp data_struct.cpp VDW_Coulomb.cu; do
      "$NVCC" $FLAGS -c "$f" -o "${f%.*}.o" &
    done
    wait
    "$NVCC" -arch=sm_100 -rdc=true -Xcompiler -fopenmp main.o read_data.o axpy.o data_struct.o VDW_Coulomb.o -o "$OUT/graspa_ref_cuda_dump.x"
  )
  rm -rf "$SCR"
fi
if [ "$WHAT" = "overlay" ]; then
  # the reference's own program with its hot-path call sites bound to graspa_b200's C ABI (oracle/overlay/): scratch copy patched by
  # overlay_patch.py, compiled with nvcc like the stock build, linked with libgraspa_b200.so (rpath relative to the binary).
  # Carries the one-line move trace of the `trace` build, so that tests/test_gpu_trace_parity.py can compare it move by move.
  echo "[build_ref] reference drivers bound to libgraspa_b200.so (sm_100)"
  SCR="$(mktemp -d /tmp/graspa_ref_overlay.XXXXXX)"
  cp -r "$REF/src_clean/." "$SCR/"
  chmod -R u+w "$SCR"
  sed -i '268s/{OLDComponent, OLDMolInComponent}/{(int) OLDComponent, (int) OLDMolInComponent}/' "$SCR/mc_swap_moves.h"
  grep -n "SystemComponents.deltaE += DeltaE;" "$SCR/axpy.cu" | head -1 | grep -q "^297:" || { echo "axpy.cu:297 is not the deltaE accumulation any more"; exit 1; }
  sed -i '297i\  { static FILE* gtf = getenv("GRASPA_TRACE") ? fopen(getenv("GRASPA_TRACE"), "w") : nullptr; if(gtf) fprintf(gtf, "%zu %d %.12e\\n", comp, MoveType, DeltaE.total()); }' "$SCR/axpy.cu"
  python "$HERE/overlay/overlay_patch.py" "$SCR"
  FLAGS="-O3 -std=c++20 -arch=sm_100 --expt-relaxed-constexpr -w -Xcompiler -fopenmp -rdc=true -x cu -I$HERE/../include -I$HERE/overlay"
  ( cd "$SCR"
    for f in axpy.cu main.cpp read_data.cpp data_struct.cpp VDW_Coulomb.cu; do
      "$NVCC" $FLAGS -c "$f" -o "${f%.*}.o" &
    done
    wait
    "$NVCC" -arch=sm_100 -rdc=true -Xcompiler -fopenmp main.o read_data.o axpy.o data_struct.o VDW_Coulomb.o -L"$HERE/../graspa_b200" -lgraspa_b200 \
        -Xlinker -rpath -Xlinker '$ORIGIN/../../graspa_b200' -o "$OUT/graspa_ref_overlay.x"
  )
  rm -rf "$SCR"
fi
echo "[build_ref] done: $(ls "$OUT")"
