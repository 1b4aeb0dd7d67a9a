#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Instruments a SCRATCH COPY of the reference sources (never /root/reference itself) so that the
reference program writes what its CBMC insertion computes on the way, to the file named by $GRASPA_DUMP, for the first
$GRASPA_DUMP_N (default 300) calls of Insertion_Body:

  BEGIN
  FB <movetype> <ntrials> <pool offset> ; then per trial: R <3 pool randoms>  P <x y z of the trial first bead>      (get_random_trial_position)
  FBE <nsurv> ; per surviving trial: T <trial index> <log Boltzmann factor> <HGVDW HGReal GGVDW GGReal>                (Host_sum_Widom_HGGG_SEPARATE)
  FBS <Goodconstruction> <SelectedTrial> <Rosenbluth sum> <uniform drawn by SelectTrialPosition or -1>               (CBMC_FirstBead_Finish)
  CH <norient> <chainsize> <pool offset> <first bead x y z> ; per orientation: R <3 randoms> ; per chain atom: P <x y z>  (get_random_trial_orientation)
  CHE / CHS  as FBE / FBS for the chain stage                                                                        (Widom_Move_Chain_PARTIAL)
  INS <Rosenbluth> <HGVDW HGReal GGVDW GGReal GGEwaldE HGEwaldE TailE>                                                (end of Insertion_Body)
  BP <component> <x y z> <blocked>     every BlockedPocket() call                                                     (read_data.cpp:3466-3640)

Used by oracle/build_ref.sh dump; the binary runs on the GPU box once (scripts/make_ref_dump.sh) and tests/golden/make_ref_dump.py
turns the text into the committed fixture tests/golden/ref_dump_*.npz that pins the oracle's trial generation, Boltzmann selection
and Rosenbluth arithmetic (tests/test_oracle_vs_reference_dump.py)."""
import sys


def patch(path, edits):
    s = open(path).read()
    for old, new, count in edits:
        assert s.count(old) >= 1, (path, old[:70])
        if count == "last":
            k = s.rfind(old)
            s = s[:k] + new + s[k + len(old):]
        else:
            s = s.replace(old, new, count)
    open(path, "w").write(s)


HELPERS = r'''
#include <cstdio>
#include <cstdlib>
inline FILE* gdump_file() { static FILE* f = getenv("GRASPA_DUMP") ? fopen(getenv("GRASPA_DUMP"), "w") : nullptr; return f; }
inline long gdump_n() { static long n = getenv("GRASPA_DUMP_N") ? atol(getenv("GRASPA_DUMP_N")) : 300; return n; }
inline long& gdump_idx() { static long i = 0; return i; }          // Insertion_Body calls seen so far
inline double& gdump_u() { static double u = -1.0; return u; }
inline bool gdump_on() { return gdump_file() != nullptr && gdump_idx() >= 1 && gdump_idx() <= gdump_n(); }
'''


def main(scr):
    # helpers + the BlockedPocket wrapper declaration
    patch(f"{scr}/read_data.h", [
        ("bool BlockedPocket(Components& SystemComponents, size_t component, const double3& pos, Boxsize& Box);",
         HELPERS + "bool BlockedPocket(Components& SystemComponents, size_t component, const double3& pos, Boxsize& Box);", 1)])
    patch(f"{scr}/read_data.cpp", [
        ("bool BlockedPocket(Components& SystemComponents, size_t component, const double3& pos, Boxsize& Box)\n{",
         "static bool BlockedPocket_impl(Components& SystemComponents, size_t component, const double3& pos, Boxsize& Box);\n"
         "bool BlockedPocket(Components& SystemComponents, size_t component, const double3& pos, Boxsize& Box)\n{\n"
         "  const bool r = BlockedPocket_impl(SystemComponents, component, pos, Box);\n"
         "  static long nbp = 0;\n"
         "  if(gdump_file() && nbp < 4000) { fprintf(gdump_file(), \"BP %zu %.17g %.17g %.17g %d\\n\", component, pos.x, pos.y, pos.z, r ? 1 : 0); nbp++; }\n"
         "  return r;\n}\n"
         "static bool BlockedPocket_impl(Components& SystemComponents, size_t component, const double3& pos, Boxsize& Box)\n{", 1)])
    fb_dump = r'''
  if(gdump_on() && MoveType == CBMC_INSERTION)
  {
    std::vector<double3> gp(NumberOfTrials);
    cudaMemcpy(gp.data(), Sims.New.pos, NumberOfTrials * sizeof(double3), cudaMemcpyDeviceToHost);
    fprintf(gdump_file(), "FB %d %zu %zu\n", MoveType, NumberOfTrials, Random.offset - NumberOfTrials);
    for(size_t i = 0; i < NumberOfTrials; i++)
    {
      const double3 r = Random.host_random[Random.offset - NumberOfTrials + i];
      fprintf(gdump_file(), "R %.17g %.17g %.17g P %.17g %.17g %.17g\n", r.x, r.y, r.z, gp[i].x, gp[i].y, gp[i].z);
    }
  }
'''
    surv_dump = r'''
  if(gdump_on() && MoveType == CBMC_INSERTION)
  {
    fprintf(gdump_file(), "%s %zu\n", "TAG", Rosen.size());
    for(size_t a = 0; a < Rosen.size(); a++)
      fprintf(gdump_file(), "T %zu %.17g %.17g %.17g %.17g %.17g\n", Trialindex[a], Rosen[a], energies[a].HGVDW, energies[a].HGReal, energies[a].GGVDW, energies[a].GGReal);
    gdump_u() = -1.0;
  }
'''
    sel_dump = r'''
  if(gdump_on() && MoveType == CBMC_INSERTION)
    fprintf(gdump_file(), "%s %d %zu %.17g %.17g\n", "TAG", Goodconstruction ? 1 : 0, SelectedTrial, Rosenbluth, gdump_u());
'''
    ch_dump = r'''
  if(gdump_on() && MoveType == CBMC_INSERTION)
  {
    const size_t no = Widom.NumberWidomTrialsOrientations;
    std::vector<double3> gp(no * chainsize); double3 fbp;
    cudaMemcpy(gp.data(), Sims.New.pos, no * chainsize * sizeof(double3), cudaMemcpyDeviceToHost);
    cudaMemcpy(&fbp, Sims.Old.pos, sizeof(double3), cudaMemcpyDeviceToHost);
    fprintf(gdump_file(), "CH %zu %zu %zu %.17g %.17g %.17g\n", no, chainsize, Random.offset - no, fbp.x, fbp.y, fbp.z);
    for(size_t i = 0; i < no; i++)
    {
      const double3 r = Random.host_random[Random.offset - no + i];
      fprintf(gdump_file(), "R %.17g %.17g %.17g\n", r.x, r.y, r.z);
    }
    for(size_t i = 0; i < no * chainsize; i++) fprintf(gdump_file(), "P %.17g %.17g %.17g\n", gp[i].x, gp[i].y, gp[i].z);
  }
'''
    patch(f"{scr}/mc_widom.h", [
        ("    double ws = Get_Uniform_Random() * SumShiftedBoltzmannFactors;",
         "    double gdu = Get_Uniform_Random(); gdump_u() = gdu; double ws = gdu * SumShiftedBoltzmannFactors;", 1),
        ("  Random.Update(NumberOfTrials);\n", "  Random.Update(NumberOfTrials);\n" + fb_dump, 1),
        # CBMC_FirstBead_Finish: survivors before the switch, outcome after it
        ("  double averagedRosen = 0.0;\n  size_t REALselected  = 0;\n",
         "  double averagedRosen = 0.0;\n  size_t REALselected  = 0;\n" + surv_dump.replace("TAG", "FBE"), 1),
        ("  if(!Goodconstruction) return;\n  REALselected = Trialindex[SelectedTrial];",
         sel_dump.replace("TAG", "FBS") + "  if(!Goodconstruction) return;\n  REALselected = Trialindex[SelectedTrial];", 1),
        # chain
        ("  Random.Update(Widom.NumberWidomTrialsOrientations);\n", "  Random.Update(Widom.NumberWidomTrialsOrientations);\n" + ch_dump, 1),
        ("  double averagedRosen= 0.0; \n", surv_dump.replace("TAG", "CHE") + "  double averagedRosen= 0.0; \n", 1),
        ("  if(!Goodconstruction)\n  {\n    CBMC.Rosenbluth = 0.0;\n    return;\n  }\n  REALselected = Trialindex[SelectedTrial];",
         sel_dump.replace("TAG", "CHS") + "  if(!Goodconstruction)\n  {\n    CBMC.Rosenbluth = 0.0;\n    return;\n  }\n  REALselected = Trialindex[SelectedTrial];", 1),
    ])
    patch(f"{scr}/mc_swap_utilities.h", [
        ("  CBMC.MoveType = CBMC_INSERTION; //Insertion//\n",
         "  CBMC.MoveType = CBMC_INSERTION; //Insertion//\n  gdump_idx()++;\n  if(gdump_on()) fprintf(gdump_file(), \"BEGIN\\n\");\n", 1),
        ("  //printf(\"Insertion energy summary: \"); energy.print();\n  return energy;",
         "  if(gdump_on()) fprintf(gdump_file(), \"INS %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\\n\", Rosenbluth, energy.HGVDW, energy.HGReal, energy.GGVDW, energy.GGReal, energy.GGEwaldE, energy.HGEwaldE, energy.TailE);\n"
         "  if(gdump_file() && gdump_idx() == gdump_n()) fflush(gdump_file());\n"
         "  //printf(\"Insertion energy summary: \"); energy.print();\n  return energy;", 1),
    ])


if __name__ == "__main__":
    main(sys.argv[1])
