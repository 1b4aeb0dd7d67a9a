#!/usr/bin/env python
"""TEST INFRASTRUCTURE / INTEGRATION DEMONSTRATION.  Binds a SCRATCH COPY of the reference's sources (never /root/reference
itself) to graspa_b200's C ABI: the call sites of SURVEY section 8(b) are redirected to oracle/overlay/graspa_b200_adapter.h, the
move drivers stay the reference's own code.  oracle/build_ref.sh overlay compiles the result with nvcc and links it with
libgraspa_b200.so -> oracle/_ref/graspa_ref_overlay.x; tests/test_gpu_trace_parity.py runs it against the stock reference.

What is redirected (file:line of the reference):
  mc_widom.h:385-507            Widom_Move_FirstBead_PARTIAL  -> gb_cbmc_first_bead
  mc_widom.h:509-614            Widom_Move_Chain_PARTIAL      -> gb_cbmc_chain
  Ewald_Energy_Functions.h:438  GPU_EwaldDifference_General   -> gb_ewald_delta
  Ewald_Energy_Functions.h:423  Update_Vector_Ewald           -> (inside the gb_accept_* calls)
  mc_single_particle.h:79       get_new_position<<<>>>        -> gb_single_body_propose
  mc_single_particle.h:174-200  Calculate_Single_Body_Energy_VDWReal<<<>>> + host sum of Blocksum -> gb_single_body_delta
  mc_utilities.h:294-417        update_translation_position / Update_insertion_data_Parallel / Update_deletion_data_Parallel<<<>>> -> gb_accept_*
  move_struct.h:271,371         StoreNewLocation_Reinsertion / Update_Reinsertion_data<<<>>> -> gb_reinsertion_store / gb_accept_reinsertion
  data_struct.h:1300-1320       RandomNumber::ResetRandom     -> + gb_upload_random_pool
  data_struct.cpp:6-11          Get_Uniform_Random            -> + Peek_Uniform_Random (one value of look-ahead)
  main.cpp:340, :427            before the CREATE_MOLECULE and FINAL energy checks -> engine state copied back into Sims.d_a (the reference's own CPU + GPU
                                                                 total-energy routines then judge the run: ENERGY DRIFT)
  axpy.cu:297                   the one-line move trace of `build_ref.sh trace`
  mc_single_particle.h:83-119, mc_swap_utilities.h:46-80, move_struct.h:219-262, mc_swap_moves.h:302-330   host-side BlockedPocket tests of
                                proposed / grown molecules -> skipped; the engine's kernels run the test (gb_set_block_pockets at set-up)
  mc_swap_moves.h:266,298,355,393-412  IdentitySwapMove: copy_firstbead_to_new / StoreNewLocation_Reinsertion<<<>>> -> (engine) / gb_reinsertion_store;
                                GPU_EwaldDifference_IdentitySwap -> gb_ewald_delta_identity_swap; Update_deletion_data + Update_IdentitySwap_Insertion_data /
                                Update_Reinsertion_data<<<>>> -> gb_accept_identity_swap
  mc_box.h:236-252, 292, 306    VolumeMove: Total_VDW_Coulomb_Energy + overlap flag + Ewald_TotalEnergy of the scaled system -> gb_volume_move_trial;
                                CopyScaledPositions / Revert_Boxsize<<<>>> -> + gb_volume_move_finish(1 / 0)
  mc_cbcfc.h:39-104             Prepare_LambdaChange / Calculate_Single_Body_Energy_VDWReal_LambdaChange<<<>>> + host sum -> gb_lambda_change_delta;
                                GPU_EwaldDifference_LambdaChange -> gb_ewald_delta_lambda_change
  mc_cbcfc.h:312,359,427,487    update_CBCF_scale / Revert_CBCF_Insertion<<<>>> -> gb_cbcf_set_scale (provisional) / gb_accept_lambda_change (accepted)
  mc_cbcfc.h:382,441            Update_deletion_data_fractional / Revert_CBCF_Deletion<<<>>> -> gb_cbcf_deletion_stage
  mc_cbcfc.h:343                Update_insertion_data_Parallel<<<>>> of an accepted CBCF insertion -> gb_accept_insertion
Scope: the moves of the CO2-MFI deck (translation, rotation, CBMC insertion / deletion, reinsertion), the identity swap of the
XeKr-Mixture deck, the Widom move of the Henrys_coefficient deck (it is Insertion_Body), the CO2_NaX_Zeolite deck (moves of a separated
framework component, block pockets), the NPT volume move of the NPTMC deck and, with CBCFProbability added to
that deck, the CB/CFC move (lambda change; the first steps and the reversal of a fractional insertion / deletion -- the reference's
CBCFMove never accepts those two: it tests a local SuccessConstruction that nothing sets, mc_cbcfc.h:307-318, :372-393)."""
import sys


def patch(path, edits):
    s = open(path).read()
    for ed in edits:
        old, new = ed[0], ed[1]
        want = ed[2] if len(ed) > 2 else 1
        assert s.count(old) == want, (path, s.count(old), old[:80])
        s = s.replace(old, new)
    open(path, "w").write(s)


def main(scr):
    patch(f"{scr}/data_struct.cpp", [(
        "double Get_Uniform_Random()\n{\n  //return (double) (rand()/RAND_MAX);\n  //std::srand(3.0);\n  return static_cast<double>(std::rand()) / RAND_MAX;\n}",
        "static bool g_uniform_pending = false; static double g_uniform_value = 0.0;\n"
        "double Peek_Uniform_Random()\n{\n  if(!g_uniform_pending) { g_uniform_value = static_cast<double>(std::rand()) / RAND_MAX; g_uniform_pending = true; }\n  return g_uniform_value;\n}\n"
        "double Get_Uniform_Random()\n{\n  if(g_uniform_pending) { g_uniform_pending = false; return g_uniform_value; }\n  return static_cast<double>(std::rand()) / RAND_MAX;\n}")])
    s = open(f"{scr}/data_struct.h").read()
    assert s.count("struct RandomNumber\n{") == 1
    s = s.replace("struct RandomNumber\n{", "inline void b200_pool_refreshed(const double3* host_random, size_t n);\nstruct RandomNumber\n{")
    old = "    cudaMemcpy(device_random, host_random, randomsize * sizeof(double3), cudaMemcpyHostToDevice);\n    Rounds ++;"
    assert s.count(old) == 1
    s = s.replace(old, old + "\n    b200_pool_refreshed(host_random, randomsize);")
    s += '\n#include "graspa_b200_adapter.h"\n'
    open(f"{scr}/data_struct.h", "w").write(s)
    patch(f"{scr}/mc_widom.h", [
        ("inline void Widom_Move_FirstBead_PARTIAL(Variables& Vars, size_t systemId, CBMC_Variables& CBMC)\n{",
         "inline void Widom_Move_FirstBead_PARTIAL(Variables& Vars, size_t systemId, CBMC_Variables& CBMC)\n{\n  b200_first_bead(Vars, systemId, CBMC); return;"),
        ("inline void Widom_Move_Chain_PARTIAL(Variables& Vars, size_t systemId, CBMC_Variables& CBMC)\n{",
         "inline void Widom_Move_Chain_PARTIAL(Variables& Vars, size_t systemId, CBMC_Variables& CBMC)\n{\n  b200_chain(Vars, systemId, CBMC); return;")])
    patch(f"{scr}/Ewald_Energy_Functions.h", [
        ("double2 GPU_EwaldDifference_General(Simulations& Sim, ForceField& FF, Components& SystemComponents, size_t SelectedComponent, int MoveType, size_t Location, double2 Scale)\n{",
         "double2 GPU_EwaldDifference_General(Simulations& Sim, ForceField& FF, Components& SystemComponents, size_t SelectedComponent, int MoveType, size_t Location, double2 Scale)\n{\n"
         "  if(b200().e) return b200_ewald_delta(SystemComponents, SelectedComponent, MoveType, Location, Scale);"),
        ("void Update_Vector_Ewald(Boxsize& Box, bool CPU, Components& SystemComponents, size_t SelectedComponent)\n{",
         "void Update_Vector_Ewald(Boxsize& Box, bool CPU, Components& SystemComponents, size_t SelectedComponent)\n{\n  if(b200().e) return;     // the gb_accept_* calls swap the engine's vectors")])
    patch(f"{scr}/mc_single_particle.h", [
        ("  get_new_position<<<1, Molsize>>>(Sims, FF, start_position, SelectedComponent, MaxChange, Random.device_random, Random.offset, MoveType);",
         "  b200_propose(Vars, systemId, MoveType, SelectedComponent, SelectedMolInComponent, MaxChange, Random.offset);"),
        ("    Calculate_Single_Body_Energy_VDWReal<<<Total_Nblock, Nthread, Nthread * 2 * sizeof(double)>>>(Sims.Box, Sims.d_a, Sims.Old, Sims.New, FF, Sims.Blocksum, SelectedComponent, Atomsize, Molsize, Sims.device_flag, NBlocks, Do_New, Do_Old, SystemComponents.NComponents);\n\n"
         "    SystemComponents.flag = Sims.device_flag;\n    cudaDeviceSynchronize();",
         "    b200_single_body_delta(SystemComponents, SelectedComponent, Do_New, Do_Old);"),
        ("    // Calculate Ewald //\n", "    b200_take_single_body(tot);\n    // Calculate Ewald //\n")])
    patch(f"{scr}/mc_utilities.h", [
        ("  update_translation_position<<<1,Molsize>>>(Sims.d_a, Sims.New, start_position, SelectedComponent);",
         "  b200_accept_translation(SelectedComponent);"),
        ("    Update_insertion_data_Parallel<<<1,SystemComponents.Moleculesize[SelectedComponent]>>>(Sims.d_a, Sims.Old, Sims.New, SelectedTrial, SelectedComponent, UpdateLocation, (int) SystemComponents.Moleculesize[SelectedComponent]);",
         "    b200_accept_insertion(SelectedComponent);"),
        ("  Update_deletion_data_Parallel<<<1,SystemComponents.Moleculesize[SelectedComponent]>>>(Sims.d_a, SelectedComponent, UpdateLocation, (int) SystemComponents.Moleculesize[SelectedComponent], LastLocation);",
         "  b200_accept_deletion(SelectedComponent, UpdateLocation / SystemComponents.Moleculesize[SelectedComponent]);")])
    patch(f"{scr}/move_struct.h", [
        ("      StoreNewLocation_Reinsertion<<<1,SystemComponents.Moleculesize[SelectedComponent]>>>(Sims.Old, Sims.New, SystemComponents.tempMolStorage, SelectedTrial, SystemComponents.Moleculesize[SelectedComponent]);",
         "      b200_reinsertion_store(SelectedComponent);"),
        ("    Update_Reinsertion_data<<<1,SystemComponents.Moleculesize[SelectedComponent]>>>(Sims.d_a, SystemComponents.tempMolStorage, SelectedComponent, UpdateLocation); checkCUDAError(\"error Updating Reinsertion data\");",
         "    b200_accept_reinsertion(SelectedComponent, SystemComponents.TempVal.molecule);")])
    # block pockets: gb_set_block_pockets hands the replicated list to the engine, whose stage kernels run BlockedPocket themselves (first-bead
    # trials, grown molecules, proposed positions); the drivers' host-side copies of those tests read Sims.Old / Sims.New, which the bound
    # program no longer fills, and are skipped while the engine is bound (their statistics counters stay at zero)
    POCK = "SystemComponents.UseBlockPockets[SelectedComponent]"
    patch(f"{scr}/mc_single_particle.h", [(f"     {POCK})\n  {{", f"     {POCK} && !b200().e)\n  {{")])
    patch(f"{scr}/mc_swap_utilities.h", [(f"       {POCK} &&\n", f"       {POCK} && !b200().e &&\n")])
    patch(f"{scr}/move_struct.h", [(f"       {POCK} &&\n", f"       {POCK} && !b200().e &&\n")])
    patch(f"{scr}/mc_swap_moves.h", [("     SystemComponents.UseBlockPockets[NEWComponent])\n  {", "     SystemComponents.UseBlockPockets[NEWComponent] && !b200().e)\n  {")])
    # IdentitySwapMove (mc_swap_moves.h:199-431): the first bead of the new species is read by the engine from the molecule named in
    # Sims.ExcludeList[0] (b200_first_bead passes the list entry on), the grown molecule is kept aside, both commits are one call
    patch(f"{scr}/mc_swap_moves.h", [
        ("  copy_firstbead_to_new<<<1,1>>>(Sims.New, Sims.d_a, OLDComponent, OLDMolInComponent * SystemComponents.Moleculesize[OLDComponent]);",
         "  b200_engine(Vars, systemId);"),
        ("  StoreNewLocation_Reinsertion<<<1,SystemComponents.Moleculesize[NEWComponent]>>>(Sims.Old, Sims.New, SystemComponents.tempMolStorage, SelectedTrial, SystemComponents.Moleculesize[NEWComponent]);",
         "  b200_reinsertion_store(NEWComponent);"),
        ("    double2 EwaldE = GPU_EwaldDifference_IdentitySwap(Sims.Box, Sims.d_a, Sims.Old, SystemComponents.tempMolStorage, FF, Sims.Blocksum, SystemComponents, OLDComponent, NEWComponent, UpdateLocation);",
         "    double2 EwaldE = b200_ewald_delta_identity_swap(OLDComponent, NEWComponent, UpdateLocation);"),
        ("      Update_deletion_data<<<1,1>>>(Sims.d_a, OLDComponent, UpdateLocation, (int) SystemComponents.Moleculesize[OLDComponent], LastLocation);",
         "      b200_accept_identity_swap(OLDComponent, OLDMolInComponent, NEWComponent);"),
        ("      Update_IdentitySwap_Insertion_data<<<1,1>>>(Sims.d_a, SystemComponents.tempMolStorage, NEWComponent, UpdateLocation, NEWMolInComponent, SystemComponents.Moleculesize[NEWComponent]); checkCUDAError(\"error Updating Identity Swap Insertion data\");",
         "      // (the engine appended the new molecule in b200_accept_identity_swap)"),
        ("      Update_Reinsertion_data<<<1,SystemComponents.Moleculesize[OLDComponent]>>>(Sims.d_a, SystemComponents.tempMolStorage, OLDComponent, UpdateLocation);",
         "      b200_accept_identity_swap(OLDComponent, OLDMolInComponent, NEWComponent);")])
    MS = "SystemComponents.Moleculesize[SelectedComponent]"
    patch(f"{scr}/mc_cbcfc.h", [
        ("  Prepare_LambdaChange<<<1, Molsize>>>(Sims.d_a, Sims.Old, Sims, FF, start_position, SelectedComponent, Sims.device_flag);",
         "  b200_engine(Vars, systemId);"),
        ("  Calculate_Single_Body_Energy_VDWReal_LambdaChange<<<Total_Nblock, Nthread, Nthread * 2 * sizeof(double)>>>(Sims.Box, Sims.d_a, Sims.Old, Sims.New, FF, Sims.Blocksum, SelectedComponent, Atomsize, Molsize, Sims.device_flag, NBlocks, Do_New, Do_Old, SystemComponents.NComponents, newScale);\n\n"
         "  cudaMemcpy(SystemComponents.flag, Sims.device_flag, sizeof(bool), cudaMemcpyDeviceToHost);",
         "  b200_lambda_change_delta(SystemComponents, SelectedComponent, SelectedMolInComponent, newScale);"),
        ("    if(!FF.noCharges && SystemComponents.hasPartialCharge[SelectedComponent])\n    {\n"
         "      //Zhao's note: since we changed it from using translation/rotation functions to its own, this needs to be changed as well//\n"
         "      double2 EwaldE = GPU_EwaldDifference_LambdaChange(Sims.Box, Sims.d_a, Sims.Old, FF, Sims.Blocksum, SystemComponents, SelectedComponent, oldScale, newScale, MoveType);",
         "    b200_take_single_body(tot);\n    if(!FF.noCharges && SystemComponents.hasPartialCharge[SelectedComponent])\n    {\n"
         "      double2 EwaldE = b200_ewald_delta_lambda_change(SelectedComponent, oldScale, newScale, MoveType);"),
        (f"    update_CBCF_scale<<<1,{MS}>>>(Sims.d_a, start_position, SelectedComponent, InterScale);",
         f"    b200_cbcf_set_scale(SelectedComponent, start_position / {MS}, InterScale);"),
        (f"    if(!Accepted) Revert_CBCF_Insertion<<<1, {MS}>>>(Sims.d_a, SelectedComponent, start_position, oldScale);",
         f"    if(!Accepted) b200_cbcf_set_scale(SelectedComponent, start_position / {MS}, oldScale);"),
        (f"        Update_insertion_data_Parallel<<<1,{MS}>>>(Sims.d_a, Sims.Old, Sims.New, SelectedTrial, SelectedComponent, UpdateLocation, (int) {MS});",
         "        b200_accept_insertion(SelectedComponent);"),
        (f"    Update_deletion_data_fractional<<<1,1>>>(Sims.d_a, SelectedComponent, UpdateLocation, (int) {MS}, LastLocation);",
         f"    b200_cbcf_deletion_stage(SelectedComponent, UpdateLocation / {MS}, false);"),
        (f"      Revert_CBCF_Deletion<<<1,1>>>(Sims.d_a, Sims.New, SelectedComponent, UpdateLocation, (int) {MS}, LastLocation);",
         f"      b200_cbcf_deletion_stage(SelectedComponent, UpdateLocation / {MS}, true);"),
        # accepted lambda change (:487) and accepted CBCF deletion (:427): scaling factors + swap of the structure-factor vectors
        (f"update_CBCF_scale<<<1,{MS}>>>(Sims.d_a, start_position, SelectedComponent, newScale);",
         f"b200_accept_lambda_change(SelectedComponent, start_position / {MS}, newScale);", 2)])
    # VolumeMove (mc_box.h:196-320): ScalePositions<<<>>> still scales the reference's own Sim.Box (cell, inverse cell, kmax, reciprocal cutoff);
    # the engine takes that box, scales its molecules and evaluates both totals in one call; acceptance / rejection are passed on
    patch(f"{scr}/mc_box.h", [
        ("  NewE = Total_VDW_Coulomb_Energy(Sim, SystemComponents, FF, UseOffset);",
         "  bool b200_overlap = false;\n  if(b200().e) NewE = b200_volume_trial(Sim, Scale, b200_overlap); else NewE = Total_VDW_Coulomb_Energy(Sim, SystemComponents, FF, UseOffset);"),
        ("  cudaMemcpy(SystemComponents.flag, Sim.device_flag, sizeof(bool), cudaMemcpyDeviceToHost);",
         "  cudaMemcpy(SystemComponents.flag, Sim.device_flag, sizeof(bool), cudaMemcpyDeviceToHost);\n  if(b200().e) SystemComponents.flag[0] = b200_overlap;"),
        ("    NewE += Ewald_TotalEnergy(Sim, SystemComponents, UseOffSet);",
         "    if(!b200().e) NewE += Ewald_TotalEnergy(Sim, SystemComponents, UseOffSet);     // the engine's call returned the Fourier totals as well"),
        ("    CopyScaledPositions<<<Nblock, Nthread>>>(Sim.d_a, SystemComponents.NComponents.x, ScaleFirstComponentFramework, totMol);",
         "    CopyScaledPositions<<<Nblock, Nthread>>>(Sim.d_a, SystemComponents.NComponents.x, ScaleFirstComponentFramework, totMol);\n    b200_volume_finish(true);"),
        ("    Revert_Boxsize<<<1,1>>>(Sim.Box, Scale, FF.noCharges, OldV);",
         "    Revert_Boxsize<<<1,1>>>(Sim.Box, Scale, FF.noCharges, OldV);\n    b200_volume_finish(false);")])
    patch(f"{scr}/main.cpp", [("    check_energy_wrapper(Vars, i);\n    //Report Random Number Summary", "    b200_sync_back(Vars, i);\n    check_energy_wrapper(Vars, i);\n    //Report Random Number Summary"),
                              # the CREATE_MOLECULE stage check reads Sims.d_a as well (molecules created through the engine)
                              ("    Check_Simulation_Energy(Vars.Box[a], Vars.SystemComponents[a].HostSystem, Vars.FF, Vars.device_FF, Vars.SystemComponents[a], CREATEMOL, a, Vars.Sims[a], true);",
                               "    b200_sync_back(Vars, a, false);\n    Check_Simulation_Energy(Vars.Box[a], Vars.SystemComponents[a].HostSystem, Vars.FF, Vars.device_FF, Vars.SystemComponents[a], CREATEMOL, a, Vars.Sims[a], true);")])


if __name__ == "__main__":
    main(sys.argv[1])
