// TEST INFRASTRUCTURE / INTEGRATION DEMONSTRATION -- the reference-side binding of graspa_b200's C ABI.
//
// This header is what a gRASPA maintainer adds to src_clean/ (INTEGRATION.md): it is compiled INSIDE a scratch copy of the
// reference's own sources (oracle/build_ref.sh overlay appends `#include "graspa_b200_adapter.h"` to data_struct.h) against the
// reference's real structs (Variables, Components, Simulations, Atoms, ForceField, Boxsize, RandomNumber, CBMC_Variables,
// MoveEnergy: data_struct.h) and linked with libgraspa_b200.so.  oracle/overlay/overlay_patch.py then redirects the call sites
// of SURVEY section 8(b) to the functions below; the move drivers themselves (Insertion_Body, Deletion_Body, ReinsertionMove,
// SingleBodyMove, RunMoves, the acceptance tests, every random-number draw) stay the reference's code, line for line.
#pragma once
#include "graspa_b200.h"
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

double Get_Uniform_Random();
double Peek_Uniform_Random();            // the next value Get_Uniform_Random() will return (added to data_struct.cpp by the overlay)

struct B200Binding
{
  gb_engine* e = nullptr;
  Variables* vars = nullptr;
  size_t sim = 0;
  bool flag[4] = {false, false, false, false};          // what SystemComponents.flag points at (mc_single_particle.h:176)
  gb_move_energy sb{};                                  // last single-body delta
  long calls = 0;
};
inline B200Binding& b200() { static B200Binding g; return g; }
inline void GB_CHECK(int rc) { if(rc != GB_OK) { fprintf(stderr, "graspa_b200: %s\n", gb_last_error()); throw std::runtime_error(gb_last_error()); } }   // the reference throws on logic errors

// RandomNumber::ResetRandom (data_struct.h:1300-1320) calls this after refilling the host pool: the engine's device pool follows
inline void b200_pool_refreshed(const double3* host_random, size_t n)
{
  if(b200().e) GB_CHECK(gb_upload_random_pool(b200().e, reinterpret_cast<const double*>(host_random), (int64_t) n));
}

// main.cpp:104-334 + fxn_main.h:46-280 (Copy_ForceField_to_GPU, Copy_Atom_data_to_device, Prepare_Widom, Allocate_Copy_Ewald_Vector,
// RandomNumber::DeviceRandom) collapse into one set-up sequence, from the reference's HOST copies.  Called on first use.
inline gb_engine* b200_engine(Variables& Vars, size_t sim)
{
  B200Binding& G = b200();
  if(G.e) return G.e;
  G.vars = &Vars; G.sim = sim;
  Components& SC = Vars.SystemComponents[sim]; Boxsize& Box = Vars.Box[sim]; ForceField& FF = Vars.FF;
  GB_CHECK(gb_engine_create(&G.e, -1));
  gb_forcefield ff{};
  ff.size = (int32_t) FF.size; ff.epsilon = FF.epsilon; ff.sigma = FF.sigma; ff.z = FF.z; ff.shift = FF.shift; ff.c10 = FF.C10;
  ff.cutoff_vdw_sq = FF.CutOffVDW; ff.cutoff_coul_sq = FF.CutOffCoul; ff.overlap_criteria = FF.OverlapCriteria;
  ff.no_charges = FF.noCharges; ff.vdw_real_bias = FF.VDWRealBias; ff.use1264 = FF.Use1264;
  std::vector<int32_t> use_tail(FF.size * FF.size); std::vector<double> tail_e(FF.size * FF.size);
  for(size_t i = 0; i < FF.size * FF.size; i++) { use_tail[i] = SC.TailCorrection[i].UseTail; tail_e[i] = SC.TailCorrection[i].Energy; }
  gb_tail_table tail{use_tail.data(), tail_e.data(), (int32_t) FF.size, 0};
  GB_CHECK(gb_upload_forcefield(G.e, &ff, &tail));
  gb_box b{};
  for(int i = 0; i < 9; i++) { b.cell[i] = Box.Cell[i]; b.inverse_cell[i] = Box.InverseCell[i]; }
  b.cubic = Box.Cubic; b.volume = Box.Volume; b.alpha = Box.Alpha; b.prefactor = Box.Prefactor;
  b.kmax[0] = Box.kmax.x; b.kmax[1] = Box.kmax.y; b.kmax[2] = Box.kmax.z; b.reciprocal_cutoff = Box.ReciprocalCutOff;
  b.use_lammps_ewald = Box.UseLAMMPSEwald;
  GB_CHECK(gb_upload_box(G.e, &b));
  GB_CHECK(gb_set_components(G.e, (int32_t) SC.NComponents.x, (int32_t) SC.NComponents.y));
  for(size_t c = 0; c < (size_t) SC.NComponents.x; c++)
  {
    Atoms& A = SC.HostSystem[c];
    gb_atoms a{}; a.n_live = (int64_t) A.size; a.n_upload = (int64_t) std::max(A.size, A.Molsize); a.n_alloc = (int64_t) A.Allocate_size;
    a.molsize = (int64_t) A.Molsize;
    a.pos = reinterpret_cast<const double*>(A.pos); a.scale = A.scale; a.charge = A.charge; a.scale_coul = A.scaleCoul;
    a.type = reinterpret_cast<const uint64_t*>(A.Type); a.molid = reinterpret_cast<const uint64_t*>(A.MolID);
    GB_CHECK(gb_upload_atoms(G.e, (int32_t) c, &a));
    // the exclusion constants exist only when the deck has charges (ewald_preparation.h:261-366 fills them)
    const double xi = c < SC.ExclusionIntra.size() ? SC.ExclusionIntra[c] : 0.0, xa = c < SC.ExclusionAtom.size() ? SC.ExclusionAtom[c] : 0.0;
    GB_CHECK(gb_set_exclusion_constants(G.e, (int32_t) c, xi, xa, SC.rigid[c], SC.hasPartialCharge[c]));
  }
  // block pockets as ReplicateBlockPockets left them (main.cpp:286-292, read_data.cpp:3290-3454): Cartesian centres and radii per component
  for(size_t c = 0; c < (size_t) SC.NComponents.x; c++)
    if(c < SC.UseBlockPockets.size() && SC.UseBlockPockets[c] && c < SC.BlockPocketCenters.size() && !SC.BlockPocketCenters[c].empty())
    {
      const std::vector<double3>& ctr = SC.BlockPocketCenters[c];
      std::vector<double> xyz(3 * ctr.size());
      for(size_t i = 0; i < ctr.size(); i++) { xyz[3 * i] = ctr[i].x; xyz[3 * i + 1] = ctr[i].y; xyz[3 * i + 2] = ctr[i].z; }
      const bool invert = c < SC.InvertBlockPockets.size() && SC.InvertBlockPockets[c];
      GB_CHECK(gb_set_block_pockets(G.e, (int32_t) c, (int32_t) ctr.size(), xyz.data(), SC.BlockPocketRadii[c].data(), invert ? 1 : 0));
    }
  GB_CHECK(gb_set_cbmc(G.e, (int32_t) Vars.Widom[sim].NumberWidomTrials, (int32_t) Vars.Widom[sim].NumberWidomTrialsOrientations, SC.Beta));
  // the CPU Ewald_Total of the reference (ewald_preparation.h:5-259) is replaced by a build of the structure factors on the GPU
  gb_move_energy E; GB_CHECK(gb_total_ewald(G.e, 1, &E));
  GB_CHECK(gb_upload_random_pool(G.e, reinterpret_cast<const double*>(Vars.Random.host_random), (int64_t) Vars.Random.randomsize));
  return G.e;
}

inline void b200_fill(MoveEnergy& E, const gb_cbmc_result& r) { E.HGVDW = r.energy[0]; E.HGReal = r.energy[1]; E.GGVDW = r.energy[2]; E.GGReal = r.energy[3]; }

// replaces the body of Widom_Move_FirstBead_PARTIAL (mc_widom.h:385-507): get_random_trial_position<<<>>>, CBMC_PairwiseInteractions
// (Calculate_Multiple_Trial_Energy_VDWReal<<<>>> + cudaDeviceSynchronize + Host_sum_Widom_HGGG_SEPARATE) and CBMC_FirstBead_Finish
inline void b200_first_bead(Variables& Vars, size_t systemId, CBMC_Variables& CBMC)
{
  gb_engine* e = b200_engine(Vars, systemId);
  Components& SC = Vars.SystemComponents[systemId]; RandomNumber& Random = Vars.Random; WidomStruct& Widom = Vars.Widom[systemId];
  const size_t comp = SC.TempVal.component; const int type = CBMC.MoveType;
  size_t ntr = Widom.NumberWidomTrials;
  if(type == REINSERTION_RETRACE || type == IDENTITY_SWAP_OLD || type == IDENTITY_SWAP_NEW) ntr = 1;
  Random.Check(ntr);
  const double scale[2] = {SC.TempVal.Scale.x, SC.TempVal.Scale.y};
  gb_cbmc_result r; int32_t used = 0;
  // Sims.ExcludeList[0] (managed memory, written by IdentitySwapMove mc_swap_moves.h:268): the molecule left out of the pair sums; for
  // IDENTITY_SWAP_NEW the engine also takes the single trial position from its first atom (copy_firstbead_to_new, :178-181)
  const int2 ex = Vars.Sims[systemId].ExcludeList[0];
  GB_CHECK(gb_cbmc_first_bead(e, type, (int32_t) comp, (int64_t) SC.TempVal.molecule, (int64_t) Random.offset, Peek_Uniform_Random(), scale, CBMC.StoredR,
                              ex.x, ex.y, nullptr, &r, &used));
  Random.Update(ntr);
  if(used) Get_Uniform_Random();                        // the draw SelectTrialPosition would have made (mc_widom.h:14-39)
  SC.Rosen.clear(); SC.TrialEnergies.clear(); SC.Trialindex.clear();
  CBMC.SuccessConstruction = r.success != 0;
  if(!r.success) return;                                // CBMC_FirstBead_Finish returns before it touches CBMC (:362)
  CBMC.selectedTrial = (size_t) r.selected;
  if(type == REINSERTION_INSERTION) CBMC.StoredR = r.stored_r;
  b200_fill(CBMC.FirstBeadEnergy, r);
  CBMC.Rosenbluth = r.rosenbluth;
}

// replaces the body of Widom_Move_Chain_PARTIAL (mc_widom.h:509-614)
inline void b200_chain(Variables& Vars, size_t systemId, CBMC_Variables& CBMC)
{
  gb_engine* e = b200_engine(Vars, systemId);
  Components& SC = Vars.SystemComponents[systemId]; RandomNumber& Random = Vars.Random; WidomStruct& Widom = Vars.Widom[systemId];
  const size_t comp = SC.TempVal.component; const int type = CBMC.MoveType;
  Random.Check(Widom.NumberWidomTrialsOrientations);
  gb_cbmc_result r; int32_t used = 0;
  const int2 ex = Vars.Sims[systemId].ExcludeList[0];
  GB_CHECK(gb_cbmc_chain(e, type, (int32_t) comp, (int64_t) SC.TempVal.molecule, (int64_t) Random.offset, Peek_Uniform_Random(), ex.x, ex.y, &r, &used));
  Random.Update(Widom.NumberWidomTrialsOrientations);
  if(used) Get_Uniform_Random();
  SC.Rosen.clear(); SC.TrialEnergies.clear(); SC.Trialindex.clear();
  if(!r.success) { CBMC.Rosenbluth = 0.0; CBMC.SuccessConstruction = false; return; }      // :586-590 (SuccessConstruction keeps its value there; the callers test Rosenbluth)
  CBMC.selectedTrialOrientation = (size_t) r.selected;
  CBMC.SuccessConstruction = true;
  b200_fill(CBMC.ChainEnergy, r);
  CBMC.Rosenbluth *= r.rosenbluth;
}

// replaces GPU_EwaldDifference_General (Ewald_Energy_Functions.h:438-580)
inline double2 b200_ewald_delta(Components& SC, size_t comp, int MoveType, size_t Location, double2 Scale)
{
  const double sc[2] = {Scale.x, Scale.y}; double out[2] = {0.0, 0.0};
  GB_CHECK(gb_ewald_delta(b200().e, (int32_t) comp, MoveType, (int64_t) Location, sc, out));
  return {out[0], out[1]};
}

// replaces get_new_position<<<1, Molsize>>> (mc_utilities.h:485-606) in SingleBody_Prepare
inline void b200_propose(Variables& Vars, size_t systemId, int MoveType, size_t comp, size_t molecule, double3 MaxChange, size_t pool_index)
{
  gb_engine* e = b200_engine(Vars, systemId);
  const double mc[3] = {MaxChange.x, MaxChange.y, MaxChange.z};
  GB_CHECK(gb_single_body_propose(e, MoveType, (int32_t) comp, (int64_t) molecule, mc, (int64_t) pool_index, nullptr));
}

// replaces Calculate_Single_Body_Energy_VDWReal<<<>>> + cudaDeviceSynchronize + the host sum of Blocksum (mc_single_particle.h:174-200)
inline void b200_single_body_delta(Components& SC, size_t comp, bool Do_New, bool Do_Old)
{
  B200Binding& G = b200(); int32_t ov = 0;
  GB_CHECK(gb_single_body_delta(G.e, (int32_t) comp, Do_New ? 1 : 0, Do_Old ? 1 : 0, &G.sb, &ov));
  G.flag[0] = ov != 0;
  SC.flag = G.flag;
}
inline void b200_take_single_body(MoveEnergy& tot)
{
  const gb_move_energy& d = b200().sb;
  tot.HHVDW = d.HHVDW; tot.HHReal = d.HHReal; tot.HGVDW = d.HGVDW; tot.HGReal = d.HGReal; tot.GGVDW = d.GGVDW; tot.GGReal = d.GGReal;
}

// CB/CFC (mc_cbcfc.h): replaces Prepare_LambdaChange<<<>>> + Calculate_Single_Body_Energy_VDWReal_LambdaChange<<<>>> + the copy of the
// overlap flag (:39-69); the host sum of Blocksum that follows is overwritten by b200_take_single_body
inline void b200_lambda_change_delta(Components& SC, size_t comp, size_t molecule, double2 newScale)
{
  B200Binding& G = b200(); int32_t ov = 0; const double sc[2] = {newScale.x, newScale.y};
  GB_CHECK(gb_lambda_change_delta(G.e, (int32_t) comp, (int64_t) molecule, sc, &G.sb, &ov));
  G.flag[0] = ov != 0;
  SC.flag = G.flag;
}
// replaces GPU_EwaldDifference_LambdaChange (Ewald_Energy_Functions.h:637-788): a CBCF deletion's lambda change continues from tempEik (:761-764)
inline double2 b200_ewald_delta_lambda_change(size_t comp, double2 oldScale, double2 newScale, int MoveType)
{
  const double so[2] = {oldScale.x, oldScale.y}, sn[2] = {newScale.x, newScale.y}; double out[2] = {0.0, 0.0};
  GB_CHECK(gb_ewald_delta_lambda_change(b200().e, (int32_t) comp, so, sn, MoveType == CBCF_DELETION ? 1 : 0, out));
  return {out[0], out[1]};
}
// update_CBCF_scale<<<>>> before the acceptance test / Revert_CBCF_Insertion<<<>>> (mc_cbcfc.h:312, :359); accepted: :427, :487
inline void b200_cbcf_set_scale(size_t comp, size_t molecule, double2 scale)
{ const double sc[2] = {scale.x, scale.y}; GB_CHECK(gb_cbcf_set_scale(b200().e, (int32_t) comp, (int64_t) molecule, sc)); }
inline void b200_accept_lambda_change(size_t comp, size_t molecule, double2 scale)
{ const double sc[2] = {scale.x, scale.y}; GB_CHECK(gb_accept_lambda_change(b200().e, (int32_t) comp, (int64_t) molecule, sc)); }
// Update_deletion_data_fractional<<<>>> / Revert_CBCF_Deletion<<<>>> (mc_cbcfc.h:382, :441)
inline void b200_cbcf_deletion_stage(size_t comp, size_t molecule, bool revert)
{ GB_CHECK(gb_cbcf_deletion_stage(b200().e, (int32_t) comp, (int64_t) molecule, revert ? 1 : 0)); }

// IdentitySwapMove (mc_swap_moves.h:355, :393-412)
inline double2 b200_ewald_delta_identity_swap(size_t oldc, size_t newc, size_t update_location)
{
  double out[2] = {0.0, 0.0};
  GB_CHECK(gb_ewald_delta_identity_swap(b200().e, (int32_t) oldc, (int32_t) newc, (int64_t) update_location, out));
  return {out[0], out[1]};
}
inline void b200_accept_identity_swap(size_t oldc, size_t oldmol, size_t newc)
{ GB_CHECK(gb_accept_identity_swap(b200().e, (int32_t) oldc, (int64_t) oldmol, (int32_t) newc)); }

// VolumeMove (mc_box.h:196-320).  ScalePositions<<<>>> has just scaled the reference's Sim.Box (cell, inverse cell, volume, kmax and reciprocal
// cutoff, :66-94); that box goes to the engine, which moves every adsorbate molecule with its first atom, rebuilds its wave-vector table and
// evaluates Total_VDW_Coulomb_Energy + Ewald_TotalEnergy of the scaled system in one call (:236-252).  The old state stays on the device
// until b200_volume_finish tells the decision (CopyScaledPositions + the swap of the structure factors / Revert_Boxsize, :292-306).
inline MoveEnergy b200_volume_trial(Simulations& Sim, double Scale, bool& overlap)
{
  cudaDeviceSynchronize();
  gb_box b{};
  cudaMemcpy(b.cell, Sim.Box.Cell, 9 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(b.inverse_cell, Sim.Box.InverseCell, 9 * sizeof(double), cudaMemcpyDeviceToHost);
  b.cubic = Sim.Box.Cubic; b.volume = Sim.Box.Volume; b.alpha = Sim.Box.Alpha; b.prefactor = Sim.Box.Prefactor;
  b.kmax[0] = Sim.Box.kmax.x; b.kmax[1] = Sim.Box.kmax.y; b.kmax[2] = Sim.Box.kmax.z; b.reciprocal_cutoff = Sim.Box.ReciprocalCutOff;
  b.use_lammps_ewald = Sim.Box.UseLAMMPSEwald;
  gb_move_energy m; int32_t ov = 0;
  GB_CHECK(gb_volume_move_trial(b200().e, &b, Scale, &m, &ov));
  overlap = ov != 0;
  MoveEnergy E;
  E.HHVDW = m.HHVDW; E.HGVDW = m.HGVDW; E.GGVDW = m.GGVDW; E.HHReal = m.HHReal; E.HGReal = m.HGReal; E.GGReal = m.GGReal;
  E.HHEwaldE = m.HHEwaldE; E.HGEwaldE = m.HGEwaldE; E.GGEwaldE = m.GGEwaldE;
  return E;
}
inline void b200_volume_finish(bool accept) { if(b200().e) GB_CHECK(gb_volume_move_finish(b200().e, accept ? 1 : 0)); }

// state commits (mc_utilities.h:294-417, move_struct.h:271,371): each includes the swap of the structure-factor vectors
inline void b200_accept_translation(size_t comp) { GB_CHECK(gb_accept_translation(b200().e, (int32_t) comp)); }
inline void b200_accept_insertion(size_t comp) { GB_CHECK(gb_accept_insertion(b200().e, (int32_t) comp)); }
inline void b200_accept_deletion(size_t comp, size_t molecule) { GB_CHECK(gb_accept_deletion(b200().e, (int32_t) comp, (int64_t) molecule)); }
inline void b200_reinsertion_store(size_t comp) { GB_CHECK(gb_reinsertion_store(b200().e, (int32_t) comp)); }
inline void b200_accept_reinsertion(size_t comp, size_t molecule) { GB_CHECK(gb_accept_reinsertion(b200().e, (int32_t) comp, (int64_t) molecule)); }

// before the reference's own FINAL energy check (fxn_main.h:282-404, which reads Sims.d_a): the engine's state goes back into the
// reference's device arrays, so that the reference's CPU and GPU total-energy routines judge the run (ENERGY DRIFT, fxn_main.h:467-468)
inline void b200_sync_back(Variables& Vars, size_t sim, bool report = true)
{
  B200Binding& G = b200();
  if(!G.e) return;
  Components& SC = Vars.SystemComponents[sim]; Simulations& Sims = Vars.Sims[sim];
  const size_t nc = (size_t) SC.NComponents.x;
  std::vector<Atoms> dev(nc);
  cudaMemcpy(dev.data(), Sims.d_a, nc * sizeof(Atoms), cudaMemcpyDeviceToHost);
  for(size_t c = 0; c < nc; c++)
  {
    const size_t n = dev[c].Allocate_size;
    std::vector<double> pos(3 * n), scale(n), charge(n), scoul(n); std::vector<uint64_t> type(n), molid(n); int64_t live = 0;
    GB_CHECK(gb_download_atoms(G.e, (int32_t) c, pos.data(), scale.data(), charge.data(), scoul.data(), type.data(), molid.data(), &live));
    cudaMemcpy(dev[c].pos, pos.data(), 3 * n * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(dev[c].scale, scale.data(), n * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(dev[c].charge, charge.data(), n * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(dev[c].scaleCoul, scoul.data(), n * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(dev[c].Type, type.data(), n * sizeof(size_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dev[c].MolID, molid.data(), n * sizeof(size_t), cudaMemcpyHostToDevice);
    dev[c].size = (size_t) live;
  }
  cudaMemcpy(Sims.d_a, dev.data(), nc * sizeof(Atoms), cudaMemcpyHostToDevice);
  if(!Vars.FF.noCharges)       // a deck without charges has no structure-factor arrays (Allocate_Copy_Ewald_Vector is not reached)
  {
    const size_t nvec = (size_t) (Sims.Box.kmax.x + 1) * (2 * Sims.Box.kmax.y + 1) * (2 * Sims.Box.kmax.z + 1);
    std::vector<double> ads(2 * nvec), fw(2 * nvec);
    GB_CHECK(gb_download_structure_factors(G.e, ads.data(), fw.data(), nullptr));
    cudaMemcpy(Sims.Box.AdsorbateEik, ads.data(), 2 * nvec * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(Sims.Box.FrameworkEik, fw.data(), 2 * nvec * sizeof(double), cudaMemcpyHostToDevice);
  }
  cudaDeviceSynchronize();
  int64_t launches = 0; gb_launch_count(G.e, &launches, 0);
  if(report) fprintf(stderr, "graspa_b200 overlay: %lld engine kernel launches served the reference's drivers\n", (long long) launches);
}
