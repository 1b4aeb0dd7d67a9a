/* graspa_b200 -- C ABI of the B200-native Monte Carlo energy engine.
 *
 * This is the drop-in boundary for gRASPA's data-parallel hot path.
 * The reference has no FFI; its seam is the set of free functions / kernels that the kept
 * move drivers (mc_widom.h, mc_swap_utilities.h, mc_swap_moves.h, mc_single_particle.h)
 * call.  Every entry point below names the reference interface it replaces
 * (file:line relative to /root/reference/src_clean).  INTEGRATION.md shows the binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all structs are POD; no C++/torch types.
 *   - every call returns an int status (GB_OK == 0); gb_last_error() gives the text.
 *     The reference printf+exit()s on CUDA errors (VDW_Coulomb.cuh:58-66); a binding keeps
 *     that behaviour by checking the status.
 *   - calls are blocking unless their name ends in _async; one CUDA stream per engine;
 *     no global state; one engine per GPU (one process per GPU in multi-GPU runs).
 *   - host pointers may be pageable or pinned; arguments named d_* are DEVICE pointers.
 *   - there is NO CPU fallback: without a CUDA device gb_engine_create fails.
 *   - energies are in the reference's internal units (10 J/mol), lengths in Angstrom.
 *   - FP64 throughout.
 */
#ifndef GRASPA_B200_H
#define GRASPA_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define GB_ABI_VERSION 9

enum {
  GB_OK = 0,
  GB_ERR_CUDA = 1,          /* a CUDA runtime call failed */
  GB_ERR_ARG = 2,           /* bad argument */
  GB_ERR_STATE = 3,         /* call made before the required upload */
  GB_ERR_CAPACITY = 4,      /* component out of allocated space (mc_utilities.h:111-115 throws) */
  GB_ERR_UNIMPLEMENTED = 5
};

/* MoveTypes, data_struct.h:20 */
enum { GB_TRANSLATION = 0, GB_ROTATION, GB_SINGLE_INSERTION, GB_SINGLE_DELETION, GB_SPECIAL_ROTATION,
       GB_INSERTION, GB_DELETION, GB_REINSERTION, GB_CBCF_LAMBDACHANGE, GB_CBCF_INSERTION,
       GB_CBCF_DELETION, GB_IDENTITY_SWAP, GB_WIDOM };
/* CBMC_Types, data_struct.h:22 */
enum { GB_CBMC_INSERTION = 0, GB_CBMC_DELETION, GB_REINSERTION_INSERTION, GB_REINSERTION_RETRACE,
       GB_IDENTITY_SWAP_NEW, GB_IDENTITY_SWAP_OLD };

typedef struct gb_engine gb_engine;   /* opaque */

/* Boxsize, data_struct.h:865-886 (scalar members; the device pointers of the reference struct are engine-owned) */
typedef struct {
  double cell[9];            /* rows = lattice vectors */
  double inverse_cell[9];
  double volume;
  double alpha;
  double prefactor;          /* 138935.483496 */
  double reciprocal_cutoff;
  int32_t kmax[3];
  int32_t cubic;
  int32_t use_lammps_ewald;
  int32_t reserved;
} gb_box;

/* ForceField, data_struct.h:838-855.  Tables have size*size entries, row = typeA*size + typeB. */
typedef struct {
  const double* epsilon; const double* sigma; const double* z; const double* shift; const double* c10;
  double cutoff_vdw_sq;      /* FF.CutOffVDW (already squared) */
  double cutoff_coul_sq;
  double overlap_criteria;
  int32_t size;
  int32_t no_charges;
  int32_t vdw_real_bias;
  int32_t use1264;
} gb_forcefield;

/* Tail table, Components::TailCorrection (data_struct.h:720-724, :1166), size*size */
typedef struct { const int32_t* use_tail; const double* energy; int32_t size; int32_t reserved; } gb_tail_table;

/* One component's Atoms, data_struct.h:788-799 (host side of Copy_Atom_data_to_device, fxn_main.h:46-96).
 * The first n_live slots are live; slots up to n_alloc are uploaded as given (slot 0 of an adsorbate
 * always holds the template molecule, read_data.cpp:2122-2147). */
typedef struct {
  const double* pos;         /* 3*n_upload, xyz interleaved (double3) */
  const double* scale; const double* charge; const double* scale_coul;
  const uint64_t* type; const uint64_t* molid;   /* size_t in the reference */
  int64_t n_upload;          /* slots given in the arrays above (>= n_live) */
  int64_t n_live;            /* Atoms.size */
  int64_t n_alloc;           /* Atoms.Allocate_size */
  int64_t molsize;           /* Atoms.Molsize */
} gb_atoms;

/* MoveEnergy, data_struct.h:416-431, same member order */
typedef struct {
  double storedHGVDW, storedHGReal, storedHGEwaldE;
  double HHVDW, HGVDW, GGVDW;
  double HHReal, HGReal, GGReal;
  double HHEwaldE, HGEwaldE, GGEwaldE;
  double TailE, DNN_E;
} gb_move_energy;

/* Result of one CBMC stage = what CBMC_FirstBead_Finish / the tail of Widom_Move_Chain_PARTIAL leave in
 * CBMC_Variables (mc_widom.h:305-383, 568-611) */
typedef struct {
  double rosenbluth;         /* stage Rosenbluth weight (already / NumberOfTrials, with the VDWRealBias correction) */
  double stored_r;           /* REINSERTION_INSERTION: Rosenbluth - Rosen[selected]  (mc_widom.h:365) */
  double energy[4];          /* selected trial: HGVDW, HGReal, GGVDW, GGReal */
  double selected_pos[3];    /* first-bead stage: position of the selected first bead */
  int32_t success;           /* Goodconstruction */
  int32_t selected;          /* REALselected: index among all trials (Trialindex[SelectedTrial]) */
  int32_t n_survivors;       /* trials without overlap */
  int32_t reserved;
} gb_cbmc_result;

/* ------------------------------------------------------------------------------------------------
 * life cycle
 * ------------------------------------------------------------------------------------------------ */
int  gb_abi_version(void);
const char* gb_last_error(void);
/* device < 0: current device.  Fails (GB_ERR_CUDA) when no CUDA device is present -- there is no CPU path. */
int  gb_engine_create(gb_engine** out, int device);
int  gb_engine_destroy(gb_engine* e);
int  gb_device_info(gb_engine* e, int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_per_block_optin);
int  gb_synchronize(gb_engine* e);
void* gb_stream(gb_engine* e);                       /* cudaStream_t, for callers that time on it */

/* ------------------------------------------------------------------------------------------------
 * set-up  (replaces fxn_main.h:12-280: Copy_ForceField_to_GPU, Copy_Atom_data_to_device,
 *          Setup_Temporary_Atoms_Structure, Prepare_Widom, Allocate_Copy_Ewald_Vector; main.cpp:97-104, 243-300)
 * ------------------------------------------------------------------------------------------------ */
int  gb_upload_forcefield(gb_engine* e, const gb_forcefield* ff, const gb_tail_table* tail /* may be NULL */);
int  gb_upload_box(gb_engine* e, const gb_box* box);
/* declares the component table: n_total components, the first n_host of them framework components */
int  gb_set_components(gb_engine* e, int32_t n_total, int32_t n_host);
int  gb_upload_atoms(gb_engine* e, int32_t component, const gb_atoms* atoms);
/* Atoms of one component back to the host (restart writers, write_data.h:109-263).  Arrays sized n_alloc. */
int  gb_download_atoms(gb_engine* e, int32_t component, double* pos, double* scale, double* charge, double* scale_coul,
                       uint64_t* type, uint64_t* molid, int64_t* n_live);
/* Snapshot of `count` LIVE molecules starting at molecule `first` (positions 3 x n, charge / scale / scale_coul n, any may be
 * NULL): what the RASPA-2 restart and LAMMPS movie writers need (write_data.h:109-263, axpy.cu:25-69) without copying the
 * component's whole Allocate_size back. */
int  gb_snapshot_molecules(gb_engine* e, int32_t component, int64_t first, int64_t count, double* pos, double* charge, double* scale, double* scale_coul);
/* stored structure factors, nvec complex (re,im); Allocate_Copy_Ewald_Vector fxn_main.h:235-280 */
int  gb_upload_structure_factors(gb_engine* e, const double* adsorbate_eik, const double* framework_eik);
int  gb_download_structure_factors(gb_engine* e, double* adsorbate_eik, double* framework_eik, double* temp_eik);
/* rigid exclusion constants per component, Calculate_Exclusion_Energy_Rigid ewald_preparation.h:351-366 */
int  gb_set_exclusion_constants(gb_engine* e, int32_t component, double exclusion_intra, double exclusion_atom, int32_t rigid, int32_t has_partial_charge);
/* the device random pool, RandomNumber::DeviceRandom/ResetRandom data_struct.h:1300-1327: n double3 */
int  gb_upload_random_pool(gb_engine* e, const double* random3, int64_t n);
/* block pockets of one component: n spheres, Cartesian centres (3n) and radii (n) as ReadBlockingPockets /
 * ReplicateBlockPockets leave them (read_data.cpp:3290-3454), invert = InvertBlockPockets.  Replaces the host-side
 * BlockedPocket() calls of the kept drivers (read_data.cpp:3466-3640; mc_widom.h:445-497 first-bead trials,
 * mc_swap_utilities.h:35-78 / move_struct.h:208-250 / mc_swap_moves.h:299-330 grown molecules,
 * mc_single_particle.h:83-119 translation / rotation proposals): the move kernels run the test on the device and report
 * a blocked trial as an overlap, a blocked grown molecule as a failed construction.  n = 0 removes the list. */
int  gb_set_block_pockets(gb_engine* e, int32_t component, int32_t n, const double* centers, const double* radii, int32_t invert);
/* Widom/CBMC parameters, WidomStruct data_struct.h:1271-1278 and Components::Beta */
int  gb_set_cbmc(gb_engine* e, int32_t n_trial_positions, int32_t n_trial_orientations, double beta);
/* host mirror of Components::NumberOfPseudoAtoms (live atoms of each force-field type), kept by the engine
 * and updated by the accept calls; this returns it (size = ff.size) */
int  gb_get_pseudo_atom_counts(gb_engine* e, int64_t* counts);

/* ------------------------------------------------------------------------------------------------
 * CBMC stages  (replaces get_random_trial_position<<<>>> mc_widom.h:122-213,
 *   get_random_trial_orientation<<<>>> :215-303, CBMC_PairwiseInteractions :89-119 with
 *   Calculate_Multiple_Trial_Energy_VDWReal<<<>>> VDW_Coulomb.cu:1183-1352 and Host_sum_Widom_HGGG_SEPARATE
 *   mc_widom.h:42-87, and the Boltzmann/selection/Rosenbluth host code :14-39, 305-383, 568-611).
 *   One launch per stage, one small result read; trial coordinates stay on the device between stages.
 * ------------------------------------------------------------------------------------------------ */
/* first bead.  molecule = SelectedMolInComponent (ignored for CBMC_INSERTION / IDENTITY_SWAP_NEW);
 * pool_offset = Random.offset; uniform = the Get_Uniform_Random() SelectTrialPosition would draw (consumed only
 * for CBMC_INSERTION / REINSERTION_INSERTION with >=1 survivor -- *uniform_used reports it);
 * scale = proposed_scale {vdw, coulomb}; stored_r = CBMC.StoredR (input for REINSERTION_RETRACE);
 * exclude_{comp,mol} = Sims.ExcludeList[0] (-1 for none).
 * For IDENTITY_SWAP_NEW the single trial position is preset_pos, or with preset_pos NULL the first atom of molecule
 * exclude_mol of component exclude_comp read on the device (copy_firstbead_to_new, mc_swap_moves.h:178-181, :267). */
int  gb_cbmc_first_bead(gb_engine* e, int32_t cbmc_type, int32_t component, int64_t molecule, int64_t pool_offset,
                        double uniform, const double scale[2], double stored_r, int32_t exclude_comp, int64_t exclude_mol,
                        const double* preset_pos, gb_cbmc_result* result, int32_t* uniform_used);
/* chain growth from the first bead selected by the previous gb_cbmc_first_bead call (FirstBeadTrial) */
int  gb_cbmc_chain(gb_engine* e, int32_t cbmc_type, int32_t component, int64_t molecule, int64_t pool_offset,
                   double uniform, int32_t exclude_comp, int64_t exclude_mol, gb_cbmc_result* result, int32_t* uniform_used);
/* positions of the molecule grown by the last first_bead(+chain) pair: 3*molsize doubles (block-pocket checks,
 * mc_swap_utilities.h:46-80, read them) */
int  gb_cbmc_grown_positions(gb_engine* e, int32_t component, double* pos);
/* reinsertion: keep the molecule just grown (REINSERTION_INSERTION stages) aside while the old one is retraced
 * (StoreNewLocation_Reinsertion<<<>>> into tempMolStorage, mc_swap_moves.h:27-41, move_struct.h:271) */
int  gb_reinsertion_store(gb_engine* e, int32_t component);
/* lower level: energies of caller-supplied trial atoms (n_trials groups of chainsize atoms), for drivers that
 * generate trials themselves.  out_energy[t*4 + {HGVDW,HGReal,GGVDW,GGReal}], out_flag[t]. */
int  gb_trial_energies(gb_engine* e, int32_t n_trials, int32_t chainsize, const double* pos, const double* scale,
                       const double* charge, const double* scale_coul, const uint64_t* type,
                       int32_t new_comp, int64_t new_molid, int32_t exclude_comp, int64_t exclude_mol,
                       double* out_energy, int32_t* out_flag);

/* ------------------------------------------------------------------------------------------------
 * fused moves: everything the kept drivers compute between choosing a move and drawing its acceptance random number,
 * in ONE call with ONE host round trip (the reference needs 3-6 synchronisations per move).  Stages chain on the
 * device; a failed stage turns the following ones into no-ops, exactly like the early returns of the drivers.
 * The caller passes the uniforms the stages WOULD draw (peeked from its stream) and advances its stream by
 * uniforms_used and its pool offset by pool_used afterwards.  A pool refill (RandomNumber::Check) must not fall
 * inside the move: across a refill use the stage calls above.
 *   gb_move_insertion     Insertion_Body            mc_swap_utilities.h:3-133   (swap insertion and Widom)
 *   gb_move_deletion      Deletion_Body             mc_swap_utilities.h:135-225
 *   gb_move_reinsertion   ReinsertionMove::Calculate_Insertion/_Deletion/_AdjustRosenbluth  move_struct.h:186-338
 *   gb_move_single_body   SingleBody_Prepare + SingleBody_Calculation  mc_single_particle.h:10-241
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  gb_cbmc_result first_bead, chain;            /* insertion leg (or the only leg) */
  gb_cbmc_result old_first_bead, old_chain;    /* retrace leg of a reinsertion */
  double ewald[2];                             /* {same, 2*cross} with the exclusion term, as GPU_EwaldDifference_General returns */
  double tail;                                 /* TailCorrectionDifference */
  gb_move_energy delta;                        /* single body: new - old pair energies */
  int32_t overlap;                             /* single body: flag[0] */
  int32_t uniforms_used;                       /* how many of the uniforms passed in the reference would have drawn */
  int32_t pool_used;                           /* pool entries the reference would have consumed (Random.Update) */
  int32_t success;                             /* growth succeeded (Rosenbluth > 1e-150) / no overlap */
} gb_move_result;

int  gb_move_insertion(gb_engine* e, int32_t component, int64_t pool_offset, const double uniforms[2], const double scale[2], gb_move_result* out);
int  gb_move_deletion(gb_engine* e, int32_t component, int64_t molecule, int64_t pool_offset, const double scale[2], gb_move_result* out);
int  gb_move_reinsertion(gb_engine* e, int32_t component, int64_t molecule, int64_t pool_offset, const double uniforms[2], gb_move_result* out);
/* the energy part of IdentitySwapMove (mc_swap_moves.h:256-362) in one kernel: result slots first_bead / chain = growth of the
 * new species, old_first_bead / old_chain = retrace of the molecule that leaves; ewald and tail as the reference's
 * GPU_EwaldDifference_IdentitySwap / TailCorrectionIdentitySwap return them; the grown molecule is left in tempMolStorage
 * for gb_accept_identity_swap.  uniform = the draw of the new species' orientation selection (see uniforms_used). */
int  gb_move_identity_swap(gb_engine* e, int32_t old_component, int64_t old_molecule, int32_t new_component, int64_t pool_offset,
                           double uniform, gb_move_result* out);
int  gb_move_single_body(gb_engine* e, int32_t move_type, int32_t component, int64_t molecule, const double max_change[3],
                         int64_t pool_offset, gb_move_result* out);

/* ------------------------------------------------------------------------------------------------
 * translation / rotation  (replaces get_new_position<<<>>> mc_utilities.h:485-606 and
 *   Calculate_Single_Body_Energy_VDWReal<<<>>> VDW_Coulomb.cu:626-841 + the host sum mc_single_particle.h:183-200)
 * ------------------------------------------------------------------------------------------------ */
int  gb_single_body_propose(gb_engine* e, int32_t move_type, int32_t component, int64_t molecule,
                            const double max_change[3], int64_t pool_offset, double* new_pos /* 3*molsize, may be NULL */);
/* energy delta (new - old) of the proposal; overlap = flag[0] */
int  gb_single_body_delta(gb_engine* e, int32_t component, int32_t do_new, int32_t do_old, gb_move_energy* delta, int32_t* overlap);
/* caller-supplied old/new atoms instead of a stored proposal (n atoms each) */
int  gb_single_body_delta_explicit(gb_engine* e, int32_t component, int64_t molid, int32_t n,
                                   const double* old_pos, const double* new_pos, const double* scale, const double* charge,
                                   const double* scale_coul, const uint64_t* type, int32_t do_new, int32_t do_old,
                                   gb_move_energy* delta, int32_t* overlap);

/* ------------------------------------------------------------------------------------------------
 * Ewald Fourier deltas  (replaces GPU_EwaldDifference_General Ewald_Energy_Functions.h:438-580 with
 *   Initialize_WaveVector_General :162-185 and Fourier_Ewald_Diff :280-397 fused into one launch;
 *   GPU_EwaldDifference_IdentitySwap :582-632; Update_Vector_Ewald :423-433)
 * ------------------------------------------------------------------------------------------------ */
/* returns {same-type, 2*cross-type} exactly like the reference's double2, exclusion constants applied.
 * location = the reference's Location argument (selected trial for INSERTION, UpdateLocation for DELETION/REINSERTION).
 * GB_CBCF_INSERTION (the second step of a fractional insertion) reads the same-type structure factors from tempEik, where the lambda
 * change of the first step left them, and leaves the sum of both steps there (UseTempVector, :362-377, :510-517);
 * GB_CBCF_DELETION is a DELETION whose exclusion term carries scale[1]^2 (:570-575). */
int  gb_ewald_delta(gb_engine* e, int32_t component, int32_t move_type, int64_t location, const double scale[2], double out[2]);
/* old molecule = slots [update_location, +molsize) of old_component; new molecule = tempMolStorage, i.e. the molecule kept by
 * gb_reinsertion_store(new_component) after its IDENTITY_SWAP_NEW growth (mc_swap_moves.h:355) */
int  gb_ewald_delta_identity_swap(gb_engine* e, int32_t old_component, int32_t new_component, int64_t update_location, double out[2]);
/* explicit atoms: n_old old atoms followed by n_new new atoms (Sims.Old layout) */
int  gb_ewald_delta_explicit(gb_engine* e, int32_t component_is_framework, int32_t n_old, int32_t n_new, const double* pos,
                             const double* charge, const double* scale_coul, double out[2]);
int  gb_ewald_commit(gb_engine* e, int32_t component);                 /* Update_Vector_Ewald */

/* ------------------------------------------------------------------------------------------------
 * CB/CFC lambda change of one (fractional) molecule  (replaces, inside CBCF_LambdaChange mc_cbcfc.h:20-140:
 *   Prepare_LambdaChange<<<>>> + Calculate_Single_Body_Energy_VDWReal_LambdaChange<<<>>> VDW_Coulomb.cu:843-1036 + the host
 *   sum of Blocksum; GPU_EwaldDifference_LambdaChange Ewald_Energy_Functions.h:637-788; and the acceptance's scale update.
 *   The Wang-Landau / lambda-bin bookkeeping of mc_cbcfc.h stays with the kept driver.)
 * ------------------------------------------------------------------------------------------------ */
/* delta = E(new_scale) - E(stored scale) of the molecule against every other atom; new_scale = {vdw, coulomb} as
 * Lambda.SET_SCALE returns them; overlap = flag[0] (from the NEW lambda only, :993-994) */
int  gb_lambda_change_delta(gb_engine* e, int32_t component, int64_t molecule, const double new_scale[2], gb_move_energy* delta, int32_t* overlap);
/* {same-type, 2*cross-type}, rigid exclusion x (new^2 - old^2) taken out (:778-780); use_temp_vector: the second step of a
 * CBCF deletion continues from tempEik (:713-716).  Call after gb_lambda_change_delta of the same molecule. */
int  gb_ewald_delta_lambda_change(gb_engine* e, int32_t component, const double old_scale[2], const double new_scale[2],
                                  int32_t use_temp_vector, double out[2]);
int  gb_accept_lambda_change(gb_engine* e, int32_t component, int64_t molecule, const double new_scale[2]);
/* CBCF insertion and deletion (CBCFMove, mc_cbcfc.h:296-447) are two-step moves that put the system into the intermediate state
 * BEFORE the acceptance test and take it back on rejection; the Fourier state of the first step stays in tempEik (UseTempVector).
 *   insertion: gb_lambda_change_delta(fractional molecule -> 1) + gb_ewald_delta_lambda_change(use_temp_vector 0);
 *              gb_cbcf_set_scale(molecule, {1,1})                      -- update_CBCF_scale<<<>>> :312
 *              gb_cbmc_first_bead / gb_cbmc_chain(scale = new lambda); gb_ewald_delta(GB_CBCF_INSERTION) continues from tempEik
 *              accepted: gb_accept_insertion;  rejected: gb_cbcf_set_scale(molecule, old scale)   -- Revert_CBCF_Insertion<<<>>> :359
 *   deletion:  gb_cbmc_first_bead / gb_cbmc_chain(GB_CBMC_DELETION, scale = old lambda); gb_ewald_delta(GB_CBCF_DELETION);
 *              gb_cbcf_deletion_stage(molecule, 0)                     -- Update_deletion_data_fractional<<<>>> + Update_NumberOfMolecules :382-386
 *              gb_lambda_change_delta(new fractional molecule, new lambda) + gb_ewald_delta_lambda_change(use_temp_vector 1)
 *              accepted: gb_accept_lambda_change;  rejected: gb_cbcf_deletion_stage(molecule, 1)  -- Revert_CBCF_Deletion<<<>>> :441-446 */
int  gb_cbcf_set_scale(gb_engine* e, int32_t component, int64_t molecule, const double scale[2]);
int  gb_cbcf_deletion_stage(gb_engine* e, int32_t component, int64_t molecule, int32_t revert);

/* ------------------------------------------------------------------------------------------------
 * tail corrections  (replaces TailCorrection_Energy_Functions.h:3-113)
 * ------------------------------------------------------------------------------------------------ */
int  gb_tail_total(gb_engine* e, double* out);
int  gb_tail_difference(gb_engine* e, int32_t component, int32_t move_type, double* out);
int  gb_tail_identity_swap(gb_engine* e, int32_t new_component, int32_t old_component, double* out);

/* ------------------------------------------------------------------------------------------------
 * state commit  (replaces AcceptTranslation/AcceptInsertion/AcceptDeletion mc_utilities.h:294-417 with
 *   update_translation_position, Update_insertion_data_Parallel, Update_deletion_data_Parallel,
 *   Update_NumberOfMolecules; reinsertion mc_swap_moves.h:27-49)
 * ------------------------------------------------------------------------------------------------ */
int  gb_accept_translation(gb_engine* e, int32_t component);          /* commits the stored single-body proposal (+ Ewald swap) */
int  gb_accept_insertion(gb_engine* e, int32_t component);            /* commits the molecule grown by the last CBMC insertion (+ Ewald swap) */
int  gb_accept_deletion(gb_engine* e, int32_t component, int64_t molecule);
int  gb_accept_reinsertion(gb_engine* e, int32_t component, int64_t molecule);
/* IdentitySwapMove accept branch (mc_swap_moves.h:393-422: Update_deletion_data + Update_IdentitySwap_Insertion_data +
 * Update_NumberOfMolecules x2, or Update_Reinsertion_data when both species are the same; + Update_Vector_Ewald).
 * The new molecule is the one stored by gb_reinsertion_store(new_component) after its IDENTITY_SWAP_NEW growth. */
int  gb_accept_identity_swap(gb_engine* e, int32_t old_component, int64_t old_molecule, int32_t new_component);
/* append a caller-supplied molecule (restart ingestion, CreateMolecule) */
int  gb_append_molecule(gb_engine* e, int32_t component, const double* pos, const double* scale, const double* charge,
                        const double* scale_coul, const uint64_t* type);
int  gb_number_of_molecules(gb_engine* e, int32_t component, int64_t* n);

/* ------------------------------------------------------------------------------------------------
 * totals  (replaces Total_VDW_Coulomb_Energy VDW_Coulomb.cu:1580-1665, Ewald_TotalEnergy
 *   Ewald_Energy_Functions.h:1272-1430 and the CPU Ewald_Total ewald_preparation.h:5-259 used at init)
 * ------------------------------------------------------------------------------------------------ */
int  gb_total_vdw_real(gb_engine* e, gb_move_energy* out);
/* total Fourier energy incl. self and intra-molecular exclusion; store_structure_factors != 0 also (re)initialises
 * AdsorbateEik / FrameworkEik from the current positions (what Ewald_Total + Allocate_Copy_Ewald_Vector do at init) */
int  gb_total_ewald(gb_engine* e, int32_t store_structure_factors, gb_move_energy* out);

/* NPT volume move (VolumeMove, mc_box.h:196-320).  gb_volume_move_trial does what the reference does between drawing the new
 * volume and the acceptance test: ScalePositions (every molecule of components >= 1 follows its first atom, which scales by
 * `scale` = cbrt(V_new / V_old); the other atoms keep their offset), the new box (`new_box`: cell * scale, inverse / scale,
 * volume, kmax and reciprocal cut-off recomputed by the caller as mc_box.h:84-94 does), Total_VDW_Coulomb_Energy with the
 * overlap flag, and Ewald_TotalEnergy in the DEVICE routine's convention (HH, HG, GG separate; Ewald_Energy_Functions.h:1366-1428),
 * whose structure factors become the stored ones.  The tail term is gb_tail_total afterwards (it reads the new volume).
 * gb_volume_move_finish(accept != 0) keeps that state (CopyScaledPositions + the swap of the structure factors);
 * accept == 0 puts positions, box and structure factors back (Revert_Boxsize).  Until finish is called no other state-changing
 * entry point may be used. */
int  gb_volume_move_trial(gb_engine* e, const gb_box* new_box, double scale, gb_move_energy* new_total, int32_t* overlap);
int  gb_volume_move_finish(gb_engine* e, int32_t accept);

/* ------------------------------------------------------------------------------------------------
 * batched Widom insertions  (the whole of Insertion_Body mc_swap_utilities.h:3-133 for n independent ghost
 *   insertions of `component`, as axpy.cu:163-186 issues them one by one)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  /* randoms: a pool of double3 plus, per insertion, the pool index of its first-bead block (n_trial_positions
   * entries) and of its orientation block (n_trial_orientations entries).  fb_index/or_index NULL = packed layout:
   * insertion i uses pool[(i*(ntp+nto)) ...] then pool[(i*(ntp+nto)+ntp) ...].  uniforms: 2 per insertion
   * (first-bead selection, orientation selection). */
  const double* pool3;  int64_t n_pool;
  const int64_t* fb_index; const int64_t* or_index;
  const double* uniforms;
  int32_t inputs_on_device;      /* 0: the four pointers above are host memory (copied inside the call) */
  int32_t n_blocks;              /* block-average bins (Components::Nblock = 5); insertion i -> bin i*n_blocks/n */
  /* sharding (one process per GPU): this call evaluates insertions [global_first, global_first + n) of a
   * job of global_n insertions; bins are assigned on the GLOBAL index so that the all-reduced sums equal the
   * single-GPU ones.  global_n = 0: the call is the whole job. */
  int64_t global_first; int64_t global_n;
  /* optional DEVICE pointer to n_blocks*12 doubles: the block sums of this call are ADDED to it on the engine's stream.  A job
   * evaluated in several calls accumulates there, and a multi-GPU run all-reduces from there (NCCL on the same stream, SURVEY
   * section 5) without staging through the host.  NULL: not used.  (ABI 6) */
  double* sums_device;
  /* RNG-exact replay (ABI 8).  pool3 == NULL (inputs_on_device 0): the pool is the one gb_upload_random_pool left on the device -- no
   * second upload.  resume_first_bead != 0: the first-bead trials of this batch were already evaluated by
   * gb_widom_first_bead_success on that same pool (called with pool3 == NULL and consecutive blocks, fb_index[k] = fb_index[0] +
   * k * n_trial_positions -- the whole pool, or this GPU's share of it; nothing but
   * Widom calls since); their energies are still on the device, so the batch starts at the first-bead SELECTION: every first bead is
   * evaluated once.  Every fb_index[i] must be the start of one of those blocks.  GB_ERR_STATE when the kept energies are not valid. */
  int32_t resume_first_bead; int32_t reserved;
} gb_widom_inputs;

/* per insertion outputs (any may be NULL): out8[i*8 + {W, HGVDW, HGReal, GGVDW, GGReal, GGEwaldE, HGEwaldE, TailE}],
 * stage[i] (0 ok, 1 first bead failed, 2 chain failed, 3 chain failed with no surviving orientation).
 * sums[(bin)*12 + {sumW, sumW2, count, sum(W*E) for the 7 energy terms, n_failed, reserved}] on the host (may be NULL), reduced on
 * the device (RecordRosen data_struct.h:627-652 and widom_energy += E*W axpy.cu:177-185).
 * Batches with >= 64 first-bead trials per 2 A cell of the box (about 70 000 insertions in config E; and every batch of a component
 * with block pockets) take the cell-sorted pair stage, smaller ones the warp-per-insertion kernel; both give the same insertions
 * to summation order. */
int  gb_widom_batch(gb_engine* e, int32_t component, int64_t n, const gb_widom_inputs* in,
                    double* out8, int32_t* stage, int32_t outputs_on_device, double* sums);

/* First-bead stage only, for n pool blocks: code[i] = 1 the first bead at pool3[fb_index[i] ...] succeeds
 * (>= 1 trial without overlap, Rosenbluth >= 1e-150), 0 it fails although a trial survived, 2 no trial survived.
 * This is what a host needs to replay the reference's random-number stream exactly for batched Widom insertions:
 * whether an insertion consumes its orientation block and its second uniform depends only on these codes
 * (mc_widom.h:332-341, mc_swap_utilities.h:19-27).  Host buffers.  pool3 == NULL: the pool gb_upload_random_pool left on the device
 * (n_pool ignored); with consecutive blocks (fb_index[k] = fb_index[0] + k * n_trial_positions) the per-trial energies stay on the device for a following
 * gb_widom_batch(resume_first_bead). */
int  gb_widom_first_bead_success(gb_engine* e, int32_t component, int64_t n, const double* pool3, int64_t n_pool,
                                 const int64_t* fb_index, int32_t* code);

/* ------------------------------------------------------------------------------------------------
 * instrumentation
 * ------------------------------------------------------------------------------------------------ */
/* kernels launched by this engine since creation / last reset (bench.py's gpu_launches) */
int  gb_launch_count(gb_engine* e, int64_t* n, int32_t reset);
/* elapsed device milliseconds of the engine's kernels of one family since last reset, measured with CUDA events
 * on the engine stream: family 0 = pair kernels (for gb_widom_batch: its whole pair stage), 1 = Ewald kernels, 2 = all,
 * 3 = the energy kernel of the cell-sorted Widom pair stage alone (its launches are also inside family 0).  Timing must be
 * enabled first; a timed call synchronises the host after every timed group. */
int  gb_timing_enable(gb_engine* e, int32_t on);
int  gb_timing_read(gb_engine* e, int32_t family, double* ms, int64_t* launches, int32_t reset);
/* The resident move server.  The gb_move_* calls do not launch a kernel per move: the first one starts ONE cooperative kernel (one
 * CTA per SM) that stays resident and executes move after move from a mailbox in pinned host memory; the gb_accept_* calls queue
 * their state change for the next move's command; gb_tail_*, gb_ewald_commit and gb_upload_random_pool run beside it; every other
 * call stops it first (and it leaves by itself after GB_MOVE_SERVER_IDLE_MS, default 200, without a command), so the calling
 * sequence of a driver does not change.  It replaces the per-move launches of the reference's move drivers (mc_swap_utilities.h:3-225,
 * move_struct.h:186-406, mc_single_particle.h:123-241, mc_swap_moves.h:199-431: 3-6 launches + synchronisations per move).  Used only
 * while one engine lives in the process (one engine per GPU) and the device supports cooperative launches; otherwise, with on = 0,
 * or with GB_MOVE_SERVER=0 in the environment, every move is one k_move launch.  Decisions, selections and committed positions are the
 * same either way; energies agree to rounding (<= 1e-12 relative: the server runs out-of-line copies of the stage routines, in which
 * the compiler contracts a few multiply-adds differently).
 * on: 1 / 0 = enable / disable (disabling stops a resident server), -1 = leave unchanged.  starts / commands (may be NULL):
 * server launches and moves executed by it since the engine was created. */
int  gb_move_server(gb_engine* e, int32_t on, int64_t* starts, int64_t* commands);
/* FP64 FMA peak microbenchmark on this device: returns TFLOP/s (2 flop per DFMA) */
int  gb_measure_fp64_peak(gb_engine* e, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* GRASPA_B200_H */
