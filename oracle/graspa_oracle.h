/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the gRASPA hot path.  It is the
 * checker for the CUDA engine: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libgraspa_b200.so)
 * never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_vs_ref.py,
 * tests/test_golden.py) against
 *   - the reference's own routines compiled from /root/reference/src_clean
 *     (oracle/_ref/libgraspa_ref_host.so, built by oracle/build_ref.sh), here, and
 *   - golden vectors generated from those routines (tests/golden/make_golden.py), everywhere.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src_clean).
 */
#ifndef GRASPA_ORACLE_H
#define GRASPA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* data_struct.h:865-886 (Boxsize), host mirror */
typedef struct {
  double cell[9];        /* rows = lattice vectors a,b,c (lower triangular), read_data.cpp:1545-1547 */
  double inv[9];         /* inverse_matrix(cell), maths.cuh:38-56 */
  double volume;
  double alpha;
  double prefactor;      /* 138935.483496, read_data.cpp:611 */
  double recip_cutoff;
  int32_t kmax[3];
  int32_t cubic;
  int32_t use_lammps_ewald;
  int32_t pad_;
} orc_box;

/* data_struct.h:838-855 (ForceField), host mirror; tables are ntypes*ntypes, row = typeA*ntypes+typeB */
typedef struct {
  const double* epsilon; const double* sigma; const double* z; const double* shift; const double* c10;
  double cutoff_vdw_sq; double cutoff_coul_sq; double overlap;
  int32_t ntypes; int32_t no_charges; int32_t vdw_real_bias; int32_t use1264;
} orc_ff;

/* data_struct.h:788-799 (Atoms), all components concatenated; component c owns
 * slots [sum(alloc[0..c)), +alloc[c]), of which the first natoms[c] are live */
typedef struct {
  const int64_t* natoms; const int64_t* molsize;
  const int64_t* alloc;  /* slots per component (>= natoms); NULL = natoms.  Offsets follow alloc. */
  const double* pos;     /* 3*sum(alloc) */
  const double* scale; const double* charge; const double* scale_coul;
  const int64_t* type; const int64_t* molid;
  int32_t ncomp; int32_t nhost;
} orc_system;

/* a set of trial / moved atoms (Sims.New / Sims.Old in the reference) */
typedef struct {
  const double* pos; const double* scale; const double* charge; const double* scale_coul;
  const int64_t* type;
  int64_t n;
} orc_atoms;

enum { ORC_CBMC_INSERTION = 0, ORC_CBMC_DELETION, ORC_REINSERTION_INSERTION, ORC_REINSERTION_RETRACE,
       ORC_IDENTITY_SWAP_NEW, ORC_IDENTITY_SWAP_OLD };                          /* data_struct.h:22 */
enum { ORC_TRANSLATION = 0, ORC_ROTATION, ORC_SINGLE_INSERTION, ORC_SINGLE_DELETION, ORC_SPECIAL_ROTATION,
       ORC_INSERTION, ORC_DELETION, ORC_REINSERTION, ORC_CBCF_LAMBDACHANGE, ORC_CBCF_INSERTION,
       ORC_CBCF_DELETION, ORC_IDENTITY_SWAP, ORC_WIDOM };                       /* data_struct.h:20 */

/* ---- setup-time host arithmetic ---- */
void   orc_inverse_cell(const double* cell, double* inv, double* det);          /* maths.cuh:28-56 */
void   orc_cell_from_cif(double a, double b, double c, double alpha_deg, double beta_deg, double gamma_deg,
                         int nx, int ny, int nz, double* cell);                 /* read_data.cpp:1517-1547 */
void   orc_ewald_setup(double cutoff_coul, double precision, orc_box* box);     /* read_data.cpp:693-702 */
void   orc_ff_mix(int ntypes, const double* eps_in, const double* sig_in, const int* shifted, const int* tail,
                  double cutoff_vdw_sq, double* eps, double* sigma, double* shift, int* use_tail, double* tail_e); /* read_data.cpp:1179-1247, 833-873 */

/* ---- pair primitives ---- */
void   orc_pbc(double* v, const orc_box* box);
int    orc_blocked_pocket(const orc_box* box, const double* pockets4, int n, int invert, const double* pos); /* read_data.cpp:3466-3640 */                                  /* maths.cuh:427-450 */
void   orc_vdw(const double* ffarg, double rr, double scaling, int use1264, double* result); /* maths.cuh:452-494 */
double orc_coulomb_real(double qa, double qb, double r, double scaling, double prefactor, double alpha); /* maths.cuh:496-500 */

/* ---- CBMC trial generation (mc_widom.h:122-303, mc_utilities.h:423-457) ---- */
void   orc_rotate_quaternions(double* vec, const double* rnd3);
void   orc_trial_positions(const orc_box* box, const orc_system* sys, int movetype, int comp, int64_t start_position,
                           int ntrials, const double* rnd3, double scale, double scale_coul,
                           double* tpos, double* tscale, double* tcharge, double* tscale_coul, int64_t* ttype);
void   orc_trial_orientations(const orc_system* sys, int movetype, int comp, int64_t start_position, int chainsize,
                              int norient, const double* rnd3, const double* first_bead_pos,
                              double fb_scale, double fb_scale_coul,
                              double* tpos, double* tscale, double* tcharge, double* tscale_coul, int64_t* ttype);

/* ---- trial-batch pair energies (VDW_Coulomb.cu:1183-1352 + mc_widom.h:42-119) ----
 * out_energy[t*4 + {0 HGvdw, 1 HGreal, 2 GGvdw, 3 GGreal}], out_flag[t]; counts = {pairs, in-vdw, in-coul, in-either}.
 * Summation order = the reference's: one value per (atom, trial atom) thread, 128-wide tree per block, blocks summed serially. */
void   orc_trial_energies(const orc_box* box, const orc_ff* ff, const orc_system* sys,
                          int ntrials, int chainsize, const orc_atoms* trial,
                          int new_comp, int64_t new_molid, int excl_comp, int64_t excl_mol,
                          double* out_energy, int32_t* out_flag, int64_t* counts);

/* ---- Rosenbluth / Boltzmann (mc_widom.h:14-39, 305-383, 568-611) ---- */
int    orc_select_trial(const double* log_boltz, int n, double uniform);
/* returns success flag; rosen[] holds -beta*U of the surviving trials on input */
int    orc_cbmc_finish(int movetype, int is_chain, double* rosen, int nsurv, int ntrials_norm, double uniform,
                       double stored_r_in, double* stored_r_out, int* selected, double* rosenbluth);

/* ---- Ewald (Ewald_Energy_Functions.h:97-185, 280-397, 438-580; ewald_preparation.h:5-366) ---- */
int64_t orc_nvec(const orc_box* box);
/* same_sf / cross_sf / temp_sf: nvec complex (re,im interleaved).  old atoms first, then new (Sims.Old layout).
 * returns {same, 2*cross} before the exclusion term (GPU_EwaldDifference_General :542-543, :579) */
void   orc_ewald_delta(const orc_box* box, const double* pos, const double* charge, const double* scale_coul,
                       int nold, int nnew, const double* same_sf, const double* cross_sf, double* temp_sf,
                       double* out_same_cross, int64_t* n_active);
void   orc_ewald_total(const orc_box* box, const orc_system* sys, int no_charges,
                       double* out_E /* GG, HH, HG as Ewald_Total returns them */, double* sf_ads, double* sf_fw);
void   orc_exclusion_rigid(const orc_box* box, int molsize, const double* pos, const double* charge,
                           const double* scale_coul, double* intra, double* self);

/* ---- tail (TailCorrection_Energy_Functions.h:3-113) ---- */
double orc_tail_total(int ntypes, const int64_t* npseudo, const int32_t* use_tail, const double* tail_e, double volume);
double orc_tail_difference(int ntypes, const int64_t* npseudo, const int32_t* use_tail, const double* tail_e, double volume,
                           const int32_t* species_counts /* ntypes, atoms of each type in one molecule of comp */, int sign);
double orc_tail_identity_swap(int ntypes, const int64_t* npseudo, const int32_t* use_tail, const double* tail_e, double volume,
                              const int32_t* new_counts, const int32_t* old_counts);

/* ---- single-body delta (VDW_Coulomb.cu:626-841 + mc_single_particle.h:183-200) ----
 * out[6] = HHvdw, HHreal, HGvdw, HGreal, GGvdw, GGreal (new - old); flag = overlap of NEW */
void   orc_single_body_delta(const orc_box* box, const orc_ff* ff, const orc_system* sys, int comp, int64_t molid,
                             const orc_atoms* oldm, const orc_atoms* newm, int do_new, int do_old,
                             double* out, int32_t* flag);

/* ---- totals (VDW_Coulomb.cu:34-226) ---- out[6] = HHvdw, HHreal, HGvdw, HGreal, GGvdw, GGreal */
void   orc_total_vdw_real(const orc_box* box, const orc_ff* ff, const orc_system* sys, double* out);

/* ---- one Widom / CBMC insertion (mc_swap_utilities.h:3-133) ----
 * rnd_fb: ntrials double3, rnd_or: norient double3, u_fb/u_or: the two SelectTrialPosition uniforms.
 * out[8] = W, HGvdw, HGreal, GGvdw, GGreal, GGewald, HGewald, tail; out_stage: 0 ok, 1 first bead failed, 2 chain failed.
 * out_sel[2] = selected first-bead trial, selected orientation; out_pos = 3*molsize selected positions. */
typedef struct {
  const double* sf_ads; const double* sf_fw;     /* stored structure factors */
  double excl_intra; double excl_self;           /* rigid exclusion constants of the inserted component */
  double beta;
  int32_t ntrials; int32_t norient; int32_t comp; int32_t has_charge;
  /* tail */
  int32_t ntypes; int32_t has_tail;
  const int64_t* npseudo; const int32_t* use_tail; const double* tail_e; const int32_t* species_counts;
} orc_widom_cfg;
void   orc_widom_insertion(const orc_box* box, const orc_ff* ff, const orc_system* sys, const orc_widom_cfg* cfg,
                           const double* rnd_fb, double u_fb, const double* rnd_or, double u_or,
                           double* out, int32_t* out_stage, int32_t* out_sel, double* out_pos, int64_t* counts);
/* n independent insertions; rnd = n*(ntrials+norient) double3 (first-bead block then orientation block per insertion),
 * uni = n*2.  OpenMP over insertions (nthreads<=0: all).  out = n*8, stage = n.  counts = {pairs, vdw, coul, either, active_k*atoms}. */
void   orc_widom_batch(const orc_box* box, const orc_ff* ff, const orc_system* sys, const orc_widom_cfg* cfg,
                       int64_t n, const double* rnd, const double* uni, int nthreads,
                       double* out, int32_t* stage, int64_t* counts);
int    orc_max_threads(void);

/* ---- the reference's RNG stream (data_struct.cpp:6-11, data_struct.h:1287-1346): libc srand/rand ---- */
void   orc_uniform_stream(int seed, int64_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif
