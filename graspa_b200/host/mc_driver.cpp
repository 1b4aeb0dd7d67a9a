// graspa_b200 host layer -- Monte Carlo driver on top of the C ABI (include/graspa_b200.h).
//
// It keeps the reference's move drivers and simulation loop semantics, including the order in which every random
// number is consumed, so that with the same RandomSeed it takes the same accept/reject decisions as the reference's
// CUDA program:
//   RunMoves / Select_Box_Component_Molecule            axpy.cu:74-298
//   Determine_Number_Of_Steps, Run_Simulation_ForOneBox axpy.cu:580-641
//   InsertionMove / DeletionMove / ReinsertionMove       move_struct.h:4-406
//   Insertion_Body / Deletion_Body                       mc_swap_utilities.h:3-225
//   SingleBodyMove                                       mc_single_particle.h:10-314
//   RandomNumber (device pool of 333 334 double3)        data_struct.h:1287-1346
//   GetPrefactor, Update_Max_Translation/Rotation        mc_utilities.h:315-351, 612-666
//   Print_Widom_Statistics                               print_statistics.cuh:100-200
// All energies come from the engine; this file only sequences calls, draws random numbers and keeps statistics.
//
// Two Widom paths: the reference's one-insertion-at-a-time sequence through the stage calls, and (for decks whose
// only move is Widom insertion) an RNG-exact batched replay: per random pool the engine first evaluates the first-bead
// success of every pool decade, a host walk then assigns pool blocks and uniforms to insertions exactly as the
// sequential program would, and one gb_widom_batch call evaluates all insertions of the pool.
#include "../../include/graspa_b200.h"
#include "deck.hpp"
#include "glibc_rand.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <memory>
#include <thread>

namespace {

[[noreturn]] void die(const std::string& what) { std::fprintf(stderr, "graspa_b200_mc: %s: %s\n", what.c_str(), gb_last_error()); std::exit(EXIT_FAILURE); }
#define GB(call) do { if((call) != GB_OK) die(#call); } while(0)

struct MoveCount { long total = 0, accepted = 0; };

struct Energy   // MoveEnergy, data_struct.h:416-431 (terms this driver touches)
{
  double HHVDW = 0, HGVDW = 0, GGVDW = 0, HHReal = 0, HGReal = 0, GGReal = 0, HHEwald = 0, HGEwald = 0, GGEwald = 0, Tail = 0;
  double total() const { return HHVDW + HGVDW + GGVDW + HHReal + HGReal + GGReal + HHEwald + HGEwald + GGEwald + Tail; }
  void add(const Energy& o, double s = 1.0)
  {
    HHVDW += s * o.HHVDW; HGVDW += s * o.HGVDW; GGVDW += s * o.GGVDW; HHReal += s * o.HHReal; HGReal += s * o.HGReal; GGReal += s * o.GGReal;
    HHEwald += s * o.HHEwald; HGEwald += s * o.HGEwald; GGEwald += s * o.GGEwald; Tail += s * o.Tail;
  }
};

struct CompState
{
  // cumulative move probabilities, Move_Statistics::NormalizeProbabilities data_struct.h:569-608
  double cTrans = 0, cRot = 0, cSpecial = 0, cWidom = 0, cReins = 0, cIdentity = 0, cCBCF = 0, cSwap = 0, cVolume = 0, cGibbsXfer = 0, cGibbsVolume = 0, total_prob = 0;
  double max_trans[3] = {1, 1, 1}, max_rot[3] = {0, 0, 0};
  MoveCount trans, rot, ins, del, reins, widom, idswap_add, idswap_remove;
  std::vector<MoveCount> idswap_to;             // IdentitySwap_Total_TO / _Acc_TO per destination component
  MoveCount trans_window, rot_window;           // TranslationTotal/Accepted are reset every 500 cycles
  MoveCount trans_cum, rot_cum;
  double load_sum = 0.0; long load_n = 0;       // production average of the number of molecules (one sample per cycle)                 // CumTranslationTotal/...: what the reference prints (print_statistics.cuh:41-44)
  long nmol = 0;
  bool has_charge = false;
  // Rosenbluth statistics per block: sum W, sum W^2, count; W-weighted widom energies
  std::vector<double> rw, rw2, rn; std::vector<Energy> wE;
};

struct CallClock
{
  double& acc; long& n; std::chrono::steady_clock::time_point t0;
  CallClock(double& a, long& c) : acc(a), n(c), t0(std::chrono::steady_clock::now()) {}
  ~CallClock() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); n++; }
};

struct Sim;
// what the boxes of one run share: ONE glibc rand() stream and ONE random pool (Vars.Random), the trace, the Gibbs statistics
struct Shared
{
  GlibcRand rng;
  std::vector<double> pool; size_t pool_size = 333334, pool_off = 0; long pool_rounds = 0;
  std::FILE* trace = nullptr; long trace_lines = 0;   // one trace line per RunMoves call (moves that do nothing included)
  long moves_done = 0;
  std::vector<Sim*> boxes;
  // --gpus N: replicas of box 0 on the other GPUs; the batched Widom replay cuts every pool into one share per engine (axpy.cu:163-186
  // issues the insertions one by one; they are independent)
  std::vector<gb_engine*> replicas;
  // Gibbs, data_struct.h:66-79
  MoveCount gibbs_vol_window, gibbs_vol_total, gibbs_xfer; double gibbs_max_change = 0.1, gibbs_total_volume = 0.0;
};

struct Sim                                        // one simulation box
{
  Shared& sh;
  GlibcRand& rng; std::vector<double>& pool; size_t& pool_size; size_t& pool_off; long& pool_rounds;
  std::FILE*& trace; long& trace_lines; long& moves_done;
  explicit Sim(Shared& s) : sh(s), rng(s.rng), pool(s.pool), pool_size(s.pool_size), pool_off(s.pool_off), pool_rounds(s.pool_rounds),
                            trace(s.trace), trace_lines(s.trace_lines), moves_done(s.moves_done) { s.boxes.push_back(this); }
  Sim(const Sim&) = delete; Sim& operator=(const Sim&) = delete;
  int box_index = 0;
  deck::Deck d;
  gb_engine* e = nullptr;
  int ncomp = 0;                                  // total components incl. framework (component 0)
  int nhost = 1;                                  // framework components (NComponents.y): 0 = the rigid rest, 1.. = separated, movable ones
  std::vector<CompState> C;
  long total_molecules = 1;                       // TotalNumberOfMolecules counts the framework as one
  Energy running;                                 // SystemComponents.deltaE
  double initial_framework_ewald = 0.0;           // SystemComponents.InitialFrameworkEwald
  Energy createmol_energy;                        // SystemComponents.CreateMol_Energy: the energy the running deltas start from
  MoveCount vol_window, vol_total; double vol_max_change = 0.025;   // VolumeMoveAttempts/Accepted, VolumeMoveMaxChange (data_struct.h:1028-1034)
  int nblock = 5; long block_size = 1; bool production = false;
  int device = -1;                                // CUDA device of the engine (-1: the current one)
  bool fused = true;                              // one host round trip per move (gb_move_*); false: the stage calls
  double widom_s[5] = {0, 0, 0, 0, 0};            // batched Widom replay: host seconds in {pool refill (rand + upload), classification, batch calls, averages, walk}
  double call_s[5] = {0, 0, 0, 0, 0}; long call_n[5] = {0, 0, 0, 0, 0};   // host time inside gb_move_{insertion,deletion,reinsertion,single_body,identity_swap}
};

inline int comp_ms(const Sim& S, int c) { return c == 0 ? 0 : (c < S.nhost ? S.d.fw[c - 1].molsize : S.d.comps[c - S.nhost].ms()); }
inline const char* comp_name(const Sim& S, int c) { return c == 0 ? S.d.framework_name.c_str() : (c < S.nhost ? S.d.fw[c - 1].name.c_str() : S.d.comps[c - S.nhost].name.c_str()); }

void pool_reset_host(Sim& S)                      // RandomNumber::ResetRandom, the host side: the next pool from the random stream
{
  S.pool_off = 0;
  for(size_t i = 0; i < S.pool_size; i++) { S.pool[3 * i] = S.rng.uniform(); S.pool[3 * i + 1] = S.rng.uniform(); S.pool[3 * i + 2] = S.rng.uniform(); }
  for(size_t i = S.pool_size * 3; i < 1000000; i++) S.rng.uniform();
  S.pool_rounds++;
}

void pool_reset(Sim& S)                           // ... and its copy to every engine
{
  pool_reset_host(S); S.pool_rounds--;
  {
    std::vector<std::thread> up;                           // the replicas' copies travel next to box 0's
    for(gb_engine* r : S.sh.replicas) up.emplace_back([&S, r] { GB(gb_upload_random_pool(r, S.pool.data(), (int64_t) S.pool_size)); });
    for(Sim* b : S.sh.boxes) if(b->e) GB(gb_upload_random_pool(b->e, S.pool.data(), (int64_t) S.pool_size));     // every box reads the same pool
    for(auto& t : up) t.join();
  }
  S.pool_rounds++;
}
inline void pool_check(Sim& S, size_t change) { if(S.pool_off + change >= S.pool_size) pool_reset(S); }
inline void pool_update(Sim& S, size_t change) { S.pool_off += change; }

void setup_engine(Sim& S)
{
  deck::Deck& d = S.d;
  GB(gb_engine_create(&S.e, S.device));
  const int n = d.ntypes();
  gb_forcefield ff{d.eps.data(), d.sigma.data(), d.z.data(), d.shift.data(), d.c10.data(), d.cutoff_vdw * d.cutoff_vdw, d.cutoff_coul * d.cutoff_coul,
                   d.overlap, n, d.no_charges ? 1 : 0, 1 /* VDWRealBias stays true: SURVEY section 5 */, d.use1264 ? 1 : 0};
  gb_tail_table tail{d.use_tail.data(), d.tail_energy.data(), n, 0};
  GB(gb_upload_forcefield(S.e, &ff, &tail));
  gb_box box; std::memset(&box, 0, sizeof(box));
  for(int i = 0; i < 9; i++) { box.cell[i] = d.cell[i]; box.inverse_cell[i] = d.inv[i]; }
  box.volume = d.volume; box.alpha = d.alpha; box.prefactor = d.prefactor; box.reciprocal_cutoff = d.recip_cutoff;
  for(int k = 0; k < 3; k++) box.kmax[k] = d.kmax[k];
  box.cubic = !((std::fabs(d.cell[3]) + std::fabs(d.cell[6]) + std::fabs(d.cell[7])) > 1e-10);
  box.use_lammps_ewald = d.lammps_ewald ? 1 : 0;
  GB(gb_upload_box(S.e, &box));
  S.nhost = 1 + (int) d.fw.size();
  S.ncomp = S.nhost + (int) d.comps.size();
  GB(gb_set_components(S.e, S.ncomp, S.nhost));
  {
    const size_t nf = d.ftype.size();
    std::vector<uint64_t> ty(nf), mol(nf, 0); std::vector<double> one(nf, 1.0);
    for(size_t i = 0; i < nf; i++) ty[i] = (uint64_t) d.ftype[i];
    gb_atoms a{d.fpos.data(), one.data(), d.fcharge.data(), one.data(), ty.data(), mol.data(), (int64_t) nf, (int64_t) nf, (int64_t) nf, (int64_t) nf};
    GB(gb_upload_atoms(S.e, 0, &a));
  }
  for(size_t f = 0; f < d.fw.size(); f++)
  {
    // separated framework component: its own slot range, MolID per molecule, movable (CheckFrameworkCIF read_data.cpp:1700-1790)
    const deck::Deck::FrameworkComponent& F = d.fw[f];
    const size_t n = F.type.size();
    std::vector<uint64_t> ty(n), mol(n); std::vector<double> one(n, 1.0);
    for(size_t i = 0; i < n; i++) { ty[i] = (uint64_t) F.type[i]; mol[i] = (uint64_t) F.molid[i]; }
    gb_atoms a{F.pos.data(), one.data(), F.charge.data(), one.data(), ty.data(), mol.data(), (int64_t) n, (int64_t) n, (int64_t) n, (int64_t) F.molsize};
    GB(gb_upload_atoms(S.e, (int32_t) (f + 1), &a));
    S.C[f + 1].nmol = (long) (n / (size_t) F.molsize);
    bool charged = false;
    for(size_t i = 0; i < n; i++) if(std::fabs(F.charge[i]) > 1e-10) charged = true;
    S.C[f + 1].has_charge = charged;
    S.total_molecules += S.C[f + 1].nmol;
  }
  for(size_t c = 0; c < d.comps.size(); c++)
  {
    const deck::Component& M = d.comps[c];
    const int ms = M.ms();
    std::vector<uint64_t> ty(ms), mol(ms, 0); std::vector<double> one(ms, 1.0);
    for(int i = 0; i < ms; i++) ty[i] = (uint64_t) M.type[i];
    // slot 0 holds the template molecule (read_data.cpp:2122-2147); Allocate_size = AdsorbateAllocateSpace (fxn_main.h:46-62)
    const long nrest = (long) (M.restart_charge.size() / (size_t) std::max(ms, 1));
    if(nrest > 0)
    {
      // RestartFileParser (read_data.cpp:3000-3221): the molecules of the restart file fill the slots from 0
      const size_t n = (size_t) nrest * ms;
      std::vector<uint64_t> rty(n), rmol(n); std::vector<double> rone(n, 1.0);
      for(size_t i = 0; i < n; i++) { rty[i] = (uint64_t) M.type[i % ms]; rmol[i] = (uint64_t) (i / ms); }
      gb_atoms a{M.restart_pos.data(), rone.data(), M.restart_charge.data(), rone.data(), rty.data(), rmol.data(), (int64_t) n, (int64_t) n,
                 (int64_t) std::max<long>(d.adsorbate_allocate, (long) n), ms};
      GB(gb_upload_atoms(S.e, (int32_t) (c + S.nhost), &a));
      S.C[c + S.nhost].nmol = nrest; S.total_molecules += nrest;
    }
    else
    {
      gb_atoms a{M.pos.data(), one.data(), M.charge.data(), one.data(), ty.data(), mol.data(), ms, 0, (int64_t) std::max<long>(d.adsorbate_allocate, ms), ms};
      GB(gb_upload_atoms(S.e, (int32_t) (c + S.nhost), &a));
    }
  }
  GB(gb_set_cbmc(S.e, d.n_trial_positions, d.n_trial_orientations, d.beta));
  // rigid exclusion constants from the template molecule, Calculate_Exclusion_Energy_Rigid ewald_preparation.h:261-298, 351-366
  for(size_t c = 0; c < d.comps.size(); c++)
  {
    const deck::Component& M = d.comps[c];
    double intra = 0.0, self = 0.0; bool charged = false;
    // the routine looks at the component's FIRST molecule as it sits in the system: the molecule definition, or molecule 0 of the
    // restart / LAMMPS data file when the run starts from one (its geometry and charges carry the rounding of that file)
    const bool from_restart = M.restart_charge.size() >= (size_t) M.ms() && M.ms() > 0;
    const double* mp = from_restart ? M.restart_pos.data() : M.pos.data();
    const double* mq = from_restart ? M.restart_charge.data() : M.charge.data();
    if(!d.no_charges)
    {
      for(int i = 0; i + 1 < M.ms(); i++)
        for(int j = i + 1; j < M.ms(); j++)
        {
          double v[3] = {mp[3 * i] - mp[3 * j], mp[3 * i + 1] - mp[3 * j + 1], mp[3 * i + 2] - mp[3 * j + 2]};
          // PBC(), maths.cuh:427-450
          const double* I = d.inv; const double* Cc = d.cell;
          if(box.cubic)
          {
            v[0] -= static_cast<int>(v[0] * I[0] + ((v[0] >= 0.0) ? 0.5 : -0.5)) * Cc[0];
            v[1] -= static_cast<int>(v[1] * I[4] + ((v[1] >= 0.0) ? 0.5 : -0.5)) * Cc[4];
            v[2] -= static_cast<int>(v[2] * I[8] + ((v[2] >= 0.0) ? 0.5 : -0.5)) * Cc[8];
          }
          else
          {
            double sx = I[0] * v[0] + I[3] * v[1] + I[6] * v[2], sy = I[1] * v[0] + I[4] * v[1] + I[7] * v[2], sz = I[2] * v[0] + I[5] * v[1] + I[8] * v[2];
            sx -= static_cast<int>(sx + ((sx >= 0.0) ? 0.5 : -0.5)); sy -= static_cast<int>(sy + ((sy >= 0.0) ? 0.5 : -0.5)); sz -= static_cast<int>(sz + ((sz >= 0.0) ? 0.5 : -0.5));
            v[0] = Cc[0] * sx + Cc[3] * sy + Cc[6] * sz; v[1] = Cc[1] * sx + Cc[4] * sy + Cc[7] * sz; v[2] = Cc[2] * sx + Cc[5] * sy + Cc[8] * sz;
          }
          const double r = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
          intra += d.prefactor * mq[i] * mq[j] * std::erf(d.alpha * r) / r;
        }
      const double ps = d.prefactor * d.alpha / std::sqrt(3.14159265358979323846);
      for(int i = 0; i < M.ms(); i++) { self += ps * mq[i] * mq[i]; if(std::fabs(mq[i]) > 1e-10) charged = true; }
    }
    GB(gb_set_exclusion_constants(S.e, (int32_t) (c + S.nhost), intra, self, 1, charged ? 1 : 0));
    S.C[c + S.nhost].has_charge = charged;
    if(M.use_pockets && !M.pocket_radii.empty())
      GB(gb_set_block_pockets(S.e, (int32_t) (c + S.nhost), (int32_t) M.pocket_radii.size(), M.pocket_centers.data(), M.pocket_radii.data(), M.invert_pockets ? 1 : 0));
  }
}

// ------------------------------------------------------------------------------------------------ probabilities
void setup_probabilities(Sim& S)
{
  for(size_t f = 0; f < S.d.fw.size(); f++)        // ReadFrameworkComponentMoves: translation, rotation, special rotation, reinsertion
  {
    const deck::Deck::FrameworkComponent& F = S.d.fw[f]; CompState& X = S.C[f + 1];
    double t = F.p_translation, r = F.p_rotation, sp = F.p_special, re = F.p_reinsertion;
    double tot = t + r + sp + re;
    if(tot > 1e-10) { t /= tot; r /= tot; sp /= tot; re /= tot; tot = 1.0; }
    X.total_prob = tot;
    X.cTrans = t; X.cRot = r + X.cTrans; X.cSpecial = sp + X.cRot; X.cWidom = X.cSpecial; X.cReins = re + X.cWidom;
    X.cIdentity = X.cReins; X.cCBCF = X.cIdentity; X.cSwap = X.cCBCF;
    X.max_trans[0] = S.d.cell[0] * 0.1; X.max_trans[1] = S.d.cell[4] * 0.1; X.max_trans[2] = S.d.cell[8] * 0.1;
    for(int k = 0; k < 3; k++) X.max_rot[k] = 30.0 / (180 / 3.1415);
    if(sp > 0.0 || re > 0.0) { std::fprintf(stderr, "graspa_b200_mc: special rotation / reinsertion of framework components is not driven by this host program\n"); std::exit(2); }
  }
  for(size_t c = 0; c < S.d.comps.size(); c++)
  {
    const deck::Component& M = S.d.comps[c]; CompState& X = S.C[c + S.nhost];
    double t = M.p_translation, r = M.p_rotation, sp = 0.0, w = M.p_widom, re = M.p_reinsertion, id = M.p_identity, sw = M.p_swap, cb = 0.0;
    // the volume-move probability enters the total but is itself NOT divided by it (NormalizeProbabilities, data_struct.h:569-608):
    // the window [cSwap, 1) it ends up with is vol / total all the same
    const double vol = S.d.volume_move_prob;
    double gx = M.p_gibbs_xfer, gv = S.d.gibbs_volume_prob;
    double tot = t + r + sp + w + re + id + sw + cb + vol + gx + gv;
    if(tot > 1e-10) { t /= tot; r /= tot; sp /= tot; w /= tot; sw /= tot; cb /= tot; re /= tot; id /= tot; gx /= tot; gv /= tot; tot = 1.0; }
    X.total_prob = tot;
    X.cTrans = t; X.cRot = r + X.cTrans; X.cSpecial = sp + X.cRot; X.cWidom = w + X.cSpecial; X.cReins = re + X.cWidom;
    X.cIdentity = id + X.cReins; X.cCBCF = cb + X.cIdentity; X.cSwap = sw + X.cCBCF; X.cVolume = vol + X.cSwap;
    X.cGibbsXfer = gx + X.cVolume; X.cGibbsVolume = gv + X.cGibbsXfer;
    // InitializeMaxTranslationRotation fxn_main.h:151-160 then Prepare... :222-227 (0.1 x box lengths, 30 degrees)
    X.max_trans[0] = S.d.cell[0] * 0.1; X.max_trans[1] = S.d.cell[4] * 0.1; X.max_trans[2] = S.d.cell[8] * 0.1;
    for(int k = 0; k < 3; k++) X.max_rot[k] = 30.0 / (180 / 3.1415);
    X.rw.assign(S.nblock, 0.0); X.rw2.assign(S.nblock, 0.0); X.rn.assign(S.nblock, 0.0); X.wE.assign(S.nblock, Energy());
  }
}

double prefactor(const Sim& S, int comp, bool insertion)          // GetPrefactor, mc_utilities.h:315-351
{
  const deck::Component& M = S.d.comps[comp - S.nhost];
  const double N = (double) S.C[comp].nmol;
  if(insertion) return S.d.beta * M.mol_fraction * S.d.pressure * M.fugacity_coeff * S.d.volume / (1.0 + N);
  return N / (S.d.beta * M.mol_fraction * S.d.pressure * M.fugacity_coeff * S.d.volume);
}

// ------------------------------------------------------------------------------------------------ CBMC growth
struct Growth { bool success = false; double W = 0.0; Energy E; int sel_fb = 0, sel_or = 0; double stored_r = 0.0; };

// Insertion_Body (mc_swap_utilities.h:3-133) with MoveType INSERTION or WIDOM
Growth insertion_body(Sim& S, int comp)
{
  Growth G;
  const int ms = S.d.comps[comp - S.nhost].ms();
  const double scale[2] = {1.0, 1.0};
  gb_cbmc_result r; int32_t used = 0;
  pool_check(S, S.d.n_trial_positions);
  if(S.fused && !(ms > 1 && S.pool_off + S.d.n_trial_positions + S.d.n_trial_orientations >= S.pool_size))
  {
    // one round trip for the whole Insertion_Body
    const double u[2] = {S.rng.peek(0), S.rng.peek(1)};
    gb_move_result m;
    { CallClock cc(S.call_s[0], S.call_n[0]); GB(gb_move_insertion(S.e, comp, (int64_t) S.pool_off, u, scale, &m)); }
    S.rng.advance(m.uniforms_used); pool_update(S, m.pool_used);
    if(!m.success) return G;
    G.sel_fb = m.first_bead.selected; G.sel_or = m.chain.selected;
    double W = m.first_bead.rosenbluth;
    G.E.HGVDW = m.first_bead.energy[0]; G.E.HGReal = m.first_bead.energy[1]; G.E.GGVDW = m.first_bead.energy[2]; G.E.GGReal = m.first_bead.energy[3];
    if(ms > 1)
    {
      W *= m.chain.rosenbluth;
      G.E.HGVDW += m.chain.energy[0]; G.E.HGReal += m.chain.energy[1]; G.E.GGVDW += m.chain.energy[2]; G.E.GGReal += m.chain.energy[3];
    }
    if(!S.d.no_charges && S.C[comp].has_charge) { G.E.GGEwald = m.ewald[0]; G.E.HGEwald = m.ewald[1]; W *= std::exp(-S.d.beta * (m.ewald[0] + m.ewald[1])); }
    W *= std::exp(-S.d.beta * m.tail);
    G.E.Tail = m.tail; G.W = W; G.success = true;
    return G;
  }
  GB(gb_cbmc_first_bead(S.e, GB_CBMC_INSERTION, comp, 0, (int64_t) S.pool_off, S.rng.peek(0), scale, 0.0, -1, -1, nullptr, &r, &used));
  pool_update(S, S.d.n_trial_positions);
  S.rng.advance(used);
  double W = r.rosenbluth;
  if(!r.success || W <= 1e-150) return G;
  G.sel_fb = r.selected;
  G.E.HGVDW = r.energy[0]; G.E.HGReal = r.energy[1]; G.E.GGVDW = r.energy[2]; G.E.GGReal = r.energy[3];
  int sel = r.selected;
  if(ms > 1)
  {
    pool_check(S, S.d.n_trial_orientations);
    GB(gb_cbmc_chain(S.e, GB_CBMC_INSERTION, comp, 0, (int64_t) S.pool_off, S.rng.peek(0), -1, -1, &r, &used));
    pool_update(S, S.d.n_trial_orientations);
    S.rng.advance(used);
    if(!r.success) return G;
    W *= r.rosenbluth;
    if(W <= 1e-150) return G;
    G.sel_or = r.selected; sel = r.selected;
    G.E.HGVDW += r.energy[0]; G.E.HGReal += r.energy[1]; G.E.GGVDW += r.energy[2]; G.E.GGReal += r.energy[3];
  }
  if(!S.d.no_charges && S.C[comp].has_charge)
  {
    double ew[2];
    GB(gb_ewald_delta(S.e, comp, GB_INSERTION, sel, scale, ew));
    G.E.GGEwald = ew[0]; G.E.HGEwald = ew[1];
    W *= std::exp(-S.d.beta * (ew[0] + ew[1]));
  }
  double tail = 0.0;
  GB(gb_tail_difference(S.e, comp, GB_INSERTION, &tail));
  W *= std::exp(-S.d.beta * tail);
  G.E.Tail = tail;
  G.W = W; G.success = true;
  return G;
}

void record_rosen(Sim& S, int comp, double W, const Energy& E, long cycle)      // axpy.cu:175-185, data_struct.h:627-652
{
  long b = cycle / S.block_size; if(b >= S.nblock) b--;
  CompState& X = S.C[comp];
  X.rw[b] += W; X.rw2[b] += W * W; X.rn[b] += 1.0;
  X.wE[b].add(E, W);
}

void trace_move(Sim& S, const char* kind, int comp, long mol, int accepted, double dE)
{
  if(S.trace) std::fprintf(S.trace, "%ld %s %d %ld %d %.12e\n", S.moves_done, kind, comp, mol, accepted, dE);
  S.trace_lines++;
}

// ------------------------------------------------------------------------------------------------ moves
void move_widom(Sim& S, int comp, long cycle)
{
  S.C[comp].widom.total++;
  Growth G = insertion_body(S, comp);
  if(S.production) record_rosen(S, comp, G.success ? G.W : 0.0, G.success ? G.E : Energy(), cycle);
  trace_move(S, "widom", comp, 0, G.success, G.W);
}

void move_insertion(Sim& S, int comp)              // InsertionMove::Run, move_struct.h:71-89
{
  CompState& X = S.C[comp];
  X.ins.total++;
  Growth G = insertion_body(S, comp);
  if(!G.success) { trace_move(S, "insertion", comp, X.nmol, 0, 0.0); return; }
  const double pacc = prefactor(S, comp, true) * G.W / S.d.comps[comp - S.nhost].ideal_rosenbluth;
  const double R = S.rng.uniform();
  if(R < pacc)
  {
    GB(gb_accept_insertion(S.e, comp));
    X.nmol++; S.total_molecules++; X.ins.accepted++;
    S.running.add(G.E);
    trace_move(S, "insertion", comp, X.nmol - 1, 1, G.E.total());
  }
  else trace_move(S, "insertion", comp, X.nmol, 0, 0.0);
}

// Deletion_Body, mc_swap_utilities.h:135-225: retrace of molecule `mol`; W and the molecule's energy terms (sign of an existing molecule)
Growth deletion_body(Sim& S, int comp, long mol)
{
  Growth G;
  CompState& X = S.C[comp];
  const int ms = S.d.comps[comp - S.nhost].ms();
  const double scale[2] = {1.0, 1.0};
  gb_cbmc_result r; int32_t used = 0;
  pool_check(S, S.d.n_trial_positions);
  double W = 0.0; Energy E; double ew[2] = {0.0, 0.0}; double tail = 0.0;
  const bool fused = S.fused && !(ms > 1 && S.pool_off + S.d.n_trial_positions + S.d.n_trial_orientations >= S.pool_size);
  if(fused)
  {
    gb_move_result m;
    { CallClock cc(S.call_s[1], S.call_n[1]); GB(gb_move_deletion(S.e, comp, mol, (int64_t) S.pool_off, scale, &m)); }
    pool_update(S, m.pool_used);
    if(!m.success) return G;
    W = m.first_bead.rosenbluth;
    E.HGVDW = m.first_bead.energy[0]; E.HGReal = m.first_bead.energy[1]; E.GGVDW = m.first_bead.energy[2]; E.GGReal = m.first_bead.energy[3];
    if(ms > 1) { W *= m.chain.rosenbluth; E.HGVDW += m.chain.energy[0]; E.HGReal += m.chain.energy[1]; E.GGVDW += m.chain.energy[2]; E.GGReal += m.chain.energy[3]; }
    ew[0] = m.ewald[0]; ew[1] = m.ewald[1]; tail = m.tail;
  }
  else
  {
    GB(gb_cbmc_first_bead(S.e, GB_CBMC_DELETION, comp, mol, (int64_t) S.pool_off, 0.0, scale, 0.0, -1, -1, nullptr, &r, &used));
    pool_update(S, S.d.n_trial_positions);
    W = r.rosenbluth;
    if(!r.success || W <= 1e-150) return G;
    E.HGVDW = r.energy[0]; E.HGReal = r.energy[1]; E.GGVDW = r.energy[2]; E.GGReal = r.energy[3];
    if(ms > 1)
    {
      pool_check(S, S.d.n_trial_orientations);
      GB(gb_cbmc_chain(S.e, GB_CBMC_DELETION, comp, mol, (int64_t) S.pool_off, 0.0, -1, -1, &r, &used));
      pool_update(S, S.d.n_trial_orientations);
      W *= r.rosenbluth;
      E.HGVDW += r.energy[0]; E.HGReal += r.energy[1]; E.GGVDW += r.energy[2]; E.GGReal += r.energy[3];
    }
    if(W <= 1e-150) return G;
    if(!S.d.no_charges && X.has_charge) GB(gb_ewald_delta(S.e, comp, GB_DELETION, mol * ms, scale, ew));
    GB(gb_tail_difference(S.e, comp, GB_DELETION, &tail));
  }
  if(!S.d.no_charges && X.has_charge)
  {
    W /= std::exp(-S.d.beta * (ew[0] + ew[1]));
    E.GGEwald = -1.0 * ew[0]; E.HGEwald = -1.0 * ew[1];
  }
  W /= std::exp(-S.d.beta * tail);
  E.Tail = -tail;
  G.success = true; G.W = W; G.E = E;
  return G;
}

void move_deletion(Sim& S, int comp, long mol)     // DeletionMove::Run, move_struct.h:148-167
{
  CompState& X = S.C[comp];
  X.del.total++;
  const Growth G = deletion_body(S, comp, mol);
  if(!G.success) { trace_move(S, "deletion", comp, mol, 0, 0.0); return; }
  const double pacc = prefactor(S, comp, false) * S.d.comps[comp - S.nhost].ideal_rosenbluth / G.W;
  const double R = S.rng.uniform();
  if(R < pacc)
  {
    GB(gb_accept_deletion(S.e, comp, mol));
    X.nmol--; S.total_molecules--; X.del.accepted++;
    S.running.add(G.E, -1.0);                                   // energy.take_negative()
    trace_move(S, "deletion", comp, mol, 1, -G.E.total());
  }
  else trace_move(S, "deletion", comp, mol, 0, 0.0);
}

void move_reinsertion(Sim& S, int comp, long mol)  // ReinsertionMove::Run, move_struct.h:169-406
{
  CompState& X = S.C[comp];
  X.reins.total++;
  const int ms = S.d.comps[comp - S.nhost].ms();
  const double scale[2] = {1.0, 1.0};
  gb_cbmc_result r; int32_t used = 0;
  // insertion leg
  pool_check(S, S.d.n_trial_positions);
  if(S.fused && !(S.pool_off + S.d.n_trial_positions + 1 + (ms > 1 ? 2 * S.d.n_trial_orientations : 0) >= S.pool_size))
  {
    const double u[2] = {S.rng.peek(0), S.rng.peek(1)};
    gb_move_result m;
    { CallClock cc(S.call_s[2], S.call_n[2]); GB(gb_move_reinsertion(S.e, comp, mol, (int64_t) S.pool_off, u, &m)); }
    S.rng.advance(m.uniforms_used); pool_update(S, m.pool_used);
    if(!m.success) { trace_move(S, "reinsertion", comp, mol, 0, 0.0); return; }
    double Wn = m.first_bead.rosenbluth, Wo = m.old_first_bead.rosenbluth;
    Energy En, Eo;
    En.HGVDW = m.first_bead.energy[0]; En.HGReal = m.first_bead.energy[1]; En.GGVDW = m.first_bead.energy[2]; En.GGReal = m.first_bead.energy[3];
    Eo.HGVDW = m.old_first_bead.energy[0]; Eo.HGReal = m.old_first_bead.energy[1]; Eo.GGVDW = m.old_first_bead.energy[2]; Eo.GGReal = m.old_first_bead.energy[3];
    if(ms > 1)
    {
      Wn *= m.chain.rosenbluth; Wo *= m.old_chain.rosenbluth;
      En.HGVDW += m.chain.energy[0]; En.HGReal += m.chain.energy[1]; En.GGVDW += m.chain.energy[2]; En.GGReal += m.chain.energy[3];
      Eo.HGVDW += m.old_chain.energy[0]; Eo.HGReal += m.old_chain.energy[1]; Eo.GGVDW += m.old_chain.energy[2]; Eo.GGReal += m.old_chain.energy[3];
    }
    Energy E = En; E.add(Eo, -1.0);
    if(!S.d.no_charges && X.has_charge) { E.GGEwald = m.ewald[0]; E.HGEwald = m.ewald[1]; Wn *= std::exp(-S.d.beta * (m.ewald[0] + m.ewald[1])); }
    const double R = S.rng.uniform();
    if(!(R >= Wn / Wo))
    {
      GB(gb_accept_reinsertion(S.e, comp, mol));
      X.reins.accepted++; S.running.add(E);
      trace_move(S, "reinsertion", comp, mol, 1, E.total());
    }
    else trace_move(S, "reinsertion", comp, mol, 0, 0.0);
    return;
  }
  GB(gb_cbmc_first_bead(S.e, GB_REINSERTION_INSERTION, comp, mol, (int64_t) S.pool_off, S.rng.peek(0), scale, 0.0, -1, -1, nullptr, &r, &used));
  pool_update(S, S.d.n_trial_positions);
  S.rng.advance(used);
  double Wn = r.rosenbluth; const double stored_r = r.stored_r;
  if(!r.success || Wn <= 1e-150) { trace_move(S, "reinsertion", comp, mol, 0, 0.0); return; }
  Energy En; En.HGVDW = r.energy[0]; En.HGReal = r.energy[1]; En.GGVDW = r.energy[2]; En.GGReal = r.energy[3];
  if(ms > 1)
  {
    pool_check(S, S.d.n_trial_orientations);
    GB(gb_cbmc_chain(S.e, GB_REINSERTION_INSERTION, comp, mol, (int64_t) S.pool_off, S.rng.peek(0), -1, -1, &r, &used));
    pool_update(S, S.d.n_trial_orientations);
    S.rng.advance(used);
    if(!r.success) { trace_move(S, "reinsertion", comp, mol, 0, 0.0); return; }
    Wn *= r.rosenbluth;
    if(Wn <= 1e-150) { trace_move(S, "reinsertion", comp, mol, 0, 0.0); return; }
    En.HGVDW += r.energy[0]; En.HGReal += r.energy[1]; En.GGVDW += r.energy[2]; En.GGReal += r.energy[3];
  }
  GB(gb_reinsertion_store(S.e, comp));
  // retrace leg: one first-bead trial + StoredR (mc_widom.h:365-366, 421-426)
  pool_check(S, 1);
  GB(gb_cbmc_first_bead(S.e, GB_REINSERTION_RETRACE, comp, mol, (int64_t) S.pool_off, 0.0, scale, stored_r, -1, -1, nullptr, &r, &used));
  pool_update(S, 1);
  double Wo = r.rosenbluth;
  Energy Eo; Eo.HGVDW = r.energy[0]; Eo.HGReal = r.energy[1]; Eo.GGVDW = r.energy[2]; Eo.GGReal = r.energy[3];
  if(ms > 1)
  {
    pool_check(S, S.d.n_trial_orientations);
    GB(gb_cbmc_chain(S.e, GB_REINSERTION_RETRACE, comp, mol, (int64_t) S.pool_off, 0.0, -1, -1, &r, &used));
    pool_update(S, S.d.n_trial_orientations);
    Wo *= r.rosenbluth;
    Eo.HGVDW += r.energy[0]; Eo.HGReal += r.energy[1]; Eo.GGVDW += r.energy[2]; Eo.GGReal += r.energy[3];
  }
  Energy E = En; E.add(Eo, -1.0);
  if(!S.d.no_charges && X.has_charge)
  {
    double ew[2];
    GB(gb_ewald_delta(S.e, comp, GB_REINSERTION, mol * ms, scale, ew));
    E.GGEwald = ew[0]; E.HGEwald = ew[1];
    Wn *= std::exp(-S.d.beta * (ew[0] + ew[1]));
  }
  const double R = S.rng.uniform();                          // drawn before the probability test (move_struct.h:350)
  const double pacc = Wn / Wo;
  if(!(R >= pacc))
  {
    GB(gb_accept_reinsertion(S.e, comp, mol));
    X.reins.accepted++;
    S.running.add(E);
    trace_move(S, "reinsertion", comp, mol, 1, E.total());
  }
  else trace_move(S, "reinsertion", comp, mol, 0, 0.0);
}

void move_single_body(Sim& S, int comp, long mol, int move_type)   // SingleBodyMove, mc_single_particle.h:10-314
{
  CompState& X = S.C[comp];
  MoveCount& cnt = move_type == GB_TRANSLATION ? X.trans : X.rot;
  MoveCount& win = move_type == GB_TRANSLATION ? X.trans_window : X.rot_window;
  cnt.total++; win.total++;
  const int ms = comp_ms(S, comp);
  const double* maxc = move_type == GB_TRANSLATION ? X.max_trans : X.max_rot;
  pool_check(S, ms);
  gb_move_energy d; int32_t overlap = 0; double ew[2] = {0.0, 0.0};
  if(S.fused)
  {
    gb_move_result m;
    { CallClock cc(S.call_s[3], S.call_n[3]); GB(gb_move_single_body(S.e, move_type, comp, mol, maxc, (int64_t) S.pool_off, &m)); }
    pool_update(S, ms);
    d = m.delta; overlap = m.overlap; ew[0] = m.ewald[0]; ew[1] = m.ewald[1];
  }
  else
  {
    GB(gb_single_body_propose(S.e, move_type, comp, mol, maxc, (int64_t) S.pool_off, nullptr));
    pool_update(S, ms);
    GB(gb_single_body_delta(S.e, comp, 1, 1, &d, &overlap));
    if(!overlap && !S.d.no_charges && X.has_charge) { const double scale[2] = {1.0, 1.0}; GB(gb_ewald_delta(S.e, comp, move_type, 0, scale, ew)); }
  }
  if(overlap) { trace_move(S, move_type == GB_TRANSLATION ? "translation" : "rotation", comp, mol, 0, 0.0); return; }
  Energy E; E.HHVDW = d.HHVDW; E.HGVDW = d.HGVDW; E.GGVDW = d.GGVDW; E.HHReal = d.HHReal; E.HGReal = d.HGReal; E.GGReal = d.GGReal;
  if(!S.d.no_charges && X.has_charge)
  {
    // {same type, cross}: a moved framework component books the same-type term as host-host (mc_single_particle.h:214-224)
    if(comp < S.nhost) { E.HHEwald = ew[0]; E.HGEwald = ew[1]; } else { E.GGEwald = ew[0]; E.HGEwald = ew[1]; }
  }
  const double pacc = 1.0 * std::exp(-S.d.beta * E.total());
  const double R = S.rng.uniform();
  if(R < pacc)
  {
    GB(gb_accept_translation(S.e, comp));
    cnt.accepted++; win.accepted++;
    S.running.add(E);
    trace_move(S, move_type == GB_TRANSLATION ? "translation" : "rotation", comp, mol, 1, E.total());
  }
  else trace_move(S, move_type == GB_TRANSLATION ? "translation" : "rotation", comp, mol, 0, 0.0);
}

// IdentitySwapMove, mc_swap_moves.h:199-431: a molecule of OLDComponent is regrown in place as a molecule of NEWComponent
void move_identity_swap(Sim& S)
{
  long adsorbates = 0;
  for(int c = S.nhost; c < S.ncomp; c++) adsorbates += S.C[c].nmol;
  if(adsorbates == 0) return;                                       // :217-221 (TotalNumberOfMolecules - NumberOfFrameworks == 0)
  int oldc = 0, newc = 0; long nold = 0;
  while(oldc == 0 || oldc >= S.ncomp || newc == 0 || newc >= S.ncomp || nold == 0)
  {
    oldc = (int) (size_t) (S.rng.uniform() * (double) (S.ncomp - S.nhost)) + S.nhost;
    newc = (int) (size_t) (S.rng.uniform() * (double) (S.ncomp - S.nhost)) + S.nhost;
    nold = S.C[oldc].nmol;
  }
  const long new_mol = S.C[newc].nmol;
  const long old_mol = (long) (size_t) (S.rng.uniform() * (double) S.C[oldc].nmol);
  CompState& XO = S.C[oldc]; CompState& XN = S.C[newc];
  XO.idswap_remove.total++; XN.idswap_add.total++;
  if(XO.idswap_to.empty()) XO.idswap_to.assign(S.ncomp, MoveCount());
  XO.idswap_to[newc].total++;
  const int ms_new = S.d.comps[newc - S.nhost].ms(), ms_old = S.d.comps[oldc - S.nhost].ms();
  const double scale[2] = {1.0, 1.0};
  gb_cbmc_result r; int32_t used = 0;
  const size_t need = 2 + (ms_new > 1 ? (size_t) S.d.n_trial_orientations : 0) + (ms_old > 1 ? (size_t) S.d.n_trial_orientations : 0);
  if(S.fused && !(S.pool_off + need >= S.pool_size))
  {
    // one kernel, one host round trip: growth at the old molecule's first atom + retrace + Ewald + tail
    gb_move_result m;
    { CallClock cc(S.call_s[4], S.call_n[4]); GB(gb_move_identity_swap(S.e, oldc, old_mol, newc, (int64_t) S.pool_off, S.rng.peek(0), &m)); }
    S.rng.advance(m.uniforms_used); pool_update(S, m.pool_used);
    if(!m.success) { trace_move(S, "identity_swap", oldc, old_mol, 0, 0.0); return; }
    double Wn = m.first_bead.rosenbluth, Wo = m.old_first_bead.rosenbluth;
    Energy En, Eo;
    En.HGVDW = m.first_bead.energy[0]; En.HGReal = m.first_bead.energy[1]; En.GGVDW = m.first_bead.energy[2]; En.GGReal = m.first_bead.energy[3];
    Eo.HGVDW = m.old_first_bead.energy[0]; Eo.HGReal = m.old_first_bead.energy[1]; Eo.GGVDW = m.old_first_bead.energy[2]; Eo.GGReal = m.old_first_bead.energy[3];
    if(ms_new > 1) { Wn *= m.chain.rosenbluth; En.HGVDW += m.chain.energy[0]; En.HGReal += m.chain.energy[1]; En.GGVDW += m.chain.energy[2]; En.GGReal += m.chain.energy[3]; }
    if(ms_old > 1) { Wo *= m.old_chain.rosenbluth; Eo.HGVDW += m.old_chain.energy[0]; Eo.HGReal += m.old_chain.energy[1]; Eo.GGVDW += m.old_chain.energy[2]; Eo.GGReal += m.old_chain.energy[3]; }
    Energy E = En; E.add(Eo, -1.0);
    if(!S.d.no_charges) { E.GGEwald = m.ewald[0]; E.HGEwald = m.ewald[1]; Wn *= std::exp(-S.d.beta * (m.ewald[0] + m.ewald[1])); }
    E.Tail = m.tail; Wn *= std::exp(-S.d.beta * m.tail);
    const double pre = prefactor(S, newc, true) * prefactor(S, oldc, false);
    const double pacc = pre * (Wn / S.d.comps[newc - S.nhost].ideal_rosenbluth) / (Wo / S.d.comps[oldc - S.nhost].ideal_rosenbluth);
    const double R = S.rng.uniform();
    if(R < pacc)
    {
      GB(gb_accept_identity_swap(S.e, oldc, old_mol, newc));
      XN.idswap_add.accepted++; XO.idswap_remove.accepted++; XO.idswap_to[newc].accepted++;
      if(newc != oldc) { XN.nmol++; XO.nmol--; }
      S.running.add(E);
      trace_move(S, "identity_swap", oldc, old_mol, 1, E.total());
    }
    else trace_move(S, "identity_swap", oldc, old_mol, 0, 0.0);
    return;
  }
  // ---- insertion leg: first bead preset to the old molecule's first atom, old molecule excluded (:260-296)
  pool_check(S, 1);
  GB(gb_cbmc_first_bead(S.e, GB_IDENTITY_SWAP_NEW, newc, new_mol, (int64_t) S.pool_off, 0.0, scale, 0.0, oldc, old_mol, nullptr, &r, &used));
  pool_update(S, 1);
  double Wn = r.rosenbluth;
  if(!r.success || Wn <= 1e-150) { trace_move(S, "identity_swap", oldc, old_mol, 0, 0.0); return; }
  Energy En; En.HGVDW = r.energy[0]; En.HGReal = r.energy[1]; En.GGVDW = r.energy[2]; En.GGReal = r.energy[3];
  if(ms_new > 1)
  {
    pool_check(S, S.d.n_trial_orientations);
    GB(gb_cbmc_chain(S.e, GB_IDENTITY_SWAP_NEW, newc, new_mol, (int64_t) S.pool_off, S.rng.peek(0), oldc, old_mol, &r, &used));
    pool_update(S, S.d.n_trial_orientations);
    S.rng.advance(used);
    if(!r.success) { trace_move(S, "identity_swap", oldc, old_mol, 0, 0.0); return; }
    Wn *= r.rosenbluth;
    if(Wn <= 1e-150) { trace_move(S, "identity_swap", oldc, old_mol, 0, 0.0); return; }
    En.HGVDW += r.energy[0]; En.HGReal += r.energy[1]; En.GGVDW += r.energy[2]; En.GGReal += r.energy[3];
  }
  GB(gb_reinsertion_store(S.e, newc));                               // StoreNewLocation_Reinsertion -> tempMolStorage (:299)
  // ---- retrace leg (:336-351)
  pool_check(S, 1);
  GB(gb_cbmc_first_bead(S.e, GB_IDENTITY_SWAP_OLD, oldc, old_mol, (int64_t) S.pool_off, 0.0, scale, 0.0, -1, -1, nullptr, &r, &used));
  pool_update(S, 1);
  double Wo = r.rosenbluth;
  Energy Eo; Eo.HGVDW = r.energy[0]; Eo.HGReal = r.energy[1]; Eo.GGVDW = r.energy[2]; Eo.GGReal = r.energy[3];
  if(ms_old > 1)
  {
    pool_check(S, S.d.n_trial_orientations);
    GB(gb_cbmc_chain(S.e, GB_IDENTITY_SWAP_OLD, oldc, old_mol, (int64_t) S.pool_off, 0.0, -1, -1, &r, &used));
    pool_update(S, S.d.n_trial_orientations);
    Wo *= r.rosenbluth;
    Eo.HGVDW += r.energy[0]; Eo.HGReal += r.energy[1]; Eo.GGVDW += r.energy[2]; Eo.GGReal += r.energy[3];
  }
  Energy E = En; E.add(Eo, -1.0);
  if(!S.d.no_charges)                                               // :354-360
  {
    double ew[2];
    GB(gb_ewald_delta_identity_swap(S.e, oldc, newc, old_mol * ms_old, ew));
    E.GGEwald = ew[0]; E.HGEwald = ew[1];
    Wn *= std::exp(-S.d.beta * (ew[0] + ew[1]));
  }
  double tail = 0.0;
  GB(gb_tail_identity_swap(S.e, newc, oldc, &tail));                 // :361-362
  E.Tail = tail;
  Wn *= std::exp(-S.d.beta * tail);
  const double pre = prefactor(S, newc, true) * prefactor(S, oldc, false);
  const double pacc = pre * (Wn / S.d.comps[newc - S.nhost].ideal_rosenbluth) / (Wo / S.d.comps[oldc - S.nhost].ideal_rosenbluth);
  const double R = S.rng.uniform();
  if(R < pacc)
  {
    GB(gb_accept_identity_swap(S.e, oldc, old_mol, newc));
    XN.idswap_add.accepted++; XO.idswap_remove.accepted++; XO.idswap_to[newc].accepted++;
    if(newc != oldc) { XN.nmol++; XO.nmol--; }
    S.running.add(E);
    trace_move(S, "identity_swap", oldc, old_mol, 1, E.total());
  }
  else trace_move(S, "identity_swap", oldc, old_mol, 0, 0.0);
}

// RunMoves, axpy.cu:102-298
// VolumeMove, mc_box.h:196-320: ln V random walk, molecules follow their first atom, total energies of the scaled system
void move_volume(Sim& S, int comp)
{
  deck::Deck& d = S.d;
  S.vol_window.total++;
  const double oldV = d.volume;
  const double newV = std::exp(std::log(oldV) + S.vol_max_change * 2.0 * (S.rng.uniform() - 0.5));
  const double scale = std::cbrt(newV / oldV), inv_scale = 1.0 / scale;
  gb_box box; std::memset(&box, 0, sizeof(box));
  for(int i = 0; i < 9; i++) { box.cell[i] = d.cell[i] * scale; box.inverse_cell[i] = d.inv[i] * inv_scale; }
  box.volume = newV; box.alpha = d.alpha; box.prefactor = d.prefactor;
  box.cubic = !((std::fabs(d.cell[3]) + std::fabs(d.cell[6]) + std::fabs(d.cell[7])) > 1e-10);
  box.use_lammps_ewald = d.lammps_ewald ? 1 : 0;
  for(int k = 0; k < 3; k++) box.kmax[k] = d.kmax[k];
  box.reciprocal_cutoff = d.recip_cutoff;
  if(!d.no_charges)
  {
    // ScalePositions, mc_box.h:84-94 (1 / pi as the literal the reference multiplies by)
    for(int k = 0; k < 3; k++) box.kmax[k] = (int) std::round(0.25 + box.cell[4 * k] * d.alpha * d.ewald_tol1 * 0.31830988618);
    box.reciprocal_cutoff = std::pow(1.05 * (double) std::max(box.kmax[0], std::max(box.kmax[1], box.kmax[2])), 2);
  }
  gb_move_energy m; int32_t overlap = 0;
  GB(gb_volume_move_trial(S.e, &box, scale, &m, &overlap));
  Energy N; N.HHVDW = m.HHVDW; N.HGVDW = m.HGVDW; N.GGVDW = m.GGVDW; N.HHReal = m.HHReal; N.HGReal = m.HGReal; N.GGReal = m.GGReal;
  N.HHEwald = m.HHEwaldE; N.HGEwald = m.HGEwaldE; N.GGEwald = m.GGEwaldE;
  GB(gb_tail_total(S.e, &N.Tail));
  bool accept = false; Energy D;
  if(!overlap)
  {
    const double nmol = (double) (S.total_molecules - S.nhost);     // Get_TotalNumberOfMolecule_In_Box: no framework "molecules"
    D = N; D.add(S.createmol_energy, -1.0); D.add(S.running, -1.0);
    const double pacc = std::exp((nmol + 1.0) * std::log(newV / oldV) - (D.total() + d.pressure * (newV - oldV)) * d.beta);
    if(S.rng.uniform() < pacc) accept = true;
  }
  GB(gb_volume_move_finish(S.e, accept ? 1 : 0));
  if(accept)
  {
    S.vol_window.accepted++;
    S.running.add(D);
    for(int i = 0; i < 9; i++) { d.cell[i] = box.cell[i]; d.inv[i] = box.inverse_cell[i]; }
    d.volume = newV; d.recip_cutoff = box.reciprocal_cutoff;
    for(int k = 0; k < 3; k++) d.kmax[k] = box.kmax[k];
  }
  // RunMoves books no energy change for this move itself (VolumeMove adds to deltaE on its own, mc_box.h:288): the trace line carries zeros
  trace_move(S, "volume", comp, 0, accept ? 1 : 0, 0.0);
}

void update_max_volume_change(Sim& S, long cycle)     // Update_Max_VolumeChange, mc_utilities.h:691-712
{
  if(S.vol_window.total == 0) return;
  const double ratio = (double) S.vol_window.accepted / (double) S.vol_window.total;
  double c = ratio / 0.5;
  if(c > 1.5) c = 1.5; else if(c < 0.5) c = 0.5;
  S.vol_max_change *= c;
  if(S.vol_max_change < 0.0005) S.vol_max_change = 0.0005;
  if(S.vol_max_change > 0.5) S.vol_max_change = 0.5;
  std::printf("CYCLE: %ld, AccRatio: %.5f, compare_to_target_ratio: %.5f, MaxVolumeChange: %.5f\n", cycle, ratio, c, S.vol_max_change);
  S.vol_total.total += S.vol_window.total; S.vol_total.accepted += S.vol_window.accepted;
  S.vol_window = MoveCount();
}

// CreateMolecule_InOneBox, axpy.cu:300-372: CBMC insertions that are taken whenever they can be built (RANDOM = 1e-100,
// move_struct.h:57-59: no acceptance uniform is drawn) until CreateNumberOfMolecules of each species are in the box
void create_molecules(Sim& S)
{
  for(int comp = S.nhost; comp < S.ncomp; comp++)
  {
    long todo = S.d.comps[comp - S.nhost].create_molecules, fails = 0;
    CompState& X = S.C[comp];
    while(todo > 0)
    {
      Growth G = insertion_body(S, comp);
      const double pacc = G.success ? prefactor(S, comp, true) * G.W / S.d.comps[comp - S.nhost].ideal_rosenbluth : 0.0;
      if(G.success && 1e-100 < pacc)
      {
        GB(gb_accept_insertion(S.e, comp));
        X.nmol++; S.total_molecules++; todo--;
        S.running.add(G.E);
      }
      else if(++fails > 10000000000L) die("bad insertions when creating molecules");
    }
  }
}

// the box and k table of `S` scaled to volume newV (ScalePositions, mc_box.h:66-94)
gb_box scaled_box(const Sim& S, double newV, double& scale)
{
  const deck::Deck& d = S.d;
  scale = std::cbrt(newV / d.volume);
  const double inv_scale = 1.0 / scale;
  gb_box box; std::memset(&box, 0, sizeof(box));
  for(int i = 0; i < 9; i++) { box.cell[i] = d.cell[i] * scale; box.inverse_cell[i] = d.inv[i] * inv_scale; }
  box.volume = newV; box.alpha = d.alpha; box.prefactor = d.prefactor;
  box.cubic = !((std::fabs(d.cell[3]) + std::fabs(d.cell[6]) + std::fabs(d.cell[7])) > 1e-10);
  box.use_lammps_ewald = d.lammps_ewald ? 1 : 0;
  for(int k = 0; k < 3; k++) box.kmax[k] = d.kmax[k];
  box.reciprocal_cutoff = d.recip_cutoff;
  if(!d.no_charges)
  {
    for(int k = 0; k < 3; k++) box.kmax[k] = (int) std::round(0.25 + box.cell[4 * k] * d.alpha * d.ewald_tol1 * 0.31830988618);
    box.reciprocal_cutoff = std::pow(1.05 * (double) std::max(box.kmax[0], std::max(box.kmax[1], box.kmax[2])), 2);
  }
  return box;
}

void adopt_box(Sim& S, const gb_box& box)
{
  deck::Deck& d = S.d;
  for(int i = 0; i < 9; i++) { d.cell[i] = box.cell[i]; d.inv[i] = box.inverse_cell[i]; }
  d.volume = box.volume; d.recip_cutoff = box.reciprocal_cutoff;
  for(int k = 0; k < 3; k++) d.kmax[k] = box.kmax[k];
}

inline double molecules_in_box(const Sim& S) { return (double) (S.total_molecules - S.nhost); }      // Get_TotalNumberOfMolecule_In_Box

// NVTGibbsMove, mc_box.h:322-478: the two boxes exchange volume at constant total volume
void move_gibbs_volume(Sim& S0, int comp)
{
  Shared& H = S0.sh;
  if(H.boxes.size() != 2) { trace_move(S0, "gibbs_volume", comp, 0, 0, 0.0); return; }
  int sel = 0, oth = 1;
  if(S0.rng.uniform() > 0.5) { sel = 1; oth = 0; }
  Sim& A = *H.boxes[sel]; Sim& B = *H.boxes[oth];
  H.gibbs_vol_window.total++;
  const double oldVA = A.d.volume, oldVB = B.d.volume, totalV = oldVA + oldVB;
  const double expdV = std::exp(std::log(oldVA / oldVB) + H.gibbs_max_change * 2.0 * (S0.rng.uniform() - 0.5));
  const double newVA = expdV * totalV / (1.0 + expdV), newVB = totalV - newVA;
  const double cut = std::max(S0.d.cutoff_vdw * S0.d.cutoff_vdw, S0.d.cutoff_coul * S0.d.cutoff_coul);
  if(std::pow(std::cbrt(newVA), 2) < 4.0 * cut || std::pow(std::cbrt(newVB), 2) < 4.0 * cut) { trace_move(S0, "gibbs_volume", comp, 0, 0, 0.0); return; }
  // boxes are scaled in index order; after an overlap in the first one the second is still scaled but not evaluated (mc_box.h:386-389)
  Sim* bx[2] = {H.boxes[0], H.boxes[1]};
  const double newV[2] = {sel == 0 ? newVA : newVB, sel == 0 ? newVB : newVA};
  gb_box nb[2]; Energy N[2]; bool overlap = false;
  for(int k = 0; k < 2; k++)
  {
    double scale = 1.0;
    nb[k] = scaled_box(*bx[k], newV[k], scale);
    gb_move_energy m; int32_t ov = 0;
    GB(gb_volume_move_trial(bx[k]->e, &nb[k], scale, &m, &ov));
    if(overlap) continue;
    if(ov) overlap = true;
    N[k].HHVDW = m.HHVDW; N[k].HGVDW = m.HGVDW; N[k].GGVDW = m.GGVDW; N[k].HHReal = m.HHReal; N[k].HGReal = m.HGReal; N[k].GGReal = m.GGReal;
    N[k].HHEwald = m.HHEwaldE; N[k].HGEwald = m.HGEwaldE; N[k].GGEwald = m.GGEwaldE;
    GB(gb_tail_total(bx[k]->e, &N[k].Tail));
  }
  bool accept = false; Energy D[2];
  if(!overlap)
  {
    for(int k = 0; k < 2; k++) { D[k] = N[k]; D[k].add(bx[k]->createmol_energy, -1.0); D[k].add(bx[k]->running, -1.0); }
    const double pacc = std::exp(-A.d.beta * (D[0].total() + D[1].total()) + (molecules_in_box(A) + 1.0) * std::log(newVA / oldVA)
                                 + (molecules_in_box(B) + 1.0) * std::log(newVB / oldVB));
    if(S0.rng.uniform() < pacc) accept = true;
  }
  for(int k = 0; k < 2; k++)
  {
    GB(gb_volume_move_finish(bx[k]->e, accept ? 1 : 0));
    if(accept) { bx[k]->running.add(D[k]); adopt_box(*bx[k], nb[k]); }
  }
  if(accept) H.gibbs_vol_window.accepted++;
  if(std::fabs(H.boxes[0]->d.volume + H.boxes[1]->d.volume - H.gibbs_total_volume) > 0.1) die("Gibbs volume move: the total volume drifted");
  trace_move(S0, "gibbs_volume", comp, 0, accept ? 1 : 0, 0.0);
}

void update_max_gibbs_volume(Shared& H)               // Update_Max_GibbsVolume, mc_box.h:480-498
{
  if(H.gibbs_vol_window.total > 0)
  {
    const double ratio = (double) H.gibbs_vol_window.accepted / (double) H.gibbs_vol_window.total;
    double v = ratio / 0.5;
    if(v > 1.5) v = 1.5; else if(ratio < 0.5) v = 0.5;               // the lower clamp tests the raw ratio, as the reference does
    H.gibbs_max_change *= v;
    if(H.gibbs_max_change < 0.0005) H.gibbs_max_change = 0.0005;
    if(H.gibbs_max_change > 0.5) H.gibbs_max_change = 0.5;
  }
  H.gibbs_vol_total.total += H.gibbs_vol_window.total; H.gibbs_vol_total.accepted += H.gibbs_vol_window.accepted;
  H.gibbs_vol_window = MoveCount();
}

Growth deletion_body(Sim& S, int comp, long mol);

// GibbsParticleXferMove, move_struct.h:408-515: CBMC insertion in one box, CBMC deletion of a random molecule of the other
void move_gibbs_transfer(Sim& S0, int comp)
{
  Shared& H = S0.sh;
  if(H.boxes.size() != 2) { trace_move(S0, "gibbs_transfer", comp, 0, 0, 0.0); return; }
  H.gibbs_xfer.total++;
  int sel = 0, oth = 1;
  if(S0.rng.uniform() > 0.5) { sel = 1; oth = 0; }
  Sim& A = *H.boxes[sel]; Sim& B = *H.boxes[oth];
  long del_mol = 0;
  if(B.C[comp].nmol >= 1)
  {
    S0.rng.uniform();                                                              // InsertionSelectedMol: drawn (move_struct.h:438), not used by the growth
    del_mol = (long) (size_t) (S0.rng.uniform() * (double) B.C[comp].nmol);
  }
  else die("Gibbs particle transfer out of an empty box is not driven by this host program");
  Growth GI = insertion_body(A, comp);
  if(!GI.success) { trace_move(S0, "gibbs_transfer", comp, del_mol, 0, 0.0); return; }
  Growth GD = deletion_body(B, comp, del_mol);
  if(!GD.success) { trace_move(S0, "gibbs_transfer", comp, del_mol, 0, 0.0); return; }
  long nA = 0, nB = 0;
  for(int c = A.nhost; c < A.ncomp; c++) nA += A.C[c].nmol;
  for(int c = B.nhost; c < B.ncomp; c++) nB += B.C[c].nmol;
  const double pacc = (GI.W * (double) nB * A.d.volume) / (GD.W * (double) (nA + 1) * B.d.volume);
  if(S0.rng.uniform() < pacc)
  {
    GB(gb_accept_insertion(A.e, comp)); A.C[comp].nmol++; A.total_molecules++; A.running.add(GI.E);
    GB(gb_accept_deletion(B.e, comp, del_mol)); B.C[comp].nmol--; B.total_molecules--; B.running.add(GD.E, -1.0);
    H.gibbs_xfer.accepted++;
    trace_move(S0, "gibbs_transfer", comp, del_mol, 1, 0.0);
  }
  else trace_move(S0, "gibbs_transfer", comp, del_mol, 0, 0.0);
}

void run_move(Sim& S, long cycle)
{
  int comp = 0;
  while(S.C[comp].total_prob < 1e-10) comp = (int) (size_t) (S.rng.uniform() * S.ncomp);
  const long mol = (long) (size_t) (S.rng.uniform() * (double) S.C[comp].nmol);
  const double R = S.rng.uniform();
  const CompState& X = S.C[comp];
  S.moves_done++;
  const long lines_before = S.trace_lines;
  if(R < X.cTrans) { if(X.nmol > 0) move_single_body(S, comp, mol, GB_TRANSLATION); }
  else if(R < X.cRot) { if(X.nmol > 0) move_single_body(S, comp, mol, GB_ROTATION); }
  else if(R < X.cSpecial) { }
  else if(R < X.cWidom) move_widom(S, comp, cycle);
  else if(R < X.cReins) { if(X.nmol > 0) move_reinsertion(S, comp, mol); }
  else if(R < X.cIdentity) move_identity_swap(S);
  else if(R < X.cCBCF) { }
  else if(R < X.cSwap)
  {
    if(S.rng.uniform() < 0.5) move_insertion(S, comp);
    else if(X.nmol > 0) move_deletion(S, comp, mol);
    else S.C[comp].del.total += 0;
  }
  else if(R < X.cVolume) move_volume(S, comp);
  else if(R < X.cGibbsXfer) move_gibbs_transfer(S, comp);
  else if(R < X.cGibbsVolume) move_gibbs_volume(S, comp);
  if(S.trace_lines == lines_before) trace_move(S, "none", comp, mol, 0, 0.0);     // the selected move had nothing to act on
  static const bool sync_every_move = std::getenv("GB_SYNC_EVERY_MOVE") != nullptr;   // debugging aid: drain every engine's stream after each move
  if(sync_every_move) for(Sim* b : S.sh.boxes) GB(gb_synchronize(b->e));
}

void update_max(double* m, MoveCount& w, MoveCount& cum, double cap)           // Update_Max_Translation / Update_Max_Rotation
{
  if(w.total == 0) return;
  cum.total += w.total; cum.accepted += w.accepted;
  const double ratio = (double) w.accepted / (double) w.total;
  for(int k = 0; k < 3; k++) { m[k] *= (ratio > 0.5) ? 1.05 : 0.95; if(m[k] < 0.01) m[k] = 0.01; if(m[k] > cap) m[k] = cap; }
  w.total = 0; w.accepted = 0;
}

void run_phase(Sim& S, long cycles, bool production)
{
  S.production = production;
  if(production) S.block_size = std::max<long>(1, cycles / S.nblock);
  for(long i = 0; i < cycles; i++)
  {
    long steps = 20;
    if(steps < S.total_molecules) steps = S.total_molecules;
    if(S.d.use_max_step && steps > S.d.max_step_per_cycle) steps = S.d.max_step_per_cycle;
    for(long j = 0; j < steps; j++) run_move(S, i);
    if(production) for(int c = S.nhost; c < S.ncomp; c++) { S.C[c].load_sum += (double) S.C[c].nmol; S.C[c].load_n++; }
    if(i % 500 == 0)
    {
      for(int c = 1; c < S.ncomp; c++) { update_max(S.C[c].max_trans, S.C[c].trans_window, S.C[c].trans_cum, 5.0); update_max(S.C[c].max_rot, S.C[c].rot_window, S.C[c].rot_cum, 3.14); }
      update_max_volume_change(S, i);
    }
  }
}

// ------------------------------------------------------------------------------------------------ RNG-exact batched Widom
// Valid when Widom insertion of one component is the only move of the deck (Henrys_coefficient): pool offsets then
// advance in decades.  Returns false (nothing consumed) when the deck does not qualify.
bool widom_only(const Sim& S, int& comp)
{
  int found = -1;
  for(int c = 1; c < S.ncomp; c++)
  {
    const CompState& X = S.C[c];
    if(X.total_prob < 1e-10) continue;
    if(found >= 0) return false;
    if(std::fabs(X.cWidom - X.cSpecial - 1.0) > 1e-12) return false;
    found = c;
  }
  comp = found;
  if(found >= 0 && (found < S.nhost || S.d.comps[found - S.nhost].use_pockets)) return false;
  return found >= 0 && S.d.n_trial_positions == S.d.n_trial_orientations && S.d.comps[found - S.nhost].ms() > 1 && S.d.use_max_step && S.d.max_step_per_cycle == 1;
}

} // namespace

// The batched walk in its final form: pools are generated ahead on the host (they only depend on the random stream),
// so the walk never has to undo a draw.  The host keeps the previous pool alive until its queue is evaluated.
namespace {

struct Queue { std::vector<int64_t> fb, orr; std::vector<double> uni; std::vector<long> cyc; };

// the pool of the walk is the one pool_reset left on the device (gb_upload_random_pool): the Widom calls are told to use it
// (pool3 = NULL) instead of receiving it again, and the batch resumes from the first-bead energies the classification pass kept
// (resume_first_bead), so every first bead is evaluated once
// shares of the pool blocks [0, ndec): engine g classifies -- and later resumes from -- blocks [ndec g / G, ndec (g + 1) / G)
inline size_t share_begin(size_t ndec, size_t g, size_t G) { return ndec * g / G; }

// One lane of the replay: a set of engines (one per GPU) that holds ONE pool at a time -- upload, classify, (walk on the host), evaluate.
// Two lanes alternate pools, so that the evaluation of pool k (a background thread) runs beside the generation, upload and
// classification of pool k + 1 and the host's walk through it.
struct WidomLane
{
  std::vector<gb_engine*> eng;
  Queue Q; std::vector<int32_t> ok; std::vector<int64_t> idx;
  std::vector<double> out8; std::vector<int32_t> stage;
  std::thread th; bool pending = false;
};

void lane_upload_pool(Sim& S, WidomLane& L)
{
  std::vector<std::thread> up;
  for(size_t g = 1; g < L.eng.size(); g++) up.emplace_back([&S, &L, g] { GB(gb_upload_random_pool(L.eng[g], S.pool.data(), (int64_t) S.pool_size)); });
  GB(gb_upload_random_pool(L.eng[0], S.pool.data(), (int64_t) S.pool_size));
  for(auto& t : up) t.join();
}

// first-bead classification of every block of the pool the lane holds, one share per engine
void lane_classify(Sim& S, int comp, WidomLane& L)
{
  long dummy = 0; CallClock cc(S.widom_s[1], dummy);
  const size_t ntp = (size_t) S.d.n_trial_positions, ndec = S.pool_size / ntp, G = L.eng.size();
  L.ok.assign(ndec, 0); L.idx.resize(ndec);
  for(size_t k = 0; k < ndec; k++) L.idx[k] = (int64_t) (k * ntp);
  auto share = [&](size_t g)
  {
    const size_t a = share_begin(ndec, g, G), b = share_begin(ndec, g + 1, G);
    if(b > a) GB(gb_widom_first_bead_success(L.eng[g], comp, (int64_t) (b - a), nullptr, 0, L.idx.data() + a, L.ok.data() + a));
  };
  std::vector<std::thread> th;
  for(size_t g = 1; g < G; g++) th.emplace_back(share, g);
  share(0);
  for(auto& t : th) t.join();
}

// the queued insertions of the lane's pool: the Widom calls use the pool the lane's engines hold (pool3 = NULL) and resume from the
// first-bead energies the classification kept, so every first bead is evaluated once.  Touches nothing but the lane: may run in a thread.
void lane_evaluate(Sim& S, int comp, WidomLane& L)
{
  const size_t n = L.Q.fb.size();
  if(n == 0) return;
  L.out8.resize(n * 8); L.stage.resize(n);
  const size_t G = L.eng.size(), ntp = (size_t) S.d.n_trial_positions, ndec = S.pool_size / ntp;
  // the queue is in pool order: the insertions whose first-bead block lies in engine g's share are one contiguous piece of it
  std::vector<size_t> cut(G + 1, n);
  cut[0] = 0;
  for(size_t g = 1; g < G; g++)
    cut[g] = (size_t) (std::lower_bound(L.Q.fb.begin(), L.Q.fb.end(), (int64_t) (share_begin(ndec, g, G) * ntp)) - L.Q.fb.begin());
  auto piece = [&](size_t g)
  {
    const size_t a = cut[g], m = cut[g + 1] - cut[g];
    if(m == 0) return;
    double sums[12];
    gb_widom_inputs in; std::memset(&in, 0, sizeof(in));
    in.pool3 = nullptr; in.n_pool = 0; in.fb_index = L.Q.fb.data() + a; in.or_index = L.Q.orr.data() + a; in.uniforms = L.Q.uni.data() + 2 * a;
    in.inputs_on_device = 0; in.n_blocks = 1; in.resume_first_bead = 1;
    int rc = gb_widom_batch(L.eng[g], comp, (int64_t) m, &in, L.out8.data() + 8 * a, L.stage.data() + a, 0, sums);
    if(rc == GB_ERR_STATE) { in.resume_first_bead = 0; rc = gb_widom_batch(L.eng[g], comp, (int64_t) m, &in, L.out8.data() + 8 * a, L.stage.data() + a, 0, sums); }
    GB(rc);
  };
  std::vector<std::thread> th;
  for(size_t g = 1; g < G; g++) th.emplace_back(piece, g);
  piece(0);
  for(auto& t : th) t.join();
}

// the averages are taken on the host in cycle order, whatever the number of engines and lanes: the same sums to the last bit
void lane_record(Sim& S, int comp, WidomLane& L)
{
  long dummy = 0; CallClock cc(S.widom_s[3], dummy);
  const size_t n = L.Q.fb.size();
  for(size_t i = 0; i < n; i++)
  {
    Energy E; const double* o = &L.out8[8 * i];
    E.HGVDW = o[1]; E.HGReal = o[2]; E.GGVDW = o[3]; E.GGReal = o[4]; E.GGEwald = o[5]; E.HGEwald = o[6]; E.Tail = o[7];
    S.C[comp].widom.total++;
    if(L.stage[i] == 3)
    {
      std::fprintf(stderr, "graspa_b200_mc: a chain stage lost every orientation to the overlap criterion; the batched replay cannot know that in advance. Re-run with --sequential-widom.\n");
      std::exit(3);
    }
    record_rosen(S, comp, L.stage[i] == 0 ? o[0] : 0.0, L.stage[i] == 0 ? E : Energy(), L.Q.cyc[i]);
  }
  L.Q.fb.clear(); L.Q.orr.clear(); L.Q.uni.clear(); L.Q.cyc.clear();
}

void lane_finish(Sim& S, int comp, WidomLane& L)               // wait for a background evaluation and book its results
{
  if(!L.pending) return;
  { long dummy = 0; CallClock cc(S.widom_s[2], dummy); L.th.join(); }
  L.pending = false;
  lane_record(S, comp, L);
}

void run_widom_batched_v2(Sim& S, int comp, long cycles, std::vector<gb_engine*> second_lane)
{
  const size_t ntp = (size_t) S.d.n_trial_positions;
  S.block_size = std::max<long>(1, cycles / S.nblock);
  S.production = true;
  WidomLane lanes[2];
  lanes[0].eng.push_back(S.e); lanes[0].eng.insert(lanes[0].eng.end(), S.sh.replicas.begin(), S.sh.replicas.end());
  lanes[1].eng = second_lane;
  const bool two = !lanes[1].eng.empty();
  int cur = 0;
  lane_classify(S, comp, lanes[cur]);                           // the first pool is on every engine already
  auto next_pool = [&](int lane)
  {
    { long dummy = 0; CallClock cc(S.widom_s[0], dummy); pool_reset_host(S); lane_upload_pool(S, lanes[lane]); }
    lane_classify(S, comp, lanes[lane]);
  };
  for(long cycle = 0; cycle < cycles; cycle++)
  {
    // Select_Box_Component_Molecule
    int c = 0;
    while(S.C[c].total_prob < 1e-10) c = (int) (size_t) (S.rng.uniform() * S.ncomp);
    S.rng.uniform(); S.rng.uniform();
    S.moves_done++;
    // first bead: Random.Check(ntp)
    if(S.pool_off + ntp >= S.pool_size)
    {
      WidomLane& L = lanes[cur];
      if(two)
      {
        // pool k is evaluated in the background by this lane while the other lane takes pool k + 1; the results of pool k - 1 are booked first
        L.pending = true; L.th = std::thread([&S, comp, &L] { lane_evaluate(S, comp, L); });
        cur = 1 - cur;
        lane_finish(S, comp, lanes[cur]);
      }
      else { { long dummy = 0; CallClock cc(S.widom_s[2], dummy); lane_evaluate(S, comp, L); } lane_record(S, comp, L); }
      next_pool(cur);
    }
    WidomLane& L = lanes[cur];
    const size_t dfb = S.pool_off / ntp;
    S.pool_off += ntp;
    const double u1 = (L.ok[dfb] != 2) ? S.rng.uniform() : 0.5;   // SelectTrialPosition draws only when a trial survived (mc_widom.h:334-335)
    if(L.ok[dfb] != 1) { L.Q.fb.push_back((int64_t) (dfb * ntp)); L.Q.orr.push_back((int64_t) (dfb * ntp)); L.Q.uni.push_back(u1); L.Q.uni.push_back(0.5); L.Q.cyc.push_back(cycle); continue; }
    // chain: Random.Check(nto)
    if(S.pool_off + ntp >= S.pool_size)
    {
      // the orientation block lies in the NEXT pool: finish this insertion with the stage calls (once per pool), on this lane's first
      // engine, which then takes the next pool as well; everything queued so far is evaluated and booked first
      lane_finish(S, comp, lanes[1 - cur]);
      { long dummy = 0; CallClock cc(S.widom_s[2], dummy); lane_evaluate(S, comp, L); }
      lane_record(S, comp, L);
      gb_engine* e0 = L.eng[0];
      gb_cbmc_result r; int32_t used = 0; const double scale[2] = {1.0, 1.0};
      GB(gb_cbmc_first_bead(e0, GB_CBMC_INSERTION, comp, 0, (int64_t) (dfb * ntp), u1, scale, 0.0, -1, -1, nullptr, &r, &used));
      double W = r.rosenbluth; Energy E; E.HGVDW = r.energy[0]; E.HGReal = r.energy[1]; E.GGVDW = r.energy[2]; E.GGReal = r.energy[3];
      { long dummy = 0; CallClock cc(S.widom_s[0], dummy); pool_reset_host(S); lane_upload_pool(S, L); }
      GB(gb_cbmc_chain(e0, GB_CBMC_INSERTION, comp, 0, 0, S.rng.peek(0), -1, -1, &r, &used));
      S.pool_off += ntp; S.rng.advance(used);
      bool good = r.success != 0; W *= r.rosenbluth; if(W <= 1e-150) good = false;
      E.HGVDW += r.energy[0]; E.HGReal += r.energy[1]; E.GGVDW += r.energy[2]; E.GGReal += r.energy[3];
      if(good)
      {
        double ew[2] = {0, 0}, tail = 0.0;
        if(!S.d.no_charges && S.C[comp].has_charge) GB(gb_ewald_delta(e0, comp, GB_INSERTION, r.selected, scale, ew));
        GB(gb_tail_difference(e0, comp, GB_INSERTION, &tail));
        E.GGEwald = ew[0]; E.HGEwald = ew[1]; E.Tail = tail;
        W *= std::exp(-S.d.beta * (ew[0] + ew[1])); W *= std::exp(-S.d.beta * tail);
      }
      S.C[comp].widom.total++;
      record_rosen(S, comp, good ? W : 0.0, good ? E : Energy(), cycle);
      lane_classify(S, comp, L);                                  // after the stage calls: what it keeps stays valid for the batch of this pool
      continue;
    }
    const size_t dor = S.pool_off / ntp;
    S.pool_off += ntp;
    const double u2 = S.rng.uniform();
    L.Q.fb.push_back((int64_t) (dfb * ntp)); L.Q.orr.push_back((int64_t) (dor * ntp)); L.Q.uni.push_back(u1); L.Q.uni.push_back(u2); L.Q.cyc.push_back(cycle);
  }
  lane_finish(S, comp, lanes[1 - cur]);
  { long dummy = 0; CallClock cc(S.widom_s[2], dummy); lane_evaluate(S, comp, lanes[cur]); }
  lane_record(S, comp, lanes[cur]);
}

void print_widom(Sim& S, int comp)
{
  const CompState& X = S.C[comp];
  const double R = 8.314462618, NA = 6.02214076e23;
  const long ncell = (long) S.d.unitcells[0] * S.d.unitcells[1] * S.d.unitcells[2];
  const double rho = S.d.framework_mass * ncell * 1.0e-3 / (NA * S.d.volume * 1.0e-30);
  double tw = 0, tw2 = 0, th = 0, th2 = 0;
  std::printf("=====================Rosenbluth Summary For Component [%d] (%s)=====================\n", comp, S.d.comps[comp - S.nhost].name.c_str());
  for(int b = 0; b < S.nblock; b++)
  {
    std::printf("=====BLOCK %d=====\nWidom Performed: %.1f\n", b, X.rn[b]);
    if(X.rn[b] > 0)
    {
      const double w = X.rw[b] / X.rn[b], h = w / (R * S.d.temperature * rho);
      std::printf("(Total) Averaged Rosenbluth Weight: %.10f\n", w);
      std::printf("(Total) Averaged Excess Mu: %.10f\n", 1.2027242847 * -(1.0 / S.d.beta) * std::log(w));
      std::printf("(Total) Averaged Henry Coefficient: %.10f\n", h);
      const Energy& E = X.wE[b];
      std::printf("AVG WIDOM %d HGVDW: %.5f, HGReal: %.5f, GGVDW: %.5f, GGReal: %.5f, HGEwaldE: %.5f, GGEwaldE: %.5f, TailE: %.5f\n", b,
                  E.HGVDW / X.rn[b] / w, E.HGReal / X.rn[b] / w, E.GGVDW / X.rn[b] / w, E.GGReal / X.rn[b] / w, E.HGEwald / X.rn[b] / w, E.GGEwald / X.rn[b] / w, E.Tail / X.rn[b] / w);
      tw += w; tw2 += w * w; th += h; th2 += h * h;
    }
  }
  const double aw = tw / S.nblock, aw2 = tw2 / S.nblock, ah = th / S.nblock, ah2 = th2 / S.nblock;
  std::printf("=========================AVERAGE========================\n");
  std::printf("Averaged Rosenbluth Weight: %.10f +/- %.10f\n", aw, 2.0 * std::pow(std::fmax(aw2 - aw * aw, 0.0), 0.5));
  std::printf("Averaged Henry Coefficient [mol/kg/Pa]: %.10g +/- %.10g\n", ah, 2.0 * std::pow(std::fmax(ah2 - ah * ah, 0.0), 0.5));
}

// RASPA-2 restart file of the final configuration (the layout of write_data.h:109-263: cell, move maxima, components, then
// per component the position / velocity / force / charge / scaling / fixed blocks), from one device-to-host snapshot of the
// LIVE atoms of each adsorbate component.
void write_restart(Sim& S, const std::string& path)
{
  std::FILE* f = std::fopen(path.c_str(), "w");
  if(!f) { std::fprintf(stderr, "graspa_b200_mc: cannot write %s\n", path.c_str()); return; }
  const double* C = S.d.cell;
  std::fprintf(f, "Cell info:\n========================================================================\nnumber-of-unit-cells: 1 1 1\n");
  const char* nm[3] = {"a", "b", "c"};
  for(int k = 0; k < 3; k++) std::fprintf(f, "unit-cell-vector-%s: %.15g %.15g %.15g\n", nm[k], C[3 * k], C[3 * k + 1], C[3 * k + 2]);
  std::fprintf(f, "\n");
  for(int k = 0; k < 3; k++) std::fprintf(f, "cell-vector-%s: %.15g %.15g %.15g\n", nm[k], C[3 * k], C[3 * k + 1], C[3 * k + 2]);
  std::fprintf(f, "\ncell-lengths: %.15g %.15g %.15g \ncell-angles: 90 90 90 \n\n\n", C[0], C[4], C[8]);
  std::fprintf(f, "Maximum changes for MC-moves:\n========================================================================\n"
                  "Maximum-volume-change: 0.006250\nMaximum-Gibbs-volume-change: 0.025000\n"
                  "Maximum-box-shape-change: 0.100000 0.100000 0.100000, 0.100000 0.100000 0.100000, 0.100000 0.100000 0.100000\n\n\n");
  std::fprintf(f, "Acceptance targets for MC-moves:\n========================================================================\n"
                  "Target-volume-change: 0.500000\nTarget-box-shape-change: 0.500000\nTarget-Gibbs-volume-change: 0.500000\n\n\n");
  long nads = 0;
  for(int c = S.nhost; c < S.ncomp; c++) nads += S.C[c].nmol;
  std::fprintf(f, "Components: %d (Adsorbates %ld, Cations 0)\n========================================================================\n", S.ncomp - S.nhost, nads);
  for(int c = S.nhost; c < S.ncomp; c++)
  {
    const int k = c - S.nhost; const CompState& X = S.C[c];
    std::fprintf(f, "Components %d (%s) \n\n", k, comp_name(S, c));
    std::fprintf(f, "Maximum-translation-change component %d: %.6f %.6f %.6f\n", k, X.max_trans[0], X.max_trans[1], X.max_trans[2]);
    std::fprintf(f, "Maximum-translation-in-plane-change component %d: 0.000000,0.000000,0.000000\n", k);
    std::fprintf(f, "Maximum-rotation-change component %d: %.6f %.6f %.6f\n\n", k, X.max_rot[0], X.max_rot[1], X.max_rot[2]);
  }
  std::fprintf(f, "Reactions: 0\n");
  long prev = 0;
  for(int c = S.nhost; c < S.ncomp; c++)
  {
    const int ms = comp_ms(S, c); const long nmol = S.C[c].nmol; const size_t n = (size_t) nmol * ms;
    std::vector<double> pos(3 * std::max<size_t>(n, 1)), sc(std::max<size_t>(n, 1)), q(std::max<size_t>(n, 1));
    GB(gb_snapshot_molecules(S.e, c, 0, nmol, pos.data(), q.data(), sc.data(), nullptr));     // live molecules only
    std::fprintf(f, "\nComponent: %d   Adsorbate %ld molecules of %s\n------------------------------------------------------------------------\n",
                 c - S.nhost, nmol, comp_name(S, c));
    for(size_t j = 0; j < n; j++) std::fprintf(f, "Adsorbate-atom-position: %ld %zu %.15g  %.15g  %.15g\n", prev + (long) (j / ms), j % ms, pos[3 * j], pos[3 * j + 1], pos[3 * j + 2]);
    for(size_t j = 0; j < n; j++) std::fprintf(f, "Adsorbate-atom-velocity: %ld %zu 0  0  0\n", prev + (long) (j / ms), j % ms);
    for(size_t j = 0; j < n; j++) std::fprintf(f, "Adsorbate-atom-force: %ld %zu 0  0  0\n", prev + (long) (j / ms), j % ms);
    for(size_t j = 0; j < n; j++) std::fprintf(f, "Adsorbate-atom-charge: %ld %zu %.15g\n", prev + (long) (j / ms), j % ms, q[j]);
    for(size_t j = 0; j < n; j++) std::fprintf(f, "Adsorbate-atom-scaling: %ld %zu %.15g\n", prev + (long) (j / ms), j % ms, sc[j]);
    for(size_t j = 0; j < n; j++) std::fprintf(f, "Adsorbate-atom-fixed: %ld %zu 0 0 0\n", prev + (long) (j / ms), j % ms);
    prev += nmol;
  }
  std::fclose(f);
}

Energy total_energy(Sim& S)
{
  gb_move_energy v, w; double tail = 0.0;
  GB(gb_total_vdw_real(S.e, &v)); GB(gb_total_ewald(S.e, 0, &w)); GB(gb_tail_total(S.e, &tail));
  Energy E; E.HHVDW = v.HHVDW; E.HGVDW = v.HGVDW; E.GGVDW = v.GGVDW; E.HHReal = v.HHReal; E.HGReal = v.HGReal; E.GGReal = v.GGReal;
  // Ewald_Total books the framework's own term into the guest-guest sum as well (ewald_preparation.h:174); the stage
  // summary takes it out again (fxn_main.h:318) and reports host-host relative to the initial framework (:324-326)
  E.HHEwald = w.HHEwaldE - S.initial_framework_ewald; E.HGEwald = w.HGEwaldE; E.GGEwald = w.GGEwaldE - w.HHEwaldE; E.Tail = tail;
  return E;
}

void print_energy(const char* tag, const Energy& E)
{
  std::printf("%s VDW [Host-Host]: %.5f, VDW [Host-Guest]: %.5f, VDW [Guest-Guest]: %.5f, Real [Host-Host]: %.5f, Real [Host-Guest]: %.5f, Real [Guest-Guest]: %.5f, "
              "Ewald [Host-Host]: %.5f, Ewald [Host-Guest]: %.5f, Ewald [Guest-Guest]: %.5f, Tail: %.5f, Total: %.5f\n",
              tag, E.HHVDW, E.HGVDW, E.GGVDW, E.HHReal, E.HGReal, E.GGReal, E.HHEwald, E.HGEwald, E.GGEwald, E.Tail, E.total());
}

// Check_Simulation_Energy(INITIAL) -> CreateMolecule_InOneBox -> Check_Simulation_Energy(CREATEMOL) for one box (main.cpp:333-340);
// -> the energy the running deltas of this box start from
Energy initial_state(Sim& S)
{
  // initial energies + structure factors (Check_Simulation_Energy(INITIAL), fxn_main.h:282-404)
  { gb_move_energy w; GB(gb_total_ewald(S.e, 1, &w)); S.initial_framework_ewald = w.HHEwaldE; }
  Energy E0 = total_energy(S);
  print_energy("INITIAL", E0);
  long ncreate = 0; for(const auto& M : S.d.comps) ncreate += M.create_molecules;
  if(ncreate > 0)
  {
    create_molecules(S);
    // Check_Simulation_Energy(CREATEMOL) (main.cpp:340): the energies are computed afresh and the running deltas restart from
    // them; the stored structure factors stay the incrementally updated ones (Allocate_Copy_Ewald_Vector runs at INITIAL only)
    const Energy created = total_energy(S);
    Energy dC = created; dC.add(E0, -1.0);
    std::printf("CREATE MOLECULE: %ld molecules, running %.5f, recomputed %.5f\n", ncreate, S.running.total(), dC.total());
    E0 = created; S.running = Energy();
    print_energy("CREATED", E0);
  }
  S.createmol_energy = E0;
  return E0;
}

// Run_Simulation_MultipleBoxes, axpy.cu:593-625: per cycle and per box slot a box is DRAWN, and runs max(20, N) moves
void run_phase_boxes(Shared& H, long cycles, bool production)
{
  const size_t nb = H.boxes.size();
  for(Sim* b : H.boxes) { b->production = production; if(production) b->block_size = std::max<long>(1, cycles / b->nblock); }
  for(long i = 0; i < cycles; i++)
  {
    for(size_t k = 0; k < nb; k++)
    {
      Sim& S = *H.boxes[(size_t) (H.rng.uniform() * (double) nb)];
      long steps = 20;
      if(steps < S.total_molecules) steps = S.total_molecules;
      if(S.d.use_max_step && steps > S.d.max_step_per_cycle) steps = S.d.max_step_per_cycle;
      for(long j = 0; j < steps; j++) run_move(S, i);
    }
    for(Sim* b : H.boxes)
    {
      Sim& S = *b;
      if(production) for(int c = S.nhost; c < S.ncomp; c++) { S.C[c].load_sum += (double) S.C[c].nmol; S.C[c].load_n++; }
      if(i % 500 == 0)
      {
        for(int c = 1; c < S.ncomp; c++) { update_max(S.C[c].max_trans, S.C[c].trans_window, S.C[c].trans_cum, 5.0); update_max(S.C[c].max_rot, S.C[c].rot_window, S.C[c].rot_cum, 3.14); }
        update_max_volume_change(S, i);
      }
    }
    if(i > 0 && i % 500 == 0) update_max_gibbs_volume(H);
  }
}

} // namespace

int main(int argc, char** argv)
{
  // self-test modes that need no GPU (tests/test_host_driver.py)
  if(argc >= 4 && std::string(argv[1]) == "--rng-dump")
  {
    // first N uniforms of Get_Uniform_Random() after std::srand(seed): draws and look-ahead must agree with libc
    GlibcRand g((unsigned) std::atol(argv[2])); const long n = std::atol(argv[3]);
    for(long i = 0; i < n; i++)
    {
      const double ahead = (i % 7 == 0) ? g.peek(3) : -1.0;      // exercise the look-ahead buffer
      const double u = g.uniform();
      std::printf("%.17g\n", u);
      if(ahead >= 0.0) { const double a = g.peek(2); if(a != ahead) { std::fprintf(stderr, "peek mismatch at %ld\n", i); return 3; } }
    }
    return 0;
  }
  if(argc >= 3 && std::string(argv[1]) == "--dump-deck")
  {
    // the parsed deck as JSON (no GPU needed): what set_up_engine() uploads, for tests/golden/make_golden.py to build the
    // config C fixture (separated framework components, mixing-rule overrides, shifted potentials, block pockets)
    try
    {
      deck::Deck d = deck::load(argv[2]);
      auto arr = [](const char* key, const double* v, size_t n, bool last = false) {
        std::printf("\"%s\": [", key); for(size_t i = 0; i < n; i++) std::printf("%s%.17g", i ? ", " : "", v[i]); std::printf("]%s\n", last ? "" : ",");
      };
      auto iarr = [](const char* key, const int* v, size_t n, bool last = false) {
        std::printf("\"%s\": [", key); for(size_t i = 0; i < n; i++) std::printf("%s%d", i ? ", " : "", v[i]); std::printf("]%s\n", last ? "" : ",");
      };
      const size_t n2 = d.names.size() * d.names.size();
      std::printf("{\n");
      arr("cell", d.cell, 9); arr("inv", d.inv, 9); iarr("kmax", d.kmax, 3);
      std::printf("\"volume\": %.17g, \"alpha\": %.17g, \"prefactor\": %.17g, \"recip_cutoff\": %.17g, \"beta\": %.17g, \"temperature\": %.17g,\n",
                  d.volume, d.alpha, d.prefactor, d.recip_cutoff, d.beta, d.temperature);
      std::printf("\"cutoff_vdw\": %.17g, \"cutoff_coul\": %.17g, \"overlap\": %.17g, \"no_charges\": %d, \"ntrials\": %d, \"norient\": %d, \"adsorbate_allocate\": %ld,\n",
                  d.cutoff_vdw, d.cutoff_coul, d.overlap, d.no_charges ? 1 : 0, d.n_trial_positions, d.n_trial_orientations, d.adsorbate_allocate);
      std::printf("\"names\": ["); for(size_t i = 0; i < d.names.size(); i++) std::printf("%s\"%s\"", i ? ", " : "", d.names[i].c_str()); std::printf("],\n");
      arr("eps", d.eps.data(), n2); arr("sigma", d.sigma.data(), n2); arr("shift", d.shift.data(), n2);
      iarr("use_tail", d.use_tail.data(), n2); arr("tail_energy", d.tail_energy.data(), n2);
      std::printf("\"framework\": [\n");
      std::printf("{\"molsize\": %zu,\n", d.ftype.size()); arr("pos", d.fpos.data(), d.fpos.size()); arr("charge", d.fcharge.data(), d.fcharge.size());
      iarr("type", d.ftype.data(), d.ftype.size(), true); std::printf("}");
      for(const auto& F : d.fw)
      {
        std::printf(",\n{\"molsize\": %d,\n", F.molsize); arr("pos", F.pos.data(), F.pos.size()); arr("charge", F.charge.data(), F.charge.size());
        iarr("molid", F.molid.data(), F.molid.size()); iarr("type", F.type.data(), F.type.size(), true); std::printf("}");
      }
      std::printf("],\n\"adsorbates\": [\n");
      for(size_t c = 0; c < d.comps.size(); c++)
      {
        const deck::Component& M = d.comps[c];
        std::printf("%s{\"name\": \"%s\", \"invert_pockets\": %d,\n", c ? ",\n" : "", M.name.c_str(), M.invert_pockets ? 1 : 0);
        arr("pos", M.pos.data(), M.pos.size()); arr("charge", M.charge.data(), M.charge.size());
        arr("pocket_centers", M.pocket_centers.data(), M.pocket_centers.size()); arr("pocket_radii", M.pocket_radii.data(), M.pocket_radii.size());
        iarr("type", M.type.data(), M.type.size(), true); std::printf("}");
      }
      std::printf("]\n}\n");
    }
    catch(const std::exception& ex) { std::fprintf(stderr, "graspa_b200_mc: %s\n", ex.what()); return 1; }
    return 0;
  }
  if(argc >= 3 && std::string(argv[1]) == "--parse-only")
  {
    try
    {
      deck::Deck d = deck::load(argv[2]);
      std::printf("{\"framework_atoms\": %zu, \"components\": %zu, \"alpha\": %.9f, \"kmax\": [%d, %d, %d], \"volume\": %.6f, \"beta\": %.10f, "
                  "\"init_cycles\": %ld, \"equil_cycles\": %ld, \"prod_cycles\": %ld, \"seed\": %ld, \"fugacity_coeff\": [",
                  d.ftype.size(), d.comps.size(), d.alpha, d.kmax[0], d.kmax[1], d.kmax[2], d.volume, d.beta,
                  (long) d.init_cycles, (long) d.equil_cycles, (long) d.prod_cycles, (long) d.random_seed);
      for(size_t c = 0; c < d.comps.size(); c++) std::printf("%s%.10f", c ? ", " : "", d.comps[c].fugacity_coeff);
      std::printf("], \"framework_components\": [%zu", d.ftype.size());
      for(const auto& F : d.fw) std::printf(", %zu", F.type.size());
      std::printf("], \"block_pockets\": [");
      for(size_t c = 0; c < d.comps.size(); c++) std::printf("%s%zu", c ? ", " : "", d.comps[c].pocket_radii.size());
      std::printf("]");
      if(d.n_simulations > 1 && !d.single_simulation)
      {
        // boxes run together (Gibbs ensemble): the per-box columns of the deck
        std::printf(", \"boxes\": [");
        for(int b = 0; b < d.n_simulations; b++)
        {
          const deck::Deck db = b == 0 ? d : deck::load(argv[2], -1.0, -1.0, b);
          std::printf("%s{\"volume\": %.6f, \"kmax\": [%d, %d, %d], \"create\": [", b ? ", " : "", db.volume, db.kmax[0], db.kmax[1], db.kmax[2]);
          for(size_t c = 0; c < db.comps.size(); c++) std::printf("%s%d", c ? ", " : "", db.comps[c].create_molecules);
          std::printf("]}");
        }
        std::printf("], \"gibbs_volume_prob\": %.6f, \"gibbs_xfer_prob\": [", d.gibbs_volume_prob);
        for(size_t c = 0; c < d.comps.size(); c++) std::printf("%s%.6f", c ? ", " : "", d.comps[c].p_gibbs_xfer);
        std::printf("]");
      }
      if(d.volume_move_prob > 0.0) std::printf(", \"npt_volume_prob\": %.6f", d.volume_move_prob);
      if(d.restart_file)
      {
        // molecules taken from the restart / LAMMPS data file, and where atom 1 of the first one sits
        std::printf(", \"restart\": {\"format\": \"%s\", \"molecules\": [", d.restart_lammps ? "LAMMPS" : "RASPA");
        for(size_t c = 0; c < d.comps.size(); c++) std::printf("%s%zu", c ? ", " : "", d.comps[c].restart_charge.size() / (size_t) std::max(1, d.comps[c].ms()));
        std::printf("]");
        if(!d.comps.empty() && d.comps[0].restart_pos.size() >= 6)
          std::printf(", \"first_molecule\": [[%.9f, %.9f, %.9f], [%.9f, %.9f, %.9f]]", d.comps[0].restart_pos[0], d.comps[0].restart_pos[1], d.comps[0].restart_pos[2],
                      d.comps[0].restart_pos[3], d.comps[0].restart_pos[4], d.comps[0].restart_pos[5]);
        std::printf("}");
      }
      if(d.use1264)
      {
        // pairs with an r^-4 term: [name i, name j, C12, C6, C4, shift] in internal units
        std::printf(", \"lj1264\": [");
        const int n = (int) d.names.size(); bool first = true;
        for(int i = 0; i < n; i++)
          for(int j = i; j < n; j++)
            if(d.z[i * n + j] != 0.0)
            {
              std::printf("%s[\"%s\", \"%s\", %.12e, %.12e, %.12e, %.12e]", first ? "" : ", ", d.names[i].c_str(), d.names[j].c_str(),
                          d.eps[i * n + j], d.sigma[i * n + j], d.z[i * n + j], d.shift[i * n + j]);
              first = false;
            }
        std::printf("]");
      }
      std::printf("}\n");
    }
    catch(const std::exception& ex) { std::fprintf(stderr, "graspa_b200_mc: %s\n", ex.what()); return 1; }
    return 0;
  }
  if(argc < 2) { std::fprintf(stderr, "usage: graspa_b200_mc <deck directory> [--sequential-widom] [--staged] [--no-server] [--gpus N] [--no-pipeline] [--timing] [--trace file] [--init N] [--equil N] [--prod N]\n"); return 2; }
  const std::string dir = argv[1];
  bool sequential_widom = false, staged = false, timing = false, no_server = false, no_pipeline = false; const char* trace_path = nullptr; const char* restart_out = nullptr;
  long o_init = -1, o_equil = -1, o_prod = -1; double o_pressure = -1.0, o_temperature = -1.0; int o_device = -1, o_gpus = 1; long o_seed = -1;
  for(int i = 2; i < argc; i++)
  {
    const std::string a = argv[i];
    if(a == "--sequential-widom") sequential_widom = true;
    else if(a == "--staged") staged = true;
    else if(a == "--no-pipeline") no_pipeline = true;        // batched Widom replay: one lane (no overlap of a pool's evaluation with the next pool)
    else if(a == "--no-server") no_server = true;            // one k_move launch per move instead of the resident move server
    else if(a == "--timing") timing = true;
    else if(a == "--trace" && i + 1 < argc) trace_path = argv[++i];
    else if(a == "--init" && i + 1 < argc) o_init = std::atol(argv[++i]);
    else if(a == "--equil" && i + 1 < argc) o_equil = std::atol(argv[++i]);
    else if(a == "--prod" && i + 1 < argc) o_prod = std::atol(argv[++i]);
    else if(a == "--pressure" && i + 1 < argc) o_pressure = std::atof(argv[++i]);        // Pa: one isotherm point per process / GPU
    else if(a == "--temperature" && i + 1 < argc) o_temperature = std::atof(argv[++i]);
    else if(a == "--device" && i + 1 < argc) o_device = std::atoi(argv[++i]);
    else if(a == "--gpus" && i + 1 < argc) o_gpus = std::atoi(argv[++i]);       // batched Widom replay: one share of every pool per GPU
    else if(a == "--seed" && i + 1 < argc) o_seed = std::atol(argv[++i]);
    else if(a == "--write-restart" && i + 1 < argc) restart_out = argv[++i];
  }
  Shared SH; Sim S(SH);
  try { S.d = deck::load(dir, o_pressure, o_temperature); } catch(const std::exception& ex) { std::fprintf(stderr, "graspa_b200_mc: %s\n", ex.what()); return 1; }
  if(o_seed >= 0) S.d.random_seed = (int) o_seed;
  S.device = o_device;
  if(o_init >= 0) S.d.init_cycles = o_init;
  if(o_equil >= 0) S.d.equil_cycles = o_equil;
  if(o_prod >= 0) S.d.prod_cycles = o_prod;
  S.C.assign(1 + S.d.fw.size() + S.d.comps.size(), CompState());
  if(trace_path) S.trace = std::fopen(trace_path, "w");
  S.fused = !staged;
  setup_engine(S);
  setup_probabilities(S);
  // --gpus N: replicas of the system on the next N - 1 devices (same deck, same uploads); only the batched Widom replay uses them.
  // A long replay (more than a few pools) gets a second engine per GPU as well: two lanes that alternate pools, so that the
  // evaluation of one pool overlaps the generation, upload, classification and walk of the next (--no-pipeline: one lane)
  std::vector<std::unique_ptr<Shared>> replica_shared;
  std::vector<gb_engine*> second_lane;
  {
    int wc = -1;
    const bool replay = !sequential_widom && widom_only(S, wc) && S.d.init_cycles == 0 && S.d.equil_cycles == 0;
    const bool pipeline = replay && !no_pipeline && S.d.prod_cycles >= 100000;
    auto replica = [&](int device) -> gb_engine*
    {
      replica_shared.emplace_back(new Shared());
      Sim R(*replica_shared.back());
      R.d = S.d; R.device = device; R.fused = S.fused;
      R.C.assign(1 + R.d.fw.size() + R.d.comps.size(), CompState());
      setup_engine(R);
      return R.e;
    };
    for(int g = 1; g < o_gpus; g++) SH.replicas.push_back(replica(std::max(o_device, 0) + g));
    if(pipeline) for(int g = 0; g < o_gpus; g++) second_lane.push_back(replica(std::max(o_device, 0) + g));
  }
  // two boxes run together (NumberOfSimulations 2, SingleSimulation no): the Gibbs ensemble of the reference's examples
  std::unique_ptr<Sim> S2;
  if(S.d.n_simulations == 2 && !S.d.single_simulation)
  {
    S2.reset(new Sim(SH)); S2->box_index = 1;
    try { S2->d = deck::load(dir, o_pressure, o_temperature, 1); } catch(const std::exception& ex) { std::fprintf(stderr, "graspa_b200_mc: box 1: %s\n", ex.what()); return 1; }
    S2->d.random_seed = S.d.random_seed; S2->d.init_cycles = S.d.init_cycles; S2->d.equil_cycles = S.d.equil_cycles; S2->d.prod_cycles = S.d.prod_cycles;
    S2->device = o_device; S2->fused = !staged;
    S2->C.assign(1 + S2->d.fw.size() + S2->d.comps.size(), CompState());
    setup_engine(*S2);
    setup_probabilities(*S2);
    std::printf("graspa_b200_mc: box 1: %zu framework atoms, kmax %d %d %d, volume %.5f\n", S2->d.ftype.size(), S2->d.kmax[0], S2->d.kmax[1], S2->d.kmax[2], S2->d.volume);
  }
  else if(S.d.n_simulations > 1 && !S.d.single_simulation) { std::fprintf(stderr, "graspa_b200_mc: %d boxes run together are not driven by this host program\n", S.d.n_simulations); return 2; }
  std::printf("graspa_b200_mc: %s, %zu framework atoms, %zu adsorbate component(s), alpha %.6f, kmax %d %d %d, volume %.5f, beta %.8f\n",
              dir.c_str(), S.d.ftype.size(), S.d.comps.size(), S.d.alpha, S.d.kmax[0], S.d.kmax[1], S.d.kmax[2], S.d.volume, S.d.beta);
  // Random.Setup(333334): std::srand(RANDOMSEED) then the first pool (main.cpp:145, data_struct.h:1338-1346)
  S.rng.reseed((unsigned) S.d.random_seed);
  S.pool.assign(3 * S.pool_size, 0.0);
  pool_reset(S); S.pool_rounds = 0;
  // RandomNumber::DeviceRandom runs the debug kernel Aaccess_device_random<<<1,1>>> after the first fill, which overwrites
  // element 0 of the device pool with {2.3, 4.5, 6.7} and copies the pool back to the host (data_struct.h:1280-1285, 1322-1328):
  // the very first trial position of a run comes from those three numbers.
  S.pool[0] = 2.3; S.pool[1] = 4.5; S.pool[2] = 6.7;
  for(Sim* b : SH.boxes) GB(gb_upload_random_pool(b->e, S.pool.data(), (int64_t) S.pool_size));
  for(gb_engine* r : SH.replicas) GB(gb_upload_random_pool(r, S.pool.data(), (int64_t) S.pool_size));
  for(gb_engine* r : second_lane) GB(gb_upload_random_pool(r, S.pool.data(), (int64_t) S.pool_size));
  Energy E0 = initial_state(S);
  for(gb_engine* r : SH.replicas) { gb_move_energy w; GB(gb_total_ewald(r, 1, &w)); }          // the stored structure factors of the Fourier stage
  for(gb_engine* r : second_lane) { gb_move_energy w; GB(gb_total_ewald(r, 1, &w)); }
  Energy E0b;
  if(S2) { std::printf("--- box 1\n"); E0b = initial_state(*S2); SH.gibbs_total_volume = S.d.volume + S2->d.volume; }

  if(timing) GB(gb_timing_enable(S.e, 1));
  if(no_server) { GB(gb_move_server(S.e, 0, nullptr, nullptr)); if(S2) GB(gb_move_server(S2->e, 0, nullptr, nullptr)); }
  const auto t0 = std::chrono::steady_clock::now();
  int wcomp = -1;
  const bool batched = !sequential_widom && widom_only(S, wcomp) && S.d.init_cycles == 0 && S.d.equil_cycles == 0;
  if(S2)
  {
    run_phase_boxes(SH, S.d.init_cycles, false);
    run_phase_boxes(SH, S.d.equil_cycles, false);
    run_phase_boxes(SH, S.d.prod_cycles, true);
    GB(gb_synchronize(S2->e));
  }
  else if(batched) run_widom_batched_v2(S, wcomp, S.d.prod_cycles, second_lane);
  else
  {
    run_phase(S, S.d.init_cycles, false);
    run_phase(S, S.d.equil_cycles, false);
    run_phase(S, S.d.prod_cycles, true);
  }
  GB(gb_synchronize(S.e));
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  const Energy E1 = total_energy(S);
  print_energy("FINAL  ", E1);
  if(restart_out) write_restart(S, restart_out);
  Energy D = E1; D.add(E0, -1.0);
  print_energy("RUNNING", S.running);
  std::printf("ENERGY DRIFT (FINAL - INITIAL - RUNNING) Total Energy: %.6e\n", D.total() - S.running.total());
  for(int c = 1; c < S.ncomp; c++)
  {
    const CompState& X = S.C[c];
    std::printf("Component %d (%s): molecules %ld | translation %ld/%ld rotation %ld/%ld insertion %ld/%ld deletion %ld/%ld reinsertion %ld/%ld widom %ld\n",
                c, comp_name(S, c), X.nmol, X.trans.accepted, X.trans.total, X.rot.accepted, X.rot.total, X.ins.accepted, X.ins.total,
                X.del.accepted, X.del.total, X.reins.accepted, X.reins.total, X.widom.total);
    // the reference prints max(cumulative up to the last 500-cycle update, current window)
    std::printf("Translation Performed: %ld\nTranslation Accepted: %ld\nRotation Performed: %ld\nRotation Accepted: %ld\n",
                std::max(X.trans_cum.total, X.trans_window.total), std::max(X.trans_cum.accepted, X.trans_window.accepted),
                std::max(X.rot_cum.total, X.rot_window.total), std::max(X.rot_cum.accepted, X.rot_window.accepted));
    if(X.idswap_add.total + X.idswap_remove.total > 0)
      for(int t = S.nhost; t < S.ncomp && !X.idswap_to.empty(); t++)
        std::printf("Identity Swap Performed, FROM [%s (%d)] TO [%s (%d)]: %ld (%ld Accepted)\n", comp_name(S, c), c, comp_name(S, t), t,
                    X.idswap_to[t].total, X.idswap_to[t].accepted);
    if(X.widom.total > 0) print_widom(S, c);
  }
  if(S.d.volume_move_prob > 0.0)
    std::printf("Volume Move: %ld/%ld accepted, final volume %.5f, cell %.5f %.5f %.5f, kmax %d %d %d, MaxVolumeChange %.5f\n",
                S.vol_total.accepted + S.vol_window.accepted, S.vol_total.total + S.vol_window.total, S.d.volume, S.d.cell[0], S.d.cell[4], S.d.cell[8],
                S.d.kmax[0], S.d.kmax[1], S.d.kmax[2], S.vol_max_change);
  if(S2)
  {
    Sim& B = *S2;
    const Energy F1 = total_energy(B);
    std::printf("--- box 1\n");
    print_energy("FINAL  ", F1);
    Energy Db = F1; Db.add(E0b, -1.0);
    print_energy("RUNNING", B.running);
    std::printf("ENERGY DRIFT (FINAL - INITIAL - RUNNING) Total Energy: %.6e\n", Db.total() - B.running.total());
    for(int c = 1; c < B.ncomp; c++)
    {
      const CompState& X = B.C[c];
      std::printf("Component %d (%s): molecules %ld | translation %ld/%ld rotation %ld/%ld insertion %ld/%ld deletion %ld/%ld reinsertion %ld/%ld widom %ld\n",
                  c, comp_name(B, c), X.nmol, X.trans.accepted, X.trans.total, X.rot.accepted, X.rot.total, X.ins.accepted, X.ins.total,
                  X.del.accepted, X.del.total, X.reins.accepted, X.reins.total, X.widom.total);
    }
    if(B.d.volume_move_prob > 0.0)
      std::printf("Volume Move: %ld/%ld accepted, final volume %.5f, MaxVolumeChange %.5f\n", B.vol_total.accepted + B.vol_window.accepted,
                  B.vol_total.total + B.vol_window.total, B.d.volume, B.vol_max_change);
    std::printf("Gibbs Volume Move: %ld/%ld accepted, MaxGibbsBoxChange %.5f; Gibbs Particle Transfer: %ld/%ld accepted\n",
                SH.gibbs_vol_total.accepted + SH.gibbs_vol_window.accepted, SH.gibbs_vol_total.total + SH.gibbs_vol_window.total, SH.gibbs_max_change,
                SH.gibbs_xfer.accepted, SH.gibbs_xfer.total);
    std::printf("Box volumes: %.5f %.5f; molecules:", S.d.volume, B.d.volume);
    for(int c = S.nhost; c < S.ncomp; c++) std::printf(" %ld/%ld", S.C[c].nmol, B.C[c].nmol);
    std::printf("\n");
    std::printf("{\"box\": 1, \"loading\": [");
    for(int c = B.nhost; c < B.ncomp; c++)
      std::printf("%s{\"component\": \"%s\", \"molecules\": %ld, \"production_average\": %.6f}", c > B.nhost ? ", " : "", comp_name(B, c), B.C[c].nmol,
                  B.C[c].load_n ? B.C[c].load_sum / (double) B.C[c].load_n : (double) B.C[c].nmol);
    std::printf("]}\n");
  }
  const long cycles = S.d.init_cycles + S.d.equil_cycles + S.d.prod_cycles;
  int64_t launches = 0; gb_launch_count(S.e, &launches, 0);
  if(S2) { int64_t l2 = 0; gb_launch_count(S2->e, &l2, 0); launches += l2; }
  std::printf("Work took %.6f seconds\n", secs);
  std::printf("{\"pressure_pa\": %.6g, \"temperature\": %.6g, \"loading\": [", S.d.pressure_pa, S.d.temperature);
  for(int c = S.nhost; c < S.ncomp; c++)
    std::printf("%s{\"component\": \"%s\", \"molecules\": %ld, \"production_average\": %.6f}", c > S.nhost ? ", " : "", comp_name(S, c), S.C[c].nmol,
                S.C[c].load_n ? S.C[c].load_sum / (double) S.C[c].load_n : (double) S.C[c].nmol);
  std::printf("]}\n");
  int64_t srv_starts = 0, srv_moves = 0; gb_move_server(S.e, -1, &srv_starts, &srv_moves);
  std::printf("{\"moves\": %ld, \"cycles\": %ld, \"seconds\": %.6f, \"moves_per_s\": %.3f, \"cycles_per_s\": %.3f, \"widom_path\": \"%s\", \"move_calls\": \"%s\", \"rng_draws\": %llu, \"pool_refills\": %ld, \"kernel_launches\": %lld, \"server_starts\": %lld, \"server_moves\": %lld, \"gpus\": %d}\n",
              S.moves_done, cycles, secs, S.moves_done / secs, cycles / secs, batched ? "batched-exact" : "sequential", S.fused ? "fused" : "staged",
              (unsigned long long) S.rng.consumed(), S.pool_rounds, (long long) launches, (long long) srv_starts, (long long) srv_moves, 1 + (int) SH.replicas.size());
  if(batched)
    std::printf("batched Widom replay, host seconds: pool refills %.3f, first-bead classification %.3f, batch calls %.3f, averages %.3f, walk and the rest %.3f\n",
                S.widom_s[0], S.widom_s[1], S.widom_s[2], S.widom_s[3], secs - S.widom_s[0] - S.widom_s[1] - S.widom_s[2] - S.widom_s[3]);
  if(S.fused)
  {
    const char* nm[5] = {"insertion", "deletion", "reinsertion", "translation/rotation", "identity swap"};
    double tot = 0.0;
    for(int k = 0; k < 5; k++) tot += S.call_s[k];
    std::printf("host time inside the move calls: %.3f s of %.3f s;", tot, secs);
    for(int k = 0; k < 5; k++) if(S.call_n[k]) std::printf(" %s %.2f us x %ld;", nm[k], 1e6 * S.call_s[k] / S.call_n[k], S.call_n[k]);
    std::printf("\n");
  }
  if(timing)
  {
    double ms = 0.0; int64_t n = 0; gb_timing_read(S.e, 0, &ms, &n, 0);
    std::printf("device time of the move kernels (CUDA events, serialises the run): %.3f ms over %lld launches = %.2f us each\n", ms, (long long) n, n ? 1e3 * ms / n : 0.0);
  }
  if(S.trace) std::fclose(S.trace);
  if(S2) gb_engine_destroy(S2->e);
  for(gb_engine* r : SH.replicas) gb_engine_destroy(r);
  for(gb_engine* r : second_lane) gb_engine_destroy(r);
  gb_engine_destroy(S.e);
  return 0;
}
