// graspa_b200 host layer -- reader of a gRASPA / RASPA-2 style input deck (the reference's input surface, kept as is):
// simulation.input, force_field_mixing_rules.def, force_field.def, pseudo_atoms.def, <molecule>.def, <framework>.cif.
// Each routine names the reference parser it mirrors (read_data.cpp); numbers that reach the engine are computed with
// the reference's expressions so that energies agree to round-off.
//
// Supported subset (what SURVEY section 8's configs A, B, D, E need): one rigid framework component from a P1 CIF,
// rigid adsorbates, Lennard-Jones + Lorentz-Berthelot mixing, shifted/truncated, tail corrections incl. the
// "rules to overwrite" of force_field.def, Ewald with the RASPA-2 heuristic or the LAMMPS-style explicit set-up.
// Not read (the engine has no use for them yet): separated framework components, block pockets, CBCF/TMMC keywords.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace deck {

inline std::vector<std::string> terms(const std::string& line)      // Split_Tab_Space, read_data.cpp:85-94 (empty tokens dropped)
{
  std::vector<std::string> t; std::string cur;
  for(char c : line)
  {
    if(c == ' ' || c == '\t' || c == '\r' || c == ',') { if(!cur.empty()) { t.push_back(cur); cur.clear(); } }
    else cur.push_back(c);
  }
  if(!cur.empty()) t.push_back(cur);
  return t;
}
inline bool ieq(const std::string& a, const std::string& b)
{
  if(a.size() != b.size()) return false;
  for(size_t i = 0; i < a.size(); i++) if(std::tolower((unsigned char) a[i]) != std::tolower((unsigned char) b[i])) return false;
  return true;
}
inline std::vector<std::string> read_lines(const std::string& path)
{
  std::ifstream f(path);
  if(!f) throw std::runtime_error("cannot open " + path);
  std::vector<std::string> L; std::string s;
  while(std::getline(f, s)) L.push_back(s);
  return L;
}

struct Component
{
  std::string name;
  double ideal_rosenbluth = 1.0, fugacity_coeff = 1.0, mol_fraction = 1.0;
  double p_translation = 0, p_rotation = 0, p_widom = 0, p_reinsertion = 0, p_identity = 0, p_swap = 0, p_gibbs_xfer = 0;   // raw inputs
  int create_molecules = 0;
  bool use_pr_eos = false;
  // molecules read from RestartInitial/System_0/restartfile (RestartFile yes): positions 3 x n, charges n; molecule-major
  std::vector<double> restart_pos, restart_charge;
  // block pockets (BlockPockets yes / BlockPocketsFilename X / InvertBlockPockets, read_data.cpp:2532-2575)
  bool use_pockets = false, invert_pockets = false; std::string pocket_file;
  std::vector<double> pocket_centers, pocket_radii;           // replicated, Cartesian (ReplicateBlockPockets :3362-3454)
  // molecule definition
  std::vector<double> pos;      // 3*ms
  std::vector<int> type; std::vector<double> charge;
  double tc = 0, pc = 0, acentric = 0, mass = 0;
  int ms() const { return (int) type.size(); }
};

struct Deck
{
  // simulation.input
  long init_cycles = 0, equil_cycles = 0, prod_cycles = 0;
  bool use_max_step = false; long max_step_per_cycle = 1;
  int random_seed = 0;
  int n_trial_positions = 8, n_trial_orientations = 8;            // WidomStruct defaults, data_struct.h:1275-1276
  long adsorbate_allocate = 10240;
  std::string framework_name; int unitcells[3] = {1, 1, 1};
  bool use_cif_charges = false, no_charges = true, restart_file = false;
  double temperature = 300.0, pressure_pa = 0.0;
  double overlap = 1e5, cutoff_vdw = 12.0, cutoff_coul = 12.0, ewald_precision = 1e-6;
  bool lammps_ewald = false; double lammps_alpha = 0.0; int lammps_kmax[3] = {0, 0, 0};
  long movies_every = 5000, print_every = 5000;
  std::vector<Component> comps;                                     // adsorbates, in input order
  // SeparateFrameworkComponents: framework components 1.. (Framework_Component_<i>.def + the Framework_Component_ blocks
  // of simulation.input, read_data.cpp:530-607, 1860-1930); component 0 keeps every atom not listed there
  struct FrameworkComponent
  {
    std::string name; int nmol = 0, molsize = 0; std::vector<int> atom_index;      // unit-cell atom indices, molecule-major
    double p_translation = 0, p_rotation = 0, p_special = 0, p_reinsertion = 0;
    std::vector<double> pos; std::vector<int> type; std::vector<double> charge; std::vector<int> molid;   // supercell atoms
  };
  bool separate_framework = false; int n_framework_components = 1;
  std::vector<FrameworkComponent> fw;                               // fw[i - 1] = framework component i
  // force field
  std::vector<std::string> names; std::vector<double> eps_in, sig_in, mass, pseudo_charge;
  bool shifted = false, tail = false;
  std::vector<double> eps, sigma, shift, tail_energy; std::vector<int> use_tail;   // n*n
  // UseLJ1264 yes: eps / sigma hold the polynomial coefficients C12 / C6, z the r^-4 coefficient, c10 the r^-10 one
  // (ForceField_Processing read_data.cpp:1196-1230, VDW maths.cuh:452-476)
  bool use1264 = false; std::vector<double> c4_in, z, c10;
  double volume_move_prob = 0.0;           // NPTVolumeChangeProbability (Components::VolumeMoveProbability)
  double gibbs_volume_prob = 0.0;          // GibbsVolumeChangeProbability (Gibbs::GibbsBoxProb)
  int n_simulations = 1; bool single_simulation = true;
  bool restart_lammps = false, lmp_read_box = false; int lmp_start_component = 0;      // RestartInputFileType LAMMPS
  double ewald_tol1 = 0.0;                 // Boxsize::tol1: kmax follows the box in a volume move (mc_box.h:84-94)
  // framework
  double cell[9] = {0}, inv[9] = {0}, volume = 0;
  std::vector<double> fpos; std::vector<int> ftype; std::vector<double> fcharge;  // supercell atoms
  double framework_mass = 0.0;
  // Ewald
  double alpha = 0.0, prefactor = 138935.483496, recip_cutoff = 0.0; int kmax[3] = {0, 0, 0};
  // derived
  double beta = 0.0, pressure = 0.0;
  int ntypes() const { return (int) names.size(); }
};

inline int type_of(const Deck& d, const std::string& name)
{
  for(size_t j = 0; j < d.names.size(); j++) if(d.names[j] == name) return (int) j;
  throw std::runtime_error("Atom type not found in pseudo atoms definitions: " + name);
}

// VDW(), maths.cuh:477-492, for the shift value (Get_Shifted_Value)
inline double lj_energy(double eps, double sig, double rr)
{
  const double arg1 = 4.0 * eps, arg2 = sig * sig;
  const double temp = rr / arg2, temp3 = temp * temp * temp, rri3 = 1.0 / (temp3 + 0.0);
  return arg1 * (rri3 * (rri3 - 1.0));
}
// GetTailCorrectionValue, read_data.cpp:833-846
inline double tail_value(double eps, double sig, double cutsq)
{
  const double arg2 = sig * sig * sig, rr = std::sqrt(cutsq);
  const double term1 = std::pow(arg2, 4) / (9.0 * std::pow(rr, 9)), term2 = std::pow(arg2, 2) / (3.0 * std::pow(rr, 3));
  return 16.0 * 3.14159265358979323846 / 2.0 * eps * (term1 - term2);
}

inline void read_simulation_input(Deck& d, const std::string& dir, int box = 0)
{
  const size_t bx = (size_t) box;      // per-box columns: termsScannedLined[1 + BoxIndex] (read_data.cpp:2530-2570)
  auto L = read_lines(dir + "/simulation.input");
  Component* cur = nullptr;
  int fw_block = 0;
  for(auto& line : L)
  {
    auto t = terms(line);
    if(t.empty() || t[0][0] == '#') continue;
    auto has = [&](const char* k) { return line.find(k) != std::string::npos; };      // the reference matches by substring
    // ReadFrameworkComponentMoves, read_data.cpp:530-607
    if(has("END_OF_Framework_Component_")) { fw_block = 0; continue; }
    if(t[0] == "Framework_Component_" && t.size() >= 2)
    {
      fw_block = std::stoi(t[1]);
      if(fw_block >= 1 && (int) d.fw.size() < fw_block) d.fw.resize(fw_block);
      continue;
    }
    if(fw_block >= 1)
    {
      Deck::FrameworkComponent& F = d.fw[fw_block - 1];
      if(has("TranslationProbability")) F.p_translation = std::stod(t[1]);
      if(has("RotationProbability")) F.p_rotation = std::stod(t[1]);
      if(has("RotationSpecialProbability")) F.p_special = std::stod(t[1]);
      if(has("ReinsertionProbability")) F.p_reinsertion = std::stod(t[1]);
      continue;
    }
    if(t[0] == "Component" && t.size() >= 4) { d.comps.emplace_back(); cur = &d.comps.back(); cur->name = t[3]; continue; }
    if(cur)
    {
      if(has("IdealGasRosenbluthWeight")) cur->ideal_rosenbluth = std::stod(t[1]);
      else if(has("TranslationProbability")) cur->p_translation = std::stod(t[1]);
      else if(has("RotationProbability")) cur->p_rotation = std::stod(t[1]);
      else if(has("WidomProbability")) cur->p_widom = std::stod(t[1]);
      else if(has("ReinsertionProbability")) cur->p_reinsertion = std::stod(t[1]);
      else if(has("IdentityChangeProbability")) cur->p_identity = std::stod(t[1]);
      else if(has("GibbsParticleXferProbability")) cur->p_gibbs_xfer = std::stod(t[1]);            // read_data.cpp:2386-2390
      else if(has("SwapProbability")) cur->p_swap = std::stod(t[1]);
      else if(has("FugacityCoefficient")) { if(ieq(t[1], "PR-EOS")) { cur->use_pr_eos = true; cur->fugacity_coeff = -1.0; } else cur->fugacity_coeff = std::stod(t[1]); }
      else if(has("MolFraction")) cur->mol_fraction = std::stod(t[1]);
      else if(has("CreateNumberOfMolecules")) cur->create_molecules = std::stoi(t.at(1 + bx));
      else if(has("BlockPocketsFilename")) cur->pocket_file = t[1] + ".block";
      else if(has("InvertBlockPockets")) cur->invert_pockets = ieq(t[1], "yes");
      else if(has("BlockPockets")) { if(ieq(t[1], "yes")) cur->use_pockets = true; }
      continue;
    }
    if(has("NumberOfInitializationCycles")) d.init_cycles = std::stol(t[1]);
    else if(has("NumberOfEquilibrationCycles")) d.equil_cycles = std::stol(t[1]);
    else if(has("NumberOfProductionCycles")) d.prod_cycles = std::stol(t[1]);
    else if(has("SeparateFrameworkComponents")) d.separate_framework = ieq(t[1], "yes");
    else if(has("NumberofFrameworkComponents")) d.n_framework_components = std::stoi(t[1]);
    else if(has("RestartInputFileType")) d.restart_lammps = ieq(t[1], "LAMMPS");                     // read_data.cpp:2750-2767
    else if(has("LMPData_Comp_to_Start_with")) d.lmp_start_component = std::stoi(t[1]);
    else if(has("Read_Boxsize")) d.lmp_read_box = ieq(t[1], "yes");
    else if(has("RestartFile")) d.restart_file = ieq(t[1], "yes");
    else if(has("UseLJ1264")) d.use1264 = ieq(t[1], "yes");
    else if(has("NPTVolumeChangeProbability")) { const double v = std::stod(t[1]); if(v > 0) d.volume_move_prob = v; }      // read_data.cpp:404-413
    else if(has("UseMaxStep")) d.use_max_step = ieq(t[1], "yes");
    else if(has("MaxStepPerCycle")) d.max_step_per_cycle = std::stol(t[1]);
    else if(has("RandomSeed")) d.random_seed = std::stoi(t[1]);
    else if(has("NumberOfTrialPositions")) d.n_trial_positions = std::stoi(t[1]);
    else if(has("NumberOfTrialOrientations")) d.n_trial_orientations = std::stoi(t[1]);
    else if(has("AdsorbateAllocateSpace")) d.adsorbate_allocate = std::stol(t[1]);
    else if(has("FrameworkName")) d.framework_name = t.at(1 + bx);
    else if(has("UnitCells") && t.size() >= 5) { if(std::stoi(t[1]) == box) { d.unitcells[0] = std::stoi(t[2]); d.unitcells[1] = std::stoi(t[3]); d.unitcells[2] = std::stoi(t[4]); } }
    else if(has("NumberOfSimulations")) d.n_simulations = std::stoi(t[1]);
    else if(has("SingleSimulation")) d.single_simulation = ieq(t[1], "yes");
    else if(has("GibbsVolumeChangeProbability")) { const double v = std::stod(t[1]); if(v > 0) d.gibbs_volume_prob = v; }     // read_data.cpp:389-403
    else if(has("UseChargesFromCIFFile")) d.use_cif_charges = ieq(t[1], "yes");
    else if(has("ChargeMethod")) d.no_charges = !ieq(t[1], "Ewald");
    else if(has("Temperature")) d.temperature = std::stod(t[1]);
    else if(has("Pressure")) d.pressure_pa = std::stod(t[1]);
    else if(has("OverlapCriteria")) d.overlap = std::stod(t[1]);
    else if(has("CutOffVDW")) d.cutoff_vdw = std::stod(t[1]);
    else if(has("CutOffCoulomb")) d.cutoff_coul = std::stod(t[1]);
    else if(has("EwaldPrecision")) d.ewald_precision = std::stod(t[1]);
    else if(has("Ewald_UseLAMMPS_Setup")) d.lammps_ewald = ieq(t[1], "yes");
    else if(has("Ewald_Alpha")) d.lammps_alpha = std::stod(t[1]);
    else if(has("Ewald_kvectors") && t.size() >= 4) { for(int k = 0; k < 3; k++) d.lammps_kmax[k] = std::stoi(t[1 + k]); }
  }
}

// ForceFieldParser :772-834, PseudoAtomParser, ForceField_Processing :1179-1247, OverWriteTailCorrection :1132-1176
inline void read_force_field(Deck& d, const std::string& dir)
{
  auto L = read_lines(dir + "/force_field_mixing_rules.def");
  d.shifted = terms(L.at(1)).at(0) == "shifted";
  d.tail = terms(L.at(3)).at(0) == "yes";
  const int n = std::stoi(terms(L.at(5)).at(0));
  for(int i = 0; i < n; i++)
  {
    auto t = terms(L.at(7 + i));
    d.names.push_back(t.at(0)); d.eps_in.push_back(std::stod(t.at(2))); d.sig_in.push_back(std::stod(t.at(3)));
    d.c4_in.push_back((t.at(1) == "lennard-jones-1264" && t.size() >= 5) ? std::stod(t[4]) : 0.0);       // read_data.cpp:819-823
  }
  auto P = read_lines(dir + "/pseudo_atoms.def");
  const int np = std::stoi(terms(P.at(1)).at(0));
  if(np != n) throw std::runtime_error("pseudo_atoms.def and force_field_mixing_rules.def list different numbers of atoms");
  d.mass.resize(n); d.pseudo_charge.resize(n);
  for(int i = 0; i < n; i++)
  {
    auto t = terms(P.at(3 + i));
    if(t.at(0) != d.names[i]) throw std::runtime_error("pseudo_atoms.def must list the force-field names in the same order (read_data.cpp:1344)");
    d.mass[i] = std::stod(t.at(5)); d.pseudo_charge[i] = std::stod(t.at(6));
  }
  const double cutsq = d.cutoff_vdw * d.cutoff_vdw;
  d.eps.assign(n * n, 0); d.sigma.assign(n * n, 0); d.shift.assign(n * n, 0); d.tail_energy.assign(n * n, 0); d.use_tail.assign(n * n, 0);
  d.z.assign(n * n, 0); d.c10.assign(n * n, 0);
  // shift of the polynomial form, Get_Shifted_Value_Coeff read_data.cpp:750-758
  auto poly_shift = [&](double C12, double C6, double C4, double C10) {
    const double ri2 = 1.0 / cutsq, ri4 = ri2 * ri2, ri6 = ri4 * ri2, ri10 = ri4 * ri6, ri12 = ri6 * ri6;
    return C12 * ri12 - C6 * ri6 + C10 * ri10 + C4 * ri4;
  };
  if(d.use1264 && d.tail) throw std::runtime_error("tail corrections with UseLJ1264 are not read by this host program");
  for(int i = 0; i < n; i++)
    for(int j = 0; j < n; j++)
    {
      const double e = std::sqrt(d.eps_in[i] * d.eps_in[j]) / 1.20272430057, s = 0.5 * (d.sig_in[i] + d.sig_in[j]);
      if(d.use1264)
      {
        const double s2 = s * s, s6 = s2 * s2 * s2, s12 = s6 * s6;
        const double C12 = 4.0 * e * s12, C6 = 4.0 * e * s6, C4 = 0.5 * (d.c4_in[i] + d.c4_in[j]) / 1.20272430057;
        d.eps[i * n + j] = C12; d.sigma[i * n + j] = C6; d.z[i * n + j] = C4;
        d.shift[i * n + j] = d.shifted ? poly_shift(C12, C6, C4, 0.0) : 0.0;
        continue;
      }
      d.eps[i * n + j] = e; d.sigma[i * n + j] = s;
      d.shift[i * n + j] = d.shifted ? lj_energy(e, s, cutsq) : 0.0;
      if(d.tail) { d.use_tail[i * n + j] = 1; d.tail_energy[i * n + j] = tail_value(e, s, cutsq); }
    }
  std::ifstream ffdef(dir + "/force_field.def");
  if(ffdef)
  {
    auto F = read_lines(dir + "/force_field.def");
    // OverWrite_Mixing_Rule, read_data.cpp:924-1130: "<I> <J> lennard-jones <eps> <sig>" lines under the FIRST
    // "mixing rules to overwrite" marker replace the mixed pair parameters (and the shift)
    {
      size_t start = 0, nmix = 0;
      for(size_t k = 0; k < F.size(); k++)
      {
        if(F[k].find("mixing rules to overwrite") != std::string::npos && start == 0) start = k;
        if(start > 0 && k == start + 1) { auto t = terms(F[k]); if(!t.empty()) nmix = (size_t) std::stol(t[0]); }
      }
      for(size_t k = start + 3; nmix > 0 && k < start + 3 + nmix && k < F.size(); k++)
      {
        auto t = terms(F[k]);
        const bool is1264 = t.size() == 6 && t[2] == "lennard-jones-1264";
        if(t.size() != 5 && !is1264) continue;
        const int i = type_of(d, t[0]), j = type_of(d, t[1]);
        const double e = std::stod(t[3]) / 1.20272430057, sg = std::stod(t[4]);
        if(d.use1264)
        {
          const double s2 = sg * sg, s6 = s2 * s2 * s2, s12 = s6 * s6, C12 = 4.0 * e * s12, C6 = 4.0 * e * s6;
          const double C4 = is1264 ? std::stod(t[5]) / 1.20272430057 : 0.0;
          d.eps[i * n + j] = d.eps[j * n + i] = C12; d.sigma[i * n + j] = d.sigma[j * n + i] = C6;
          d.z[i * n + j] = d.z[j * n + i] = C4; d.c10[i * n + j] = d.c10[j * n + i] = 0.0;
          if(d.shifted) d.shift[i * n + j] = d.shift[j * n + i] = poly_shift(C12, C6, C4, 0.0);
          continue;
        }
        d.eps[i * n + j] = d.eps[j * n + i] = e; d.sigma[i * n + j] = d.sigma[j * n + i] = sg;
        if(d.shifted) d.shift[i * n + j] = d.shift[j * n + i] = lj_energy(e, sg, cutsq);
      }
      // "defined interactions" of the GENERIC2_HC form (UseLJ1264 only, read_data.cpp:1000-1049):
      // U = p0 exp(-p1 r) - p2/r^4 - p3/r^6 - p4/r^8 - p5/r^12  ->  C12 = -p5, C6 = p3, C4 = -p2, C10 = 0
      if(d.use1264)
      {
        size_t dstart = 0, ndef = 0, done = 0;
        for(size_t k = 0; k < F.size(); k++)
        {
          if(F[k].find("number of defined interactions") != std::string::npos && dstart == 0) dstart = k;
          if(dstart > 0 && k == dstart + 1) { auto t = terms(F[k]); if(!t.empty()) ndef = (size_t) std::stol(t[0]); }
        }
        for(size_t k = dstart + 2; ndef > 0 && done < ndef && k < F.size(); k++)
        {
          auto t = terms(F[k]);
          if(t.size() < 9 || t[2] != "GENERIC2_HC") continue;
          done++;
          int i, j;
          try { i = type_of(d, t[0]); j = type_of(d, t[1]); } catch(const std::exception&) { continue; }
          const double C12 = -std::stod(t[8]) / 1.20272430057, C6 = std::stod(t[6]) / 1.20272430057, C4 = -std::stod(t[5]) / 1.20272430057;
          d.eps[i * n + j] = d.eps[j * n + i] = C12; d.sigma[i * n + j] = d.sigma[j * n + i] = C6;
          d.z[i * n + j] = d.z[j * n + i] = C4; d.c10[i * n + j] = d.c10[j * n + i] = 0.0;
          if(d.shifted) d.shift[i * n + j] = d.shift[j * n + i] = poly_shift(C12, C6, C4, 0.0);
        }
      }
    }
    if(F.size() > 1)
    {
      const int nover = std::stoi(terms(F.at(1)).at(0));
      for(int k = 0; k < nover && 3 + k < (int) F.size(); k++)
      {
        auto t = terms(F[3 + k]);
        if(t.size() == 4 && t[3] == "yes")
        {
          const int i = type_of(d, t[0]), j = type_of(d, t[1]);
          d.use_tail[i * n + j] = d.use_tail[j * n + i] = 1;
          d.tail_energy[i * n + j] = d.tail_energy[j * n + i] = tail_value(d.eps[i * n + j], d.sigma[i * n + j], cutsq);
        }
      }
    }
  }
}

// MoleculeDefinitionParser, read_data.cpp:2044-2160 (rigid molecules)
inline void read_molecule(Deck& d, Component& c, const std::string& dir)
{
  auto L = read_lines(dir + "/" + c.name + ".def");
  c.tc = std::stod(terms(L.at(1)).at(0)); c.pc = std::stod(terms(L.at(2)).at(0)); c.acentric = std::stod(terms(L.at(3)).at(0));
  const int ms = std::stoi(terms(L.at(5)).at(0));
  if(!ieq(terms(L.at(9)).at(0), "rigid")) throw std::runtime_error("Currently Not allowing flexible molecule");
  for(int a = 0; a < ms; a++)
  {
    auto t = terms(L.at(13 + a));
    const int ty = type_of(d, t.at(1));
    c.type.push_back(ty); c.charge.push_back(d.pseudo_charge[ty]); c.mass += d.mass[ty];
    if(t.size() == 5) { c.pos.push_back(std::stod(t[2])); c.pos.push_back(std::stod(t[3])); c.pos.push_back(std::stod(t[4])); }
    else if(t.size() == 2 && ms == 1) { c.pos.push_back(0.0); c.pos.push_back(0.0); c.pos.push_back(0.0); }
    else throw std::runtime_error("Flexible molecules not implemented");
  }
}

// inverse_matrix / matrix_determinant, maths.cuh:28-56
inline void invert_cell(const double* x, double* r, double& det)
{
  const double m11 = x[0], m21 = x[3], m31 = x[6], m12 = x[1], m22 = x[4], m32 = x[7], m13 = x[2], m23 = x[5], m33 = x[8];
  det = +m11 * (m22 * m33 - m23 * m32) - m12 * (m21 * m33 - m23 * m31) + m13 * (m21 * m32 - m22 * m31);
  r[0] = +(m22 * m33 - m32 * m23) / det; r[3] = -(m21 * m33 - m31 * m23) / det; r[6] = +(m21 * m32 - m31 * m22) / det;
  r[1] = -(m12 * m33 - m32 * m13) / det; r[4] = +(m11 * m33 - m31 * m13) / det; r[7] = -(m11 * m32 - m31 * m12) / det;
  r[2] = +(m12 * m23 - m22 * m13) / det; r[5] = -(m11 * m23 - m21 * m13) / det; r[8] = +(m11 * m22 - m21 * m12) / det;
}

// ReadFramework (CIF, P1), read_data.cpp:1480-1760
inline void read_framework(Deck& d, const std::string& dir)
{
  auto L = read_lines(dir + "/" + d.framework_name + ".cif");
  double a = 0, b = 0, c = 0, al = 0, be = 0, ga = 0;
  for(auto& s : L)
  {
    auto t = terms(s);
    if(t.size() < 2) continue;
    if(s.find("_cell_length_a") != std::string::npos) a = std::stod(t[1]);
    if(s.find("_cell_length_b") != std::string::npos) b = std::stod(t[1]);
    if(s.find("_cell_length_c") != std::string::npos) c = std::stod(t[1]);
    if(s.find("_cell_angle_alpha") != std::string::npos) al = std::stod(t[1]) / (180.0 / 3.14159265358979323846);
    if(s.find("_cell_angle_beta") != std::string::npos) be = std::stod(t[1]) / (180.0 / 3.14159265358979323846);
    if(s.find("_cell_angle_gamma") != std::string::npos) ga = std::stod(t[1]) / (180.0 / 3.14159265358979323846);
  }
  const int nx = d.unitcells[0], ny = d.unitcells[1], nz = d.unitcells[2];
  const double dy = b * std::sin(ga);
  const double tempd = (std::cos(al) - std::cos(ga) * std::cos(be)) / std::sin(ga);
  const double dz = c * std::sqrt(1 - std::pow(std::cos(be), 2) - std::pow(tempd, 2));
  const double bx = b * std::cos(ga), cx = c * std::cos(be), cy = c * tempd;
  d.cell[0] = nx * a;  d.cell[1] = 0.0;     d.cell[2] = 0.0;
  d.cell[3] = ny * bx; d.cell[4] = ny * dy; d.cell[5] = 0.0;
  d.cell[6] = nz * cx; d.cell[7] = nz * cy; d.cell[8] = nz * dz;
  invert_cell(d.cell, d.inv, d.volume);
  int col[5] = {-1, -1, -1, -1, -1}, count = 0, last = -1;
  for(size_t i = 0; i < L.size(); i++)
  {
    if(L[i].find("_atom_site") != std::string::npos)
    {
      const char* keys[5] = {"_atom_site_label", "_atom_site_fract_x", "_atom_site_fract_y", "_atom_site_fract_z", "_atom_site_charge"};
      for(int k = 0; k < 5; k++) if(L[i].find(keys[k]) != std::string::npos) col[k] = count;
      count++; last = (int) i;
    }
    else if(last >= 0) break;
  }
  if(col[0] < 0 || col[1] < 0 || col[2] < 0 || col[3] < 0) throw std::runtime_error("Couldn't find required columns in the CIF file! Abort.");
  std::vector<double> uf; std::vector<int> ut; std::vector<double> uq;
  for(size_t i = last + 1; i < L.size(); i++)
  {
    auto t = terms(L[i]);
    if(t.size() < 4) break;
    std::string label = t[col[0]];
    while(!label.empty() && std::isdigit((unsigned char) label.back())) label.pop_back();     // remove_number_at_the_end
    const int ty = type_of(d, label);
    uf.push_back(std::stod(t[col[1]])); uf.push_back(std::stod(t[col[2]])); uf.push_back(std::stod(t[col[3]]));
    ut.push_back(ty);
    uq.push_back((d.use_cif_charges && col[4] >= 0) ? std::stod(t[col[4]]) : d.pseudo_charge[ty]);
    d.framework_mass += d.mass[ty];
  }
  // DetermineFrameworkComponent + CheckFrameworkComponentAtomOrder (read_data.cpp:1398-1478): unit-cell atoms listed in
  // Framework_Component_<i>.def leave component 0 and are kept in the order of that file (molecule-major)
  std::vector<int> owner(ut.size(), 0);
  for(size_t f = 0; f < d.fw.size(); f++)
  {
    // an index that no CIF atom carries never matches in DetermineFrameworkComponent (the NaX example lists one): drop it
    std::vector<int> kept;
    for(int idx : d.fw[f].atom_index) if(idx >= 0 && idx < (int) ut.size()) { owner[idx] = (int) f + 1; kept.push_back(idx); }
    d.fw[f].atom_index = kept;
    if((int) kept.size() != d.fw[f].nmol * d.fw[f].molsize)
      throw std::runtime_error("In CheckFrameworkCIF function, NMol and value in FrameworkComponentDef don't match!!!!");
  }
  const double sx = (double) 1 / nx, sy = (double) 1 / ny, sz = (double) 1 / nz;
  auto place = [&](size_t A, int ix, int jy, int kz, std::vector<double>& pos) {
    const double fx = (uf[3 * A] + ix) * sx, fy = (uf[3 * A + 1] + jy) * sy, fz = (uf[3 * A + 2] + kz) * sz;
    pos.push_back(fx * d.cell[0] + fy * d.cell[3] + fz * d.cell[6]);
    pos.push_back(fx * d.cell[1] + fy * d.cell[4] + fz * d.cell[7]);
    pos.push_back(fx * d.cell[2] + fy * d.cell[5] + fz * d.cell[8]);
  };
  for(int ix = 0; ix < nx; ix++) for(int jy = 0; jy < ny; jy++) for(int kz = 0; kz < nz; kz++)
    for(size_t A = 0; A < ut.size(); A++)
    {
      if(owner[A] != 0) continue;
      place(A, ix, jy, kz, d.fpos);
      d.ftype.push_back(ut[A]); d.fcharge.push_back(uq[A]);
    }
  for(size_t f = 0; f < d.fw.size(); f++)
  {
    Deck::FrameworkComponent& F = d.fw[f];
    if(F.molsize > 1 && nx * ny * nz > 1) throw std::runtime_error("separated framework molecules with several atoms need UnitCells 1 1 1 (atoms of a molecule must stay contiguous)");
    for(int ix = 0; ix < nx; ix++) for(int jy = 0; jy < ny; jy++) for(int kz = 0; kz < nz; kz++)
    {
      const int cell_id = (ix * ny + jy) * nz + kz;
      for(size_t k = 0; k < F.atom_index.size(); k++)
      {
        const size_t A = (size_t) F.atom_index[k];
        place(A, ix, jy, kz, F.pos);
        F.type.push_back(ut[A]); F.charge.push_back(uq[A]);
        F.molid.push_back(F.nmol * cell_id + (int) (k / (size_t) F.molsize));       // :1728-1733
      }
    }
  }
}

// ReadFrameworkSpeciesDefinitions, read_data.cpp:1860-1930: Framework_Component_<i>.def
inline void read_framework_components(Deck& d, const std::string& dir)
{
  if(!d.separate_framework || d.n_framework_components <= 1) { d.fw.clear(); return; }
  d.fw.resize(d.n_framework_components - 1);
  for(int i = 1; i < d.n_framework_components; i++)
  {
    Deck::FrameworkComponent& F = d.fw[i - 1];
    auto L = read_lines(dir + "/Framework_Component_" + std::to_string(i) + ".def");
    for(auto& s : L)
    {
      auto t = terms(s);
      if(t.size() < 2) continue;
      if(s.find("Framework_Component_Name") != std::string::npos) F.name = t[1];
      else if(s.find("Number_of_Molecules_for_Framework_component") != std::string::npos) F.nmol = std::stoi(t[1]);
      else if(s.find("Number_of_atoms_for_each_molecule") != std::string::npos) F.molsize = std::stoi(t[1]);
      else if(s.find("Atom_Indices_for_Molecule") != std::string::npos)
        for(size_t k = 2; k < t.size(); k++) F.atom_index.push_back(std::stoi(t[k]));
    }
    if(F.nmol <= 0 || F.molsize <= 0) throw std::runtime_error("Framework_Component_" + std::to_string(i) + ".def: missing molecule count or size");
  }
}

// ReadBlockPockets + ReplicateBlockPockets, read_data.cpp:3320-3454
inline void read_block_pockets(Deck& d, const std::string& dir)
{
  for(auto& c : d.comps)
  {
    if(!c.use_pockets || c.pocket_file.empty()) { c.use_pockets = c.use_pockets && !c.pocket_file.empty(); continue; }
    std::ifstream f(dir + "/" + c.pocket_file);
    if(!f) throw std::runtime_error("Cannot open block pocket file: " + c.pocket_file);
    size_t n = 0; f >> n;
    std::vector<double> cen(3 * n), rad(n);
    double maxc = 0.0;
    for(size_t i = 0; i < n; i++) { f >> cen[3 * i] >> cen[3 * i + 1] >> cen[3 * i + 2] >> rad[i]; for(int k = 0; k < 3; k++) maxc = std::max(maxc, std::fabs(cen[3 * i + k])); }
    const int nx = d.unitcells[0], ny = d.unitcells[1], nz = d.unitcells[2];
    if(maxc > 1.5)       // Cartesian input: to fractional coordinates of ONE unit cell by the cell's diagonal (:3399-3411)
    {
      const double cx = d.cell[0] / nx, cy = d.cell[4] / ny, cz = d.cell[8] / nz;
      for(size_t i = 0; i < n; i++) { cen[3 * i] /= cx; cen[3 * i + 1] /= cy; cen[3 * i + 2] /= cz; }
    }
    for(size_t i = 0; i < n; i++)
      for(int j = 0; j < nx; j++) for(int k = 0; k < ny; k++) for(int l = 0; l < nz; l++)
      {
        const double vx = (cen[3 * i] + j) / nx, vy = (cen[3 * i + 1] + k) / ny, vz = (cen[3 * i + 2] + l) / nz;
        c.pocket_centers.push_back(d.cell[0] * vx + d.cell[3] * vy + d.cell[6] * vz);
        c.pocket_centers.push_back(d.cell[1] * vx + d.cell[4] * vy + d.cell[7] * vz);
        c.pocket_centers.push_back(d.cell[2] * vx + d.cell[5] * vy + d.cell[8] * vz);
        c.pocket_radii.push_back(rad[i]);
      }
  }
}

// read_Ewald_Parameters_from_input, read_data.cpp:609-702
inline void setup_ewald(Deck& d)
{
  const double PI = 3.14159265358979323846;
  if(d.lammps_ewald)
  {
    d.alpha = d.lammps_alpha; for(int k = 0; k < 3; k++) d.kmax[k] = d.lammps_kmax[k];
    const double ux = 2 * PI / d.cell[0], vy = 2 * PI / d.cell[4], wz = 2 * PI / d.cell[8];
    const double kx = d.kmax[0] * ux, ky = d.kmax[1] * vy, kz = d.kmax[2] * wz;
    d.recip_cutoff = std::fmax(kx * kx, std::fmax(ky * ky, kz * kz)) * 1.00001;
    return;
  }
  const double rc = d.cutoff_coul, p = d.ewald_precision;
  const double tol = std::sqrt(std::fabs(std::log(p * rc)));
  const double alpha = std::sqrt(std::fabs(std::log(p * rc * tol))) / rc;
  const double tol1 = std::sqrt(-std::log(p * rc * std::pow(2.0 * tol * alpha, 2)));
  d.alpha = alpha; d.ewald_tol1 = tol1;
  d.kmax[0] = (int) std::round(0.25 + d.cell[0] * alpha * tol1 / PI);
  d.kmax[1] = (int) std::round(0.25 + d.cell[4] * alpha * tol1 / PI);
  d.kmax[2] = (int) std::round(0.25 + d.cell[8] * alpha * tol1 / PI);
  const int m = std::max(d.kmax[0], std::max(d.kmax[1], d.kmax[2]));
  d.recip_cutoff = std::pow(1.05 * (double) m, 2);
}

// PBC(), maths.cuh:427-450
inline void min_image(const Deck& d, double* v)
{
  const double* I = d.inv; const double* C = d.cell;
  const bool cubic = !((std::fabs(C[3]) + std::fabs(C[6]) + std::fabs(C[7])) > 1e-10);
  if(cubic)
  {
    v[0] -= static_cast<int>(v[0] * I[0] + ((v[0] >= 0.0) ? 0.5 : -0.5)) * C[0];
    v[1] -= static_cast<int>(v[1] * I[4] + ((v[1] >= 0.0) ? 0.5 : -0.5)) * C[4];
    v[2] -= static_cast<int>(v[2] * I[8] + ((v[2] >= 0.0) ? 0.5 : -0.5)) * C[8];
    return;
  }
  double sx = I[0] * v[0] + I[3] * v[1] + I[6] * v[2], sy = I[1] * v[0] + I[4] * v[1] + I[7] * v[2], sz = I[2] * v[0] + I[5] * v[1] + I[8] * v[2];
  sx -= static_cast<int>(sx + ((sx >= 0.0) ? 0.5 : -0.5)); sy -= static_cast<int>(sy + ((sy >= 0.0) ? 0.5 : -0.5)); sz -= static_cast<int>(sz + ((sz >= 0.0) ? 0.5 : -0.5));
  v[0] = C[0] * sx + C[3] * sy + C[6] * sz; v[1] = C[1] * sx + C[4] * sy + C[7] * sz; v[2] = C[2] * sx + C[5] * sy + C[8] * sz;
}

// RestartFileParser, read_data.cpp:3000-3221 (RASPA-2 restart file): per adsorbate component the block starts two lines
// after "Component: <i>"; `interval` position lines, then velocity, force, charge and scaling blocks of the same length.
// Atoms other than the first of a molecule are re-wrapped to the nearest image of the first (:3147-3160).
inline void read_restart(Deck& d, const std::string& dir)
{
  if(!d.restart_file || d.restart_lammps) return;
  auto L = read_lines(dir + "/RestartInitial/System_0/restartfile");
  for(size_t ci = 0; ci < d.comps.size(); ci++)
  {
    Component& c = d.comps[ci];
    const std::string key = "Component: " + std::to_string(ci);
    size_t start = 0; long nmol = 0; bool found = false;
    for(size_t k = 0; k < L.size(); k++)
      if(L[k].find(key) == 0) { nmol = std::stol(terms(L[k]).at(3)); start = k + 2; found = true; break; }
    if(!found || nmol == 0) continue;
    const size_t ms = (size_t) c.ms(), interval = (size_t) nmol * ms;
    if(start + 5 * interval > L.size()) throw std::runtime_error("restart file shorter than its molecule count says");
    c.restart_pos.resize(3 * interval); c.restart_charge.resize(interval);
    for(size_t a = 0; a < interval; a++)
    {
      auto t = terms(L[start + a]);
      if(t.size() < 6 || t[0].find("Adsorbate-atom-position") != 0) throw std::runtime_error("Cannot find matching strings in the range for reading positions!");
      double p[3] = {std::stod(t[3]), std::stod(t[4]), std::stod(t[5])};
      // The reference means to place every atom at its minimum image from the molecule's first atom, but its `first_bead_pos` is
      // declared inside the line loop (read_data.cpp:3147-3157) and holds nothing when the other atoms reach it; as built with
      // nvcc / gcc it is zero, so atoms 1.. end up at their minimum image from the ORIGIN while atom 0 stays where the file has it.
      // Energies do not notice (every pair goes through PBC), rotations about atom 0 do: the accept/reject sequence of a run that
      // starts from a RASPA-2 restart file only matches with the same coordinates.
      if(std::stol(t[2]) != 0) { double v[3] = {p[0], p[1], p[2]}; min_image(d, v); for(int k = 0; k < 3; k++) p[k] = v[k]; }
      for(int k = 0; k < 3; k++) c.restart_pos[3 * a + k] = p[k];
      c.restart_charge[a] = std::stod(terms(L[start + 3 * interval + a]).at(3));
      const double lambda = std::stod(terms(L[start + 4 * interval + a]).at(3));
      if(lambda < 1.0) throw std::runtime_error("restart file holds a fractional molecule: that needs the CB/CFC driver");
    }
  }
}

// LMPDataFileParser, read_data.cpp:2811-2998: the initial configuration from LMPDataInitial/System_0/init.data.  Atom lines carry
// "id mol type q x y z # component atomname"; atoms are put in id order, grouped by component name, cut into molecules of the
// component's size; the first atom of a molecule stays where the file puts it (WrapInBox takes its argument by value, so it
// wraps nothing) and the others are placed at their minimum image from it.
inline void read_lammps_box(Deck& d, const std::string& dir)
{
  if(!(d.restart_file && d.restart_lammps && d.lmp_read_box)) return;
  auto L = read_lines(dir + "/LMPDataInitial/System_0/init.data");
  bool fx = false, fy = false, fz = false, ft = false;
  double c3 = 0, c6 = 0, c7 = 0;
  for(auto& ln : L)
  {
    auto t = terms(ln);
    auto has = [&](const char* k) { return ln.find(k) != std::string::npos; };
    if(has("xlo") && has("xhi")) { d.cell[0] = std::stod(t.at(1)) - std::stod(t.at(0)); fx = true; }
    if(has("ylo") && has("yhi")) { d.cell[4] = std::stod(t.at(1)) - std::stod(t.at(0)); fy = true; }
    if(has("zlo") && has("zhi")) { d.cell[8] = std::stod(t.at(1)) - std::stod(t.at(0)); fz = true; }
    if(has("xy") && has("xz") && has("yz")) { c3 = std::stod(t.at(0)); c6 = std::stod(t.at(1)); c7 = std::stod(t.at(2)); ft = true; }
    if(has("atom types")) break;
  }
  if(!(fx && fy && fz)) throw std::runtime_error("no box size region in LMPDataInitial/System_0/init.data");
  d.cell[1] = d.cell[2] = d.cell[5] = 0.0;
  d.cell[3] = ft ? c3 : 0.0; d.cell[6] = ft ? c6 : 0.0; d.cell[7] = ft ? c7 : 0.0;
  invert_cell(d.cell, d.inv, d.volume);
}

inline void read_lammps_data(Deck& d, const std::string& dir)
{
  if(!(d.restart_file && d.restart_lammps)) return;
  if(d.lmp_start_component < 1) throw std::runtime_error("LMPData_Comp_to_Start_with 0 (framework atoms from the data file) is not read by this host program");
  auto L = read_lines(dir + "/LMPDataInitial/System_0/init.data");
  struct Rec { long id; int type; double q, p[3]; int comp; };
  std::vector<Rec> atoms; size_t total = 0, start = 0;
  for(size_t k = 0; k < L.size(); k++)
  {
    auto t = terms(L[k]);
    if(k < 3 && L[k].find("toms") != std::string::npos && !t.empty()) total = (size_t) std::stol(t[0]);
    if(L[k].find("Atoms") != std::string::npos) { start = k + 2; continue; }
    if(start > 0 && k >= start && atoms.size() < total)
    {
      if(t.size() != 10) throw std::runtime_error("LAMMPS data file: atom line " + std::to_string(k + 1) + " needs 10 fields (id mol type q x y z # component atom)");
      Rec r; r.id = std::stol(t[0]) - 1; r.type = std::stoi(t[2]) - 1; r.q = std::stod(t[3]);
      for(int m = 0; m < 3; m++) r.p[m] = std::stod(t[4 + m]);
      r.comp = -1;
      for(size_t c = 0; c < d.comps.size(); c++) if(d.comps[c].name == t[8]) r.comp = (int) c;
      atoms.push_back(r);
    }
  }
  std::stable_sort(atoms.begin(), atoms.end(), [](const Rec& a, const Rec& b) { return a.id < b.id; });
  const int nfw = 1 + (int) d.fw.size();
  for(size_t ci = 0; ci < d.comps.size(); ci++)
  {
    if((int) ci + nfw < d.lmp_start_component) continue;
    Component& c = d.comps[ci];
    const size_t ms = (size_t) c.ms();
    std::vector<const Rec*> mine;
    for(const Rec& r : atoms) if(r.comp == (int) ci) mine.push_back(&r);
    if(mine.empty()) continue;
    if(mine.size() % ms != 0) throw std::runtime_error("LAMMPS data file: atoms of component " + c.name + " are not a whole number of molecules");
    c.restart_pos.resize(3 * mine.size()); c.restart_charge.resize(mine.size());
    double first[3] = {0, 0, 0};
    for(size_t a = 0; a < mine.size(); a++)
    {
      const Rec& r = *mine[a];
      if(r.type != c.type[a % ms]) throw std::runtime_error("LAMMPS data file: atom types of component " + c.name + " do not follow its molecule definition");
      double p[3] = {r.p[0], r.p[1], r.p[2]};
      if(a % ms == 0) { for(int m = 0; m < 3; m++) first[m] = p[m]; }
      else { double v[3] = {p[0] - first[0], p[1] - first[1], p[2] - first[2]}; min_image(d, v); for(int m = 0; m < 3; m++) p[m] = first[m] + v[m]; }
      for(int m = 0; m < 3; m++) c.restart_pos[3 * a + m] = p[m];
      c.restart_charge[a] = r.q;
    }
  }
}

// ComputeFugacity, equations_of_state.h:121-330 (Peng-Robinson mixture, no binary interaction parameters) with the
// cubic solver of :61-118.  Runs when any adsorbate asks for "FugacityCoefficient PR-EOS" and then overrides the
// coefficients of EVERY adsorbate, as the reference does (:137-147).
inline void compute_fugacity(Deck& d)
{
  bool need = false;
  for(const auto& c : d.comps) if(c.fugacity_coeff < 0.0) need = true;
  if(!need) return;
  const double Rg = 8.314, P = d.pressure_pa, T = d.temperature;
  const size_t n = d.comps.size();
  std::vector<double> x(n), a(n), b(n), A(n), B(n);
  double sum = 0.0;
  for(size_t i = 0; i < n; i++)
  {
    const Component& c = d.comps[i];
    x[i] = c.mol_fraction; sum += x[i];
    const double Tr = T / c.tc;
    const double kappa = 0.37464 + 1.54226 * c.acentric - 0.26992 * std::pow(c.acentric, 2);
    const double alpha = std::pow(1.0 + kappa * (1.0 - std::sqrt(Tr)), 2);
    a[i] = 0.45724 * alpha * std::pow(Rg * c.tc, 2) / c.pc;
    b[i] = 0.07780 * Rg * c.tc / c.pc;
    A[i] = a[i] * P / std::pow(Rg * T, 2);
    B[i] = b[i] * P / (Rg * T);
  }
  if(std::fabs(sum - 1.0) > 0.0001) throw std::runtime_error("Sum of Mol Fractions does not equal 1.0");
  double Amix = 0.0, Bmix = 0.0;
  for(size_t i = 0; i < n; i++)
  {
    Bmix += x[i] * b[i];
    for(size_t j = 0; j < n; j++) Amix += x[i] * x[j] * std::sqrt(a[i] * a[j]);
  }
  Amix *= P / std::pow(Rg * T, 2);
  Bmix *= P / (Rg * T);
  // Z^3 + (Bmix - 1) Z^2 + (Amix - 3 Bmix^2 - 2 Bmix) Z - (Amix Bmix - Bmix^2 - Bmix^3) = 0
  const double c3 = 1.0, c2 = Bmix - 1.0, c1 = Amix - 3.0 * std::pow(Bmix, 2) - 2.0 * Bmix, c0 = -(Amix * Bmix - std::pow(Bmix, 2) - std::pow(Bmix, 3));
  std::vector<double> Z;
  {
    const double PI = 3.14159265358979323846, THIRD = 1.0 / 3.0;
    const double W = c2 / c3 * THIRD;
    double Pp = std::pow(c1 / c3 * THIRD - std::pow(W, 2), 3);
    const double Q = -.5 * (2.0 * std::pow(W, 3) - (c1 * W - c0) / c3);
    double DIS = std::pow(Q, 2) + Pp;
    if(DIS < 0.0)
    {
      const double PHI = std::acos(std::max(-1.0, std::min(1.0, Q / std::sqrt(-Pp))));
      Pp = 2.0 * std::pow((-Pp), 0.5 * THIRD);
      for(int i = 0; i < 3; i++) Z.push_back(Pp * std::cos((PHI + 2.0 * (double) i * PI) * THIRD) - W);
      std::sort(Z.begin(), Z.end());
    }
    else { DIS = std::sqrt(DIS); Z.push_back(std::cbrt(Q + DIS) + std::cbrt(Q - DIS) - W); }
    for(double& z : Z) z = z - (c0 + z * (c1 + z * (c2 + z * c3))) / (c1 + z * (2.0 * c2 + z * 3.0 * c3));   // one Newton step
  }
  if(Z.size() == 3) std::sort(Z.begin(), Z.end(), [](double u, double v) { return u > v; });   // descending, :247-262
  for(size_t i = 0; i < n; i++)
  {
    std::vector<double> phi(Z.size());
    double sumAij = 0.0;
    for(size_t k = 0; k < n; k++) sumAij += 2.0 * x[k] * std::sqrt(A[i] * A[k]);
    for(size_t j = 0; j < Z.size(); j++)
      phi[j] = std::exp((B[i] / Bmix) * (Z[j] - 1.0) - std::log(Z[j] - Bmix)
                        - (Amix / (2.0 * std::sqrt(2.0) * Bmix)) * (sumAij / Amix - B[i] / Bmix) *
                        std::log((Z[j] + (1.0 + std::sqrt(2.0)) * Bmix) / (Z[j] + (1.0 - std::sqrt(2.0)) * Bmix)));
    double f = phi[0];
    if(Z.size() == 3 && Z[2] > 0.0 && phi[0] > phi[2]) f = phi[2];                               // :309-330
    d.comps[i].fugacity_coeff = f;
  }
}

inline Deck load(const std::string& dir, double pressure_override = -1.0, double temperature_override = -1.0, int box = 0)
{
  Deck d;
  read_simulation_input(d, dir, box);
  if(pressure_override >= 0.0) d.pressure_pa = pressure_override;          // one isotherm point per process (and per GPU)
  if(temperature_override >= 0.0) d.temperature = temperature_override;
  read_force_field(d, dir);
  for(auto& c : d.comps) read_molecule(d, c, dir);
  read_framework_components(d, dir);
  read_framework(d, dir);
  read_block_pockets(d, dir);
  read_lammps_box(d, dir);
  read_restart(d, dir);
  read_lammps_data(d, dir);
  if(!d.no_charges) setup_ewald(d);
  // Setup_Box_Temperature_Pressure, fxn_main.h:115-127 with Units data_struct.h:58-68
  const double kB = 1.380649e-23, mass_unit = 1.6605402e-27, length_unit = 1e-10, time_unit = 1e-12;
  d.beta = 1.0 / (kB / (mass_unit * std::pow(length_unit, 2) / std::pow(time_unit, 2)) * d.temperature);
  d.pressure = d.pressure_pa / (mass_unit / (length_unit * std::pow(time_unit, 2)));
  compute_fugacity(d);
  // prepare_MixtureStats (fxn_main.h:639-656, called for mixtures only, main.cpp:326-330): the mol fractions are divided by
  // their sum over every component but component 0 -- which includes each separated framework component with its default 1.0
  // (fxn_main.h:70).  Runs after the equation of state, which sees the fractions as given (main.cpp:240).
  if(d.comps.size() > 1)
  {
    double tot = (double) d.fw.size();
    for(const auto& c : d.comps) tot += c.mol_fraction;
    for(auto& c : d.comps) c.mol_fraction /= tot;
  }
  return d;
}

} // namespace deck
