// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Host harness around the reference's OWN routines.  It #includes the reference
// headers where they lie under /root/reference/src_clean (the include path is
// given by oracle/build_ref.sh; nothing is copied into this repository) and
// exposes a flat C ABI so that tests/ and tests/golden/make_golden.py can
//   (a) validate the plain-C restatement in oracle/graspa_oracle.c, and
//   (b) generate the committed golden vectors under tests/golden/.
//
// Reference routines called here, unmodified:
//   PBC, VDW, CoulombReal, inverse_matrix, matrix_determinant   maths.cuh:28-52,427-500
//   Ewald_Total, Calculate_Intra_Molecule_Exclusion,
//   Calculate_Self_Exclusion                                     ewald_preparation.h:5-298
//   TotalTailCorrection, TailCorrectionDifference,
//   TailCorrectionIdentitySwap                                   TailCorrection_Energy_Functions.h:3-113
// Loops that in the reference contain cudaMemcpy / kernel indexing (the per-pair
// bodies of VDW_Coulomb.cu:1270-1327 and :741-818) are restated around those
// same primitives.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <omp.h>
#include "VDW_Coulomb.cuh"
#include "maths.cuh"
#include "ewald_preparation.h"
#include "TailCorrection_Energy_Functions.h"

// data_struct.cpp:6-11 (the reference's uniform RNG; needed to satisfy the linker and
// used by ref_uniform_stream to pin the glibc rand() stream)
double Get_Uniform_Random() { return static_cast<double>(std::rand()) / RAND_MAX; }

namespace {
struct FlatSystem
{
  int ncomp, nhost;
  const long long* natoms;    // live atoms per component
  const long long* molsize;   // atoms per molecule per component
  const double* pos;          // 3*N, components concatenated
  const double* scale; const double* charge; const double* scaleCoul;
  const long long* type; const long long* molid;
};

static void fill_atoms(const FlatSystem& S, std::vector<Atoms>& A, std::vector<std::vector<double3>>& P,
                       std::vector<std::vector<size_t>>& T, std::vector<std::vector<size_t>>& M)
{
  A.resize(S.ncomp); P.resize(S.ncomp); T.resize(S.ncomp); M.resize(S.ncomp);
  size_t off = 0;
  for(int c = 0; c < S.ncomp; c++)
  {
    size_t n = (size_t) S.natoms[c];
    P[c].resize(n + 1); T[c].resize(n + 1); M[c].resize(n + 1);
    for(size_t i = 0; i < n; i++)
    {
      P[c][i] = {S.pos[3*(off+i)], S.pos[3*(off+i)+1], S.pos[3*(off+i)+2]};
      T[c][i] = (size_t) S.type[off+i]; M[c][i] = (size_t) S.molid[off+i];
    }
    A[c].pos = P[c].data();
    A[c].scale = const_cast<double*>(S.scale + off);
    A[c].charge = const_cast<double*>(S.charge + off);
    A[c].scaleCoul = const_cast<double*>(S.scaleCoul + off);
    A[c].Type = T[c].data(); A[c].MolID = M[c].data();
    A[c].Molsize = (size_t) S.molsize[c]; A[c].size = n; A[c].Allocate_size = n;
    off += n;
  }
}
} // namespace

extern "C" {

void ref_inverse_matrix(const double* cell, double* inv, double* det)
{
  double* r = nullptr; inverse_matrix(const_cast<double*>(cell), &r);
  for(int i = 0; i < 9; i++) inv[i] = r[i];
  free(r);
  *det = matrix_determinant(const_cast<double*>(cell));
}

void ref_uniform_stream(int seed, long long n, double* out)
{
  std::srand(seed);
  for(long long i = 0; i < n; i++) out[i] = Get_Uniform_Random();
}

// one pair through PBC + VDW + CoulombReal; out = {rr, Evdw, dUdlambda, Ecoul, in_vdw, in_coul}
void ref_pair(const double* cell, const double* inv, int cubic, const double* posA, const double* posB,
              const double* ffarg, double scaling, int use1264, double qA, double qB, double scalingCoul,
              double prefactor, double alpha, double cutvdw2, double cutcoul2, int nocharges, double* out)
{
  double3 posvec = {posA[0] - posB[0], posA[1] - posB[1], posA[2] - posB[2]};
  PBC(posvec, const_cast<double*>(cell), const_cast<double*>(inv), cubic != 0);
  double rr = dot(posvec, posvec);
  out[0] = rr; out[1] = out[2] = out[3] = out[4] = out[5] = 0.0;
  if(rr < cutvdw2)
  {
    double res[2] = {0.0, 0.0};
    VDW(ffarg, rr, scaling, res, use1264 != 0);
    out[1] = res[0]; out[2] = res[1]; out[4] = 1.0;
  }
  if(!nocharges && rr < cutcoul2)
  {
    double res[2] = {0.0, 0.0};
    CoulombReal(qA, qB, sqrt(rr), scalingCoul, res, prefactor, alpha);
    out[3] = res[0]; out[5] = 1.0;
  }
}

// Per-trial HG/GG VDW/real energies + overlap flags for NTrials x chainsize trial atoms.
// Pair body: VDW_Coulomb.cu:1270-1327; sums are plain sequential per trial in atom order
// (the reference's 128-thread tree order is restated in oracle/graspa_oracle.c).
// out_energy[trial*4 + {0:HGvdw,1:HGreal,2:GGvdw,3:GGreal}], out_flag[trial], counts = {Npairs, Nvdw, Ncoul}
void ref_trial_energies(int ncomp, int nhost, const long long* natoms, const long long* molsize,
                        const double* pos, const double* scale, const double* charge, const double* scaleCoul,
                        const long long* type, const long long* molid,
                        const double* cell, const double* inv, int cubic, double prefactor, double alpha,
                        int ntypes, const double* eps, const double* sigma, const double* z, const double* shift, const double* c10,
                        double cutvdw2, double cutcoul2, double overlap, int nocharges, int use1264,
                        int ntrials, int chainsize, const double* tpos, const double* tscale, const double* tcharge,
                        const double* tscaleCoul, const long long* ttype, long long new_molid, int new_comp,
                        int excl_comp, long long excl_mol,
                        double* out_energy, int* out_flag, long long* counts)
{
  counts[0] = counts[1] = counts[2] = 0;
  for(int t = 0; t < ntrials; t++)
  {
    double e[4] = {0, 0, 0, 0}; int flag = 0;
    size_t off = 0;
    for(int c = 0; c < ncomp; c++)
    {
      int kind = (c < nhost) ? 0 : 2;
      for(long long i = 0; i < natoms[c]; i++)
      {
        size_t g = off + i;
        bool consider = true;
        if(c == excl_comp && molid[g] == excl_mol) consider = false;
        if(molid[g] == new_molid && c == new_comp)   consider = false;
        if(!consider) continue;
        for(int a = 0; a < chainsize; a++)
        {
          int j = t * chainsize + a;
          double3 posvec = {pos[3*g] - tpos[3*j], pos[3*g+1] - tpos[3*j+1], pos[3*g+2] - tpos[3*j+2]};
          PBC(posvec, const_cast<double*>(cell), const_cast<double*>(inv), cubic != 0);
          const double rr = dot(posvec, posvec);
          counts[0]++;
          if(rr < cutvdw2)
          {
            double res[2] = {0.0, 0.0};
            size_t row = (size_t) type[g] * ntypes + (size_t) ttype[j];
            const double FFarg[5] = {eps[row], sigma[row], z[row], shift[row], c10[row]};
            VDW(FFarg, rr, scale[g] * tscale[j], res, use1264 != 0);
            if(res[0] > overlap) flag = 1;
            if(rr < 0.01) flag = 1;
            e[kind] += res[0]; counts[1]++;
          }
          if(!nocharges && rr < cutcoul2)
          {
            double res[2] = {0.0, 0.0};
            CoulombReal(charge[g], tcharge[j], sqrt(rr), scaleCoul[g] * tscaleCoul[j], res, prefactor, alpha);
            e[kind + 1] += res[0]; counts[2]++;
          }
        }
      }
      off += natoms[c];
    }
    for(int k = 0; k < 4; k++) out_energy[4*t + k] = e[k];
    out_flag[t] = flag;
  }
}

// The reference's CPU Ewald_Total (ewald_preparation.h:5-259), called as is.
// out_E = {GGEwaldE, HHEwaldE, HGEwaldE} (GG already includes +HH, self and intra terms exactly as the reference returns it),
// sf_ads / sf_fw = stored structure factors, nvec complex each (interleaved re,im).
void ref_ewald_total(int ncomp, int nhost, const long long* natoms, const long long* molsize,
                     const double* pos, const double* scale, const double* charge, const double* scaleCoul,
                     const long long* type, const long long* molid,
                     const double* cell, const double* inv, int cubic, double volume, double prefactor, double alpha,
                     const int* kmax, double recip_cutoff, int use_lammps, int nocharges,
                     double* out_E, double* sf_ads, double* sf_fw)
{
  FlatSystem S{ncomp, nhost, natoms, molsize, pos, scale, charge, scaleCoul, type, molid};
  std::vector<Atoms> A; std::vector<std::vector<double3>> P; std::vector<std::vector<size_t>> T, M;
  fill_atoms(S, A, P, T, M);
  Boxsize Box{};
  Box.Cell = const_cast<double*>(cell); Box.InverseCell = const_cast<double*>(inv);
  Box.Volume = volume; Box.ReciprocalCutOff = recip_cutoff; Box.Prefactor = prefactor; Box.Alpha = alpha;
  Box.Cubic = cubic != 0; Box.UseLAMMPSEwald = use_lammps != 0; Box.kmax = {kmax[0], kmax[1], kmax[2]};
  ForceField FF{}; FF.noCharges = nocharges != 0;
  Components C;
  C.NComponents = {ncomp, nhost, ncomp - nhost};
  C.NumberOfFrameworks = nhost > 0 ? 1 : 0;
  for(int c = 0; c < ncomp; c++)
  {
    C.Moleculesize.push_back((size_t) molsize[c]);
    C.NumberOfMolecule_for_Component.push_back(molsize[c] > 0 ? (size_t)(natoms[c] / molsize[c]) : 0);
  }
  C.OUTPUT = fopen("/dev/null", "w");
  MoveEnergy E;
  Atoms* Aptr = A.data();
  Ewald_Total(Box, Aptr, FF, C, E);
  fclose(C.OUTPUT);
  out_E[0] = E.GGEwaldE; out_E[1] = E.HHEwaldE; out_E[2] = E.HGEwaldE;
  for(size_t i = 0; i < C.AdsorbateEik.size(); i++)
  {
    sf_ads[2*i] = C.AdsorbateEik[i].real(); sf_ads[2*i+1] = C.AdsorbateEik[i].imag();
    sf_fw[2*i]  = C.FrameworkEik[i].real(); sf_fw[2*i+1]  = C.FrameworkEik[i].imag();
  }
}

// Rigid-molecule exclusion constants of one component (ewald_preparation.h:261-298, :351-366),
// computed on the component's first molecule exactly as the reference does.
void ref_exclusion_rigid(int molsize, const double* pos, const double* charge, const double* scaleCoul,
                         const double* cell, const double* inv, int cubic, double prefactor, double alpha,
                         double* out_intra, double* out_self)
{
  std::vector<double3> P(molsize);
  for(int i = 0; i < molsize; i++) P[i] = {pos[3*i], pos[3*i+1], pos[3*i+2]};
  Atoms A{}; A.pos = P.data(); A.charge = const_cast<double*>(charge); A.scaleCoul = const_cast<double*>(scaleCoul);
  A.Molsize = molsize; A.size = molsize;
  Boxsize Box{}; Box.Cell = const_cast<double*>(cell); Box.InverseCell = const_cast<double*>(inv); Box.Cubic = cubic != 0;
  Components C; C.Moleculesize.push_back((size_t) molsize); C.OUTPUT = fopen("/dev/null", "w");
  *out_intra = Calculate_Intra_Molecule_Exclusion(Box, &A, alpha, prefactor, C, 0);
  *out_self  = Calculate_Self_Exclusion(Box, &A, alpha, prefactor, C, 0);
  fclose(C.OUTPUT);
}

static void fill_tail(Components& C, int ntypes, const long long* npseudo, const int* use_tail, const double* tail_energy,
                      int ncomp, const int* species_nentries, const int* species_type, const int* species_count)
{
  C.HasTailCorrection = true;
  for(int i = 0; i < ntypes; i++) C.NumberOfPseudoAtoms.push_back((size_t) npseudo[i]);
  C.TailCorrection.resize((size_t) ntypes * ntypes);
  for(int i = 0; i < ntypes * ntypes; i++) { C.TailCorrection[i].UseTail = use_tail[i] != 0; C.TailCorrection[i].Energy = tail_energy[i]; }
  C.NumberOfPseudoAtomsForSpecies.resize(ncomp);
  int k = 0;
  for(int c = 0; c < ncomp; c++)
    for(int e = 0; e < species_nentries[c]; e++, k++) C.NumberOfPseudoAtomsForSpecies[c].push_back({species_type[k], species_count[k]});
}

double ref_tail_total(int ntypes, const long long* npseudo, const int* use_tail, const double* tail_energy, double volume)
{
  Components C; int zero = 0;
  fill_tail(C, ntypes, npseudo, use_tail, tail_energy, 0, &zero, nullptr, nullptr);
  return TotalTailCorrection(C, (size_t) ntypes, volume);
}

double ref_tail_difference(int ntypes, const long long* npseudo, const int* use_tail, const double* tail_energy, double volume,
                           int ncomp, const int* species_nentries, const int* species_type, const int* species_count,
                           int comp, int movetype)
{
  Components C;
  fill_tail(C, ntypes, npseudo, use_tail, tail_energy, ncomp, species_nentries, species_type, species_count);
  return TailCorrectionDifference(C, (size_t) comp, (size_t) ntypes, volume, movetype);
}

double ref_tail_identity_swap(int ntypes, const long long* npseudo, const int* use_tail, const double* tail_energy, double volume,
                              int ncomp, const int* species_nentries, const int* species_type, const int* species_count,
                              int newcomp, int oldcomp)
{
  Components C;
  fill_tail(C, ntypes, npseudo, use_tail, tail_energy, ncomp, species_nentries, species_type, species_count);
  return TailCorrectionIdentitySwap(C, (size_t) newcomp, (size_t) oldcomp, (size_t) ntypes, volume);
}

int ref_movetype_insertion() { return INSERTION; }
int ref_movetype_deletion()  { return DELETION; }

} // extern "C"
