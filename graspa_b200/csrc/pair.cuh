// graspa_b200 -- warp-level pair loop: one warp evaluates one trial group (chainsize trial atoms whose
// energies are summed) against ranges of system atoms.
//
// Replaces the per-pair thread of Calculate_Multiple_Trial_Energy_VDWReal (VDW_Coulomb.cu:1183-1352) and of
// Calculate_Single_Body_Energy_VDWReal (:626-841).  Differences in structure, not in result:
//   - lanes stride over system atoms (coalesced SoA loads from the TMA-staged shared-memory pack or from L2);
//     the trial atoms of the group are warp-uniform registers;
//   - the distance test runs for every pair, but only pairs inside a cutoff are pushed (ballot + popc
//     compaction) into a per-warp shared-memory queue; the expensive LJ / erfc body is executed on full
//     32-entry batches, so the FP64 pipe does not idle on the ~85 % of lanes that fail the cutoff;
//   - minimum image in fractional space with a magic-number round (3 DADD per axis, no F2I/I2F).
#pragma once
#include "common.cuh"

// trial atoms of the group being evaluated, per warp, in shared memory
struct TrialGroup
{
  double fx[GBK_MAX_CS], fy[GBK_MAX_CS], fz[GBK_MAX_CS];   // fractional coordinates
  double q[GBK_MAX_CS];                                    // charge * scaleCoul
  double scale[GBK_MAX_CS];
  int    type[GBK_MAX_CS];
};

struct WarpQueue
{
  double r2[GBK_QCAP];
  int    code[GBK_QCAP];     // (system atom index << 6) | trial atom
};

struct PairAcc { double vdw, real; int flag; };

// the expensive body, executed by lanes [0, n)
__device__ __forceinline__ void drain_queue(const DevParams& P, const SysView& S, const TrialGroup* T, const WarpQueue* Q,
                                            int n, PairAcc& acc)
{
  const int lane = lane_id();
  if(lane < n)
  {
    const double r2 = Q->r2[lane];
    const int code = Q->code[lane];
    const int i = code >> 6, a = code & 63;
    const int row = S.type[i] * P.ntypes + T->type[a];
    double scaling = 1.0, qq = S.q[i] * T->q[a];
    if(!P.all_unit_scale) { scaling = S.scale[i] * T->scale[a]; qq *= S.scoul[i]; }
    else scaling = T->scale[a];
    pair_energy(P, r2, row, scaling, qq, acc.vdw, acc.real, acc.flag);
  }
}

// One range of system atoms [start, start+count) against the CS trial atoms in *T.
// excl_a / excl_b: molecule ids to skip in this range (-1: none) -- VDW_Coulomb.cu:1282-1283.
// wslice/nslice: this warp handles iterations wslice, wslice+nslice, ... (nslice = 1: the whole range).
template <int CS>
__device__ __forceinline__ void pair_range(const DevParams& P, const SysView& S, int start, int count,
                                           int excl_a, int excl_b, const TrialGroup* T, int cs_dyn, WarpQueue* Q,
                                           int wslice, int nslice, PairAcc& acc)
{
  const int lane = lane_id();
  const unsigned lt_mask = (1u << lane) - 1u;
  const int cs = CS > 0 ? CS : cs_dyn;
  const bool check_excl = (excl_a >= 0) || (excl_b >= 0);
  const bool charged = !P.no_charges;
  const double cut_max = charged ? fmax(P.cut_vdw2, P.cut_coul2) : P.cut_vdw2;
  double tx[CS > 0 ? CS : 1], ty[CS > 0 ? CS : 1], tz[CS > 0 ? CS : 1];
  if(CS > 0)
  {
#pragma unroll
    for(int a = 0; a < CS; a++) { tx[a] = T->fx[a]; ty[a] = T->fy[a]; tz[a] = T->fz[a]; }
  }
  int qn = 0;
  const int end = start + count;
  for(int base = start + 32 * wslice; base < end; base += 32 * nslice)
  {
    const int i = base + lane;
    bool valid = i < end;
    double ax = 0.0, ay = 0.0, az = 0.0;
    if(valid)
    {
      ax = S.fx[i]; ay = S.fy[i]; az = S.fz[i];
      if(check_excl) { const int m = S.molid[i]; if(m == excl_a || m == excl_b) valid = false; }
    }
#pragma unroll
    for(int a = 0; a < cs; a++)
    {
      double bx, by, bz;
      if(CS > 0) { bx = tx[a]; by = ty[a]; bz = tz[a]; } else { bx = T->fx[a]; by = T->fy[a]; bz = T->fz[a]; }
      const double r2 = min_image_r2(P, ax - bx, ay - by, az - bz);
      const bool hit = valid && (r2 < cut_max);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if(m)
      {
        if(hit) { const int p = qn + __popc(m & lt_mask); Q->r2[p] = r2; Q->code[p] = (i << 6) | a; }
        qn += __popc(m);
        if(qn >= 32)
        {
          __syncwarp();
          drain_queue(P, S, T, Q, 32, acc);
          const int rest = qn - 32;
          double r2m = 0.0; int cm = 0;
          if(lane < rest) { r2m = Q->r2[32 + lane]; cm = Q->code[32 + lane]; }
          __syncwarp();
          if(lane < rest) { Q->r2[lane] = r2m; Q->code[lane] = cm; }
          qn = rest;
          __syncwarp();
        }
      }
    }
  }
  if(qn > 0) { __syncwarp(); drain_queue(P, S, T, Q, qn, acc); __syncwarp(); }
}

// all segments of one kind-class for one trial group.  out4 = {HGvdw, HGreal, GGvdw, GGreal} for trial moves;
// kinds are mapped by the caller: acc_of[kind] in {0,1,2}.
template <int CS>
__device__ __forceinline__ void pair_group(const DevParams& P, const SysView& Sg, const SysView& Ss, const SegList& L,
                                           int new_comp, int new_molid, int excl_comp, int excl_mol,
                                           const TrialGroup* T, int cs_dyn, WarpQueue* Q, int wslice, int nslice,
                                           double* e6 /* HHv,HHr,HGv,HGr,GGv,GGr lane-partial sums */, int& flag)
{
  for(int s = 0; s < L.nseg; s++)
  {
    PairAcc acc; acc.vdw = 0.0; acc.real = 0.0; acc.flag = 0;
    const int ea = (L.comp[s] == excl_comp) ? excl_mol : -1;
    const int eb = (L.comp[s] == new_comp) ? new_molid : -1;
    pair_range<CS>(P, L.staged[s] ? Ss : Sg, L.start[s], L.count[s], ea, eb, T, cs_dyn, Q, wslice, nslice, acc);
    e6[2 * L.kind[s]] += acc.vdw; e6[2 * L.kind[s] + 1] += acc.real; flag |= acc.flag;
  }
}
