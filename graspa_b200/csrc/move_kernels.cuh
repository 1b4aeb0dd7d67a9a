// graspa_b200 -- kernels of the single-move path (CBMC stages, translation/rotation, state commit).
// One launch per CBMC stage: trial generation + pair energies + Boltzmann/selection/Rosenbluth (the last CTA to
// finish runs the host code of mc_widom.h:305-383 / 568-611 on the device), one 128-byte result read.
#pragma once
#include "common.cuh"
#include "pair.cuh"

#define GBK_MV_MAXT 32
#define GBK_MV_TRIAL_SLOTS 1024       // Setup_Temporary_Atoms_Structure allocates 1024 slots (fxn_main.h:99-113)
#define GBK_MV_MOL_SLOTS 64
enum { GBK_BUF_GROWN = 0, GBK_BUF_OLD = 1, GBK_BUF_NEW = 2, GBK_BUF_TEMP = 3 };

// device scratch of the move path: trial atoms (Sims.New), and four molecule buffers:
// grown (Sims.Old[0] + selected orientation), old / new (Sims.Old / Sims.New of single-body moves), temp (tempMolStorage)
struct MoveBufs
{
  double* d; int* i;
  __host__ __device__ double* tr(int arr) const { return d + (size_t) arr * GBK_MV_TRIAL_SLOTS; }           // 0..8: x y z fx fy fz q scale scoul
  __host__ __device__ int* tr_type() const { return i; }
  __host__ __device__ double* mol(int buf, int arr) const { return d + 9 * GBK_MV_TRIAL_SLOTS + ((size_t) buf * 9 + arr) * GBK_MV_MOL_SLOTS; }
  __host__ __device__ int* mol_type(int buf) const { return i + GBK_MV_TRIAL_SLOTS + buf * GBK_MV_MOL_SLOTS; }
  __host__ __device__ double* stage_e() const { return d + 9 * GBK_MV_TRIAL_SLOTS + 36 * GBK_MV_MOL_SLOTS; }   // [32][6]
  __host__ __device__ int* stage_flag() const { return i + GBK_MV_TRIAL_SLOTS + 4 * GBK_MV_MOL_SLOTS; }      // [32]
  // result slots of 16 doubles: 0 first bead (new), 1 chain (new), 2 first bead (old/retrace), 3 chain (old/retrace),
  // 4 single-body delta, 5 Ewald {same, 2*cross}, 6-7 spare.  Fused move calls read all of them back in one copy.
  __host__ __device__ double* result(int slot = 0) const { return stage_e() + GBK_MV_MAXT * 6 + 16 * slot; }
  __host__ __device__ double* partial() const { return result(0) + 128; }                                     // 8192 doubles of CTA partials (128-byte aligned)
  static size_t doubles() { return 9 * GBK_MV_TRIAL_SLOTS + 36 * GBK_MV_MOL_SLOTS + GBK_MV_MAXT * 6 + 128 + 8192; }
  static size_t ints() { return GBK_MV_TRIAL_SLOTS + 4 * GBK_MV_MOL_SLOTS + GBK_MV_MAXT + 16; }
};

// component slot arrays (Cartesian + properties), for reading existing molecules / the template
struct CompView
{
  const double* __restrict__ x; const double* __restrict__ y; const double* __restrict__ z;
  const double* __restrict__ q; const double* __restrict__ scale; const double* __restrict__ scoul;
  const int* __restrict__ type;
  // block pockets of the component (gb_set_block_pockets): npocket x {x, y, z, radius}, Cartesian centres
  const double* __restrict__ pocket; int npocket, pocket_invert;
};

// BlockedPocket, read_data.cpp:3466-3640 (RASPA-2's rule): a point is blocked when it lies inside any block-pocket sphere
// (InvertBlockPockets: when it lies outside all of them).  The difference centre - point goes through the same
// truncating nearest-image round as the reference's apply_pbc_raspa2 lambda.
__device__ __forceinline__ bool blocked_pocket(const DevParams& P, const CompView& C, double x, double y, double z)
{
  if(C.npocket <= 0) return false;
  for(int i = 0; i < C.npocket; i++)
  {
    double dx = C.pocket[4 * i] - x, dy = C.pocket[4 * i + 1] - y, dz = C.pocket[4 * i + 2] - z;
    if(P.cubic)
    {
      dx -= P.cell[0] * (double) static_cast<int>(dx / P.cell[0] + ((dx >= 0.0) ? 0.5 : -0.5));
      dy -= P.cell[4] * (double) static_cast<int>(dy / P.cell[4] + ((dy >= 0.0) ? 0.5 : -0.5));
      dz -= P.cell[8] * (double) static_cast<int>(dz / P.cell[8] + ((dz >= 0.0) ? 0.5 : -0.5));
    }
    else
    {
      const double sx = P.inv[0] * dx + P.inv[3] * dy + P.inv[6] * dz, sy = P.inv[1] * dx + P.inv[4] * dy + P.inv[7] * dz, sz = P.inv[2] * dx + P.inv[5] * dy + P.inv[8] * dz;
      const double tx = sx - (double) static_cast<int>(sx + ((sx >= 0.0) ? 0.5 : -0.5));
      const double ty = sy - (double) static_cast<int>(sy + ((sy >= 0.0) ? 0.5 : -0.5));
      const double tz = sz - (double) static_cast<int>(sz + ((sz >= 0.0) ? 0.5 : -0.5));
      dx = P.cell[0] * tx + P.cell[3] * ty + P.cell[6] * tz; dy = P.cell[1] * tx + P.cell[4] * ty + P.cell[7] * tz; dz = P.cell[2] * tx + P.cell[5] * ty + P.cell[8] * tz;
    }
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    if(r < C.pocket[4 * i + 3]) return C.pocket_invert == 0;
  }
  return C.pocket_invert != 0;
}

struct CbmcArgs
{
  int cbmc_type, comp, ms, ntrials, norm, new_molid, excl_comp, excl_mol, first_bead_trial;
  // chaining of stages on the device (fused move calls): where this stage writes its result, which earlier stage must
  // have succeeded for this one to run (-1: none), and from which result slot StoredR is taken (-1: the stored_r argument)
  int rslot, dep_slot, stored_slot;
  long long molecule;          // SelectedMolInComponent
  long long pool_off;
  const double* __restrict__ pool3;
  double uniform, scale, scale_coul, stored_r;
  double preset[3]; int has_preset;   // 1: preset[] holds the position; 2: read it from preset_src (copy_firstbead_to_new, mc_swap_moves.h:178-181)
  const double* preset_src[3];
  CompView C;                  // slots of the component (index 0 = slot 0 of the component)
  MoveBufs B;
  unsigned int* ticket;
};

// Boltzmann factors, selection and Rosenbluth weight of one CBMC stage, run by ONE WARP (lane t = trial t) of the last CTA:
// the tail of CBMC_FirstBead_Finish (mc_widom.h:305-383) / Widom_Move_Chain_PARTIAL (:568-611) with
// Host_sum_Widom_HGGG_SEPARATE (:42-87) in front.  Sums run in trial order like the host code (rosenbluth_warp).
// result layout: r[0] rosenbluth, r[1] stored_r, r[2..5] energy, r[6..8] selected pos, r[9] success, r[10] selected,
// r[11] nsurv, r[12] uniform used, r[13] growth still alive (success and running product > 1e-150), r[14] running product
__device__ __forceinline__ void cbmc_finish_core(const DevParams& P, int ty, bool is_chain, int ntrials, int norm, double uniform, double stored_r,
                                                 const double* E, const int* F, double* r)
{
  const int lane = (int) lane_id();
  double e[6] = {0, 0, 0, 0, 0, 0}; bool surv = false;
  if(lane < ntrials)
  {
#pragma unroll
    for(int k = 0; k < 6; k++) e[k] = E[6 * lane + k];
    surv = F[lane] == 0;
  }
  // HH terms (a framework component grown against framework components) count with HG: VDW_Coulomb.cu:1232-1235
  double tot = (e[0] + e[2]) + e[4];
  if(P.vdw_real_bias) tot += (e[1] + e[3]) + e[5];
  const bool insertion_like = (ty == 0 /*CBMC_INSERTION*/ || ty == 2 /*REINSERTION_INSERTION*/ || (is_chain && ty == 4 /*IDENTITY_SWAP_NEW*/));
  const bool needs_survivor = insertion_like || ty == 4;
  const RosenResult rr = rosenbluth_warp(-P.beta * tot, surv, ntrials, uniform, insertion_like);
  const int ns = rr.nsurv;
  const bool good = needs_survivor ? (ns > 0 && !(rr.R < 1e-150)) : true;
  const int sel = rr.sel_lane;
  const double hgv = __shfl_sync(0xffffffffu, e[2] + e[0], sel), hgr = __shfl_sync(0xffffffffu, e[3] + e[1], sel);
  const double ggv = __shfl_sync(0xffffffffu, e[4], sel), ggr = __shfl_sync(0xffffffffu, e[5], sel);
  if(lane != 0) return;
  for(int k = 0; k < 16; k++) r[k] = 0.0;
  r[11] = ns; r[12] = (insertion_like && ns > 0) ? 1.0 : 0.0;
  if(!good) return;
  if(ns == 0) { r[9] = 1.0; return; }          // deletion types with every trial overlapping: Rosenbluth 0, Trialindex empty
  double avg = rr.R;
  if(!is_chain)
  {
    if(ty == 2) r[1] = rr.R_minus_sel;                                                          // StoredR, mc_widom.h:365
    if(ty == 3) avg += stored_r;                                                                // REINSERTION_RETRACE :366
    if(ty != 4 && ty != 5) avg /= (double) norm;                                                // :369-370
  }
  else avg = rr.R / (double) norm;                                                              // :601
  if(!P.vdw_real_bias) avg *= exp(-P.beta * (hgr + ggr));                                       // :373-377, :603-607
  r[0] = avg; r[2] = hgv; r[3] = hgr; r[4] = ggv; r[5] = ggr; r[9] = 1.0; r[10] = sel;
}

__device__ __forceinline__ void cbmc_finish_warp(const DevParams& P, const CbmcArgs& A, bool is_chain, double* r)
{
  const double stored = (A.stored_slot >= 0) ? A.B.result(A.stored_slot)[1] : A.stored_r;
  cbmc_finish_core(P, A.cbmc_type, is_chain, A.ntrials, A.norm, A.uniform, stored, A.B.stage_e(), A.B.stage_flag(), r);
}

// pair energies of the trial group in *T for this CTA's slice of the atom ranges; partial sums go to
// partial[(group * nsplit + split) * 8 + {HHv,HHr,HGv,HGr,GGv,GGr,flag}]
template <int CS>
__device__ __forceinline__ void cbmc_group_energy(const DevParams& P, const PairTables& W, const SysView& S, const SegList& L, const CbmcArgs& A,
                                                  TrialGroup* T, WarpQueue* Q, double* red, int cs, int group, int split, int nsplit)
{
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = (int) lane_id();
  double e6[6] = {0, 0, 0, 0, 0, 0}; int flag = 0;
  pair_group_generic<CS>(P, W, S, L, A.comp, A.new_molid, A.excl_comp, A.excl_mol, T, cs, Q + warp, split * nwarps + warp, nsplit * nwarps, e6, flag);
#pragma unroll
  for(int k = 0; k < 6; k++) e6[k] = warp_sum(e6[k]);
  flag = __any_sync(0xffffffffu, flag);
  if(lane == 0) { for(int k = 0; k < 6; k++) red[warp * 8 + k] = e6[k]; red[warp * 8 + 6] = flag ? 1.0 : 0.0; }
  __syncthreads();
  if(threadIdx.x < 7)
  {
    double s = 0.0;
    for(int w = 0; w < nwarps; w++) s += red[w * 8 + threadIdx.x];
    A.B.partial()[(size_t)(group * nsplit + split) * 8 + threadIdx.x] = s;
  }
}

// last CTA: fixed-order sum of the split partials into stage_e / stage_flag (one thread per trial)
__device__ __forceinline__ void cbmc_collect(const CbmcArgs& A, int nsplit)
{
  if(threadIdx.x < A.ntrials)
  {
    const volatile double* p = A.B.partial() + (size_t) threadIdx.x * nsplit * 8;
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for(int k = 0; k < nsplit; k++) for(int j = 0; j < 7; j++) s[j] += p[k * 8 + j];
    for(int j = 0; j < 6; j++) A.B.stage_e()[6 * threadIdx.x + j] = s[j];
    A.B.stage_flag()[threadIdx.x] = s[6] > 0.0 ? 1 : 0;
  }
  __syncthreads();
}

// first bead: get_random_trial_position (mc_widom.h:122-213) + energies + CBMC_FirstBead_Finish (:305-383).
// grid = ntrials * nsplit CTAs: CTA (t, k) evaluates trial t against slice k of the system atoms.
__global__ void __launch_bounds__(256)
k_cbmc_first_bead(DevParams P, SysView S, SegList L, CbmcArgs A, int nsplit)
{
  __shared__ TrialGroup T;
  __shared__ WarpQueue Q[8];
  __shared__ double red[8 * 8];
  __shared__ double etab[(GBK_ERFC_DEG + 1) * GBK_ERFC_NINT];
  __shared__ bool last;
  __shared__ int blocked_s;
  if(A.dep_slot >= 0 && A.B.result(A.dep_slot)[13] == 0.0)
  {
    // the stage this one depends on failed (or left a Rosenbluth weight <= 1e-150): report failure, consume nothing
    if(blockIdx.x == 0 && threadIdx.x < 16) A.B.result(A.rslot)[threadIdx.x] = 0.0;
    return;
  }
  stage_erfc_table(P, etab);
  const int t = blockIdx.x / nsplit, split = blockIdx.x % nsplit;
  if(threadIdx.x == 0)
  {
    const int ty = A.cbmc_type;
    const bool insertion_like = (ty == 0 || ty == 4);
    const long long start = insertion_like ? 0 : A.molecule * A.ms;
    double scale = A.scale, scoul = A.scale_coul;
    if(!insertion_like) { scale = A.C.scale[start]; scoul = A.C.scoul[start]; }
    double x, y, z;
    const bool existing = (ty == 1 || ty == 3 || ty == 5) && t == 0;
    if(existing) { x = A.C.x[start]; y = A.C.y[start]; z = A.C.z[start]; }
    else if(ty == 4 && A.has_preset == 2) { x = A.preset_src[0][0]; y = A.preset_src[1][0]; z = A.preset_src[2][0]; }
    else if(ty == 4 && A.has_preset) { x = A.preset[0]; y = A.preset[1]; z = A.preset[2]; }
    else { const double* r = A.pool3 + 3 * (A.pool_off + t); x = P.cell[0] * r[0]; y = P.cell[4] * r[1]; z = P.cell[8] * r[2]; }
    double fx, fy, fz; to_frac(P, x, y, z, fx, fy, fz);
    const double q = A.C.q[start]; const int type = A.C.type[start];
    if(split == 0)
    {
      A.B.tr(0)[t] = x; A.B.tr(1)[t] = y; A.B.tr(2)[t] = z; A.B.tr(3)[t] = fx; A.B.tr(4)[t] = fy; A.B.tr(5)[t] = fz;
      A.B.tr(6)[t] = q; A.B.tr(7)[t] = scale; A.B.tr(8)[t] = scoul; A.B.tr_type()[t] = type;
    }
    T.fx[0] = fx; T.fy[0] = fy; T.fz[0] = fz; T.q[0] = q * scoul; T.scale[0] = scale; T.type[0] = type; T.slot[0] = 0;
    // block pockets, mc_widom.h:445-497: growth types only; a blocked STARTING bead (trial 0) flags every trial,
    // otherwise each trial is flagged on its own
    int blk = 0;
    if((ty == 0 || ty == 2 || ty == 4) && A.C.npocket > 0)
    {
      double x0 = x, y0 = y, z0 = z;
      if(t > 0) { const double* r0 = A.pool3 + 3 * A.pool_off; x0 = P.cell[0] * r0[0]; y0 = P.cell[4] * r0[1]; z0 = P.cell[8] * r0[2]; }
      blk = blocked_pocket(P, A.C, x0, y0, z0) ? 1 : 0;
      if(!blk && t > 0) blk = blocked_pocket(P, A.C, x, y, z) ? 1 : 0;
    }
    blocked_s = blk;
  }
  __syncthreads();
  PairTables W; W.etab = etab; W.ffp = P.ffA; W.unit = false;
  cbmc_group_energy<1>(P, W, S, L, A, &T, Q, red, 1, t, split, nsplit);
  if(threadIdx.x == 6 && blocked_s) A.B.partial()[(size_t)(t * nsplit + split) * 8 + 6] = 1.0;   // same thread that stored the overlap flag
  __syncthreads();
  if(threadIdx.x == 0) { __threadfence(); last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1); }
  __syncthreads();
  if(!last) return;
  __threadfence();
  cbmc_collect(A, nsplit);
  if(threadIdx.x < 32)
  {
    double* r = A.B.result(A.rslot);
    cbmc_finish_warp(P, A, false, r);
    if(threadIdx.x == 0)
    {
      r[14] = r[0];                                                   // running Rosenbluth product of the growth
      r[13] = (r[9] != 0.0 && r[0] > 1e-150) ? 1.0 : 0.0;             // "Rosenbluth <= 1e-150 -> SuccessConstruction = false", mc_swap_utilities.h:21
      if(r[9] != 0.0 && r[11] > 0.0)
      {
        const int s = (int) r[10];
        r[6] = A.B.tr(0)[s]; r[7] = A.B.tr(1)[s]; r[8] = A.B.tr(2)[s];
        // Mol.pos[0] = NewMol.pos[FirstBeadTrial] etc., mc_widom.h:230-241
        for(int k = 0; k < 9; k++) A.B.mol(GBK_BUF_GROWN, k)[0] = A.B.tr(k)[s];
        A.B.mol_type(GBK_BUF_GROWN)[0] = A.B.tr_type()[s];
      }
      *A.ticket = 0u;
    }
  }
}

// chain: get_random_trial_orientation (mc_widom.h:215-303) + energies + the tail of Widom_Move_Chain_PARTIAL (:568-611)
__global__ void __launch_bounds__(256)
k_cbmc_chain(DevParams P, SysView S, SegList L, CbmcArgs A, int nsplit)
{
  __shared__ TrialGroup T;
  __shared__ WarpQueue Q[8];
  __shared__ double red[8 * 8];
  __shared__ double etab[(GBK_ERFC_DEG + 1) * GBK_ERFC_NINT];
  __shared__ bool last;
  __shared__ int sel_s;
  if(A.dep_slot >= 0 && A.B.result(A.dep_slot)[13] == 0.0)
  {
    if(blockIdx.x == 0 && threadIdx.x < 16) A.B.result(A.rslot)[threadIdx.x] = 0.0;
    return;
  }
  stage_erfc_table(P, etab);
  const int o = blockIdx.x / nsplit, split = blockIdx.x % nsplit, cs = A.ms - 1;
  if(threadIdx.x < cs)
  {
    const int a = threadIdx.x, ty = A.cbmc_type;
    const bool insertion_like = (ty == 0 || ty == 4);
    const long long start = (insertion_like ? 0 : A.molecule * A.ms) + 1;        // start_position, mc_widom.h:536-556
    const double fbx = A.B.mol(GBK_BUF_GROWN, 0)[0], fby = A.B.mol(GBK_BUF_GROWN, 1)[0], fbz = A.B.mol(GBK_BUF_GROWN, 2)[0];
    double vx = A.C.x[1 + a] - A.C.x[0], vy = A.C.y[1 + a] - A.C.y[0], vz = A.C.z[1 + a] - A.C.z[0];   // :256
    double x, y, z;
    if((ty == 1 || ty == 3 || ty == 5) && o == 0) { x = A.C.x[start + a]; y = A.C.y[start + a]; z = A.C.z[start + a]; }
    else
    {
      const double* r = A.pool3 + 3 * (A.pool_off + o);
      rotate_quaternion(vx, vy, vz, r[0], r[1], r[2]);
      x = fbx + vx; y = fby + vy; z = fbz + vz;
    }
    double fx, fy, fz; to_frac(P, x, y, z, fx, fy, fz);
    const double scale = A.B.mol(GBK_BUF_GROWN, 7)[0], scoul = A.B.mol(GBK_BUF_GROWN, 8)[0];             // Chosenscale(Coul)
    const double q = A.C.q[start + a]; const int type = A.C.type[start + a];
    const int j = o * cs + a;
    if(split == 0)
    {
      A.B.tr(0)[j] = x; A.B.tr(1)[j] = y; A.B.tr(2)[j] = z; A.B.tr(3)[j] = fx; A.B.tr(4)[j] = fy; A.B.tr(5)[j] = fz;
      A.B.tr(6)[j] = q; A.B.tr(7)[j] = scale; A.B.tr(8)[j] = scoul; A.B.tr_type()[j] = type;
    }
    T.fx[a] = fx; T.fy[a] = fy; T.fz[a] = fz; T.q[a] = q * scoul; T.scale[a] = scale; T.type[a] = type; T.slot[a] = 0;
  }
  __syncthreads();
  PairTables W; W.etab = etab; W.ffp = P.ffA; W.unit = false;
  if(cs == 1) cbmc_group_energy<1>(P, W, S, L, A, &T, Q, red, cs, o, split, nsplit);
  else if(cs == 2) cbmc_group_energy<2>(P, W, S, L, A, &T, Q, red, cs, o, split, nsplit);
  else cbmc_group_energy<0>(P, W, S, L, A, &T, Q, red, cs, o, split, nsplit);
  __syncthreads();
  if(threadIdx.x == 0) { __threadfence(); last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1); }
  __syncthreads();
  if(!last) return;
  __threadfence();
  cbmc_collect(A, nsplit);
  if(threadIdx.x < 32)
  {
    double* r = A.B.result(A.rslot);
    cbmc_finish_warp(P, A, true, r);
    if(threadIdx.x == 0)
    {
      r[14] = (A.dep_slot >= 0 ? A.B.result(A.dep_slot)[14] : 1.0) * r[0];   // CBMC.Rosenbluth *= averagedRosen, mc_widom.h:611
      r[13] = (r[9] != 0.0 && r[14] > 1e-150) ? 1.0 : 0.0;                  // mc_swap_utilities.h:32
      sel_s = (r[9] != 0.0 && r[11] > 0.0) ? (int) r[10] : -1;
      // block pockets: after the chain growth of an insertion / reinsertion / identity swap EVERY atom of the grown
      // molecule is tested (mc_swap_utilities.h:35-78, move_struct.h:208-250, mc_swap_moves.h:299-330); a blocked atom
      // fails the construction (Rosenbluth 0) -- the selection's random number stays consumed
      const int ty = A.cbmc_type;
      if(sel_s >= 0 && (ty == 0 || ty == 2 || ty == 4) && A.C.npocket > 0)
      {
        bool blk = blocked_pocket(P, A.C, A.B.mol(GBK_BUF_GROWN, 0)[0], A.B.mol(GBK_BUF_GROWN, 1)[0], A.B.mol(GBK_BUF_GROWN, 2)[0]);
        for(int a = 0; a < cs && !blk; a++) { const int j = sel_s * cs + a; blk = blocked_pocket(P, A.C, A.B.tr(0)[j], A.B.tr(1)[j], A.B.tr(2)[j]); }
        if(blk) { r[0] = 0.0; r[9] = 0.0; r[13] = 0.0; r[14] = 0.0; sel_s = -1; }
      }
      *A.ticket = 0u;
    }
  }
  __syncthreads();
  // selected orientation joins the first bead in the grown-molecule buffer
  if(sel_s >= 0 && threadIdx.x < cs)
  {
    const int j = sel_s * cs + threadIdx.x;
    for(int k = 0; k < 9; k++) A.B.mol(GBK_BUF_GROWN, k)[1 + threadIdx.x] = A.B.tr(k)[j];
    A.B.mol_type(GBK_BUF_GROWN)[1 + threadIdx.x] = A.B.tr_type()[j];
  }
}

// ---------------------------------------------------------------------------------------------
// translation / rotation proposal, get_new_position mc_utilities.h:485-606 (+ RotationAroundAxis :459-483)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rotate_about_axis(double& px, double& py, double& pz, double theta, double ax, double ay, double az)
{
  double s, c; sincos(theta, &s, &c);
  const double w = 1.0 - c;
  const double r00 = ax * ax * w + c,      r01 = ax * ay * w + az * s, r02 = ax * az * w - ay * s;
  const double r10 = ax * ay * w - az * s, r11 = ay * ay * w + c,      r12 = ay * az * w + ax * s;
  const double r20 = ax * az * w + ay * s, r21 = ay * az * w - ax * s, r22 = az * az * w + c;
  const double x = px * r00 + py * r01 + pz * r02;
  const double y = px * r10 + py * r11 + pz * r12;
  const double z = px * r20 + py * r21 + pz * r22;
  px = x; py = y; pz = z;
}

struct ProposeArgs
{
  int move_type, ms; long long start; long long pool_index;
  const double* __restrict__ pool3;
  double maxc[3];
  CompView C; MoveBufs B; int new_molid;
};

// one atom of a molecule in flight: Cartesian, fractional, charge, scaling factors, pseudo-atom type
struct AtomRec { double x, y, z, fx, fy, fz, q, scale, scoul; int type; };

__device__ __forceinline__ void store_atom(const MoveBufs& B, int buf, int i, const AtomRec& r)
{
  B.mol(buf, 0)[i] = r.x; B.mol(buf, 1)[i] = r.y; B.mol(buf, 2)[i] = r.z; B.mol(buf, 3)[i] = r.fx; B.mol(buf, 4)[i] = r.fy; B.mol(buf, 5)[i] = r.fz;
  B.mol(buf, 6)[i] = r.q; B.mol(buf, 7)[i] = r.scale; B.mol(buf, 8)[i] = r.scoul; B.mol_type(buf)[i] = r.type;
}

__device__ __forceinline__ void propose_atom(const DevParams& P, const ProposeArgs& A, int i, AtomRec& nw, AtomRec& od)
{
  const long long rp = A.start + i;
  const double x = A.C.x[rp], y = A.C.y[rp], z = A.C.z[rp];
  const double* R = A.pool3 + 3 * A.pool_index;
  double nx = x, ny = y, nz = z;
  const double x0 = A.C.x[A.start], y0 = A.C.y[A.start], z0 = A.C.z[A.start];
  switch(A.move_type)
  {
    case 0: // TRANSLATION
      nx = x + A.maxc[0] * 2.0 * (R[0] - 0.5); ny = y + A.maxc[1] * 2.0 * (R[1] - 0.5); nz = z + A.maxc[2] * 2.0 * (R[2] - 0.5);
      break;
    case 1: // ROTATION
    {
      nx = x - x0; ny = y - y0; nz = z - z0;
      const double ax = A.maxc[0] * 2.0 * (R[0] - 0.5), ay = A.maxc[1] * 2.0 * (R[1] - 0.5), az = A.maxc[2] * 2.0 * (R[2] - 0.5);
      rotate_about_axis(nx, ny, nz, ax, 1.0, 0.0, 0.0);
      rotate_about_axis(nx, ny, nz, ay, 0.0, 1.0, 0.0);
      rotate_about_axis(nx, ny, nz, az, 0.0, 0.0, 1.0);
      nx += x0; ny += y0; nz += z0;
      break;
    }
    case 2: // SINGLE_INSERTION
    {
      const double cx = P.cell[0] * R[0], cy = P.cell[4] * R[1], cz = P.cell[8] * R[2];
      if(i == 0) { nx = cx; ny = cy; nz = cz; }
      else { double vx = x - x0, vy = y - y0, vz = z - z0; rotate_quaternion(vx, vy, vz, R[3], R[4], R[5]); nx = vx + cx; ny = vy + cy; nz = vz + cz; }
      break;
    }
    case 3: // SINGLE_DELETION: falls through to SPECIAL_ROTATION in the reference (mc_utilities.h:561-571)
    case 4: // SPECIAL_ROTATION
    {
      nx = x - x0; ny = y - y0; nz = z - z0;
      const double ang = A.maxc[0] * 2.0 * (R[0] - 0.5);
      double ax = A.C.x[A.start + 1] - x0, ay = A.C.y[A.start + 1] - y0, az = A.C.z[A.start + 1] - z0;
      const double inv = 1.0 / sqrt(ax * ax + ay * ay + az * az);
      ax *= inv; ay *= inv; az *= inv;
      if(i != 0 && i != 1) rotate_about_axis(nx, ny, nz, 3.0 * ang, ax, ay, az);
      nx += x0; ny += y0; nz += z0;
      break;
    }
  }
  nw.q = od.q = A.C.q[rp]; nw.scale = od.scale = A.C.scale[rp]; nw.scoul = od.scoul = A.C.scoul[rp]; nw.type = od.type = A.C.type[rp];
  nw.x = nx; nw.y = ny; nw.z = nz; to_frac(P, nx, ny, nz, nw.fx, nw.fy, nw.fz);
  od.x = x; od.y = y; od.z = z; to_frac(P, x, y, z, od.fx, od.fy, od.fz);
}

// result(6)[0] = 1 when a block pocket contains an atom of the proposal: the move is then treated as an overlap
// (SingleBody_Prepare sets device_flag and returns early, mc_single_particle.h:83-119)
__global__ void k_single_propose(DevParams P, ProposeArgs A)
{
  bool blk = false;
  if((int) threadIdx.x < A.ms)
  {
    AtomRec nw, od;
    propose_atom(P, A, threadIdx.x, nw, od);
    store_atom(A.B, GBK_BUF_NEW, threadIdx.x, nw); store_atom(A.B, GBK_BUF_OLD, threadIdx.x, od);
    if(A.move_type != 3) blk = blocked_pocket(P, A.C, nw.x, nw.y, nw.z);          // Do_New moves only
  }
  const int any = __syncthreads_or(blk ? 1 : 0);
  if(threadIdx.x == 0) A.B.result(6)[0] = any ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------
// single-body delta: Calculate_Single_Body_Energy_VDWReal (VDW_Coulomb.cu:626-841) + host sum
// (mc_single_particle.h:183-200).  CTAs slice the atom ranges; each evaluates the NEW and the OLD molecule;
// the last CTA sums the CTA partials in fixed order: delta = sum(new) - sum(old).
// ---------------------------------------------------------------------------------------------
struct SingleBodyArgs { int comp, molid, ms, do_new, do_old; MoveBufs B; unsigned int* ticket; };

__global__ void __launch_bounds__(256)
k_single_body(DevParams P, SysView S, SegList L, SingleBodyArgs A)
{
  __shared__ TrialGroup T;
  __shared__ WarpQueue Q[8];
  __shared__ double red[8 * 16];
  __shared__ double etab[(GBK_ERFC_DEG + 1) * GBK_ERFC_NINT];
  __shared__ bool last;
  stage_erfc_table(P, etab);
  PairTables W; W.etab = etab; W.ffp = P.ffA; W.unit = false;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = (int) lane_id();
  double tot[14];
  for(int k = 0; k < 14; k++) tot[k] = 0.0;
  for(int pass = 0; pass < 2; pass++)
  {
    const bool run = pass == 0 ? A.do_new != 0 : A.do_old != 0;
    const int buf = pass == 0 ? GBK_BUF_NEW : GBK_BUF_OLD;
    __syncthreads();
    if(run && threadIdx.x < A.ms)
    {
      const int a = threadIdx.x;
      T.fx[a] = A.B.mol(buf, 3)[a]; T.fy[a] = A.B.mol(buf, 4)[a]; T.fz[a] = A.B.mol(buf, 5)[a];
      T.q[a] = A.B.mol(buf, 6)[a] * A.B.mol(buf, 8)[a]; T.scale[a] = A.B.mol(buf, 7)[a]; T.type[a] = A.B.mol_type(buf)[a]; T.slot[a] = 0;
    }
    __syncthreads();
    if(!run) continue;
    double e6[6] = {0, 0, 0, 0, 0, 0}; int flag = 0;
    pair_group_generic<0>(P, W, S, L, A.comp, A.molid, -1, -1, &T, A.ms, Q + warp, blockIdx.x * nwarps + warp, gridDim.x * nwarps, e6, flag);
#pragma unroll
    for(int k = 0; k < 6; k++) tot[pass * 7 + k] = warp_sum(e6[k]);
    tot[pass * 7 + 6] = __any_sync(0xffffffffu, flag) ? 1.0 : 0.0;
  }
  if(lane == 0) for(int k = 0; k < 14; k++) red[warp * 16 + k] = tot[k];
  __syncthreads();
  if(threadIdx.x == 0)
  {
    double s[14];
    for(int k = 0; k < 14; k++) { s[k] = 0.0; for(int w = 0; w < nwarps; w++) s[k] += red[w * 16 + k]; }
    for(int k = 0; k < 14; k++) A.B.partial()[blockIdx.x * 16 + k] = s[k];
    __threadfence();
    last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if(last && threadIdx.x == 0)
  {
    __threadfence();
    const volatile double* p = A.B.partial();
    double* r = A.B.result(4);
    for(int k = 0; k < 6; k++)
    {
      double n = 0.0, o = 0.0;
      for(unsigned int b = 0; b < gridDim.x; b++) { n += p[b * 16 + k]; o += p[b * 16 + 7 + k]; }
      r[k] = n - o;
    }
    double fl = 0.0;
    for(unsigned int b = 0; b < gridDim.x; b++) fl += p[b * 16 + 6];     // overlap of the NEW configuration only (:768-769)
    if(A.do_new && A.B.result(6)[0] != 0.0) fl += 1.0;                   // blocked by a block pocket (k_single_propose)
    r[6] = fl > 0.0 ? 1.0 : 0.0;
    r[7] = 1.0 - r[6];                                                   // "no overlap": dependency flag of the Ewald stage
    *A.ticket = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// state commit: update_translation_position / Update_insertion_data_Parallel / Update_deletion_data_Parallel /
// Update_Reinsertion_data (mc_utilities.h:118-292, mc_swap_moves.h:44-49)
// ---------------------------------------------------------------------------------------------
struct SlotArrays { double* x; double* y; double* z; double* fx; double* fy; double* fz; double* q; double* scale; double* scoul; int* type; int* molid; };

// copy n atoms of a molecule buffer into slots [dst, dst+n); what: 0 = positions only, 1 = pos+scale+charge+scaleCoul,
// 2 = everything incl. type and MolID (= molid)
__global__ void k_commit_from_buffer(SlotArrays S, MoveBufs B, int buf, int dst, int n, int what, int molid)
{
  const int i = threadIdx.x;
  if(i >= n) return;
  S.x[dst + i] = B.mol(buf, 0)[i]; S.y[dst + i] = B.mol(buf, 1)[i]; S.z[dst + i] = B.mol(buf, 2)[i];
  S.fx[dst + i] = B.mol(buf, 3)[i]; S.fy[dst + i] = B.mol(buf, 4)[i]; S.fz[dst + i] = B.mol(buf, 5)[i];
  if(what >= 1) { S.q[dst + i] = B.mol(buf, 6)[i]; S.scale[dst + i] = B.mol(buf, 7)[i]; S.scoul[dst + i] = B.mol(buf, 8)[i]; }
  if(what >= 2) { S.type[dst + i] = B.mol_type(buf)[i]; S.molid[dst + i] = molid; }
}

// deletion: the last molecule takes the place of the deleted one; MolID of the slot is kept (mc_utilities.h:163-189)
__global__ void k_commit_delete(SlotArrays S, int dst, int src, int n)
{
  const int i = threadIdx.x;
  if(i >= n || dst == src) return;
  S.x[dst + i] = S.x[src + i]; S.y[dst + i] = S.y[src + i]; S.z[dst + i] = S.z[src + i];
  S.fx[dst + i] = S.fx[src + i]; S.fy[dst + i] = S.fy[src + i]; S.fz[dst + i] = S.fz[src + i];
  S.q[dst + i] = S.q[src + i]; S.scale[dst + i] = S.scale[src + i]; S.scoul[dst + i] = S.scoul[src + i]; S.type[dst + i] = S.type[src + i];
}

// exchange two molecules' slots (Update_deletion_data_fractional / Revert_CBCF_Deletion, mc_cbcfc.h:133-214: the fractional molecule
// that is provisionally deleted is parked in the last slot so that a rejection can put it back); MolIDs stay where they are
__global__ void k_commit_swap(SlotArrays S, int a, int b, int n)
{
  const int i = threadIdx.x;
  if(i >= n || a == b) return;
  double t;
  t = S.x[a + i]; S.x[a + i] = S.x[b + i]; S.x[b + i] = t;  t = S.y[a + i]; S.y[a + i] = S.y[b + i]; S.y[b + i] = t;  t = S.z[a + i]; S.z[a + i] = S.z[b + i]; S.z[b + i] = t;
  t = S.fx[a + i]; S.fx[a + i] = S.fx[b + i]; S.fx[b + i] = t;  t = S.fy[a + i]; S.fy[a + i] = S.fy[b + i]; S.fy[b + i] = t;  t = S.fz[a + i]; S.fz[a + i] = S.fz[b + i]; S.fz[b + i] = t;
  t = S.q[a + i]; S.q[a + i] = S.q[b + i]; S.q[b + i] = t;  t = S.scale[a + i]; S.scale[a + i] = S.scale[b + i]; S.scale[b + i] = t;
  t = S.scoul[a + i]; S.scoul[a + i] = S.scoul[b + i]; S.scoul[b + i] = t;
  const int ty = S.type[a + i]; S.type[a + i] = S.type[b + i]; S.type[b + i] = ty;
}

// buffer <- buffer / buffer <- slots copies (StoreNewLocation_Reinsertion mc_swap_moves.h:27-41 and Ewald gathers)
__global__ void k_copy_buffer(MoveBufs B, int dst_buf, int src_buf, int n)
{
  const int i = threadIdx.x;
  if(i >= n) return;
  for(int k = 0; k < 9; k++) B.mol(dst_buf, k)[i] = B.mol(src_buf, k)[i];
  B.mol_type(dst_buf)[i] = B.mol_type(src_buf)[i];
}
// overwrite the two scaling factors of a molecule buffer (the NEW side of a lambda change: same atoms, new lambda)
__global__ void k_set_buffer_scale(MoveBufs B, int buf, int n, double scale, double scoul)
{
  const int i = threadIdx.x;
  if(i >= n) return;
  B.mol(buf, 7)[i] = scale; B.mol(buf, 8)[i] = scoul;
}
// commit of an accepted lambda change: the molecule's slots take the new scaling factors (mc_cbcfc.h acceptance)
__global__ void k_commit_scale(SlotArrays S, int dst, int n, double scale, double scoul)
{
  const int i = threadIdx.x;
  if(i >= n) return;
  S.scale[dst + i] = scale; S.scoul[dst + i] = scoul;
}
__global__ void k_load_buffer(DevParams P, MoveBufs B, int dst_buf, CompView C, long long start, int n, const double* pos3_override)
{
  const int i = threadIdx.x;
  if(i >= n) return;
  double x = C.x[start + i], y = C.y[start + i], z = C.z[start + i];
  if(pos3_override) { x = pos3_override[3 * i]; y = pos3_override[3 * i + 1]; z = pos3_override[3 * i + 2]; }
  double fx, fy, fz; to_frac(P, x, y, z, fx, fy, fz);
  B.mol(dst_buf, 0)[i] = x; B.mol(dst_buf, 1)[i] = y; B.mol(dst_buf, 2)[i] = z; B.mol(dst_buf, 3)[i] = fx; B.mol(dst_buf, 4)[i] = fy; B.mol(dst_buf, 5)[i] = fz;
  B.mol(dst_buf, 6)[i] = C.q[start + i]; B.mol(dst_buf, 7)[i] = C.scale[start + i]; B.mol(dst_buf, 8)[i] = C.scoul[start + i];
  B.mol_type(dst_buf)[i] = C.type[start + i];
}

// Ewald gather: [old atoms | new atoms] -> pos3 / qeff in the layout k_ewald_delta reads (Initialize_Copy_Positions_Together,
// Ewald_Energy_Functions.h:107-160).  Sources are molecule buffers (buf >= 0) .
__global__ void k_ewald_gather(MoveBufs B, int old_buf, int nold, int new_buf, int nnew, double* pos3, double* qeff)
{
  const int i = threadIdx.x;
  if(i >= nold + nnew) return;
  const int buf = i < nold ? old_buf : new_buf;
  const int j = i < nold ? i : i - nold;
  pos3[3 * i] = B.mol(buf, 0)[j]; pos3[3 * i + 1] = B.mol(buf, 1)[j]; pos3[3 * i + 2] = B.mol(buf, 2)[j];
  qeff[i] = B.mol(buf, 8)[j] * B.mol(buf, 6)[j];
}
