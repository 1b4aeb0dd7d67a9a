// graspa_b200 -- pair-energy kernels (included by engine.cu after misc_kernels.cuh).
#pragma once
#include "common.cuh"
#include "pair.cuh"

// ---------------------------------------------------------------------------------------------
// generic trial-group energies: one CTA per trial group, warps split the atom ranges, fixed-order reduce.
// Replaces Calculate_Multiple_Trial_Energy_VDWReal + Host_sum_Widom_HGGG_SEPARATE for caller-supplied trials
// and is the energy step of the single-move CBMC stages.
// ---------------------------------------------------------------------------------------------
struct TrialBuf            // device arrays of trial atoms (Sims.New), group-major
{
  const double* __restrict__ fx; const double* __restrict__ fy; const double* __restrict__ fz;
  const double* __restrict__ q;      // charge * scaleCoul
  const double* __restrict__ scale;
  const int*    __restrict__ type;
};

template <int CS>
__device__ __forceinline__ void group_energy_cta(const DevParams& P, const PairTables& W, const SysView& S, const SegList& L, const TrialBuf& B,
                                                 int group, int cs, int new_comp, int new_molid, int excl_comp, int excl_mol,
                                                 TrialGroup* T, WarpQueue* Qall, double* red /* [nwarps][8] */, double* out6, int* out_flag)
{
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = lane_id();
  if(threadIdx.x < cs)
  {
    const int j = group * cs + threadIdx.x;
    T->fx[threadIdx.x] = B.fx[j]; T->fy[threadIdx.x] = B.fy[j]; T->fz[threadIdx.x] = B.fz[j];
    T->q[threadIdx.x] = B.q[j]; T->scale[threadIdx.x] = B.scale[j]; T->type[threadIdx.x] = B.type[j]; T->slot[threadIdx.x] = 0;
  }
  __syncthreads();
  double e6[6] = {0, 0, 0, 0, 0, 0}; int flag = 0;
  pair_group_generic<CS>(P, W, S, L, new_comp, new_molid, excl_comp, excl_mol, T, cs, Qall + warp, warp, nwarps, e6, flag);
#pragma unroll
  for(int k = 0; k < 6; k++) e6[k] = warp_sum(e6[k]);
  flag = __any_sync(0xffffffffu, flag);
  if(lane == 0) { for(int k = 0; k < 6; k++) red[warp * 8 + k] = e6[k]; red[warp * 8 + 6] = flag ? 1.0 : 0.0; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for(int w = 0; w < nwarps; w++) for(int k = 0; k < 7; k++) s[k] += red[w * 8 + k];
    for(int k = 0; k < 6; k++) out6[k] = s[k];
    *out_flag = s[6] > 0.0 ? 1 : 0;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
k_trial_energies(DevParams P, SysView S, SegList L, TrialBuf B, int cs, int new_comp, int new_molid, int excl_comp, int excl_mol,
                 double* out6 /* [ngroups][6] */, int* out_flag)
{
  __shared__ TrialGroup T;
  __shared__ WarpQueue Q[8];
  __shared__ double red[8 * 8];
  __shared__ double etab[(GBK_ERFC_DEG + 1) * GBK_ERFC_NINT];
  stage_erfc_table(P, etab);
  __syncthreads();
  PairTables W; W.etab = etab; W.ffp = P.ffA; W.unit = false;
  const int g = blockIdx.x;
  if(cs == 1)      group_energy_cta<1>(P, W, S, L, B, g, cs, new_comp, new_molid, excl_comp, excl_mol, &T, Q, red, out6 + 6 * g, out_flag + g);
  else if(cs == 2) group_energy_cta<2>(P, W, S, L, B, g, cs, new_comp, new_molid, excl_comp, excl_mol, &T, Q, red, out6 + 6 * g, out_flag + g);
  else             group_energy_cta<0>(P, W, S, L, B, g, cs, new_comp, new_molid, excl_comp, excl_mol, &T, Q, red, out6 + 6 * g, out_flag + g);
}

// ---------------------------------------------------------------------------------------------
// total VDW + real: every live atom is a one-atom "trial group" against all live atoms with the own-molecule
// exclusion; 0.5 x the double-counted sum, exactly the reference's CPU loop structure (VDW_Coulomb.cu:94-206).
// ---------------------------------------------------------------------------------------------
struct TotalArgs { SegList L; int nhost; double* out; /* [natoms_live][6] */ int* flag; /* NULL or one word: any overlapping pair */ };

__global__ void __launch_bounds__(128)
k_total_vdw_real(DevParams P, SysView S, TotalArgs A)
{
  __shared__ TrialGroup T;
  __shared__ WarpQueue Q[4];
  __shared__ double red[4 * 8];
  __shared__ double etab[(GBK_ERFC_DEG + 1) * GBK_ERFC_NINT];
  __shared__ SegList L;
  stage_erfc_table(P, etab);
  int g = blockIdx.x, seg = 0;
  while(seg < A.L.nseg && g >= A.L.count[seg]) { g -= A.L.count[seg]; seg++; }
  const int i = A.L.start[seg] + g;
  const int mycomp = A.L.comp[seg];
  const bool mine_host = mycomp < A.nhost;
  if(threadIdx.x == 0)
  {
    T.fx[0] = S.fx[i]; T.fy[0] = S.fy[i]; T.fz[0] = S.fz[i];
    T.q[0] = S.q[i] * S.scoul[i]; T.scale[0] = S.scale[i]; T.type[0] = S.type[i]; T.slot[0] = 0;
    L = A.L;
    // kinds relative to the trial atom: host-host 0, mixed 1, guest-guest 2
    for(int s = 0; s < L.nseg; s++) { const bool oh = L.comp[s] < A.nhost; L.kind[s] = (mine_host && oh) ? 0 : ((!mine_host && !oh) ? 2 : 1); L.staged[s] = 0; }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double e6[6] = {0, 0, 0, 0, 0, 0}; int flag = 0;
  PairTables W; W.etab = etab; W.ffp = P.ffA; W.unit = false;
  pair_group_generic<1>(P, W, S, L, mycomp, S.molid[i], -1, -1, &T, 1, Q + warp, warp, nwarps, e6, flag);
#pragma unroll
  for(int k = 0; k < 6; k++) e6[k] = warp_sum(e6[k]);
  if(lane_id() == 0) for(int k = 0; k < 6; k++) red[warp * 8 + k] = e6[k];
  __syncthreads();
  if(threadIdx.x < 6)
  {
    double s = 0.0;
    for(int w = 0; w < nwarps; w++) s += red[w * 8 + threadIdx.x];
    A.out[(size_t) blockIdx.x * 6 + threadIdx.x] = 0.5 * s;
  }
  if(A.flag && __any_sync(0xffffffffu, flag) && lane_id() == 0) atomicOr(A.flag, 1);
}

// ---------------------------------------------------------------------------------------------
// Batched Widom, stage A: first bead + chain growth for n independent ghost insertions.
// One warp = one insertion at a time (static stride over insertions -> deterministic results);
// the framework pack is staged once per CTA by TMA bulk copies and stays resident in shared memory.
// Follows Widom_Move_FirstBead_PARTIAL / Widom_Move_Chain_PARTIAL (mc_widom.h:385-614) with MoveType CBMC_INSERTION.
// ---------------------------------------------------------------------------------------------
struct WidomA
{
  const double* __restrict__ pool3;           // double3 pool
  const long long* __restrict__ fb_index; const long long* __restrict__ or_index;
  const double* __restrict__ uni;             // 2 per insertion
  long long n;
  int ntrials, norient, ms, comp, new_molid;
  // template molecule = slot 0 of the component (mc_widom.h:256): Cartesian positions, charge, scaleCoul, type
  const double* __restrict__ tx; const double* __restrict__ ty; const double* __restrict__ tz;
  const double* __restrict__ tq; const double* __restrict__ tscoul; const int* __restrict__ ttype;
  const double* __restrict__ pack; int npad, ntp, pack_n; int use_pack; int stage_ff;
  int first_bead_only;                         // 1: write the first-bead success code into stage[] and stop there
  double* rec;        // per insertion: [W12, HGv, HGr, GGv, GGr, x0,y0,z0, x1,...]  stride 5 + 3*ms
  int* stage;         // 0 ok, 1 first bead failed, 2 chain failed
};

// per-warp scratch of the Widom pair kernel: trial group, queue, chain coordinates (fractional + Cartesian), first-bead Cartesian
__host__ __device__ inline size_t widom_per_warp_bytes(int norient, int cs)
{
  return (sizeof(TrialGroup) + sizeof(WarpQueue) + (size_t) norient * (cs > 0 ? cs : 1) * 6 * sizeof(double) + 4 * sizeof(double) + 15) / 16 * 16;
}

// per-warp context of the Widom pair kernel
struct WidomCtx
{
  PairTables W;
  SysView Sg;        // global slot arrays
  TileView V;        // staged, tile-sorted pack of the host components (shared memory); V.pack == nullptr: not staged
  const SegList* L;  // in shared memory
  TrialGroup* T; WarpQueue* Q;
};

// energies of the group currently in *T (Cartesian copies in tc[cs][3]) against every segment: out = {HGv, HGr, GGv, GGr}
template <int CS, int CELL>
__device__ __forceinline__ void widom_group(const DevParams& P, const WidomCtx& X, int comp, int new_molid, int cs_dyn, const double* tc,
                                            double out[4], int& flag)
{
  const SegList& L = *X.L;
  double part[4] = {0.0, 0.0, 0.0, 0.0};
  int fl = 0;
  const int cs = CS > 0 ? CS : cs_dyn;
  if(X.V.pack != nullptr)
  {
    // all staged segments are host components (kind HG) and live in ONE tile-sorted pack
    PairAcc<1> acc; acc.clear();
    pair_tiles_group<CELL>(P, X.W, X.V, X.T, cs, tc, X.Q, acc);
    part[0] += acc.vdw[0]; part[1] += acc.real[0]; fl |= acc.flag;
  }
  const int nseg = L.nseg;
  for(int g = 0; g < nseg; g++)
  {
    if(L.staged[g]) continue;
    PairAcc<1> acc; acc.clear();
    const int start = L.start[g], end = start + L.count[g];
    const bool gg = L.kind[g] == 2;
    const SysAccess<false> S = make_access<false>(X.Sg);
    const int eb = (L.comp[g] == comp) ? new_molid : -1;
    pair_range<CS, 1, CELL, false, true>(P, X.W, S, start, end, -1, eb, X.T, cs_dyn, X.Q, 0, 1, acc);
    part[0] += gg ? 0.0 : acc.vdw[0]; part[1] += gg ? 0.0 : acc.real[0];
    part[2] += gg ? acc.vdw[0] : 0.0; part[3] += gg ? acc.real[0] : 0.0;
    fl |= acc.flag;
  }
#pragma unroll
  for(int k = 0; k < 4; k++) out[k] = warp_sum(part[k]);
  flag = __any_sync(0xffffffffu, fl & 1) ? 1 : 0;
}

template <int CELL>
__global__ void __launch_bounds__(512, 1)
k_widom_pair(DevParams P, SysView Sg, SegList Lin, WidomA A)
{
  extern __shared__ __align__(16) unsigned char smem[];
  // layout: [mbarrier 16 B][erfc table][LJ table (if it fits)][SegList][pack][per-warp: TrialGroup, WarpQueue, chain_f, chain_c]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  double* etab = reinterpret_cast<double*>(smem + GBK_SMEM_TABLES_OFF);
  double4* fftab = reinterpret_cast<double4*>(smem + GBK_SMEM_TABLES_OFF + GBK_ERFC_BYTES_PAD);
  const size_t ff_bytes = A.stage_ff ? (size_t) P.ntypes * P.ntypes * sizeof(double4) : 0;
  SegList* L = reinterpret_cast<SegList*>(smem + GBK_SMEM_TABLES_OFF + GBK_ERFC_BYTES_PAD + ff_bytes);
  const size_t head = (GBK_SMEM_TABLES_OFF + GBK_ERFC_BYTES_PAD + ff_bytes + sizeof(SegList) + 15) / 16 * 16;
  double* pack = reinterpret_cast<double*>(smem + head);
  const size_t pack_bytes = A.use_pack ? gbk_pack_bytes(A.npad, A.ntp) : 0;
  unsigned char* wbase = smem + head + pack_bytes;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = lane_id();
  const int cs = A.ms - 1;
  const size_t per_warp = widom_per_warp_bytes(A.norient, cs);
  TrialGroup* T = reinterpret_cast<TrialGroup*>(wbase + warp * per_warp);
  WarpQueue* Q = reinterpret_cast<WarpQueue*>(reinterpret_cast<unsigned char*>(T) + sizeof(TrialGroup));
  double* chain_f = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(Q) + sizeof(WarpQueue));   // fractional [norient][cs][3]
  double* chain_c = chain_f + (size_t) A.norient * (cs > 0 ? cs : 1) * 3;                                  // Cartesian
  double* fb_c = chain_c + (size_t) A.norient * (cs > 0 ? cs : 1) * 3;                                     // Cartesian first-bead trial

  stage_erfc_table(P, etab);
  if(A.stage_ff) for(int i = threadIdx.x; i < P.ntypes * P.ntypes; i += blockDim.x) fftab[i] = P.ffA[i];
  if(threadIdx.x == 0) *L = Lin;
  WidomCtx X; X.W.etab = etab; X.W.ffp = A.stage_ff ? fftab : P.ffA; X.W.unit = P.all_unit_scale != 0; X.Sg = Sg; X.L = L; X.T = T; X.Q = Q;
  X.V.pack = nullptr; X.V.tile = nullptr; X.V.npad = 0; X.V.ntp = 0; X.V.n = 0; X.V.ntiles = 0;
  if(A.use_pack)
  {
    stage_bulk(pack, A.pack, (uint32_t) pack_bytes, bar);
    X.V.pack = pack; X.V.tile = reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(pack) + (size_t) A.npad * 36);
    X.V.npad = A.npad; X.V.ntp = A.ntp; X.V.n = A.pack_n; X.V.ntiles = A.npad / 32;
  }
  __syncthreads();
  const long long gw = (long long) blockIdx.x * nwarps + warp, tw = (long long) gridDim.x * nwarps;
  const int rec_stride = 5 + 3 * A.ms;
  const double q0 = A.tq[0] * A.tscoul[0]; const int type0 = A.ttype[0];

  for(long long ins = gw; ins < A.n; ins += tw)
  {
    const long long fb_off = A.fb_index ? A.fb_index[ins] : ins * (A.ntrials + A.norient);
    const long long or_off = A.or_index ? A.or_index[ins] : fb_off + A.ntrials;
    // ---------------- first bead: BoxLength o random, mc_widom.h:137-138
    double px = 0, py = 0, pz = 0, fsx = 0, fsy = 0, fsz = 0;
    if(lane < A.ntrials)
    {
      const double* r = A.pool3 + 3 * (fb_off + lane);
      px = P.cell[0] * r[0]; py = P.cell[4] * r[1]; pz = P.cell[8] * r[2];
      to_frac(P, px, py, pz, fsx, fsy, fsz);
    }
    double my_e[4] = {0, 0, 0, 0}; int my_flag = 0;
    // one tile-culled pass per first-bead trial
    for(int t = 0; t < A.ntrials; t++)
    {
      const double bx = __shfl_sync(0xffffffffu, fsx, t), by = __shfl_sync(0xffffffffu, fsy, t), bz = __shfl_sync(0xffffffffu, fsz, t);
      const double cx = __shfl_sync(0xffffffffu, px, t), cy = __shfl_sync(0xffffffffu, py, t), cz = __shfl_sync(0xffffffffu, pz, t);
      if(lane == 0)
      {
        T->fx[0] = bx; T->fy[0] = by; T->fz[0] = bz; T->q[0] = q0; T->scale[0] = 1.0; T->type[0] = type0; T->slot[0] = 0;
        fb_c[0] = cx; fb_c[1] = cy; fb_c[2] = cz;
      }
      __syncwarp();
      double e[4]; int fl;
      widom_group<0, CELL>(P, X, A.comp, A.new_molid, 1, fb_c, e, fl);
      if(lane == t) { my_e[0] = e[0]; my_e[1] = e[1]; my_e[2] = e[2]; my_e[3] = e[3]; my_flag = fl; }
      __syncwarp();
    }
    double tot = my_e[0] + my_e[2]; if(P.vdw_real_bias) tot += my_e[1] + my_e[3];
    RosenResult r1 = rosenbluth_warp(-P.beta * tot, !my_flag, A.ntrials, A.uni ? A.uni[2 * ins] : 0.5, true);
    double W = 0.0; int ok = r1.success && !(r1.R < 1e-150);
    const int sfb = r1.sel_lane;
    double efb[4];
#pragma unroll
    for(int k = 0; k < 4; k++) efb[k] = __shfl_sync(0xffffffffu, my_e[k], sfb);
    if(ok)
    {
      W = r1.R / (double) A.ntrials;
      if(!P.vdw_real_bias) W *= exp(-P.beta * (efb[1] + efb[3]));
      if(W <= 1e-150) ok = 0;
    }
    if(A.first_bead_only)
    {
      // code for the host's random-stream walk: 1 success, 0 failed with survivors, 2 no survivor
      if(lane == 0) A.stage[ins] = ok ? 1 : (r1.nsurv > 0 ? 0 : 2);
      continue;
    }
    double* rec = A.rec + (size_t) ins * rec_stride;
    if(!ok) { if(lane == 0) { A.stage[ins] = 1; rec[0] = 0.0; } continue; }
    const double fbx = __shfl_sync(0xffffffffu, px, sfb), fby = __shfl_sync(0xffffffffu, py, sfb), fbz = __shfl_sync(0xffffffffu, pz, sfb);
    double ech[4] = {0, 0, 0, 0}; int so = 0;
    // ---------------- chain: mc_widom.h:509-614
    if(cs > 0)
    {
      if(lane < A.norient)
      {
        const double* r = A.pool3 + 3 * (or_off + lane);
        for(int a = 0; a < cs; a++)
        {
          double vx = A.tx[1 + a] - A.tx[0], vy = A.ty[1 + a] - A.ty[0], vz = A.tz[1 + a] - A.tz[0];
          rotate_quaternion(vx, vy, vz, r[0], r[1], r[2]);
          const double cx = fbx + vx, cy = fby + vy, cz = fbz + vz;
          double* cc = chain_c + (size_t)(lane * cs + a) * 3; cc[0] = cx; cc[1] = cy; cc[2] = cz;
          double* cf = chain_f + (size_t)(lane * cs + a) * 3; to_frac(P, cx, cy, cz, cf[0], cf[1], cf[2]);
        }
      }
      __syncwarp();
      my_e[0] = my_e[1] = my_e[2] = my_e[3] = 0.0; my_flag = 0;
      for(int o = 0; o < A.norient; o++)
      {
        if(lane < cs)
        {
          const double* c = chain_f + (size_t)(o * cs + lane) * 3;
          T->fx[lane] = c[0]; T->fy[lane] = c[1]; T->fz[lane] = c[2];
          T->q[lane] = A.tq[1 + lane] * A.tscoul[1 + lane]; T->scale[lane] = 1.0; T->type[lane] = A.ttype[1 + lane]; T->slot[lane] = 0;
        }
        __syncwarp();
        double e[4]; int fl;
        const double* tc = chain_c + (size_t) o * cs * 3;
        // one instantiation for every chain size: the tile loop takes the atoms one by one anyway, and five inlined copies
        // of the pair loops (17 400 SASS instructions in all) cost more in instruction-cache misses (stall_no_instruction
        // 0.34 -> 0.26 per issue) than the register-resident trial atoms of the CS-specialised variants saved
        widom_group<0, CELL>(P, X, A.comp, A.new_molid, cs, tc, e, fl);
        if(lane == o) { my_e[0] = e[0]; my_e[1] = e[1]; my_e[2] = e[2]; my_e[3] = e[3]; my_flag = fl; }
        __syncwarp();
      }
      double tot2 = my_e[0] + my_e[2]; if(P.vdw_real_bias) tot2 += my_e[1] + my_e[3];
      RosenResult r2 = rosenbluth_warp(-P.beta * tot2, !my_flag, A.norient, A.uni[2 * ins + 1], true);
      int ok2 = r2.success && !(r2.R < 1e-150);
      so = r2.sel_lane;
#pragma unroll
      for(int k = 0; k < 4; k++) ech[k] = __shfl_sync(0xffffffffu, my_e[k], so);
      if(ok2)
      {
        double W2 = r2.R / (double) A.norient;
        if(!P.vdw_real_bias) W2 *= exp(-P.beta * (ech[1] + ech[3]));
        W *= W2;
        if(W <= 1e-150) ok2 = 0;
      }
      if(!ok2) { if(lane == 0) { A.stage[ins] = (r2.nsurv > 0) ? 2 : 3; rec[0] = 0.0; } continue; }
    }
    if(lane == 0)
    {
      A.stage[ins] = 0;
      rec[0] = W;
      for(int k = 0; k < 4; k++) rec[1 + k] = efb[k] + ech[k];
      rec[5] = fbx; rec[6] = fby; rec[7] = fbz;
    }
    for(int k = lane; k < 3 * cs; k += 32) rec[8 + k] = chain_c[(size_t) so * cs * 3 + k];
    __syncwarp();
  }
}
