// graspa_b200 -- ONE kernel per Monte Carlo move.
//
// A single GCMC move is a chain of small dependent stages (first bead -> selection -> chain growth -> selection ->
// Ewald delta).  The reference runs each stage as its own launch followed by cudaDeviceSynchronize and a host-side sum
// (3-6 host round trips per move); what bounds GCMC cycles/s is therefore latency, not arithmetic.  Here the whole move
// is one launch of a small co-resident grid: the stages are separated by a device-wide barrier (all CTAs are resident,
// grid <= number of SMs), selections run on the device (last stage's results stay in L2), and the host reads one 1 KB
// result block.  Covers Insertion_Body / Deletion_Body (mc_swap_utilities.h:3-225), the reinsertion growth + retrace
// (move_struct.h:186-338) and SingleBody_Prepare + SingleBody_Calculation (mc_single_particle.h:10-241).
#pragma once
#include "common.cuh"
#include "pair.cuh"
#include "ewald.cuh"
#include "move_kernels.cuh"

enum { GBF_INSERTION = 0, GBF_DELETION = 1, GBF_REINSERTION = 2, GBF_SINGLE = 3 };

struct FusedArgs
{
  int kind, comp, ms, move_type;          // move_type: GB_TRANSLATION / GB_ROTATION / GB_SPECIAL_ROTATION for GBF_SINGLE
  long long molecule, pool_off;
  double u0, u1, scale0, scale1, maxc[3];
  int ntrials, norient, nmol;             // nmol = NumberOfMolecule_for_Component (MolID of an inserted molecule)
  int do_ewald, check_overlap, framework_moved;
  const double* __restrict__ pool3;
  CompView C; MoveBufs B;
  SegList L;                              // live ranges with the kinds of THIS move
  KTable K; const double* same_sf; const double* cross_sf; double* temp_sf;
  unsigned int* bar;                      // device-wide barrier counter, left at 0 by the kernel
};

// all CTAs of the grid are co-resident (the host launches at most one CTA per SM)
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& phase)
{
  __syncthreads();
  if(threadIdx.x == 0)
  {
    phase++;
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = phase * gridDim.x;
    while(atomicAdd(counter, 0u) < target) { }
    __threadfence();
  }
  __syncthreads();
}

struct FusedSmem
{
  TrialGroup T;
  WarpQueue Q[8];
  double red[8 * 16];
  double etab[(GBK_ERFC_DEG + 1) * GBK_ERFC_NINT];
};

__device__ __forceinline__ int fused_nsplit(const SegList& L, int ngroups)
{
  int natoms = 0; for(int s = 0; s < L.nseg; s++) natoms += L.count[s];
  int ns = (natoms + 511) / 512;
  const int cap = max(1, min(4096 / (8 * max(ngroups, 1)), (int) (2 * gridDim.x) / max(ngroups, 1)));
  return max(1, min(ns, cap));
}

// one CBMC stage (first bead or chain) of growth type `type`
__device__ __forceinline__ void fused_cbmc_stage(const DevParams& P, const SysView& S, const FusedArgs& F, FusedSmem* sm, const PairTables& W,
                                                 bool chain, int type, long long pool_off, double uniform, int rslot, int dep_slot, int stored_slot,
                                                 unsigned int& phase)
{
  CbmcArgs A; memset(&A, 0, sizeof(A));
  A.cbmc_type = type; A.comp = F.comp; A.ms = F.ms; A.molecule = F.molecule;
  if(chain) { A.ntrials = F.norient; A.norm = F.norient; }
  else { A.ntrials = (type == 3 || type == 4 || type == 5) ? 1 : F.ntrials; A.norm = F.ntrials; }
  const bool insertion_like = (type == 0 || type == 4);
  A.new_molid = insertion_like ? F.nmol : (int) F.molecule;
  A.excl_comp = -1; A.excl_mol = -1;
  A.pool_off = pool_off; A.pool3 = F.pool3; A.uniform = uniform; A.scale = F.scale0; A.scale_coul = F.scale1;
  A.C = F.C; A.B = F.B; A.rslot = rslot; A.dep_slot = dep_slot; A.stored_slot = stored_slot;
  const bool run = dep_slot < 0 || F.B.result(dep_slot)[13] != 0.0;
  const int cs = chain ? F.ms - 1 : 1;
  const int nsplit = fused_nsplit(F.L, A.ntrials);
  if(run)
  {
    for(int w = blockIdx.x; w < A.ntrials * nsplit; w += gridDim.x)
    {
      const int g = w / nsplit, split = w % nsplit;
      __syncthreads();
      if(!chain)
      {
        if(threadIdx.x == 0)
        {
          const long long start = insertion_like ? 0 : A.molecule * A.ms;
          double scale = A.scale, scoul = A.scale_coul;
          if(!insertion_like) { scale = A.C.scale[start]; scoul = A.C.scoul[start]; }
          double x, y, z;
          const bool existing = (type == 1 || type == 3 || type == 5) && g == 0;
          if(existing) { x = A.C.x[start]; y = A.C.y[start]; z = A.C.z[start]; }
          else { const double* r = A.pool3 + 3 * (A.pool_off + g); x = P.cell[0] * r[0]; y = P.cell[4] * r[1]; z = P.cell[8] * r[2]; }
          double fx, fy, fz; to_frac(P, x, y, z, fx, fy, fz);
          const double q = A.C.q[start]; const int ty = A.C.type[start];
          if(split == 0)
          {
            A.B.tr(0)[g] = x; A.B.tr(1)[g] = y; A.B.tr(2)[g] = z; A.B.tr(3)[g] = fx; A.B.tr(4)[g] = fy; A.B.tr(5)[g] = fz;
            A.B.tr(6)[g] = q; A.B.tr(7)[g] = scale; A.B.tr(8)[g] = scoul; A.B.tr_type()[g] = ty;
          }
          sm->T.fx[0] = fx; sm->T.fy[0] = fy; sm->T.fz[0] = fz; sm->T.q[0] = q * scoul; sm->T.scale[0] = scale; sm->T.type[0] = ty; sm->T.slot[0] = 0;
        }
      }
      else if(threadIdx.x < cs)
      {
        const int a = threadIdx.x;
        const long long start = (insertion_like ? 0 : A.molecule * A.ms) + 1;
        const double fbx = A.B.mol(GBK_BUF_GROWN, 0)[0], fby = A.B.mol(GBK_BUF_GROWN, 1)[0], fbz = A.B.mol(GBK_BUF_GROWN, 2)[0];
        double vx = A.C.x[1 + a] - A.C.x[0], vy = A.C.y[1 + a] - A.C.y[0], vz = A.C.z[1 + a] - A.C.z[0];
        double x, y, z;
        if((type == 1 || type == 3 || type == 5) && g == 0) { x = A.C.x[start + a]; y = A.C.y[start + a]; z = A.C.z[start + a]; }
        else
        {
          const double* r = A.pool3 + 3 * (A.pool_off + g);
          rotate_quaternion(vx, vy, vz, r[0], r[1], r[2]);
          x = fbx + vx; y = fby + vy; z = fbz + vz;
        }
        double fx, fy, fz; to_frac(P, x, y, z, fx, fy, fz);
        const double scale = A.B.mol(GBK_BUF_GROWN, 7)[0], scoul = A.B.mol(GBK_BUF_GROWN, 8)[0];
        const double q = A.C.q[start + a]; const int ty = A.C.type[start + a];
        const int j = g * cs + a;
        if(split == 0)
        {
          A.B.tr(0)[j] = x; A.B.tr(1)[j] = y; A.B.tr(2)[j] = z; A.B.tr(3)[j] = fx; A.B.tr(4)[j] = fy; A.B.tr(5)[j] = fz;
          A.B.tr(6)[j] = q; A.B.tr(7)[j] = scale; A.B.tr(8)[j] = scoul; A.B.tr_type()[j] = ty;
        }
        sm->T.fx[a] = fx; sm->T.fy[a] = fy; sm->T.fz[a] = fz; sm->T.q[a] = q * scoul; sm->T.scale[a] = scale; sm->T.type[a] = ty; sm->T.slot[a] = 0;
      }
      __syncthreads();
      if(cs == 1)      cbmc_group_energy<1>(P, W, S, F.L, A, &sm->T, sm->Q, sm->red, cs, g, split, nsplit);
      else if(cs == 2) cbmc_group_energy<2>(P, W, S, F.L, A, &sm->T, sm->Q, sm->red, cs, g, split, nsplit);
      else             cbmc_group_energy<0>(P, W, S, F.L, A, &sm->T, sm->Q, sm->red, cs, g, split, nsplit);
    }
  }
  grid_barrier(F.bar, phase);
  if(blockIdx.x == 0)
  {
    double* r = A.B.result(rslot);
    if(!run) { if(threadIdx.x < 16) r[threadIdx.x] = 0.0; }
    else
    {
      cbmc_collect(A, nsplit);
      __shared__ int sel_s;
      if(threadIdx.x < 32)
      {
        cbmc_finish_warp(P, A, chain, r);
        if(threadIdx.x == 0)
        {
          r[14] = ((chain && dep_slot >= 0) ? A.B.result(dep_slot)[14] : 1.0) * r[0];
          r[13] = (r[9] != 0.0 && r[14] > 1e-150) ? 1.0 : 0.0;
          sel_s = (r[9] != 0.0 && r[11] > 0.0) ? (int) r[10] : -1;
          if(!chain && sel_s >= 0)
          {
            r[6] = A.B.tr(0)[sel_s]; r[7] = A.B.tr(1)[sel_s]; r[8] = A.B.tr(2)[sel_s];
            for(int k = 0; k < 9; k++) A.B.mol(GBK_BUF_GROWN, k)[0] = A.B.tr(k)[sel_s];
            A.B.mol_type(GBK_BUF_GROWN)[0] = A.B.tr_type()[sel_s];
          }
        }
      }
      __syncthreads();
      if(chain && sel_s >= 0 && threadIdx.x < cs)
      {
        const int j = sel_s * cs + threadIdx.x;
        for(int k = 0; k < 9; k++) A.B.mol(GBK_BUF_GROWN, k)[1 + threadIdx.x] = A.B.tr(k)[j];
        A.B.mol_type(GBK_BUF_GROWN)[1 + threadIdx.x] = A.B.tr_type()[j];
      }
      // reinsertion: the grown molecule is kept in tempMolStorage while the old one is retraced (StoreNewLocation_Reinsertion)
      __syncthreads();
      if(type == 2 && (chain || F.ms == 1) && threadIdx.x < F.ms)
      {
        for(int k = 0; k < 9; k++) A.B.mol(GBK_BUF_TEMP, k)[threadIdx.x] = A.B.mol(GBK_BUF_GROWN, k)[threadIdx.x];
        A.B.mol_type(GBK_BUF_TEMP)[threadIdx.x] = A.B.mol_type(GBK_BUF_GROWN)[threadIdx.x];
      }
    }
  }
  grid_barrier(F.bar, phase);
}

// Ewald Fourier delta of [old atoms | new atoms] taken from molecule buffers / component slots; result slot 5
__device__ __forceinline__ void fused_ewald_stage(const DevParams& P, const FusedArgs& F, unsigned char* dyn, double* red, int old_src, long long old_start,
                                                  int nold, int new_buf, int nnew, const double* dep, unsigned int& phase)
{
  double* res = F.B.result(5);
  const bool run = (dep == nullptr || dep[0] != 0.0);
  const int n = nold + nnew;
  if(run)
  {
    const int kx1 = P.kmax[0] + 1, ky1 = P.kmax[1] + 1, kz1 = P.kmax[2] + 1;
    cplx* ex = reinterpret_cast<cplx*>(dyn);
    cplx* ey = ex + (size_t) n * kx1; cplx* ez = ey + (size_t) n * ky1;
    double* qeff = reinterpret_cast<double*>(ez + (size_t) n * kz1);
    double* pos3 = qeff + n;
    __syncthreads();
    for(int i = threadIdx.x; i < n; i += blockDim.x)
    {
      if(i < nold)
      {
        if(old_src < 0) { pos3[3 * i] = F.C.x[old_start + i]; pos3[3 * i + 1] = F.C.y[old_start + i]; pos3[3 * i + 2] = F.C.z[old_start + i]; qeff[i] = F.C.scoul[old_start + i] * F.C.q[old_start + i]; }
        else { pos3[3 * i] = F.B.mol(old_src, 0)[i]; pos3[3 * i + 1] = F.B.mol(old_src, 1)[i]; pos3[3 * i + 2] = F.B.mol(old_src, 2)[i]; qeff[i] = F.B.mol(old_src, 8)[i] * F.B.mol(old_src, 6)[i]; }
      }
      else
      {
        const int j = i - nold;
        pos3[3 * i] = F.B.mol(new_buf, 0)[j]; pos3[3 * i + 1] = F.B.mol(new_buf, 1)[j]; pos3[3 * i + 2] = F.B.mol(new_buf, 2)[j]; qeff[i] = F.B.mol(new_buf, 8)[j] * F.B.mol(new_buf, 6)[j];
      }
    }
    __syncthreads();
    build_eik(P, pos3, n, ex, ey, ez, threadIdx.x, blockDim.x);
    __syncthreads();
    double same = 0.0, cross = 0.0;
    for(int kk = blockIdx.x * blockDim.x + threadIdx.x; kk < F.K.nact; kk += gridDim.x * blockDim.x)
    {
      int kx, ky, kz; unpack_k(F.K.kpack[kk], kx, ky, kz);
      const cplx co = ck_sum(ex, ey, ez, qeff, n, 0, nold, kx, ky, kz);
      const cplx cn = ck_sum(ex, ey, ez, qeff, n, nold, n, kx, ky, kz);
      const double temp = F.K.temp[kk];
      const int slot = F.K.slot[kk];
      const double ore = F.same_sf[2 * slot], oim = F.same_sf[2 * slot + 1];
      const double nre = ore + cn.re - co.re, nim = oim + cn.im - co.im;
      same += temp * (nre * nre + nim * nim);
      same -= temp * (ore * ore + oim * oim);
      F.temp_sf[2 * slot] = nre; F.temp_sf[2 * slot + 1] = nim;
      cross += temp * (F.cross_sf[2 * slot] * (cn.re - co.re) + F.cross_sf[2 * slot + 1] * (cn.im - co.im));
    }
    same = warp_sum(same); cross = warp_sum(cross);
    if(lane_id() == 0) { red[threadIdx.x >> 5] = same; red[8 + (threadIdx.x >> 5)] = cross; }
    __syncthreads();
    if(threadIdx.x == 0)
    {
      double s = 0.0, c = 0.0;
      for(int w = 0; w < (int)(blockDim.x >> 5); w++) { s += red[w]; c += red[8 + w]; }
      F.B.partial()[2 * blockIdx.x] = s; F.B.partial()[2 * blockIdx.x + 1] = c;
    }
  }
  grid_barrier(F.bar, phase);
  if(blockIdx.x == 0 && threadIdx.x == 0)
  {
    double s = 0.0, c = 0.0;
    if(run) { const volatile double* p = F.B.partial(); for(unsigned int b = 0; b < gridDim.x; b++) { s += p[2 * b]; c += p[2 * b + 1]; } }
    res[0] = s; res[1] = 2.0 * c;
  }
}

__global__ void __launch_bounds__(256)
k_move(DevParams P, SysView S, FusedArgs F)
{
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ FusedSmem sm;
  unsigned int phase = 0;
  stage_erfc_table(P, sm.etab);
  __syncthreads();
  PairTables W; W.etab = sm.etab; W.ffp = P.ffA; W.unit = false;
  const int ms = F.ms;
  if(F.kind == GBF_INSERTION || F.kind == GBF_DELETION)
  {
    const int type = F.kind == GBF_INSERTION ? 0 : 1;
    fused_cbmc_stage(P, S, F, &sm, W, false, type, F.pool_off, F.u0, 0, -1, -1, phase);
    int last = 0;
    if(ms > 1) { fused_cbmc_stage(P, S, F, &sm, W, true, type, F.pool_off + F.ntrials, F.u1, 1, 0, -1, phase); last = 1; }
    if(F.do_ewald)
    {
      if(F.kind == GBF_INSERTION) fused_ewald_stage(P, F, dyn, sm.red, 0, 0, 0, GBK_BUF_GROWN, ms, F.B.result(last) + 13, phase);
      else                        fused_ewald_stage(P, F, dyn, sm.red, -1, F.molecule * ms, ms, GBK_BUF_NEW, 0, F.B.result(last) + 13, phase);
    }
  }
  else if(F.kind == GBF_REINSERTION)
  {
    long long off = F.pool_off;
    fused_cbmc_stage(P, S, F, &sm, W, false, 2, off, F.u0, 0, -1, -1, phase); off += F.ntrials;
    int nl = 0;
    if(ms > 1) { fused_cbmc_stage(P, S, F, &sm, W, true, 2, off, F.u1, 1, 0, -1, phase); off += F.norient; nl = 1; }
    fused_cbmc_stage(P, S, F, &sm, W, false, 3, off, 0.0, 2, nl, 0, phase); off += 1;
    if(ms > 1) fused_cbmc_stage(P, S, F, &sm, W, true, 3, off, 0.0, 3, nl, -1, phase);
    if(F.do_ewald) fused_ewald_stage(P, F, dyn, sm.red, -1, F.molecule * ms, ms, GBK_BUF_TEMP, ms, F.B.result(nl) + 13, phase);
  }
  else
  {
    // ---- single body: proposal (every CTA computes it; CTA 0 publishes the buffers), new/old energies over CTA slices
    const int i = threadIdx.x;
    double* pn = reinterpret_cast<double*>(dyn);           // [ms][3] new Cartesian (kept for the Ewald stage through the buffers)
    (void) pn;
    if(i < ms)
    {
      ProposeArgs A; memset(&A, 0, sizeof(A));
      A.move_type = F.move_type; A.ms = ms; A.start = F.molecule * ms; A.pool_index = F.pool_off; A.pool3 = F.pool3;
      A.maxc[0] = F.maxc[0]; A.maxc[1] = F.maxc[1]; A.maxc[2] = F.maxc[2]; A.C = F.C; A.B = F.B;
      if(blockIdx.x == 0) propose_atom(P, A, i);
    }
    grid_barrier(F.bar, phase);
    double tot[14];
    for(int k = 0; k < 14; k++) tot[k] = 0.0;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = (int) lane_id();
    for(int pass = 0; pass < 2; pass++)
    {
      const int buf = pass == 0 ? GBK_BUF_NEW : GBK_BUF_OLD;
      __syncthreads();
      if(threadIdx.x < ms)
      {
        const int a = threadIdx.x;
        sm.T.fx[a] = F.B.mol(buf, 3)[a]; sm.T.fy[a] = F.B.mol(buf, 4)[a]; sm.T.fz[a] = F.B.mol(buf, 5)[a];
        sm.T.q[a] = F.B.mol(buf, 6)[a] * F.B.mol(buf, 8)[a]; sm.T.scale[a] = F.B.mol(buf, 7)[a]; sm.T.type[a] = F.B.mol_type(buf)[a]; sm.T.slot[a] = 0;
      }
      __syncthreads();
      double e6[6] = {0, 0, 0, 0, 0, 0}; int flag = 0;
      pair_group_generic<0>(P, W, S, F.L, F.comp, (int) F.molecule, -1, -1, &sm.T, ms, sm.Q + warp, blockIdx.x * nwarps + warp, gridDim.x * nwarps, e6, flag);
#pragma unroll
      for(int k = 0; k < 6; k++) tot[pass * 7 + k] = warp_sum(e6[k]);
      tot[pass * 7 + 6] = __any_sync(0xffffffffu, flag) ? 1.0 : 0.0;
    }
    if(lane == 0) for(int k = 0; k < 14; k++) sm.red[warp * 16 + k] = tot[k];
    __syncthreads();
    if(threadIdx.x < 14)
    {
      double s = 0.0;
      for(int w = 0; w < nwarps; w++) s += sm.red[w * 16 + threadIdx.x];
      F.B.partial()[1024 + blockIdx.x * 16 + threadIdx.x] = s;
    }
    grid_barrier(F.bar, phase);
    if(blockIdx.x == 0 && threadIdx.x == 0)
    {
      const volatile double* p = F.B.partial() + 1024;
      double* r = F.B.result(4);
      for(int k = 0; k < 6; k++)
      {
        double n = 0.0, o = 0.0;
        for(unsigned int b = 0; b < gridDim.x; b++) { n += p[b * 16 + k]; o += p[b * 16 + 7 + k]; }
        r[k] = n - o;
      }
      double fl = 0.0;
      for(unsigned int b = 0; b < gridDim.x; b++) fl += p[b * 16 + 6];
      r[6] = fl > 0.0 ? 1.0 : 0.0; r[7] = 1.0 - r[6];
    }
    grid_barrier(F.bar, phase);
    if(F.do_ewald) fused_ewald_stage(P, F, dyn, sm.red, GBK_BUF_OLD, 0, ms, GBK_BUF_NEW, ms, F.check_overlap ? F.B.result(4) + 7 : nullptr, phase);
  }
  // leave the barrier counter at zero for the next launch (every CTA has passed the last barrier once CTA 0 gets here
  // only if it is the last to arrive; so the reset is done by the last CTA through a second counter)
  __syncthreads();
  if(threadIdx.x == 0)
  {
    __threadfence();
    if(atomicAdd(F.bar + 1, 1u) == gridDim.x - 1) { F.bar[0] = 0u; F.bar[1] = 0u; }
  }
}
