// graspa_b200 -- ONE kernel per Monte Carlo move.
//
// A single GCMC move is a chain of small dependent stages (first bead -> selection -> chain growth -> selection ->
// Ewald delta).  The reference runs each stage as its own launch followed by cudaDeviceSynchronize and a host-side sum
// (3-6 host round trips per move); what bounds GCMC cycles/s is therefore latency, not arithmetic.  Here the whole move
// is one launch of a small co-resident grid (grid <= number of SMs):
//   * everything that does not depend on a selection runs in the FIRST stage: a deletion evaluates its first-bead
//     trials, its orientations and its Ewald delta together; a reinsertion evaluates the retrace of the old molecule
//     next to the first-bead trials of the new one; a translation/rotation evaluates new, old and Ewald at once;
//   * stages are separated by a flag barrier (one store + one poll per CTA, no atomics);
//   * after a barrier EVERY CTA reduces the stage's partial sums and runs the Boltzmann selection redundantly
//     (bitwise identical arithmetic), so no CTA waits for another one to publish the selected trial;
//   * CTA 0 stores the result block straight into pinned host memory and raises a sequence flag the host polls.
// Barriers per move: translation/rotation 1, deletion 1, insertion 3 (2 for single-bead molecules), reinsertion 3.
// Covers Insertion_Body / Deletion_Body (mc_swap_utilities.h:3-225), the reinsertion growth + retrace
// (move_struct.h:186-338), SingleBody_Prepare + SingleBody_Calculation (mc_single_particle.h:10-241) and the energy part of
// IdentitySwapMove (mc_swap_moves.h:199-362: nothing there depends on a selection except the Ewald delta -- the new first
// bead is the old molecule's first atom -- so growth of the new species and retrace of the old one share ONE stage).
#pragma once
#ifndef GBF_SPLIT_ATOMS
#define GBF_SPLIT_ATOMS 256
#endif
#include "common.cuh"
#include "pair.cuh"
#include "ewald.cuh"
#include "move_kernels.cuh"

enum { GBF_INSERTION = 0, GBF_DELETION = 1, GBF_REINSERTION = 2, GBF_SINGLE = 3, GBF_IDSWAP = 4, GBF_EXIT = 5 };
#define GBF_TAIL_TYPES 16           // pseudo-atom types the in-kernel tail difference handles (more: the host asks the engine separately)
#define GBF_MAX_GROUPS 96           // trial groups of one stage: ntrials + 1 + norient <= 65
#define GBF_PART_HALF 4096          // doubles per parity half of MoveBufs::partial()
#define GBF_MAX_DYN_SMEM (160 * 1024)
#define GBF_MAX_ITEMS 192           // (group, split) items of one stage; each publishes one 128-byte record
#define GBF_SPIN_LIMIT (1u << 24)      // polls of a hand-over record before the kernel gives up (__trap)
#define GBF_PART_EWALD (GBF_MAX_ITEMS * 16)   // Ewald CTA records (32 bytes each) start here inside a half

// a state commit the resident move server applies before the move that carries it (the accept calls queue these instead of
// launching k_commit_*): op 1 = k_commit_from_buffer(buf -> dst, n, what, molid), op 2 = k_commit_delete(dst <- src, n)
struct CommitOp { int op, buf, dst, src, n, what, molid, pad; };

struct FusedArgs
{
  int kind, comp, ms, move_type;          // move_type: GB_TRANSLATION / GB_ROTATION / GB_SPECIAL_ROTATION for GBF_SINGLE
  int ngrid;                              // CTAs that take part in this move (= gridDim.x of a k_move launch; <= grid of the resident server)
  int ncommit; CommitOp commit[2];        // resident server only: commits of the previous accepted move, applied first
  // tail-correction difference of the move (TailCorrectionDifference / TailCorrectionIdentitySwap, TailCorrection_Energy_Functions.h:36-113),
  // evaluated by a warp of CTA 0 beside the first stage: tail_n pseudo-atom types (0: not asked for), their occupation numbers and
  // the change the move would make; result in result slot 5, entry 2
  int tail_n; int tail_np[GBF_TAIL_TYPES]; short tail_d[GBF_TAIL_TYPES];
  const int* tail_use; const double* tail_e;
  long long molecule, pool_off;
  double u0, u1, scale0, scale1, maxc[3];
  int ntrials, norient, nmol;             // nmol = NumberOfMolecule_for_Component (MolID of an inserted molecule)
  int do_ewald, check_overlap, framework_moved;
  int natoms;                             // atoms in the live ranges of L
  const double* __restrict__ pool3;
  CompView C; MoveBufs B;
  // identity swap: F.comp / F.C / F.ms / F.nmol describe the NEW species; the molecule that leaves is `molecule` of old_comp
  int old_comp, ms2, nold_ew, nnew_ew; CompView C2;
  SegList L;                              // live ranges with the kinds of THIS move
  KTable K; const double* same_sf; const double* cross_sf; double* temp_sf;
  double* host_result;                    // pinned host memory (UVA): 128 tagged 16-byte records the host polls
  unsigned long long seq;                 // launch counter; its low 32 bits tag every record of this launch
};


// Cross-CTA hand-over without barriers or fences.  A CTA publishes a partial sum as 16-byte stores {lo, tag, hi, tag}:
// every 8-byte half carries its own validity tag (8-byte accesses are single-copy atomic), so a consumer that polls the
// 16 bytes with a volatile load and finds both tags equal to this stage's tag has the value -- no flag, no fence, one L2
// round trip.  Tags are unique per (launch, stage); the halves of MoveBufs::partial() alternate between stages, and a CTA
// only publishes stage s+1 after it has consumed ALL of stage s, so a record is never overwritten while someone still
// polls for its previous content.  All CTAs are co-resident (grid <= number of SMs), so polling cannot deadlock.
__device__ __forceinline__ void ll_store(double* rec, int j, double v, unsigned int tag)
{
  const unsigned int lo = (unsigned int) __double2loint(v), hi = (unsigned int) __double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(rec + 2 * j), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}

__device__ __forceinline__ bool ll_load(const double* rec, int j, unsigned int tag, double& v)
{
  unsigned int lo, t0, hi, t1;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(rec + 2 * j) : "memory");
  v = __hiloint2double((int) hi, (int) lo);
  return t0 == tag && t1 == tag;
}

__device__ __forceinline__ unsigned int stage_tag(unsigned long long seq, int stage) { return (unsigned int) (seq * 4ULL + (unsigned long long) stage + 1ULL); }

struct SmemMol { double a[9][GBK_MV_MOL_SLOTS]; int type[GBK_MV_MOL_SLOTS]; };   // x y z fx fy fz q scale scoul

__device__ __forceinline__ void put_atom(SmemMol& M, int i, const AtomRec& r)
{
  M.a[0][i] = r.x; M.a[1][i] = r.y; M.a[2][i] = r.z; M.a[3][i] = r.fx; M.a[4][i] = r.fy; M.a[5][i] = r.fz;
  M.a[6][i] = r.q; M.a[7][i] = r.scale; M.a[8][i] = r.scoul; M.type[i] = r.type;
}

struct FusedSmem
{
  TrialGroup T;
  WarpQueueG Q[8];
  double red[8 * 16];
  double etab[(GBK_ERFC_DEG + 1) * GBK_ERFC_NINT];
  double E[GBF_MAX_GROUPS * 6]; int Fl[GBF_MAX_GROUPS];     // collected trial energies / overlap flags of the last stage
  double Ekeep[33 * 6]; int Fkeep[33];                      // reinsertion: the retrace groups of stage 1, finished at the end
  __align__(32) double res[8 * 16];                                       // the result slots (MoveBufs::result layout)
  SmemMol mN, mO;                                           // molecule being grown / proposed, and its old image
  SmemMol tmpl, exist;                                      // template molecule (slot 0 of the component) and the selected molecule
  SmemMol tmpl2;                                            // identity swap: template of the OLD species (geometry of its retrace orientations)
  double pool[3 * (2 * 32 + 2 * 32 + 1)];                   // the random-pool entries this move consumes
  int blocked;                                              // block-pocket flag of the trial group being evaluated
  double tailv[GBF_TAIL_TYPES * (GBF_TAIL_TYPES + 1) / 2];  // terms of the tail-correction difference, in (i, j >= i) order
};

// one segment of a stage: n trial groups of one CBMC type
struct StageSeg { int type, chain, n; long long pool_off; };

__device__ __forceinline__ double* part_half(const FusedArgs& F, int par) { return F.B.partial() + (size_t) par * GBF_PART_HALF; }

__device__ __forceinline__ int stage_nsplit(const FusedArgs& F, int ngroups)
{
  const int ns = (F.natoms + GBF_SPLIT_ATOMS - 1) / GBF_SPLIT_ATOMS;   // system atoms per (group, split) item; 256 = one 32-atom iteration per warp
  const int cap = max(1, min(GBF_MAX_ITEMS / max(ngroups, 1), F.ngrid / max(ngroups, 1)));
  return max(1, min(ns, cap));
}

// get_random_trial_position, mc_widom.h:122-213 (template / selected molecule / pool entries come from shared memory)
__device__ __forceinline__ AtomRec first_bead_atom(const DevParams& P, const FusedArgs& F, const FusedSmem* sm, int type, int g, long long pool_off)
{
  const bool insertion_like = (type == 0 || type == 4);
  const SmemMol& M = insertion_like ? sm->tmpl : sm->exist;
  AtomRec r; r.scale = F.scale0; r.scoul = F.scale1;
  if(!insertion_like) { r.scale = M.a[7][0]; r.scoul = M.a[8][0]; }
  const bool existing = (type == 1 || type == 3 || type == 5) && g == 0;
  if(existing) { r.x = M.a[0][0]; r.y = M.a[1][0]; r.z = M.a[2][0]; }
  else if(type == 4) { r.x = sm->exist.a[0][0]; r.y = sm->exist.a[1][0]; r.z = sm->exist.a[2][0]; }     // copy_firstbead_to_new, mc_swap_moves.h:178-181
  else { const double* u = sm->pool + 3 * (pool_off - F.pool_off + g); r.x = P.cell[0] * u[0]; r.y = P.cell[4] * u[1]; r.z = P.cell[8] * u[2]; }
  to_frac(P, r.x, r.y, r.z, r.fx, r.fy, r.fz);
  r.q = M.a[6][0]; r.type = M.type[0];
  return r;
}

// get_random_trial_orientation, mc_widom.h:215-303: atom 1 + a of orientation g around the first bead (fbx, fby, fbz)
__device__ __forceinline__ AtomRec chain_atom(const DevParams& P, const FusedArgs& F, const FusedSmem* sm, int type, int g, int a, long long pool_off,
                                              double fbx, double fby, double fbz, double scale, double scoul)
{
  const bool insertion_like = (type == 0 || type == 4);
  const SmemMol& M = insertion_like ? sm->tmpl : sm->exist;                    // start_position, mc_widom.h:536-556
  const SmemMol& G = (F.kind == GBF_IDSWAP && type == 5) ? sm->tmpl2 : sm->tmpl;
  double vx = G.a[0][1 + a] - G.a[0][0], vy = G.a[1][1 + a] - G.a[1][0], vz = G.a[2][1 + a] - G.a[2][0];   // :256
  AtomRec r;
  if((type == 1 || type == 3 || type == 5) && g == 0) { r.x = M.a[0][1 + a]; r.y = M.a[1][1 + a]; r.z = M.a[2][1 + a]; }
  else
  {
    const double* u = sm->pool + 3 * (pool_off - F.pool_off + g);
    rotate_quaternion(vx, vy, vz, u[0], u[1], u[2]);
    r.x = fbx + vx; r.y = fby + vy; r.z = fbz + vz;
  }
  to_frac(P, r.x, r.y, r.z, r.fx, r.fy, r.fz);
  r.scale = scale; r.scoul = scoul;
  r.q = M.a[6][1 + a]; r.type = M.type[1 + a];
  return r;
}

// first bead the orientations of `type` are grown around: the selected trial (growth) or the existing atom (deletion / retrace)
__device__ __forceinline__ void chain_anchor(const FusedArgs& F, const FusedSmem* sm, int type, double& x, double& y, double& z, double& scale, double& scoul)
{
  const SmemMol& M = (type == 1 || type == 3 || type == 5) ? sm->exist : sm->mN;
  x = M.a[0][0]; y = M.a[1][0]; z = M.a[2][0]; scale = M.a[7][0]; scoul = M.a[8][0];
}

__device__ __forceinline__ void set_trial(TrialGroup& T, int a, const AtomRec& r)
{
  T.fx[a] = r.fx; T.fy[a] = r.fy; T.fz[a] = r.fz; T.q[a] = r.q * r.scoul; T.scale[a] = r.scale; T.type[a] = r.type; T.slot[a] = 0;
}

// pair energies of the trial group in sm->T against slice `split` of the live ranges; 7 partial sums to out8
template <int CS>
__device__ __forceinline__ void group_energy(const DevParams& P, const PairTables& W, const SysView& S, const FusedArgs& F, FusedSmem* sm,
                                             int new_molid, int cs, int split, int nsplit, double* rec, unsigned int tag)
{
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = (int) lane_id();
  double e6[6] = {0, 0, 0, 0, 0, 0}; int flag = 0;
  pair_group_flat<CS>(P, W, S, F.L, F.kind == GBF_IDSWAP ? F.old_comp : F.comp, new_molid, &sm->T, cs, sm->Q + warp, split * nwarps + warp, nsplit * nwarps, e6, flag);
#pragma unroll
  for(int k = 0; k < 6; k++) e6[k] = warp_sum(e6[k]);
  flag = __any_sync(0xffffffffu, flag);
  if(lane == 0) { for(int k = 0; k < 6; k++) sm->red[warp * 8 + k] = e6[k]; sm->red[warp * 8 + 6] = flag ? 1.0 : 0.0; }
  __syncthreads();
  if(threadIdx.x < 7)
  {
    double s = 0.0;
    for(int w = 0; w < nwarps; w++) s += sm->red[w * 8 + threadIdx.x];
    if(threadIdx.x == 6 && sm->blocked) s += 1.0;                       // block pockets flag the trial like an overlap
    ll_store(rec, threadIdx.x, s, tag);
  }
}

// this CTA's share of the (group, split) items of a stage made of up to 3 segments
__device__ __forceinline__ void run_stage(const DevParams& P, const SysView& S, const FusedArgs& F, FusedSmem* sm, const PairTables& W,
                                          const StageSeg* segs, int nseg, int ngroups, int nsplit, int par)
{
  double* part = part_half(F, par);
  const unsigned int tag = stage_tag(F.seq, par);
  for(int w = blockIdx.x; w < ngroups * nsplit; w += F.ngrid)
  {
    const int gg = w / nsplit, split = w % nsplit;
    int g = gg, si = 0;
    while(si + 1 < nseg && g >= segs[si].n) { g -= segs[si].n; si++; }
    const int type = segs[si].type; const bool chain = segs[si].chain != 0; const long long off = segs[si].pool_off;
    const int msg = (F.kind == GBF_IDSWAP && type == 5) ? F.ms2 : F.ms;
    const int cs = chain ? msg - 1 : 1;
    __syncthreads();
    if(!chain)
    {
      if(threadIdx.x == 0)
      {
        const AtomRec fb = first_bead_atom(P, F, sm, type, g, off);
        set_trial(sm->T, 0, fb);
        // block pockets, mc_widom.h:445-497: growth types; a blocked starting bead (trial 0) flags every trial
        int blk = 0;
        if((type == 0 || type == 2 || type == 4) && F.C.npocket > 0)
        {
          AtomRec f0 = fb;
          if(g > 0) f0 = first_bead_atom(P, F, sm, type, 0, off);
          blk = blocked_pocket(P, F.C, f0.x, f0.y, f0.z) ? 1 : 0;
          if(!blk && g > 0) blk = blocked_pocket(P, F.C, fb.x, fb.y, fb.z) ? 1 : 0;
        }
        sm->blocked = blk;
      }
    }
    else
    {
      if((int) threadIdx.x < cs)
      {
        double ax, ay, az, sc, scc; chain_anchor(F, sm, type, ax, ay, az, sc, scc);
        set_trial(sm->T, threadIdx.x, chain_atom(P, F, sm, type, g, threadIdx.x, off, ax, ay, az, sc, scc));
      }
      if(threadIdx.x == 0) sm->blocked = 0;
    }
    __syncthreads();
    // molecule skipped by the pair loops: the one being grown / retraced; identity swap: the molecule that leaves, for the
    // growth of the new species (Sims.ExcludeList[0], mc_swap_moves.h:268) and for its own retrace alike
    const int new_molid = (F.kind == GBF_IDSWAP) ? (int) F.molecule : ((type == 0 || type == 4) ? F.nmol : (int) F.molecule);
    double* rec = part + (size_t) w * 16;
    if(cs == 1)      group_energy<1>(P, W, S, F, sm, new_molid, cs, split, nsplit, rec, tag);
    else if(cs == 2) group_energy<2>(P, W, S, F, sm, new_molid, cs, split, nsplit, rec, tag);
    else             group_energy<0>(P, W, S, F, sm, new_molid, cs, split, nsplit, rec, tag);
  }
  GBK_MARK();
}

// wait for `nitems` published records of 7 sums (stage tag `tag`) and copy them to stash[item * 8 + j]
__device__ __forceinline__ void fetch_records(const double* part, int nitems, unsigned int tag, double* stash)
{
  for(int w = threadIdx.x; w < nitems; w += blockDim.x)
  {
    const double* rec = part + (size_t) w * 16;
    double v[7]; bool ok; unsigned int spins = 0;
    do
    {
      ok = true;
#pragma unroll
      for(int j = 0; j < 7; j++) ok = ll_load(rec, j, tag, v[j]) && ok;
      if(!ok && ++spins > GBF_SPIN_LIMIT) __trap();           // ~10 s of polling: a record that never comes is an error, not a hang
    } while(!ok);
#pragma unroll
    for(int j = 0; j < 7; j++) stash[w * 8 + j] = v[j];
  }
}

// fixed-order sum of the split partials of every group of the stage, by EVERY CTA: one thread per (group, split) item
// polls that item's record, then one thread per group adds the sums in split order
__device__ __forceinline__ void collect_stage(const FusedArgs& F, FusedSmem* sm, double* stash, int ngroups, int nsplit, int par)
{
  __syncthreads();
  fetch_records(part_half(F, par), ngroups * nsplit, stage_tag(F.seq, par), stash);
  __syncthreads();
  if((int) threadIdx.x < ngroups)
  {
    const double* p = stash + (size_t) threadIdx.x * nsplit * 8;
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for(int k = 0; k < nsplit; k++)
#pragma unroll
      for(int j = 0; j < 7; j++) s[j] += p[k * 8 + j];
#pragma unroll
    for(int j = 0; j < 6; j++) sm->E[6 * threadIdx.x + j] = s[j];
    sm->Fl[threadIdx.x] = s[6] > 0.0 ? 1 : 0;
  }
  __syncthreads();
  GBK_MARK();
}

// Boltzmann weights / selection / Rosenbluth factor of one segment into result slot `slot` (all threads call; warp 0 works)
__device__ __forceinline__ void finish_segment(const DevParams& P, const FusedArgs& F, FusedSmem* sm, int type, bool chain, int n, double uniform,
                                               double stored, const double* E, const int* Fl, int slot, double prev_product)
{
  double* r = sm->res + 16 * slot;
  if(threadIdx.x < 32)
  {
    cbmc_finish_core(P, type, chain, n, chain ? F.norient : F.ntrials, uniform, stored, E, Fl, r);
    if(threadIdx.x == 0)
    {
      r[14] = prev_product * r[0];                                       // CBMC.Rosenbluth *= averagedRosen, mc_widom.h:611
      r[13] = (r[9] != 0.0 && r[14] > 1e-150) ? 1.0 : 0.0;               // mc_swap_utilities.h:21,32
    }
  }
  __syncthreads();
  GBK_MARK();
}

// the selected trial joins the molecule being grown (Mol.pos[0] = NewMol.pos[FirstBeadTrial], mc_widom.h:230-241 / 590-599);
// recomputed from the pool, which gives the same bits as the trial that was evaluated
__device__ __forceinline__ void adopt_selection(const DevParams& P, const FusedArgs& F, FusedSmem* sm, int type, bool chain, long long pool_off, int slot)
{
  double* r = sm->res + 16 * slot;
  const bool have = r[9] != 0.0 && r[11] > 0.0;
  const int sel = (int) r[10];
  __syncthreads();
  if(have)
  {
    if(!chain)
    {
      if(threadIdx.x == 0)
      {
        const AtomRec a = first_bead_atom(P, F, sm, type, sel, pool_off);
        put_atom(sm->mN, 0, a);
        r[6] = a.x; r[7] = a.y; r[8] = a.z;
      }
    }
    else if((int) threadIdx.x < F.ms - 1)
    {
      double ax, ay, az, sc, scc; chain_anchor(F, sm, type, ax, ay, az, sc, scc);
      put_atom(sm->mN, 1 + threadIdx.x, chain_atom(P, F, sm, type, sel, threadIdx.x, pool_off, ax, ay, az, sc, scc));
    }
  }
  __syncthreads();
  // block pockets: every atom of the grown molecule is tested after the chain growth of an insertion / reinsertion
  // (mc_swap_utilities.h:35-78, move_struct.h:208-250); a blocked atom fails the construction, the selection's random
  // number stays consumed.  Every CTA takes the same decision from the same shared-memory molecule.
  if(have && chain && (type == 0 || type == 2 || type == 4) && F.C.npocket > 0)
  {
    if(threadIdx.x == 0)
    {
      bool blk = false;
      for(int a = 0; a < F.ms && !blk; a++) blk = blocked_pocket(P, F.C, sm->mN.a[0][a], sm->mN.a[1][a], sm->mN.a[2][a]);
      if(blk) { r[0] = 0.0; r[9] = 0.0; r[13] = 0.0; r[14] = 0.0; }
    }
    __syncthreads();
  }
  GBK_MARK();
}

// Ewald Fourier delta of [old atoms | new atoms]: the last `ne` CTAs of the grid each take 256 k-vectors; CTA partial
// {same, cross} goes to out2[2 * slice].  old atoms: component slots (oldS == nullptr) or a shared-memory molecule.
__device__ __forceinline__ int ewald_ctas(const FusedArgs& F) { return min(F.ngrid, (F.K.nact + 255) / 256); }

__device__ __forceinline__ void ewald_slice(const DevParams& P, const FusedArgs& F, unsigned char* dyn, double* red, const SmemMol* oldS,
                                            int nold, const SmemMol* newS, int nnew, double* out2)
{
  const int ne = ewald_ctas(F);
  const int slice = (int) blockIdx.x - (F.ngrid - ne);
  if(slice < 0) return;
  const int n = nold + nnew;
  const int kx1 = P.kmax[0] + 1, ky1 = P.kmax[1] + 1, kz1 = P.kmax[2] + 1;
  cplx* ex = reinterpret_cast<cplx*>(dyn);
  cplx* ey = ex + (size_t) n * kx1; cplx* ez = ey + (size_t) n * ky1;
  double* qeff = reinterpret_cast<double*>(ez + (size_t) n * kz1);
  double* pos3 = qeff + n;
  // the k-vector of this thread and its stored structure factors do not depend on the eik tables: issue those loads first
  const int kk0 = slice * blockDim.x + threadIdx.x;
  int kp0 = 0, slot0 = 0; double temp0 = 0.0, ore0 = 0.0, oim0 = 0.0, cre0 = 0.0, cim0 = 0.0;
  if(kk0 < F.K.nact)
  {
    kp0 = F.K.kpack[kk0]; temp0 = F.K.temp[kk0]; slot0 = F.K.slot[kk0];
    ore0 = F.same_sf[2 * slot0]; oim0 = F.same_sf[2 * slot0 + 1]; cre0 = F.cross_sf[2 * slot0]; cim0 = F.cross_sf[2 * slot0 + 1];
  }
  __syncthreads();
  for(int i = threadIdx.x; i < n; i += blockDim.x)
  {
    if(i < nold)
    {
      { pos3[3 * i] = oldS->a[0][i]; pos3[3 * i + 1] = oldS->a[1][i]; pos3[3 * i + 2] = oldS->a[2][i]; qeff[i] = oldS->a[8][i] * oldS->a[6][i]; }
    }
    else
    {
      const int j = i - nold;
      pos3[3 * i] = newS->a[0][j]; pos3[3 * i + 1] = newS->a[1][j]; pos3[3 * i + 2] = newS->a[2][j]; qeff[i] = newS->a[8][j] * newS->a[6][j];
    }
  }
  __syncthreads();
  build_eik(P, pos3, n, ex, ey, ez, threadIdx.x, blockDim.x);
  __syncthreads();
  double same = 0.0, cross = 0.0;
  for(int kk = kk0; kk < F.K.nact; kk += ne * blockDim.x)
  {
    int kp = kp0, slot = slot0; double temp = temp0, ore = ore0, oim = oim0, cre = cre0, cim = cim0;
    if(kk != kk0)
    {
      kp = F.K.kpack[kk]; temp = F.K.temp[kk]; slot = F.K.slot[kk];
      ore = F.same_sf[2 * slot]; oim = F.same_sf[2 * slot + 1]; cre = F.cross_sf[2 * slot]; cim = F.cross_sf[2 * slot + 1];
    }
    int kx, ky, kz; unpack_k(kp, kx, ky, kz);
    const cplx co = ck_sum(ex, ey, ez, qeff, n, 0, nold, kx, ky, kz);
    const cplx cn = ck_sum(ex, ey, ez, qeff, n, nold, n, kx, ky, kz);
    const double nre = ore + cn.re - co.re, nim = oim + cn.im - co.im;
    same += temp * (nre * nre + nim * nim);
    same -= temp * (ore * ore + oim * oim);
    F.temp_sf[2 * slot] = nre; F.temp_sf[2 * slot + 1] = nim;
    cross += temp * (cre * (cn.re - co.re) + cim * (cn.im - co.im));
  }
  same = warp_sum(same); cross = warp_sum(cross);
  __syncthreads();
  if(lane_id() == 0) { red[threadIdx.x >> 5] = same; red[8 + (threadIdx.x >> 5)] = cross; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    double s = 0.0, c = 0.0;
    for(int w = 0; w < (int)(blockDim.x >> 5); w++) { s += red[w]; c += red[8 + w]; }
    const unsigned int tag = stage_tag(F.seq, 2);
    ll_store(out2 + 4 * slice, 0, s, tag); ll_store(out2 + 4 * slice, 1, c, tag);
  }
  GBK_MARK();
}

// wait for the records of the Ewald CTAs and copy {same, cross} to stash[2 * b]
__device__ __forceinline__ void fetch_ewald(const FusedArgs& F, const double* in2, double* stash)
{
  const int ne = ewald_ctas(F);
  const unsigned int tag = stage_tag(F.seq, 2);
  for(int w = threadIdx.x; w < 2 * ne; w += blockDim.x)
  {
    double v; unsigned int spins = 0;
    while(!ll_load(in2 + 4 * (w >> 1), w & 1, tag, v)) { if(++spins > GBF_SPIN_LIMIT) __trap(); }
    stash[w] = v;
  }
}

// CTA 0: result slot 5 = {same, 2 * cross} summed over the Ewald CTAs in fixed order
__device__ __forceinline__ void ewald_total_slot(const FusedArgs& F, FusedSmem* sm, double* stash, const double* in2, bool run)
{
  const int ne = ewald_ctas(F);
  __syncthreads();
  if(run) fetch_ewald(F, in2, stash);
  __syncthreads();
  if(threadIdx.x == 0)
  {
    double s = 0.0, c = 0.0;
    if(run) for(int b = 0; b < ne; b++) { s += stash[2 * b]; c += stash[2 * b + 1]; }
    sm->res[16 * 5] = s; sm->res[16 * 5 + 1] = 2.0 * c;
  }
  __syncthreads();
}

// CTA 0: copy a shared-memory molecule into one of the device molecule buffers (read by the accept calls)
__device__ __forceinline__ void export_molecule(const FusedArgs& F, const SmemMol& M, int buf)
{
  if((int) threadIdx.x < F.ms)
  {
#pragma unroll
    for(int k = 0; k < 9; k++) F.B.mol(buf, k)[threadIdx.x] = M.a[k][threadIdx.x];
    F.B.mol_type(buf)[threadIdx.x] = M.type[threadIdx.x];
  }
}

// One kernel for every move kind.  Measured and not kept (profiles/r2_move_kernel.md): one instantiation per kind with the kind as a
// build-time constant (each carries only its own move's code) -- 10 % SLOWER on all three GCMC decks, because consecutive moves of
// different kinds then alternate between kernels and every switch starts with cold instruction caches.
// ---- out-of-line copies of the stage routines for the resident server (SRV = true).  k_move inlines every routine at every call
// site (48 800 SASS instructions, 780 KB): one launch walks through its own kind's code once, so layout does not matter there.  The
// server executes move after move of every kind on the same SMs, and what it touches must stay inside the instruction caches:
// one copy of each routine, shared by all kinds.  Arguments are references into shared memory (P, S, F live there in the server).
#define GBF_OUT __device__ __noinline__
GBF_OUT void run_stage_out(const DevParams& P, const SysView S, const FusedArgs& F, FusedSmem* sm, PairTables W, const StageSeg* segs, int nseg, int ngroups, int nsplit, int par)
{ run_stage(P, S, F, sm, W, segs, nseg, ngroups, nsplit, par); }
GBF_OUT void collect_stage_out(const FusedArgs& F, FusedSmem* sm, double* stash, int ngroups, int nsplit, int par) { collect_stage(F, sm, stash, ngroups, nsplit, par); }
GBF_OUT void finish_segment_out(const DevParams& P, const FusedArgs& F, FusedSmem* sm, int type, bool chain, int n, double uniform, double stored, const double* E, const int* Fl, int slot, double prev_product)
{ finish_segment(P, F, sm, type, chain, n, uniform, stored, E, Fl, slot, prev_product); }
GBF_OUT void adopt_selection_out(const DevParams& P, const FusedArgs& F, FusedSmem* sm, int type, bool chain, long long pool_off, int slot) { adopt_selection(P, F, sm, type, chain, pool_off, slot); }
GBF_OUT void ewald_slice_out(const DevParams& P, const FusedArgs& F, unsigned char* dyn, double* red, const SmemMol* oldS, int nold, const SmemMol* newS, int nnew, double* out2)
{ ewald_slice(P, F, dyn, red, oldS, nold, newS, nnew, out2); }
GBF_OUT void ewald_total_slot_out(const FusedArgs& F, FusedSmem* sm, double* stash, const double* in2, bool run) { ewald_total_slot(F, sm, stash, in2, run); }
GBF_OUT void export_molecule_out(const FusedArgs& F, const SmemMol& M, int buf) { export_molecule(F, M, buf); }
GBF_OUT void group_energy0_out(const DevParams& P, PairTables W, const SysView S, const FusedArgs& F, FusedSmem* sm, int new_molid, int cs, int split, int nsplit, double* rec, unsigned int tag)
{ group_energy<0>(P, W, S, F, sm, new_molid, cs, split, nsplit, rec, tag); }

template <bool SRV> __device__ __forceinline__ void run_stage_d(const DevParams& P, const SysView& S, const FusedArgs& F, FusedSmem* sm, const PairTables& W, const StageSeg* segs, int nseg, int ngroups, int nsplit, int par)
{ if(SRV) run_stage_out(P, S, F, sm, W, segs, nseg, ngroups, nsplit, par); else run_stage(P, S, F, sm, W, segs, nseg, ngroups, nsplit, par); }
template <bool SRV> __device__ __forceinline__ void collect_stage_d(const FusedArgs& F, FusedSmem* sm, double* stash, int ngroups, int nsplit, int par)
{ if(SRV) collect_stage_out(F, sm, stash, ngroups, nsplit, par); else collect_stage(F, sm, stash, ngroups, nsplit, par); }
template <bool SRV> __device__ __forceinline__ void finish_segment_d(const DevParams& P, const FusedArgs& F, FusedSmem* sm, int type, bool chain, int n, double uniform, double stored, const double* E, const int* Fl, int slot, double prev_product)
{ if(SRV) finish_segment_out(P, F, sm, type, chain, n, uniform, stored, E, Fl, slot, prev_product); else finish_segment(P, F, sm, type, chain, n, uniform, stored, E, Fl, slot, prev_product); }
template <bool SRV> __device__ __forceinline__ void adopt_selection_d(const DevParams& P, const FusedArgs& F, FusedSmem* sm, int type, bool chain, long long pool_off, int slot)
{ if(SRV) adopt_selection_out(P, F, sm, type, chain, pool_off, slot); else adopt_selection(P, F, sm, type, chain, pool_off, slot); }
template <bool SRV> __device__ __forceinline__ void ewald_slice_d(const DevParams& P, const FusedArgs& F, unsigned char* dyn, double* red, const SmemMol* oldS, int nold, const SmemMol* newS, int nnew, double* out2)
{ if(SRV) ewald_slice_out(P, F, dyn, red, oldS, nold, newS, nnew, out2); else ewald_slice(P, F, dyn, red, oldS, nold, newS, nnew, out2); }
template <bool SRV> __device__ __forceinline__ void ewald_total_slot_d(const FusedArgs& F, FusedSmem* sm, double* stash, const double* in2, bool run)
{ if(SRV) ewald_total_slot_out(F, sm, stash, in2, run); else ewald_total_slot(F, sm, stash, in2, run); }
template <bool SRV> __device__ __forceinline__ void export_molecule_d(const FusedArgs& F, const SmemMol& M, int buf)
{ if(SRV) export_molecule_out(F, M, buf); else export_molecule(F, M, buf); }
template <bool SRV> __device__ __forceinline__ void group_energy0_d(const DevParams& P, const PairTables& W, const SysView& S, const FusedArgs& F, FusedSmem* sm, int new_molid, int cs, int split, int nsplit, double* rec, unsigned int tag)
{ if(SRV) group_energy0_out(P, W, S, F, sm, new_molid, cs, split, nsplit, rec, tag); else group_energy<0>(P, W, S, F, sm, new_molid, cs, split, nsplit, rec, tag); }

// Tail-correction difference by the last warp of CTA 0: the lanes evaluate the (i, j >= i) terms, lane 0 adds them in the reference's
// loop order (a term that is not used adds an exact zero) and divides by the volume -- the arithmetic of k_tail, engine.cu.
__device__ __forceinline__ void tail_delta_warp(const DevParams& P, const FusedArgs& F, FusedSmem& sm)
{
  const int lane = (int) lane_id(), n = F.tail_n, npair = n * (n + 1) / 2;
  for(int k = lane; k < npair; k += 32)
  {
    int i = 0, rem = k;
    while(rem >= n - i) { rem -= n - i; i++; }
    const int j = i + rem;
    double v = 0.0;
    if(F.tail_use[i * n + j])
    {
      const int Ni = F.tail_np[i], Nj = F.tail_np[j], di = F.tail_d[i], dj = F.tail_d[j];
      const int dN = Ni * dj + Nj * di + di * dj;
      v = F.tail_e[i * n + j] * (double) dN;
      if(i != j) v *= 2.0;
    }
    sm.tailv[k] = v;
  }
  __syncwarp();
  if(lane == 0)
  {
    double T = 0.0;
    for(int k = 0; k < npair; k++) T += sm.tailv[k];
    sm.res[16 * 5 + 2] = T / P.volume;
  }
}

// The whole move, run by every CTA with blockIdx.x < F.ngrid.  SRV = false: the body of one k_move launch (P, S, F are kernel
// parameters).  SRV = true: one command of the resident server (k_move_server): F and S are shared-memory copies -- the compiler then
// cannot route the loads of the (mutable) slot arrays through the non-coherent path -- and the erfc table is already staged.
template <bool SRV>
__device__ __forceinline__ void move_body(const DevParams& P, const SysView& S, const FusedArgs& F, FusedSmem& sm, unsigned char* dyn)
{
#ifdef GBK_PHASE_TIMING
  if(!SRV && blockIdx.x == 0 && threadIdx.x == 0) g_nmarks = 0;
#endif
  GBK_MARK();
  const int ms = F.ms;
  const int no = ms > 1 ? F.norient : 0;
  // prologue, one round trip to L2 for everything a stage set-up needs: warp 0 fetches the template molecule, the selected
  // molecule and the pool entries of this move (or builds the translation / rotation proposal) while warps 1-7 stage the erfc table
  if(threadIdx.x < 32)
  {
    if(F.kind == GBF_SINGLE)
    {
      if((int) threadIdx.x < ms)
      {
        ProposeArgs A; memset(&A, 0, sizeof(A));
        A.move_type = F.move_type; A.ms = ms; A.start = F.molecule * ms; A.pool_index = F.pool_off; A.pool3 = F.pool3;
        A.maxc[0] = F.maxc[0]; A.maxc[1] = F.maxc[1]; A.maxc[2] = F.maxc[2]; A.C = F.C; A.B = F.B;
        AtomRec nw, od;
        propose_atom(P, A, threadIdx.x, nw, od);
        put_atom(sm.mN, threadIdx.x, nw); put_atom(sm.mO, threadIdx.x, od);
      }
    }
    else
    {
      // every global load is issued before the first shared-memory store: in the resident server F itself lives in shared memory, and
      // a store the compiler cannot tell apart from F would otherwise order the loads one after the other (one L2 round trip each)
      const int npool = F.kind == GBF_IDSWAP ? 2 + no + (F.ms2 > 1 ? F.norient : 0) : F.ntrials + no + (F.kind == GBF_REINSERTION ? 1 + no : 0);
      const CompView Cv = F.C, Cv2 = F.C2;
      const int kind = F.kind, ms2 = F.ms2, i = threadIdx.x;
      const long long mol = F.molecule;
      const double* pool_src = F.pool3 + 3 * F.pool_off;
      double pv[7]; int np = 0;                       // 3 * npool <= 3 * 129 entries over 32 lanes: at most 13 per lane, 7 in the common case
      const int n3 = 3 * npool;
      for(int k = i; k < n3 && np < 7; k += 32) pv[np++] = pool_src[k];
      double t[6] = {0, 0, 0, 0, 0, 0}, x[6] = {0, 0, 0, 0, 0, 0}, t2[3] = {0, 0, 0}; int tt = 0, xt = 0;
      const bool has_t = i < ms, has_x = has_t && kind != GBF_INSERTION && kind != GBF_IDSWAP, has_x2 = kind == GBF_IDSWAP && i < ms2;
      if(has_t) { t[0] = Cv.x[i]; t[1] = Cv.y[i]; t[2] = Cv.z[i]; t[3] = Cv.q[i]; t[4] = Cv.scale[i]; t[5] = Cv.scoul[i]; tt = Cv.type[i]; }
      if(has_x) { const long long j = mol * ms + i; x[0] = Cv.x[j]; x[1] = Cv.y[j]; x[2] = Cv.z[j]; x[3] = Cv.q[j]; x[4] = Cv.scale[j]; x[5] = Cv.scoul[j]; xt = Cv.type[j]; }
      if(has_x2)
      {
        const long long j = mol * ms2 + i;
        x[0] = Cv2.x[j]; x[1] = Cv2.y[j]; x[2] = Cv2.z[j]; x[3] = Cv2.q[j]; x[4] = Cv2.scale[j]; x[5] = Cv2.scoul[j]; xt = Cv2.type[j];
        t2[0] = Cv2.x[i]; t2[1] = Cv2.y[i]; t2[2] = Cv2.z[i];
      }
      np = 0;
      for(int k = i; k < n3 && np < 7; k += 32) sm.pool[k] = pv[np++];
      for(int k = i + 7 * 32; k < n3; k += 32) sm.pool[k] = pool_src[k];
      if(has_t)
      {
        sm.tmpl.a[0][i] = t[0]; sm.tmpl.a[1][i] = t[1]; sm.tmpl.a[2][i] = t[2];
        sm.tmpl.a[6][i] = t[3]; sm.tmpl.a[7][i] = t[4]; sm.tmpl.a[8][i] = t[5]; sm.tmpl.type[i] = tt;
      }
      if(has_x || has_x2)
      {
        sm.exist.a[0][i] = x[0]; sm.exist.a[1][i] = x[1]; sm.exist.a[2][i] = x[2];
        sm.exist.a[6][i] = x[3]; sm.exist.a[7][i] = x[4]; sm.exist.a[8][i] = x[5]; sm.exist.type[i] = xt;
      }
      if(has_x2)
      {
        sm.tmpl2.a[0][i] = t2[0]; sm.tmpl2.a[1][i] = t2[1]; sm.tmpl2.a[2][i] = t2[2];
        double f0, f1, f2; to_frac(P, x[0], x[1], x[2], f0, f1, f2);
        sm.exist.a[3][i] = f0; sm.exist.a[4][i] = f1; sm.exist.a[5][i] = f2;
      }
    }
  }
  else
  {
    if(!SRV && !P.no_charges)  // without charges no pair ever evaluates erfc: one L2 round trip less in the prologue
      for(int i = threadIdx.x - 32; i < (GBK_ERFC_DEG + 1) * GBK_ERFC_NINT; i += blockDim.x - 32) sm.etab[i] = __ldg(&P.erfc_tab[i]);
    if(threadIdx.x - 32 < 128) sm.res[threadIdx.x - 32] = 0.0;
  }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    // translation / rotation: a block pocket containing any atom of the proposal counts as an overlap
    // (SingleBody_Prepare, mc_single_particle.h:83-119); the CBMC kinds set the flag per trial group in run_stage
    int blk = 0;
    if(F.kind == GBF_SINGLE && F.C.npocket > 0)
      for(int a = 0; a < ms && !blk; a++) blk = blocked_pocket(P, F.C, sm.mN.a[0][a], sm.mN.a[1][a], sm.mN.a[2][a]) ? 1 : 0;
    sm.blocked = blk;
  }
  __syncthreads();
  if(F.tail_n > 0 && blockIdx.x == 0 && threadIdx.x >= blockDim.x - 32) tail_delta_warp(P, F, sm);
  GBK_MARK();
  PairTables W; W.etab = sm.etab; W.ffp = P.ffA; W.unit = false;

  if(F.kind == GBF_INSERTION)
  {
    StageSeg sg[1];
    sg[0].type = 0; sg[0].chain = 0; sg[0].n = F.ntrials; sg[0].pool_off = F.pool_off;
    int nsplit = stage_nsplit(F, F.ntrials);
    run_stage_d<SRV>(P, S, F, &sm, W, sg, 1, F.ntrials, nsplit, 0);
    // a single-bead molecule without a Fourier stage has nothing left that depends on the selection: only CTA 0 finishes the move
    if(ms == 1 && !F.do_ewald && blockIdx.x != 0) return;
    collect_stage_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), F.ntrials, nsplit, 0);
    finish_segment_d<SRV>(P, F, &sm, 0, false, F.ntrials, F.u0, 0.0, sm.E, sm.Fl, 0, 1.0);
    adopt_selection_d<SRV>(P, F, &sm, 0, false, F.pool_off, 0);
    bool alive = sm.res[13] != 0.0;
    if(ms > 1)
    {
      sg[0].chain = 1; sg[0].n = no; sg[0].pool_off = F.pool_off + F.ntrials;
      nsplit = stage_nsplit(F, no);
      if(alive)
      {
        run_stage_d<SRV>(P, S, F, &sm, W, sg, 1, no, nsplit, 1);
        collect_stage_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), no, nsplit, 1);
        finish_segment_d<SRV>(P, F, &sm, 0, true, no, F.u1, 0.0, sm.E, sm.Fl, 1, sm.res[14]);
        adopt_selection_d<SRV>(P, F, &sm, 0, true, F.pool_off + F.ntrials, 1);
        alive = sm.res[16 + 13] != 0.0;
      }
    }
    double* ew = part_half(F, 0) + GBF_PART_EWALD;
    if(F.do_ewald && alive) ewald_slice_d<SRV>(P, F, dyn, sm.red, &sm.exist, 0, &sm.mN, ms, ew);
    if(blockIdx.x == 0)
    {
      if(F.do_ewald) ewald_total_slot_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), ew, alive);
      export_molecule_d<SRV>(F, sm.mN, GBK_BUF_GROWN);
    }
  }
  else if(F.kind == GBF_DELETION)
  {
    // nothing in a deletion depends on a selection: trial 0 of either stage is the existing molecule
    StageSeg sg[2];
    sg[0].type = 1; sg[0].chain = 0; sg[0].n = F.ntrials; sg[0].pool_off = F.pool_off;
    sg[1].type = 1; sg[1].chain = 1; sg[1].n = no; sg[1].pool_off = F.pool_off + F.ntrials;
    const int ngroups = F.ntrials + no;
    const int nsplit = stage_nsplit(F, ngroups);
    run_stage_d<SRV>(P, S, F, &sm, W, sg, ms > 1 ? 2 : 1, ngroups, nsplit, 0);
    double* ew = part_half(F, 0) + GBF_PART_EWALD;
    if(F.do_ewald) ewald_slice_d<SRV>(P, F, dyn, sm.red, &sm.exist, ms, &sm.mN, 0, ew);
    if(blockIdx.x == 0)
    {
      collect_stage_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), ngroups, nsplit, 0);
      finish_segment_d<SRV>(P, F, &sm, 1, false, F.ntrials, 0.0, 0.0, sm.E, sm.Fl, 0, 1.0);
      bool alive = sm.res[13] != 0.0;
      if(ms > 1 && alive)
      {
        finish_segment_d<SRV>(P, F, &sm, 1, true, no, 0.0, 0.0, sm.E + 6 * F.ntrials, sm.Fl + F.ntrials, 1, sm.res[14]);
        alive = sm.res[16 + 13] != 0.0;
      }
      if(threadIdx.x == 0 && sm.res[9] != 0.0 && sm.res[11] > 0.0)
      {
        sm.res[6] = sm.exist.a[0][0]; sm.res[7] = sm.exist.a[1][0]; sm.res[8] = sm.exist.a[2][0];
      }
      if(F.do_ewald) ewald_total_slot_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), ew, alive);
    }
  }
  else if(F.kind == GBF_REINSERTION)
  {
    // stage 1: first-bead trials of the new position + the whole retrace of the old molecule
    StageSeg sg[3];
    sg[0].type = 2; sg[0].chain = 0; sg[0].n = F.ntrials; sg[0].pool_off = F.pool_off;
    sg[1].type = 3; sg[1].chain = 0; sg[1].n = 1;         sg[1].pool_off = F.pool_off + F.ntrials + no;
    sg[2].type = 3; sg[2].chain = 1; sg[2].n = no;        sg[2].pool_off = F.pool_off + F.ntrials + no + 1;
    const int ngroups = F.ntrials + 1 + no;
    int nsplit = stage_nsplit(F, ngroups);
    run_stage_d<SRV>(P, S, F, &sm, W, sg, ms > 1 ? 3 : 2, ngroups, nsplit, 0);
    if(ms == 1 && !F.do_ewald && blockIdx.x != 0) return;
    collect_stage_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), ngroups, nsplit, 0);
    if(blockIdx.x == 0)
    {
      if((int) threadIdx.x < 6 * (1 + no)) sm.Ekeep[threadIdx.x] = sm.E[6 * F.ntrials + threadIdx.x];
      if((int) threadIdx.x < 1 + no) sm.Fkeep[threadIdx.x] = sm.Fl[F.ntrials + threadIdx.x];
      __syncthreads();
    }
    finish_segment_d<SRV>(P, F, &sm, 2, false, F.ntrials, F.u0, 0.0, sm.E, sm.Fl, 0, 1.0);
    adopt_selection_d<SRV>(P, F, &sm, 2, false, F.pool_off, 0);
    bool alive = sm.res[13] != 0.0;
    int nl = 0;
    if(ms > 1)
    {
      StageSeg sc[1];
      sc[0].type = 2; sc[0].chain = 1; sc[0].n = no; sc[0].pool_off = F.pool_off + F.ntrials;
      nsplit = stage_nsplit(F, no);
      if(alive)
      {
        run_stage_d<SRV>(P, S, F, &sm, W, sc, 1, no, nsplit, 1);
        collect_stage_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), no, nsplit, 1);
        finish_segment_d<SRV>(P, F, &sm, 2, true, no, F.u1, 0.0, sm.E, sm.Fl, 1, sm.res[14]);
        adopt_selection_d<SRV>(P, F, &sm, 2, true, F.pool_off + F.ntrials, 1);
        alive = sm.res[16 + 13] != 0.0;
      }
      nl = 1;
    }
    double* ew = part_half(F, 0) + GBF_PART_EWALD;
    if(F.do_ewald && alive) ewald_slice_d<SRV>(P, F, dyn, sm.red, &sm.exist, ms, &sm.mN, ms, ew);
    if(blockIdx.x == 0)
    {
      if(alive)
      {
        // retrace: Rosenbluth weight of the old configuration, with the stored weights of the insertion's other trials
        finish_segment_d<SRV>(P, F, &sm, 3, false, 1, 0.0, sm.res[1], sm.Ekeep, sm.Fkeep, 2, 1.0);
        if(threadIdx.x == 0 && sm.res[32 + 9] != 0.0 && sm.res[32 + 11] > 0.0)
        {
          sm.res[32 + 6] = sm.exist.a[0][0]; sm.res[32 + 7] = sm.exist.a[1][0]; sm.res[32 + 8] = sm.exist.a[2][0];
        }
        if(ms > 1) finish_segment_d<SRV>(P, F, &sm, 3, true, no, 0.0, 0.0, sm.Ekeep + 6, sm.Fkeep + 1, 3, sm.res[16 * nl + 14]);
      }
      if(F.do_ewald) ewald_total_slot_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), ew, alive);
      export_molecule_d<SRV>(F, sm.mN, GBK_BUF_TEMP);            // tempMolStorage, StoreNewLocation_Reinsertion
    }
  }
  else if(F.kind == GBF_IDSWAP)
  {
    // growth of the new species at the old molecule's first atom + retrace of the old molecule: one stage
    const int no2 = F.ms2 > 1 ? F.norient : 0;
    StageSeg sg[4]; int ns = 0;
    const long long off1 = F.pool_off + 1, off2 = off1 + no, off3 = off2 + 1;
    sg[ns].type = 4; sg[ns].chain = 0; sg[ns].n = 1; sg[ns].pool_off = F.pool_off; ns++;
    if(ms > 1) { sg[ns].type = 4; sg[ns].chain = 1; sg[ns].n = no; sg[ns].pool_off = off1; ns++; }
    sg[ns].type = 5; sg[ns].chain = 0; sg[ns].n = 1; sg[ns].pool_off = off2; ns++;
    if(F.ms2 > 1) { sg[ns].type = 5; sg[ns].chain = 1; sg[ns].n = no2; sg[ns].pool_off = off3; ns++; }
    const int ngroups = 2 + no + no2;
    // the new species' first bead is known before anything is evaluated: its chain trials grow around it right away
    if(threadIdx.x == 0) { const AtomRec a = first_bead_atom(P, F, &sm, 4, 0, F.pool_off); put_atom(sm.mN, 0, a); }
    __syncthreads();
    const int nsplit = stage_nsplit(F, ngroups);
    run_stage_d<SRV>(P, S, F, &sm, W, sg, ns, ngroups, nsplit, 0);
    if(ms == 1 && F.ms2 == 1 && !F.do_ewald && blockIdx.x != 0) return;
    collect_stage_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), ngroups, nsplit, 0);
    finish_segment_d<SRV>(P, F, &sm, 4, false, 1, 0.0, 0.0, sm.E, sm.Fl, 0, 1.0);
    if(threadIdx.x == 0 && sm.res[9] != 0.0) { sm.res[6] = sm.mN.a[0][0]; sm.res[7] = sm.mN.a[1][0]; sm.res[8] = sm.mN.a[2][0]; }
    __syncthreads();
    bool alive = sm.res[13] != 0.0;
    if(ms > 1 && alive)
    {
      finish_segment_d<SRV>(P, F, &sm, 4, true, no, F.u0, 0.0, sm.E + 6, sm.Fl + 1, 1, sm.res[14]);
      adopt_selection_d<SRV>(P, F, &sm, 4, true, off1, 1);
      alive = sm.res[16 + 13] != 0.0;
    }
    double* ew = part_half(F, 0) + GBF_PART_EWALD;
    if(F.do_ewald && alive) ewald_slice_d<SRV>(P, F, dyn, sm.red, &sm.exist, F.nold_ew, &sm.mN, F.nnew_ew, ew);
    if(blockIdx.x == 0)
    {
      if(alive)
      {
        const int g2 = 1 + no;
        finish_segment_d<SRV>(P, F, &sm, 5, false, 1, 0.0, 0.0, sm.E + 6 * g2, sm.Fl + g2, 2, 1.0);
        if(threadIdx.x == 0 && sm.res[32 + 9] != 0.0) { sm.res[32 + 6] = sm.exist.a[0][0]; sm.res[32 + 7] = sm.exist.a[1][0]; sm.res[32 + 8] = sm.exist.a[2][0]; }
        if(F.ms2 > 1) finish_segment_d<SRV>(P, F, &sm, 5, true, no2, 0.0, 0.0, sm.E + 6 * (g2 + 1), sm.Fl + g2 + 1, 3, sm.res[32 + 14]);
      }
      if(F.do_ewald) ewald_total_slot_d<SRV>(F, &sm, reinterpret_cast<double*>(dyn), ew, alive);
      export_molecule_d<SRV>(F, sm.mN, GBK_BUF_TEMP);            // tempMolStorage: what gb_accept_identity_swap commits
    }
  }
  else
  {
    // ---- translation / rotation: every CTA builds the proposal; items = (new | old) x atom slices; Ewald CTAs at the grid's end
    const int ne = F.do_ewald ? ewald_ctas(F) : 0;
    const int npair = max(1, F.ngrid - ne);          // CTAs that share the pair items
    const int nslice = max(1, min((F.natoms + 255) / 256, GBF_MAX_ITEMS / 2));
    const unsigned int tag0 = stage_tag(F.seq, 0);
    double* part = part_half(F, 0);
    if((int) blockIdx.x < npair)
    {
      for(int w = blockIdx.x; w < 2 * nslice; w += npair)
      {
        const int pass = w / nslice, slice = w % nslice;
        const SmemMol& M = pass == 0 ? sm.mN : sm.mO;
        __syncthreads();
        if((int) threadIdx.x < ms)
        {
          const int a = threadIdx.x;
          sm.T.fx[a] = M.a[3][a]; sm.T.fy[a] = M.a[4][a]; sm.T.fz[a] = M.a[5][a];
          sm.T.q[a] = M.a[6][a] * M.a[8][a]; sm.T.scale[a] = M.a[7][a]; sm.T.type[a] = M.type[a]; sm.T.slot[a] = 0;
        }
        __syncthreads();
        group_energy0_d<SRV>(P, W, S, F, &sm, (int) F.molecule, ms, slice, nslice, part + (size_t) w * 16, tag0);
      }
    }
    double* ew = part + GBF_PART_EWALD;
    if(F.do_ewald) ewald_slice_d<SRV>(P, F, dyn, sm.red, &sm.mO, ms, &sm.mN, ms, ew);
    if(blockIdx.x == 0)
    {
      // delta = sum(new) - sum(old) over the slices in fixed order (mc_single_particle.h:183-200); overlap of NEW only (:768-769)
      double* stash = reinterpret_cast<double*>(dyn);
      __syncthreads();
      fetch_records(part, 2 * nslice, tag0, stash);
      if(F.do_ewald) fetch_ewald(F, ew, stash + GBF_MAX_ITEMS * 8);
      __syncthreads();
      GBK_MARK();
      if(threadIdx.x < 7)
      {
        double n = 0.0, o = 0.0;
        for(int b = 0; b < nslice; b++) { n += stash[b * 8 + threadIdx.x]; o += stash[(nslice + b) * 8 + threadIdx.x]; }
        if(threadIdx.x < 6) sm.res[64 + threadIdx.x] = n - o;
        else { sm.res[64 + 6] = n > 0.0 ? 1.0 : 0.0; sm.res[64 + 7] = n > 0.0 ? 0.0 : 1.0; }
      }
      __syncthreads();
      if(threadIdx.x == 0)
      {
        const bool run = !F.check_overlap || sm.res[64 + 6] == 0.0;
        double sS = 0.0, cS = 0.0;
        if(run) for(int b = 0; b < ne; b++) { sS += stash[GBF_MAX_ITEMS * 8 + 2 * b]; cS += stash[GBF_MAX_ITEMS * 8 + 2 * b + 1]; }
        sm.res[16 * 5] = sS; sm.res[16 * 5 + 1] = 2.0 * cS;
      }
      __syncthreads();
      export_molecule_d<SRV>(F, sm.mN, GBK_BUF_NEW);
      export_molecule_d<SRV>(F, sm.mO, GBK_BUF_OLD);
    }
  }
  // publish: CTA 0 holds every result slot; it stores them to the host and then raises the sequence flag, so the host
  // needs neither a copy nor a stream synchronisation to read the outcome of the move
  if(blockIdx.x == 0)
  {
    __syncthreads();
    GBK_MARK();
    // every double goes out as one 16-byte store {lo, tag, hi, tag}: the host polls the tags, no system-scope fence needed
    if(threadIdx.x < 128)
    {
      const double v = sm.res[threadIdx.x];
      ll_store(F.host_result, threadIdx.x, v, (unsigned int) F.seq);
      F.B.result(0)[threadIdx.x] = v;
    }
  }
#ifdef GBK_PHASE_TIMING
  GBK_MARK();
  if(blockIdx.x == 0 && threadIdx.x == 0 && F.seq >= 20000 && F.seq < 20040)
  {
    printf("kind %d grid %d:", F.kind, F.ngrid);
    for(int i = 1; i < g_nmarks; i++) printf(" %lld", g_marks[i] - g_marks[i - 1]);
    printf("\n");
  }
#endif
}

__global__ void __launch_bounds__(256, 1)
k_move(DevParams P, SysView S, FusedArgs F)
{
  // dynamic shared memory: [FusedSmem | scratch: eik tables of the Ewald stage / stash of the collect steps (>= 12 KB)]
  extern __shared__ __align__(128) unsigned char dyn_all[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(dyn_all);
  unsigned char* dyn = dyn_all + ((sizeof(FusedSmem) + 127) / 128) * 128;
  move_body<false>(P, S, F, sm, dyn);
}

