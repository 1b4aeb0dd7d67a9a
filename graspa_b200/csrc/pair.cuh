// graspa_b200 -- warp-level pair loop: one warp evaluates one trial group (up to GBK_MAX_CS trial atoms) against
// ranges of system atoms.
//
// Replaces the per-pair thread of Calculate_Multiple_Trial_Energy_VDWReal (VDW_Coulomb.cu:1183-1352) and of
// Calculate_Single_Body_Energy_VDWReal (:626-841).  Differences in structure, not in result:
//   - lanes stride over system atoms (coalesced SoA loads from the TMA-staged shared-memory pack or from L2);
//     the trial atoms of the group are warp-uniform registers; one loaded system atom is reused for every trial
//     atom of the group;
//   - the distance test runs for every pair, but only pairs inside a cutoff are pushed (ballot + popc
//     compaction) into a per-warp shared-memory queue; the expensive LJ / erfc body is executed on full
//     32-entry batches taken from the queue tail, so the FP64 pipe does not idle on the ~85 % of lanes that
//     fail the cutoff;
//   - minimum image in fractional space with a magic-number round (3 DADD per axis, no F2I/I2F), cell matrix in
//     registers with the structural zeros of lower-triangular / orthorhombic cells removed at compile time;
//   - a group may feed NS (1 or 2) accumulator slots: the two atoms of a CO2 chain trial feed one slot, two
//     different first-bead trials evaluated together feed two.
#pragma once
#include "common.cuh"

// trial atoms of the group being evaluated, per warp, in shared memory
struct TrialGroup
{
  double fx[GBK_MAX_CS], fy[GBK_MAX_CS], fz[GBK_MAX_CS];   // fractional coordinates
  double q[GBK_MAX_CS];                                    // charge * scaleCoul
  double scale[GBK_MAX_CS];
  int    type[GBK_MAX_CS];
  int    slot[GBK_MAX_CS];                                 // accumulator slot of the atom (0 or 1)
};

struct WarpQueue
{
  double r2[GBK_QCAP];
  int    code[GBK_QCAP];     // (system atom index << 6) | trial atom
};

template <int NS>
struct PairAcc
{
  double vdw[NS], real[NS];
  int flag;                  // bit s: overlap seen in slot s
  __device__ __forceinline__ void clear() { for(int s = 0; s < NS; s++) { vdw[s] = 0.0; real[s] = 0.0; } flag = 0; }
};

// tables a pair loop needs besides the atoms: erfc polynomial table (shared memory), LJ table (shared copy or
// P.ffA), and the warp-uniform "every scaling factor is 1" flag
struct PairTables { const double* etab; const double4* ffp; bool unit; };

// where the system atoms of a range live
template <bool STAGED>
struct SysAccess
{
  const double* __restrict__ fx; const double* __restrict__ fy; const double* __restrict__ fz; const double* __restrict__ q;
  const double* __restrict__ scale; const double* __restrict__ scoul;
  const int* __restrict__ type; const int* __restrict__ molid;
  __device__ __forceinline__ void hint() const
  {
    if(STAGED) { GBK_ASSUME_SHARED(fx); GBK_ASSUME_SHARED(fy); GBK_ASSUME_SHARED(fz); GBK_ASSUME_SHARED(q); GBK_ASSUME_SHARED(type); }
    else { GBK_ASSUME_GLOBAL(fx); GBK_ASSUME_GLOBAL(fy); GBK_ASSUME_GLOBAL(fz); GBK_ASSUME_GLOBAL(q); GBK_ASSUME_GLOBAL(type); }
  }
};

template <bool STAGED>
__device__ __forceinline__ SysAccess<STAGED> make_access(const SysView& S)
{
  SysAccess<STAGED> A; A.fx = S.fx; A.fy = S.fy; A.fz = S.fz; A.q = S.q; A.scale = S.scale; A.scoul = S.scoul; A.type = S.type; A.molid = S.molid;
  return A;
}

// the expensive body: lanes [0, n) take queue entries [off, off+n)
template <int NS, bool STAGED>
__device__ __forceinline__ void drain_queue(const DevParams& P, const PairTables& W, const SysAccess<STAGED>& S,
                                            const TrialGroup* T, const WarpQueue* Q, int lane, int off, int n, PairAcc<NS>& acc)
{
  if(lane < n)
  {
    const double r2 = Q->r2[off + lane];
    const int code = Q->code[off + lane];
    const int i = code >> 6, a = code & 63;
    const int row = S.type[i] * P.ntypes + T->type[a];
    double scaling = T->scale[a], qq = S.q[i] * T->q[a];      // staged pack: q already holds charge*scaleCoul
    if(!STAGED && !P.all_unit_scale) { scaling *= S.scale[i]; qq *= S.scoul[i]; }
    double ev, er; int fl;
    pair_energy(P, W.etab, W.ffp, W.unit, r2, row, scaling, qq, ev, er, fl);
    if(NS == 1) { acc.vdw[0] += ev; acc.real[0] += er; acc.flag |= fl; }
    else
    {
      const int s = T->slot[a];
      acc.vdw[0] += (s == 0) ? ev : 0.0; acc.real[0] += (s == 0) ? er : 0.0;
      acc.vdw[1] += (s == 1) ? ev : 0.0; acc.real[1] += (s == 1) ? er : 0.0;
      acc.flag |= fl << s;
    }
  }
}

// One range of system atoms [start, end) against the trial atoms in *T.
// excl_a / excl_b: molecule ids to skip in this range (-1: none) -- VDW_Coulomb.cu:1282-1283.
// wslice/nslice: this warp handles iterations wslice, wslice+nslice, ... (nslice = 1: the whole range).
template <int CS, int NS, int CELL, bool STAGED, bool EXCL>
__device__ __forceinline__ void pair_range(const DevParams& P, const PairTables& W, const SysAccess<STAGED>& S,
                                           int start, int end, int excl_a, int excl_b, const TrialGroup* T, int cs_dyn,
                                           WarpQueue* Q, int wslice, int nslice, PairAcc<NS>& acc)
{
  S.hint();
  const int lane = (int) lane_id();
  const unsigned lt_mask = (1u << lane) - 1u;
  const int cs = CS > 0 ? CS : cs_dyn;
  const double cut_max = P.no_charges ? P.cut_vdw2 : fmax(P.cut_vdw2, P.cut_coul2);
  CellRegs<CELL> C; C.load(P);
  constexpr int NR = (CS > 0 && CS <= 3) ? CS : 1;
  double tx[NR], ty[NR], tz[NR];
  if(CS > 0 && CS <= 3)
  {
#pragma unroll
    for(int a = 0; a < NR; a++) { tx[a] = T->fx[a]; ty[a] = T->fy[a]; tz[a] = T->fz[a]; }
  }
  int qn = 0;
  for(int base = start + 32 * wslice; base < end; base += 32 * nslice)
  {
    const int i = base + lane;
    bool valid = i < end;
    const int ii = STAGED ? i : (valid ? i : end - 1);        // the staged pack is padded: unconditional loads
    const double ax = S.fx[ii], ay = S.fy[ii], az = S.fz[ii];
    if(EXCL) { const int m = S.molid[ii]; valid = valid && (m != excl_a) && (m != excl_b); }
    if(CS > 0 && CS <= 3)
    {
      // all distances of the group first (independent FP64 chains), then one compaction pass and one drain check
      double r2v[NR]; unsigned mv[NR];
#pragma unroll
      for(int a = 0; a < NR; a++) r2v[a] = C.r2(ax - tx[a], ay - ty[a], az - tz[a]);
#pragma unroll
      for(int a = 0; a < NR; a++) mv[a] = __ballot_sync(0xffffffffu, valid && (r2v[a] < cut_max));
#pragma unroll
      for(int a = 0; a < NR; a++)
      {
        if((mv[a] >> lane) & 1u) { const int p = qn + __popc(mv[a] & lt_mask); Q->r2[p] = r2v[a]; Q->code[p] = (i << 6) | a; }
        qn += __popc(mv[a]);
      }
      if(qn >= 32)
      {
        __syncwarp();
        do { qn -= 32; drain_queue<NS, STAGED>(P, W, S, T, Q, lane, qn, 32, acc); } while(qn >= 32);
        __syncwarp();
      }
    }
    else
    {
      for(int a = 0; a < cs; a++)
      {
        const double r2 = C.r2(ax - T->fx[a], ay - T->fy[a], az - T->fz[a]);
        const bool hit = valid && (r2 < cut_max);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if(hit) { const int p = qn + __popc(m & lt_mask); Q->r2[p] = r2; Q->code[p] = (i << 6) | a; }
        qn += __popc(m);
        if(qn >= GBK_QCAP - 32)
        {
          __syncwarp();
          do { qn -= 32; drain_queue<NS, STAGED>(P, W, S, T, Q, lane, qn, 32, acc); } while(qn >= 32);
          __syncwarp();
        }
      }
    }
  }
  if(qn > 0)
  {
    __syncwarp();
    while(qn >= 32) { qn -= 32; drain_queue<NS, STAGED>(P, W, S, T, Q, lane, qn, 32, acc); }
    if(qn > 0) drain_queue<NS, STAGED>(P, W, S, T, Q, lane, 0, qn, acc);
    __syncwarp();
  }
}

// all segments for one trial group, generic (cold) flavour: global memory, general cell, exclusions checked.
// e6 = lane-partial sums {HHv, HHr, HGv, HGr, GGv, GGr}; kinds come from L.kind[].
template <int CS>
__device__ __forceinline__ void pair_group_generic(const DevParams& P, const PairTables& W, const SysView& Sg, const SegList& L,
                                                   int new_comp, int new_molid, int excl_comp, int excl_mol,
                                                   const TrialGroup* T, int cs_dyn, WarpQueue* Q, int wslice, int nslice,
                                                   double* e6, int& flag)
{
  const SysAccess<false> S = make_access<false>(Sg);
  for(int s = 0; s < L.nseg; s++)
  {
    PairAcc<1> acc; acc.clear();
    const int ea = (L.comp[s] == excl_comp) ? excl_mol : -1;
    const int eb = (L.comp[s] == new_comp) ? new_molid : -1;
    pair_range<CS, 1, 0, false, true>(P, W, S, L.start[s], L.start[s] + L.count[s], ea, eb, T, cs_dyn, Q, wslice, nslice, acc);
    const int k = L.kind[s];
    e6[0] += (k == 0) ? acc.vdw[0] : 0.0; e6[1] += (k == 0) ? acc.real[0] : 0.0;
    e6[2] += (k == 1) ? acc.vdw[0] : 0.0; e6[3] += (k == 1) ? acc.real[0] : 0.0;
    e6[4] += (k == 2) ? acc.vdw[0] : 0.0; e6[5] += (k == 2) ? acc.real[0] : 0.0;
    flag |= acc.flag;
  }
}

// ---------------------------------------------------------------------------------------------
// Latency-oriented flavour for the single-move kernels, where one warp sees only a handful of system atoms and
// what costs is the number of DEPENDENT memory round trips, not arithmetic:
//   * the live ranges are walked as ONE concatenated index space (one pass, not one pass per range);
//   * type, charge and scaling factors of a system atom are fetched in the same round trip as its position and
//     travel through the queue, so the drain needs no second trip to global memory.
// ---------------------------------------------------------------------------------------------
struct WarpQueueG
{
  double r2[GBK_QCAP];
  double qe[GBK_QCAP];       // charge * scaleCoul of the system atom
  double sc[GBK_QCAP];       // its LJ scaling factor
  int    code[GBK_QCAP];     // trial atom | type << 6 | kind << 20
};

__device__ __forceinline__ void drain_flat(const DevParams& P, const PairTables& W, const TrialGroup* T, const WarpQueueG* Q,
                                           int lane, int off, int n, double* e6, int& flag)
{
  if(lane < n)
  {
    const double r2 = Q->r2[off + lane];
    const int code = Q->code[off + lane];
    const int a = code & 63, type = (code >> 6) & 0x3fff, kind = code >> 20;
    const int row = type * P.ntypes + T->type[a];
    double scaling = T->scale[a];
    if(!P.all_unit_scale) scaling *= Q->sc[off + lane];
    const double qq = Q->qe[off + lane] * T->q[a];
    double ev, er; int fl;
    pair_energy(P, W.etab, W.ffp, W.unit, r2, row, scaling, qq, ev, er, fl);
    e6[0] += (kind == 0) ? ev : 0.0; e6[1] += (kind == 0) ? er : 0.0;
    e6[2] += (kind == 1) ? ev : 0.0; e6[3] += (kind == 1) ? er : 0.0;
    e6[4] += (kind == 2) ? ev : 0.0; e6[5] += (kind == 2) ? er : 0.0;
    flag |= fl;
  }
}

// e6 = lane-partial sums {HHv, HHr, HGv, HGr, GGv, GGr}; atoms of molecule new_molid of component new_comp are skipped
template <int CS>
__device__ __forceinline__ void pair_group_flat(const DevParams& P, const PairTables& W, const SysView& S, const SegList& L,
                                                int new_comp, int new_molid, const TrialGroup* T, int cs_dyn, WarpQueueG* Q,
                                                int wslice, int nslice, double* e6, int& flag)
{
  const int lane = (int) lane_id();
  const unsigned lt_mask = (1u << lane) - 1u;
  const int cs = CS > 0 ? CS : cs_dyn;
  const double cut_max = P.no_charges ? P.cut_vdw2 : fmax(P.cut_vdw2, P.cut_coul2);
  CellRegs<0> C; C.load(P);
  int total = 0;
  for(int s = 0; s < L.nseg; s++) total += L.count[s];
  int qn = 0;
  for(int base = 32 * wslice; base < total; base += 32 * nslice)
  {
    const int v = base + lane;
    bool valid = v < total;
    int off = valid ? v : total - 1, s = 0;
    while(s + 1 < L.nseg && off >= L.count[s]) { off -= L.count[s]; s++; }
    const int i = L.start[s] + off;
    const double ax = S.fx[i], ay = S.fy[i], az = S.fz[i];
    const int m = S.molid[i], type = S.type[i];
    double qe = S.q[i], sc = 1.0;
    if(!P.all_unit_scale) { qe *= S.scoul[i]; sc = S.scale[i]; }
    valid = valid && !(L.comp[s] == new_comp && m == new_molid);
    const int tag = (type << 6) | (L.kind[s] << 20);
    for(int a = 0; a < cs; a++)
    {
      const double r2 = C.r2(ax - T->fx[a], ay - T->fy[a], az - T->fz[a]);
      const bool hit = valid && (r2 < cut_max);
      const unsigned mk = __ballot_sync(0xffffffffu, hit);
      if(hit) { const int p = qn + __popc(mk & lt_mask); Q->r2[p] = r2; Q->qe[p] = qe; Q->sc[p] = sc; Q->code[p] = tag | a; }
      qn += __popc(mk);
      if(CS == 0 && qn >= GBK_QCAP - 32)
      {
        __syncwarp();
        do { qn -= 32; drain_flat(P, W, T, Q, lane, qn, 32, e6, flag); } while(qn >= 32);
        __syncwarp();
      }
    }
    if(CS > 0 && qn >= GBK_QCAP - 32 * CS)
    {
      __syncwarp();
      do { qn -= 32; drain_flat(P, W, T, Q, lane, qn, 32, e6, flag); } while(qn >= 32);
      __syncwarp();
    }
  }
  if(qn > 0)
  {
    __syncwarp();
    while(qn >= 32) { qn -= 32; drain_flat(P, W, T, Q, lane, qn, 32, e6, flag); }
    if(qn > 0) drain_flat(P, W, T, Q, lane, 0, qn, e6, flag);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Tile-culled pair loop over the staged framework pack.
// 92 % of the nominal pair flops are minimum-image distance checks against atoms that are nowhere near the trial atom.
// The pack is sorted by a k-d split into 32-atom tiles (host side, ensure_pack); each tile carries its fractional
// bounding box and a Cartesian bounding sphere.  One lane tests one tile:
//   * per fractional axis, the wrapped interval [lo - t, hi - t] must reach into [-w, w], w = r_cut * |column of the
//     inverse cell| -- a necessary condition for ANY atom of the tile to be inside the cutoff under the reference's
//     convention (maths.cuh:437-448: every fractional difference is wrapped independently);
//   * if the whole interval wraps with ONE integer shift on every axis, the tile is "uniform": the reference's image
//     of every atom of the tile is atom - n.H, so the distance is a plain Cartesian difference against the shifted
//     trial position t' = t + n.H (6 FP64 instructions per pair instead of 21) and the bounding sphere is tested
//     against t' exactly;
//   * tiles that straddle a half-box boundary take the general path (Cartesian difference -> fractional -> wrap).
// Results differ from the untiled loop by summation order and by the rounding of Cartesian versus fractional
// differences (1e-16 relative).
// ---------------------------------------------------------------------------------------------
// bytes of the staged pack: 4 double arrays + 1 int array of npad entries, then 10 tile arrays of ntp entries
__host__ __device__ inline size_t gbk_pack_bytes(int npad, int ntp) { return ((size_t) npad * 36 + (size_t) ntp * 80 + 15) / 16 * 16; }

struct TileView
{
  const double* pack;     // [x | y | z | q*scoul | type(int)] each npad long (Cartesian, wrapped into the primary cell)
  const double* tile;     // [lo x,y,z | hi x,y,z | centre x,y,z | radius] each ntp long
  int npad, ntp, n, ntiles;
  __device__ __forceinline__ double x(int i) const { return pack[i]; }
  __device__ __forceinline__ double y(int i) const { return pack[npad + i]; }
  __device__ __forceinline__ double z(int i) const { return pack[2 * npad + i]; }
  __device__ __forceinline__ double q(int i) const { return pack[3 * npad + i]; }
  __device__ __forceinline__ int type(int i) const { return reinterpret_cast<const int*>(pack + 4 * (size_t) npad)[i]; }
  __device__ __forceinline__ double lo(int k, int t) const { return tile[k * ntp + t]; }
  __device__ __forceinline__ double hi(int k, int t) const { return tile[(3 + k) * ntp + t]; }
  __device__ __forceinline__ double c(int k, int t) const { return tile[(6 + k) * ntp + t]; }
  __device__ __forceinline__ double rad(int t) const { return tile[9 * ntp + t]; }
};

template <int CELL>
struct InvRegs
{
  double i0, i1, i2, i3, i4, i5, i6, i7, i8;
  __device__ __forceinline__ void load(const DevParams& P)
  {
    i0 = P.inv[0]; i1 = P.inv[1]; i2 = P.inv[2]; i3 = P.inv[3]; i4 = P.inv[4]; i5 = P.inv[5]; i6 = P.inv[6]; i7 = P.inv[7]; i8 = P.inv[8];
  }
  __device__ __forceinline__ void frac(double dx, double dy, double dz, double& sx, double& sy, double& sz) const
  {
    if(CELL == 2) { sx = i0 * dx; sy = i4 * dy; sz = i8 * dz; }
    else if(CELL == 1) { sx = i0 * dx + i3 * dy + i6 * dz; sy = i4 * dy + i7 * dz; sz = i8 * dz; }
    else { sx = i0 * dx + i3 * dy + i6 * dz; sy = i1 * dx + i4 * dy + i7 * dz; sz = i2 * dx + i5 * dy + i8 * dz; }
  }
};

__device__ __forceinline__ double round_magic(double d)
{
  const double M = 6755399441055744.0;
  return __dsub_rn(__dadd_rn(d, M), M);
}

// drain for the tile path: queue codes are (pack index << 6) | trial atom, type and charge come from the pack
__device__ __forceinline__ void drain_tiles(const DevParams& P, const PairTables& W, const TileView& V, const TrialGroup* T, const WarpQueue* Q,
                                            int lane, int off, int n, PairAcc<1>& acc)
{
  if(lane < n)
  {
    const double r2 = Q->r2[off + lane];
    const int code = Q->code[off + lane];
    const int i = code >> 6, a = code & 63;
    const int row = V.type(i) * P.ntypes + T->type[a];
    const double qq = V.q(i) * T->q[a];
    double ev, er; int fl;
    pair_energy(P, W.etab, W.ffp, W.unit, r2, row, T->scale[a], qq, ev, er, fl);
    acc.vdw[0] += ev; acc.real[0] += er; acc.flag |= fl;
  }
}

// one trial atom `a` (fractional tf*, Cartesian tc*) against the whole pack
template <int CELL>
__device__ __forceinline__ void pair_tiles_atom(const DevParams& P, const PairTables& W, const TileView& V, const TrialGroup* T, int a,
                                                double tfx, double tfy, double tfz, double tcx, double tcy, double tcz,
                                                WarpQueue* Q, int& qn, PairAcc<1>& acc)
{
  GBK_ASSUME_SHARED(V.pack); GBK_ASSUME_SHARED(V.tile);
  const int lane = (int) lane_id();
  const unsigned lt_mask = (1u << lane) - 1u;
  const double cut_max = P.cull_rcut * P.cull_rcut;
  CellRegs<CELL> C; C.load(P);
  for(int tb = 0; tb < V.ntiles; tb += 32)
  {
    const int t = tb + lane;
    bool visit = false, fast = false;
    double px = 0.0, py = 0.0, pz = 0.0;       // shifted trial position t' of a uniform tile
    if(t < V.ntiles)
    {
      const double ax = V.lo(0, t) - tfx, bx = V.hi(0, t) - tfx;
      const double ay = V.lo(1, t) - tfy, by = V.hi(1, t) - tfy;
      const double az = V.lo(2, t) - tfz, bz = V.hi(2, t) - tfz;
      const double nax = round_magic(ax), nbx = round_magic(bx), nay = round_magic(ay), nby = round_magic(by), naz = round_magic(az), nbz = round_magic(bz);
      const bool ux = nax == nbx, uy = nay == nby, uz = naz == nbz;
      const bool okx = ux ? ((ax - nax <= P.cull_w[0]) && (bx - nax >= -P.cull_w[0])) : ((ax - nax <= P.cull_w[0]) || (bx - nbx >= -P.cull_w[0]));
      const bool oky = uy ? ((ay - nay <= P.cull_w[1]) && (by - nay >= -P.cull_w[1])) : ((ay - nay <= P.cull_w[1]) || (by - nby >= -P.cull_w[1]));
      const bool okz = uz ? ((az - naz <= P.cull_w[2]) && (bz - naz >= -P.cull_w[2])) : ((az - naz <= P.cull_w[2]) || (bz - nbz >= -P.cull_w[2]));
      visit = okx && oky && okz;
      fast = ux && uy && uz;
      if(visit && fast)
      {
        // t' = t + n.H (rows of the cell are the lattice vectors)
        px = tcx + (C.c0 * nax + C.c3 * nay + C.c6 * naz);
        py = tcy + (C.c1 * nax + C.c4 * nay + C.c7 * naz);
        pz = tcz + (C.c2 * nax + C.c5 * nay + C.c8 * naz);
        const double dx = V.c(0, t) - px, dy = V.c(1, t) - py, dz = V.c(2, t) - pz;
        const double reach = P.cull_rcut + V.rad(t);
        visit = (dx * dx + dy * dy + dz * dz) <= reach * reach * (1.0 + 1e-12);
      }
    }
    unsigned mfast = __ballot_sync(0xffffffffu, visit && fast);
    unsigned mslow = __ballot_sync(0xffffffffu, visit && !fast);
    for(; mfast; mfast &= mfast - 1)
    {
      const int l = __ffs(mfast) - 1;
      const double sx = __shfl_sync(0xffffffffu, px, l), sy = __shfl_sync(0xffffffffu, py, l), sz = __shfl_sync(0xffffffffu, pz, l);
      const int i = (tb + l) * 32 + lane;
      const double dx = V.x(i) - sx, dy = V.y(i) - sy, dz = V.z(i) - sz;
      const double r2 = dx * dx + dy * dy + dz * dz;
      const bool hit = r2 < cut_max;                       // padded atoms sit at 1e30
      const unsigned mk = __ballot_sync(0xffffffffu, hit);
      if(hit) { const int p = qn + __popc(mk & lt_mask); Q->r2[p] = r2; Q->code[p] = (i << 6) | a; }
      qn += __popc(mk);
      if(qn >= 32)
      {
        __syncwarp();
        qn -= 32; drain_tiles(P, W, V, T, Q, lane, qn, 32, acc);
        __syncwarp();
      }
    }
    if(mslow)
    {
      InvRegs<CELL> I; I.load(P);
      for(; mslow; mslow &= mslow - 1)
      {
        const int l = __ffs(mslow) - 1;
        const int i = (tb + l) * 32 + lane;
        double sx, sy, sz;
        I.frac(V.x(i) - tcx, V.y(i) - tcy, V.z(i) - tcz, sx, sy, sz);
        const double r2 = C.r2(sx, sy, sz);
        const bool hit = (i < V.n) && (r2 < cut_max);
        const unsigned mk = __ballot_sync(0xffffffffu, hit);
        if(hit) { const int p = qn + __popc(mk & lt_mask); Q->r2[p] = r2; Q->code[p] = (i << 6) | a; }
        qn += __popc(mk);
        if(qn >= 32)
        {
          __syncwarp();
          qn -= 32; drain_tiles(P, W, V, T, Q, lane, qn, 32, acc);
          __syncwarp();
        }
      }
    }
  }
}

// every trial atom of the group against the pack; the queue is drained at the end
template <int CELL>
__device__ __forceinline__ void pair_tiles_group(const DevParams& P, const PairTables& W, const TileView& V, const TrialGroup* T, int cs,
                                                 const double* tc /* Cartesian [cs][3] */, WarpQueue* Q, PairAcc<1>& acc)
{
  int qn = 0;
  for(int a = 0; a < cs; a++)
    pair_tiles_atom<CELL>(P, W, V, T, a, T->fx[a], T->fy[a], T->fz[a], tc[3 * a], tc[3 * a + 1], tc[3 * a + 2], Q, qn, acc);
  if(qn > 0)
  {
    __syncwarp();
    drain_tiles(P, W, V, T, Q, (int) lane_id(), 0, qn, acc);
    __syncwarp();
  }
}
