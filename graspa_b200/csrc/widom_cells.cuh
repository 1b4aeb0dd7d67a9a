// graspa_b200 -- batched Widom insertions, cell-sorted pair stage (included by engine.cu).
//
// Same arithmetic per pair as k_widom_pair (pair_energy, common.cuh) and the same stage logic (Widom_Move_FirstBead_PARTIAL /
// Widom_Move_Chain_PARTIAL, mc_widom.h:385-614), but the work is laid out for the batch instead of for one insertion:
//
//   * every trial atom of the batch (n x NumberOfTrials first beads, then n x NumberOfTrialOrientations x (Molsize-1) chain atoms)
//     is binned into a grid of small cells of the box (counting sort on the device: histogram, scan, scatter);
//   * one CTA takes one cell at a time and builds, ONCE for all trial atoms of the cell, the list of system atoms that can be
//     inside the cutoff of any point of the cell: atoms within r_cut + r_cell of the cell centre, stored as Cartesian vectors
//     from the centre with the periodic image already chosen (the reference's per-axis nearest image, maths.cuh:437-448, is the
//     same for every point of the cell unless the atom sits within the cell's extent of a half-box plane; those few atoms go to
//     a second list that keeps the general wrap);
//   * a warp then evaluates one trial atom against the list: 3 shared-memory loads, 3 subtractions and a dot product per
//     candidate -- no minimum image, no tile tests, no shuffles -- and ~70 % of the candidates are inside the cutoff, so the
//     LJ / erfc body runs in place without a compaction queue.
//
// Against the warp-per-insertion kernel (k_widom_pair: k-d tiles tested per trial atom, 1300 distance tests for 450 pairs, 37 % of
// them on the general minimum-image path in config E) this is ~650 distance tests per trial atom at a third of the instructions
// each.  Results per trial atom do not depend on what else is in the batch (list order = atom order), so batches can be cut
// anywhere without changing a bit.
#pragma once
#include "common.cuh"
#include "pair.cuh"
#include "misc_kernels.cuh"
#include "move_kernels.cuh"

struct WcGrid
{
  int n[3]; int ncells;
  double inv_n[3];          // 1 / n_k
  double margin[3];         // half the fractional extent of a cell + guard: a trial of the cell is within margin of the centre, per axis
  double rcell;             // largest Cartesian distance from a cell centre to a point of its cell
  int cap_fast, cap_slow;   // list capacities (shared memory)
  int chunk;                // trial atoms per work item
};

// live atoms of every component, gathered contiguously (k_wc_pack): fractional coordinates, charge * scaleCoul, type | kind << 16
struct WcAtoms { const double* __restrict__ fx; const double* __restrict__ fy; const double* __restrict__ fz; const double* __restrict__ q; const int* __restrict__ tk; int n; };

__global__ void k_wc_pack(SysView S, SegList L, int ntot, double* fx, double* fy, double* fz, double* q, int* tk)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if(v >= ntot) return;
  int off = v, s = 0;
  while(s + 1 < L.nseg && off >= L.count[s]) { off -= L.count[s]; s++; }
  const int i = L.start[s] + off;
  fx[v] = S.fx[i]; fy[v] = S.fy[i]; fz[v] = S.fz[i]; q[v] = S.q[i] * S.scoul[i];
  tk[v] = S.type[i] | ((L.kind[s] == 2 ? 1 : 0) << 16);
}

// cell of a Cartesian point and the Cartesian vector from the cell centre to (the image in the primary cell of) the point
__device__ __forceinline__ void wc_bin(const DevParams& P, const WcGrid& G, double x, double y, double z, int& cell, double& dx, double& dy, double& dz)
{
  double fx, fy, fz; to_frac(P, x, y, z, fx, fy, fz);
  fx -= floor(fx); fy -= floor(fy); fz -= floor(fz);
  int ix = (int) (fx * G.n[0]), iy = (int) (fy * G.n[1]), iz = (int) (fz * G.n[2]);
  ix = min(max(ix, 0), G.n[0] - 1); iy = min(max(iy, 0), G.n[1] - 1); iz = min(max(iz, 0), G.n[2] - 1);
  const double ex = fx - ((double) ix + 0.5) * G.inv_n[0], ey = fy - ((double) iy + 0.5) * G.inv_n[1], ez = fz - ((double) iz + 0.5) * G.inv_n[2];
  dx = P.cell[0] * ex + P.cell[3] * ey + P.cell[6] * ez;
  dy = P.cell[1] * ex + P.cell[4] * ey + P.cell[7] * ez;
  dz = P.cell[2] * ex + P.cell[5] * ey + P.cell[8] * ez;
  cell = (ix * G.n[1] + iy) * G.n[2] + iz;
}

// ---------------------------------------------------------------------------------------------- first-bead trial positions
struct WcGen
{
  const double* __restrict__ pool3; const long long* __restrict__ fb_index;
  long long n; int ntrials, norient;
  int* ucell; double* udelta; int* count;
};

// BoxLength o random (mc_widom.h:137-138), binned.  One thread per (insertion, trial).
__global__ void k_wc_gen_fb(DevParams P, WcGrid G, WcGen A)
{
  const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= A.n * A.ntrials) return;
  const long long ins = g / A.ntrials; const int t = (int) (g - ins * A.ntrials);
  const long long fb_off = A.fb_index ? A.fb_index[ins] : ins * (A.ntrials + A.norient);
  const double* r = A.pool3 + 3 * (fb_off + t);
  int cell; double dx, dy, dz;
  wc_bin(P, G, P.cell[0] * r[0], P.cell[4] * r[1], P.cell[8] * r[2], cell, dx, dy, dz);
  A.ucell[g] = cell; A.udelta[3 * g] = dx; A.udelta[3 * g + 1] = dy; A.udelta[3 * g + 2] = dz;
  atomicAdd(&A.count[cell], 1);
}

// ---------------------------------------------------------------------------------------------- counting sort: scan + work items
// one block: exclusive scan of the cell counts -> off[]; work items {cell, first sorted index, count} of at most `chunk` trial atoms
__global__ void __launch_bounds__(1024)
k_wc_scan(const int* __restrict__ count, int ncells, int chunk, int* off, int* cursor, int* items /* 3 ints each */, int* ctl /* [0] = number of items, [1] = work counter */)
{
  __shared__ int wa[32], wb[32];
  __shared__ int carry[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if(threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
  __syncthreads();
  for(int base = 0; base < ncells; base += 1024)
  {
    const int c = base + threadIdx.x;
    const int cnt = c < ncells ? count[c] : 0;
    const int nch = (cnt + chunk - 1) / chunk;
    int a = cnt, b = nch;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const int ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
      if(lane >= o) { a += ta; b += tb; }
    }
    if(lane == 31) { wa[warp] = a; wb[warp] = b; }
    __syncthreads();
    if(warp == 0)
    {
      int x = wa[lane], y = wb[lane];
#pragma unroll
      for(int o = 1; o < 32; o <<= 1)
      {
        const int tx = __shfl_up_sync(0xffffffffu, x, o), ty = __shfl_up_sync(0xffffffffu, y, o);
        if(lane >= o) { x += tx; y += ty; }
      }
      wa[lane] = x; wb[lane] = y;
    }
    __syncthreads();
    const int exa = carry[0] + (warp ? wa[warp - 1] : 0) + a - cnt;
    const int exb = carry[1] + (warp ? wb[warp - 1] : 0) + b - nch;
    if(c < ncells)
    {
      off[c] = exa; cursor[c] = 0;
      for(int k = 0; k < nch; k++) { items[3 * (exb + k)] = c; items[3 * (exb + k) + 1] = exa + k * chunk; items[3 * (exb + k) + 2] = min(chunk, cnt - k * chunk); }
    }
    __syncthreads();
    if(threadIdx.x == 0) { carry[0] += wa[31]; carry[1] += wb[31]; }
    __syncthreads();
  }
  if(threadIdx.x == 0) { off[ncells] = carry[0]; ctl[0] = carry[1]; ctl[1] = 0; }
}

// sorted records {dx, dy, dz, id}: the order inside a cell is whatever the atomics give; no result depends on it
__global__ void k_wc_scatter(const int* __restrict__ ucell, const double* __restrict__ udelta, long long ntot, const int* __restrict__ off, int* cursor, double4* srec)
{
  const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= ntot) return;
  const int c = ucell[g];
  if(c < 0) return;
  const int p = off[c] + atomicAdd(&cursor[c], 1);
  srec[p] = make_double4(udelta[3 * g], udelta[3 * g + 1], udelta[3 * g + 2], __longlong_as_double(g));
}

// ---------------------------------------------------------------------------------------------- pair energies, cell by cell
struct WcEnergy
{
  WcAtoms atoms;
  const double4* __restrict__ srec; const int* __restrict__ items; int* ctl;
  // trial atom a of the molecule: type and charge * scaleCoul (template = slot 0 of the component, mc_widom.h:256)
  const double* __restrict__ tq; const double* __restrict__ tscoul; const int* __restrict__ ttype; int ms;
  int amod, abase;          // trial atom of record id: abase + id % amod  (first beads: 0; chain atoms: 1 + id % (ms - 1))
  int stage_ff;
  double* e4; int* flag;    // per record id: {HGVDW, HGReal, GGVDW, GGReal}, overlap
  int* overflow;            // set when a list does not fit its capacity (the caller then falls back to k_widom_pair)
};

__host__ __device__ inline size_t wc_energy_smem(int ntypes, bool stage_ff, int cap_fast, int cap_slow)
{
  size_t b = GBK_SMEM_TABLES_OFF + GBK_ERFC_BYTES_PAD + (stage_ff ? (size_t) ntypes * ntypes * 32 : 0);
  b += 64 * 8 + 64 * 4 + 64 * 4;                                  // trial atom tables + warp counters
  b = (b + 15) / 16 * 16;
  b += (size_t) (cap_fast + cap_slow) * 36 + 64;
  return (b + 15) / 16 * 16;
}

// The pair body of the common case, without the run-time switches of pair_energy (common.cuh): plain 12-6 LJ with unit scaling factors,
// the LJ table and the erfc table in shared memory (LDS, not generic loads), every in-cutoff argument inside the erfc table.
// FAST 1: with real-space Coulomb, FAST 2: a system without charges, FAST 3: as 1 with CutOffVDW == CutOffCoul (every pair the caller
// found inside the cutoff gets both terms: no range tests); FAST 4 / 5: as 1 / 3 with the short erfc table (P.erfc10_ok).
// Same expressions as pair_energy's unit path.
__device__ __forceinline__ double lds_f64(uint32_t addr, int byte_off)
{
  double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr + (uint32_t) byte_off)); return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
  double2 v; asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); return v;
}

// erfc_table_eval (common.cuh) with the table given by its 32-bit shared-memory address: plain LDS with immediate offsets, no
// generic-to-shared address arithmetic per pair.  Same polynomial, same order.
// DEG / NINT: the table of common.cuh (12 / 49, arguments up to 6.06) or the short one (10 / 28, arguments up to 3.4375: as accurate
// there -- 2.3e-16 relative -- with two coefficient gathers and two multiply-adds fewer per pair; the gathers are what binds the kernel).
template <int DEG, int NINT>
__device__ __forceinline__ double erfc_table_eval_s(uint32_t tab, double x)
{
  const double M = 6755399441055744.0;
  const double y = x * GBK_ERFC_SCALE;
  double kd = __dadd_rn(y, M);
  const int k = __double2loint(kd);
  kd = __dsub_rn(kd, M);
  const double t = y - kd;
  const uint32_t a = tab + 8u * (uint32_t) k;
  double acc = lds_f64(a, 8 * NINT * DEG);
#pragma unroll
  for(int j = DEG - 1; j >= 0; j--) acc = fma(acc, t, lds_f64(a, 8 * NINT * j));
  return acc;
}

template <int FAST>
__device__ __forceinline__ void pair_energy_fast(const DevParams& P, uint32_t etab_s, uint32_t ff_s,
                                                 double r2, int row, double qq, double& e_vdw, double& e_real, int& flag)
{
  double rinv = rsqrt(r2);
  asm volatile("" : "+d"(rinv));                       // ONE reciprocal square root for both terms (the compiler otherwise sinks a copy into each branch)
  const double rinv2 = rinv * rinv;
  e_vdw = 0.0; e_real = 0.0; flag = 0;
  constexpr bool SAMECUT = (FAST == 3 || FAST == 5), SHORT_TABLE = (FAST == 4 || FAST == 5);
  if(SAMECUT || r2 < P.cut_vdw2)
  {
    const double2 f01 = lds_f64x2(ff_s + 32u * (uint32_t) row);            // {4 eps, sigma^2}
    const double fz = lds_f64(ff_s + 32u * (uint32_t) row, 16);            // shift
    const double x = f01.y * rinv2; const double rri3 = x * x * x;
    const double e = f01.x * (rri3 * (rri3 - 1.0)) - fz;
    if(e > P.overlap) flag = 1;
    if(r2 < 0.01) flag = 1;
    e_vdw = e;
  }
  if(SAMECUT || ((FAST == 1 || FAST == 4) && r2 < P.cut_coul2))
  {
    const double r = r2 * rinv;
    const double ec = SHORT_TABLE ? erfc_table_eval_s<GBK_ERFC10_DEG, GBK_ERFC10_NINT>(etab_s, P.alpha * r)
                                  : erfc_table_eval_s<GBK_ERFC_DEG, GBK_ERFC_NINT>(etab_s, P.alpha * r);
    e_real = P.prefactor * qq * ec * rinv;
  }
}

// FAST 0: pair_energy with all its switches (12-6-4 potential, LJ table in global memory, arguments beyond the erfc table)
template <int FAST>
__device__ __forceinline__ void wc_pair(const DevParams& P, const double* __restrict__ etab, const double4* __restrict__ ffp, uint32_t etab_s, uint32_t ff_s,
                                        double r2, int row, double qq, double& ev, double& er, int& f)
{
  if(FAST == 0) pair_energy(P, etab, ffp, true, r2, row, 1.0, qq, ev, er, f);
  else          pair_energy_fast<FAST>(P, etab_s, ff_s, r2, row, qq, ev, er, f);
}

// MODE 0: a warp per trial atom, lanes over the candidates (warp reduction at the end).
// MODE 1: a lane per trial atom, the warp walks the candidates in lockstep: every candidate load is a shared-memory BROADCAST
//         (one wavefront for 32 pairs, two candidates per 16-byte load) and nothing is reduced across lanes.
template <int CELL, bool HAS_GG, int MODE, int FAST>
__device__ __forceinline__ void wc_energy_body(const DevParams& P, const WcGrid& G, const WcEnergy& A)
{
  extern __shared__ __align__(16) unsigned char smem[];
  double* etab = reinterpret_cast<double*>(smem + GBK_SMEM_TABLES_OFF);
  double4* fftab = reinterpret_cast<double4*>(smem + GBK_SMEM_TABLES_OFF + GBK_ERFC_BYTES_PAD);
  unsigned char* p = smem + GBK_SMEM_TABLES_OFF + GBK_ERFC_BYTES_PAD + (A.stage_ff ? (size_t) P.ntypes * P.ntypes * 32 : 0);
  double* Tq = reinterpret_cast<double*>(p); p += 64 * 8;
  int* Ttype = reinterpret_cast<int*>(p); p += 64 * 4;
  int* wcnt = reinterpret_cast<int*>(p); p += 64 * 4;
  p = smem + ((size_t) (p - smem) + 15) / 16 * 16;
  const int CF = G.cap_fast, CS = G.cap_slow;
  double* Fx = reinterpret_cast<double*>(p); double* Fy = Fx + CF; double* Fz = Fy + CF; double* Fq = Fz + CF;
  double* Sx = Fq + CF; double* Sy = Sx + CS; double* Sz = Sy + CS; double* Sq = Sz + CS;
  int* Ftk = reinterpret_cast<int*>(Sq + CS); int* Stk = Ftk + CF;
  __shared__ int s_item;

  const int lane = (int) lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  if(FAST == 4 || FAST == 5) { for(int i = threadIdx.x; i < (GBK_ERFC10_DEG + 1) * GBK_ERFC10_NINT; i += blockDim.x) etab[i] = __ldg(&P.erfc_tab10[i]); }
  else stage_erfc_table(P, etab);
  if(A.stage_ff) for(int i = threadIdx.x; i < P.ntypes * P.ntypes; i += blockDim.x) fftab[i] = P.ffA[i];
  for(int a = threadIdx.x; a < A.ms && a < 64; a += blockDim.x) { Tq[a] = A.tq[a] * A.tscoul[a]; Ttype[a] = A.ttype[a]; }
  const double4* ffp = (FAST > 0 || A.stage_ff) ? fftab : P.ffA;      // FAST kernels are launched only with the staged table
  uint32_t etab_s = smem_u32(etab), ff_s = smem_u32(fftab);
  asm volatile("" : "+r"(etab_s), "+r"(ff_s));       // kept in registers: the compiler otherwise re-derives the shared window base at every use
  GBK_ASSUME_SHARED(Fx); GBK_ASSUME_SHARED(Fy); GBK_ASSUME_SHARED(Fz); GBK_ASSUME_SHARED(Fq); GBK_ASSUME_SHARED(Ftk);
  GBK_ASSUME_SHARED(Sx); GBK_ASSUME_SHARED(Sy); GBK_ASSUME_SHARED(Sz); GBK_ASSUME_SHARED(Sq); GBK_ASSUME_SHARED(Stk);
  CellRegs<CELL> C; C.load(P);
  const double cut_max = P.cull_rcut * P.cull_rcut;
  const double reach = P.cull_rcut + G.rcell;
  const double R2 = reach * reach * (1.0 + 1e-12);
  const int nitems = A.ctl[0];
  __syncthreads();

  for(;;)
  {
    if(threadIdx.x == 0) s_item = atomicAdd(&A.ctl[1], 1);
    __syncthreads();
    const int item = s_item;
    if(item >= nitems) break;
    const int cell = A.items[3 * item], first = A.items[3 * item + 1], cnt = A.items[3 * item + 2];
    const int iz = cell % G.n[2], iy = (cell / G.n[2]) % G.n[1], ix = cell / (G.n[2] * G.n[1]);
    const double fcx = ((double) ix + 0.5) * G.inv_n[0], fcy = ((double) iy + 0.5) * G.inv_n[1], fcz = ((double) iz + 0.5) * G.inv_n[2];
    // ---- candidate lists of the cell, in atom order (deterministic): ballot per warp, counts through shared memory
    int nfast = 0, nslow = 0;
    for(int i0 = 0; i0 < A.atoms.n; i0 += blockDim.x)
    {
      const int i = i0 + threadIdx.x;
      int cls = 0; double v0 = 0.0, v1 = 0.0, v2 = 0.0;
      if(i < A.atoms.n)
      {
        const double d0 = frac_wrap(A.atoms.fx[i] - fcx), d1 = frac_wrap(A.atoms.fy[i] - fcy), d2 = frac_wrap(A.atoms.fz[i] - fcz);
        const bool n0 = fabs(d0) > 0.5 - G.margin[0], n1 = fabs(d1) > 0.5 - G.margin[1], n2 = fabs(d2) > 0.5 - G.margin[2];
        if(!(n0 || n1 || n2))
        {
          // the image is the same for every point of the cell: keep the Cartesian vector from the cell centre
          double x, y, z;
          if(CELL == 2) { x = C.c0 * d0; y = C.c4 * d1; z = C.c8 * d2; }
          else if(CELL == 1) { x = C.c0 * d0 + C.c3 * d1 + C.c6 * d2; y = C.c4 * d1 + C.c7 * d2; z = C.c8 * d2; }
          else { x = C.c0 * d0 + C.c3 * d1 + C.c6 * d2; y = C.c1 * d0 + C.c4 * d1 + C.c7 * d2; z = C.c2 * d0 + C.c5 * d1 + C.c8 * d2; }
          if(x * x + y * y + z * z <= R2) { cls = 1; v0 = x; v1 = y; v2 = z; }
        }
        else
        {
          // within the cell's extent of a half-box plane on some axis: the image depends on the trial.  Necessary per axis for any
          // trial of the cell to see the atom inside the cutoff: the wrapped fractional difference reaches into [-w, w].
          const bool ok0 = (fabs(d0) <= P.cull_w[0] + G.margin[0]) || (n0 && P.cull_w[0] + 2.0 * G.margin[0] >= 0.5);
          const bool ok1 = (fabs(d1) <= P.cull_w[1] + G.margin[1]) || (n1 && P.cull_w[1] + 2.0 * G.margin[1] >= 0.5);
          const bool ok2 = (fabs(d2) <= P.cull_w[2] + G.margin[2]) || (n2 && P.cull_w[2] + 2.0 * G.margin[2] >= 0.5);
          if(ok0 && ok1 && ok2) { cls = 2; v0 = d0; v1 = d1; v2 = d2; }
        }
      }
      const unsigned mf = __ballot_sync(0xffffffffu, cls == 1), msl = __ballot_sync(0xffffffffu, cls == 2);
      if(lane == 0) { wcnt[warp] = __popc(mf); wcnt[32 + warp] = __popc(msl); }
      __syncthreads();
      int pf = nfast, ps = nslow;
      for(int w = 0; w < nwarps; w++)
      {
        const int cf = wcnt[w], cs = wcnt[32 + w];
        if(w < warp) { pf += cf; ps += cs; }
        nfast += cf; nslow += cs;
      }
      if(cls == 1)
      {
        const int q = pf + __popc(mf & lt_mask);
        if(q < CF) { Fx[q] = v0; Fy[q] = v1; Fz[q] = v2; Fq[q] = A.atoms.q[i]; Ftk[q] = A.atoms.tk[i]; }
      }
      else if(cls == 2)
      {
        const int q = ps + __popc(msl & lt_mask);
        if(q < CS) { Sx[q] = v0; Sy[q] = v1; Sz[q] = v2; Sq[q] = A.atoms.q[i]; Stk[q] = A.atoms.tk[i]; }
      }
      __syncthreads();
    }
    if(nfast > CF || nslow > CS)
    {
      if(threadIdx.x == 0) *A.overflow = 1;
      nfast = min(nfast, CF); nslow = min(nslow, CS);
    }
    if(MODE == 1)
    {
      // pad both lists to an even length with a candidate that is never inside the cutoff: the loops take two per step
      if(threadIdx.x == 0)
      {
        if(nfast & 1) { Fx[nfast] = 1e30; Fy[nfast] = 1e30; Fz[nfast] = 1e30; Fq[nfast] = 0.0; Ftk[nfast] = 0; }
      }
      __syncthreads();
      const int nf2 = (nfast + 1) & ~1;
      for(int t0 = warp * 32; t0 < cnt; t0 += nwarps * 32)
      {
        const int t = t0 + lane;
        const bool valid = t < cnt;
        const double4 rec = A.srec[first + (valid ? t : cnt - 1)];
        const long long id = __double_as_longlong(rec.w);
        const int a = A.abase + (int) (id % A.amod);
        const int ttype = Ttype[a]; const double tq = Tq[a];
        double ev0 = 0.0, er0 = 0.0, ev1 = 0.0, er1 = 0.0; int fl = 0;
        for(int j = 0; j < nf2; j += 2)
        {
          const double2 X = *reinterpret_cast<const double2*>(Fx + j), Y = *reinterpret_cast<const double2*>(Fy + j), Z = *reinterpret_cast<const double2*>(Fz + j);
          const double ax = X.x - rec.x, ay = Y.x - rec.y, az = Z.x - rec.z;
          const double bx = X.y - rec.x, by = Y.y - rec.y, bz = Z.y - rec.z;
          const double ra = ax * ax + ay * ay + az * az, rb = bx * bx + by * by + bz * bz;
          // charge and type of both candidates in one broadcast load each (the loop takes two candidates per step): 1 % on the kernel
          const double2 Qp = *reinterpret_cast<const double2*>(Fq + j); const int2 Tp = *reinterpret_cast<const int2*>(Ftk + j);
          if(ra < cut_max)
          {
            const int tk = Tp.x; const double qa = Qp.x;
            double ev, er; int f;
            wc_pair<FAST>(P, etab, ffp, etab_s, ff_s, ra, (tk & 0xffff) * P.ntypes + ttype, qa * tq, ev, er, f);
            if(HAS_GG) { const bool gg = (tk >> 16) != 0; ev0 += gg ? 0.0 : ev; er0 += gg ? 0.0 : er; ev1 += gg ? ev : 0.0; er1 += gg ? er : 0.0; }
            else { ev0 += ev; er0 += er; }
            fl |= f;
          }
          if(rb < cut_max)
          {
            const int tk = Tp.y; const double qb = Qp.y;
            double ev, er; int f;
            wc_pair<FAST>(P, etab, ffp, etab_s, ff_s, rb, (tk & 0xffff) * P.ntypes + ttype, qb * tq, ev, er, f);
            if(HAS_GG) { const bool gg = (tk >> 16) != 0; ev0 += gg ? 0.0 : ev; er0 += gg ? 0.0 : er; ev1 += gg ? ev : 0.0; er1 += gg ? er : 0.0; }
            else { ev0 += ev; er0 += er; }
            fl |= f;
          }
        }
        if(nslow > 0)
        {
          double ox, oy, oz; to_frac(P, rec.x, rec.y, rec.z, ox, oy, oz);
          for(int j = 0; j < nslow; j++)
          {
            const double r2 = C.r2(Sx[j] - ox, Sy[j] - oy, Sz[j] - oz);
            if(r2 < cut_max)
            {
              const int tk = Stk[j];
              double ev, er; int f;
              wc_pair<FAST>(P, etab, ffp, etab_s, ff_s, r2, (tk & 0xffff) * P.ntypes + ttype, Sq[j] * tq, ev, er, f);
              if(HAS_GG) { const bool gg = (tk >> 16) != 0; ev0 += gg ? 0.0 : ev; er0 += gg ? 0.0 : er; ev1 += gg ? ev : 0.0; er1 += gg ? er : 0.0; }
              else { ev0 += ev; er0 += er; }
              fl |= f;
            }
          }
        }
        if(valid)
        {
          double4* o = reinterpret_cast<double4*>(A.e4 + 4 * id);
          *o = make_double4(ev0, er0, ev1, er1);
          A.flag[id] = fl;
        }
      }
    }
    else
    {
  // ---- one warp per trial atom of the item
      for(int t = warp; t < cnt; t += nwarps)
      {
        const double4 rec = A.srec[first + t];
        const long long id = __double_as_longlong(rec.w);
        const int a = A.abase + (int) (id % A.amod);
        const int ttype = Ttype[a]; const double tq = Tq[a];
        double ev0 = 0.0, er0 = 0.0, ev1 = 0.0, er1 = 0.0; int fl = 0;
#pragma unroll 2
        for(int j = lane; j < nfast; j += 32)
        {
          const double dx = Fx[j] - rec.x, dy = Fy[j] - rec.y, dz = Fz[j] - rec.z;
          const double r2 = dx * dx + dy * dy + dz * dz;
          if(r2 < cut_max)
          {
            const int tk = Ftk[j];
            double ev, er; int f;
            pair_energy(P, etab, ffp, true, r2, (tk & 0xffff) * P.ntypes + ttype, 1.0, Fq[j] * tq, ev, er, f);
            if(HAS_GG) { const bool gg = (tk >> 16) != 0; ev0 += gg ? 0.0 : ev; er0 += gg ? 0.0 : er; ev1 += gg ? ev : 0.0; er1 += gg ? er : 0.0; }
            else { ev0 += ev; er0 += er; }
            fl |= f;
          }
        }
        if(nslow > 0)
        {
          // fractional offset of the trial from the cell centre: the general wrap (CellRegs::r2) for the atoms near a half-box plane
          double ox, oy, oz; to_frac(P, rec.x, rec.y, rec.z, ox, oy, oz);
          for(int j = lane; j < nslow; j += 32)
          {
            const double r2 = C.r2(Sx[j] - ox, Sy[j] - oy, Sz[j] - oz);
            if(r2 < cut_max)
            {
              const int tk = Stk[j];
              double ev, er; int f;
              pair_energy(P, etab, ffp, true, r2, (tk & 0xffff) * P.ntypes + ttype, 1.0, Sq[j] * tq, ev, er, f);
              if(HAS_GG) { const bool gg = (tk >> 16) != 0; ev0 += gg ? 0.0 : ev; er0 += gg ? 0.0 : er; ev1 += gg ? ev : 0.0; er1 += gg ? er : 0.0; }
              else { ev0 += ev; er0 += er; }
              fl |= f;
            }
          }
        }
        ev0 = warp_sum(ev0); er0 = warp_sum(er0);
        if(HAS_GG) { ev1 = warp_sum(ev1); er1 = warp_sum(er1); }
        fl = __any_sync(0xffffffffu, fl) ? 1 : 0;
        if(lane == 0)
        {
          double4* o = reinterpret_cast<double4*>(A.e4 + 4 * id);
          *o = make_double4(ev0, er0, ev1, er1);
          A.flag[id] = fl;
        }
      }
}
    __syncthreads();          // the lists are rebuilt by the next item
  }
}

template <int CELL, bool HAS_GG>
__global__ void __launch_bounds__(768, 1)
k_wc_energy(DevParams P, WcGrid G, WcEnergy A) { wc_energy_body<CELL, HAS_GG, 0, 0>(P, G, A); }

template <int CELL, bool HAS_GG, int FAST>
__global__ void __launch_bounds__(256, (FAST == 1 || FAST >= 3) ? 4 : 3)      // the bodies without switches fit 64 registers: 32 warps per SM
k_wc_energy_lt(DevParams P, WcGrid G, WcEnergy A) { wc_energy_body<CELL, HAS_GG, 1, FAST>(P, G, A); }

// ---------------------------------------------------------------------------------------------- caller-supplied trial atoms
// gb_trial_energies with a large batch of trial groups (every group the same molecule, unit scaling factors, nothing of the system
// excluded): the trial atoms are binned like the Widom ones (fractional coordinates in, one thread per atom) and evaluated by the same
// energy kernel; k_wc_sum_groups then adds the atoms of a group in atom order.
__global__ void k_wc_gen_explicit(DevParams P, WcGrid G, const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz, long long n,
                                  int* ucell, double* udelta, int* count)
{
  const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= n) return;
  double a = fx[g], b = fy[g], c = fz[g];
  a -= floor(a); b -= floor(b); c -= floor(c);
  int ix = (int) (a * G.n[0]), iy = (int) (b * G.n[1]), iz = (int) (c * G.n[2]);
  ix = min(max(ix, 0), G.n[0] - 1); iy = min(max(iy, 0), G.n[1] - 1); iz = min(max(iz, 0), G.n[2] - 1);
  const double ex = a - ((double) ix + 0.5) * G.inv_n[0], ey = b - ((double) iy + 0.5) * G.inv_n[1], ez = c - ((double) iz + 0.5) * G.inv_n[2];
  const int cell = (ix * G.n[1] + iy) * G.n[2] + iz;
  ucell[g] = cell;
  udelta[3 * g] = P.cell[0] * ex + P.cell[3] * ey + P.cell[6] * ez;
  udelta[3 * g + 1] = P.cell[1] * ex + P.cell[4] * ey + P.cell[7] * ez;
  udelta[3 * g + 2] = P.cell[2] * ex + P.cell[5] * ey + P.cell[8] * ez;
  atomicAdd(&count[cell], 1);
}

// out6[t] = {0, 0, HGVDW, HGReal, GGVDW, GGReal} (the layout of k_trial_energies; host-host terms do not occur on this route), flag[t]
__global__ void k_wc_sum_groups(const double* __restrict__ e4, const int* __restrict__ flag, long long ngroups, int cs, double* out6, int* out_flag)
{
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= ngroups) return;
  double s[4] = {0, 0, 0, 0}; int fl = 0;
  for(int a = 0; a < cs; a++)
  {
    const double4 v = *reinterpret_cast<const double4*>(e4 + 4 * (t * cs + a));
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w; fl |= flag[t * cs + a];
  }
  double* o = out6 + 6 * t;
  o[0] = 0.0; o[1] = 0.0; o[2] = s[0]; o[3] = s[1]; o[4] = s[2]; o[5] = s[3];
  out_flag[t] = fl ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------- selection stages
struct WcSel
{
  const double* __restrict__ pool3; const long long* __restrict__ fb_index; const long long* __restrict__ or_index;
  const double* __restrict__ uni; long long n;
  int ntrials, norient, ms;
  const double* __restrict__ tx; const double* __restrict__ ty; const double* __restrict__ tz;       // template molecule (Cartesian)
  CompView C;                                   // only the block-pocket fields are used
  const double* __restrict__ e4; const int* __restrict__ flag;
  int e4_by_row;                                // 1: e4 / flag are indexed by pool row (kept by gb_widom_first_bead_success), 0: by insertion * ntrials
  double* fbres;                                // per insertion {W, HGv, HGr, GGv, GGr, x, y, z} of the selected first bead
  double* rec; int* stage;                      // as k_widom_pair leaves them for k_widom_ewald
  int first_bead_only;
  int* ucell; double* udelta; int* count;       // chain atoms, binned (first-bead stage)
};

// Boltzmann selection of the first bead (CBMC_FirstBead_Finish, mc_widom.h:305-383) + the chain trial atoms
// (get_random_trial_orientation, :215-303), binned for the second energy pass.  One warp per insertion.
__global__ void __launch_bounds__(256)
k_wc_select_fb(DevParams P, WcGrid G, WcSel A)
{
  const int lane = (int) lane_id();
  const long long ins = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if(ins >= A.n) return;
  const int cs = A.ms - 1, rec_stride = 5 + 3 * A.ms;
  const long long fb_off = A.fb_index ? A.fb_index[ins] : ins * (A.ntrials + A.norient);
  const long long or_off = A.or_index ? A.or_index[ins] : fb_off + A.ntrials;
  double px = 0, py = 0, pz = 0; double e[4] = {0, 0, 0, 0}; int fl = 1;
  if(lane < A.ntrials)
  {
    const double* r = A.pool3 + 3 * (fb_off + lane);
    px = P.cell[0] * r[0]; py = P.cell[4] * r[1]; pz = P.cell[8] * r[2];
    const long long g = (A.e4_by_row ? fb_off : ins * A.ntrials) + lane;
    const double4 v = *reinterpret_cast<const double4*>(A.e4 + 4 * g);
    e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w; fl = A.flag[g];
  }
  if(A.C.npocket > 0)
  {
    // mc_widom.h:463-498: a blocked STARTING bead (trial 0) flags every trial, otherwise every blocked trial is flagged on its own
    const double x0 = __shfl_sync(0xffffffffu, px, 0), y0 = __shfl_sync(0xffffffffu, py, 0), z0 = __shfl_sync(0xffffffffu, pz, 0);
    bool blk = blocked_pocket(P, A.C, x0, y0, z0);
    if(!blk && lane > 0 && lane < A.ntrials) blk = blocked_pocket(P, A.C, px, py, pz);
    if(blk) fl = 1;
  }
  double tot = e[0] + e[2]; if(P.vdw_real_bias) tot += e[1] + e[3];
  RosenResult r1 = rosenbluth_warp(-P.beta * tot, !fl, A.ntrials, A.uni ? A.uni[2 * ins] : 0.5, true);
  double W = 0.0; int ok = r1.success && !(r1.R < 1e-150);
  const int sfb = r1.sel_lane;
  double efb[4];
#pragma unroll
  for(int k = 0; k < 4; k++) efb[k] = __shfl_sync(0xffffffffu, e[k], sfb);
  if(ok)
  {
    W = r1.R / (double) A.ntrials;
    if(!P.vdw_real_bias) W *= exp(-P.beta * (efb[1] + efb[3]));
    if(W <= 1e-150) ok = 0;
  }
  if(A.first_bead_only) { if(lane == 0) A.stage[ins] = ok ? 1 : (r1.nsurv > 0 ? 0 : 2); return; }
  double* rec = A.rec + (size_t) ins * rec_stride;
  if(!ok) { if(lane == 0) { A.stage[ins] = 1; rec[0] = 0.0; } return; }
  const double fbx = __shfl_sync(0xffffffffu, px, sfb), fby = __shfl_sync(0xffffffffu, py, sfb), fbz = __shfl_sync(0xffffffffu, pz, sfb);
  if(cs == 0)
  {
    if(lane == 0) { A.stage[ins] = 0; rec[0] = W; for(int k = 0; k < 4; k++) rec[1 + k] = efb[k]; rec[5] = fbx; rec[6] = fby; rec[7] = fbz; }
    return;
  }
  if(lane == 0)
  {
    double* f = A.fbres + 8 * ins;
    f[0] = W; f[1] = efb[0]; f[2] = efb[1]; f[3] = efb[2]; f[4] = efb[3]; f[5] = fbx; f[6] = fby; f[7] = fbz;
    A.stage[ins] = -1;                          // chain stage pending
  }
  if(lane < A.norient)
  {
    const double* r = A.pool3 + 3 * (or_off + lane);
    for(int a = 0; a < cs; a++)
    {
      double vx = A.tx[1 + a] - A.tx[0], vy = A.ty[1 + a] - A.ty[0], vz = A.tz[1 + a] - A.tz[0];
      rotate_quaternion(vx, vy, vz, r[0], r[1], r[2]);
      int cell; double dx, dy, dz;
      wc_bin(P, G, fbx + vx, fby + vy, fbz + vz, cell, dx, dy, dz);
      const long long g = (ins * A.norient + lane) * cs + a;
      A.ucell[g] = cell; A.udelta[3 * g] = dx; A.udelta[3 * g + 1] = dy; A.udelta[3 * g + 2] = dz;
      atomicAdd(&A.count[cell], 1);
    }
  }
}

// Boltzmann selection of the orientation (tail of Widom_Move_Chain_PARTIAL, mc_widom.h:568-611) and the record k_widom_ewald reads
__global__ void __launch_bounds__(256)
k_wc_select_chain(DevParams P, WcSel A)
{
  const int lane = (int) lane_id();
  const long long ins = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if(ins >= A.n) return;
  if(A.stage[ins] != -1) return;
  const int cs = A.ms - 1, rec_stride = 5 + 3 * A.ms;
  const long long fb_off = A.fb_index ? A.fb_index[ins] : ins * (A.ntrials + A.norient);
  const long long or_off = A.or_index ? A.or_index[ins] : fb_off + A.ntrials;
  const double* f = A.fbres + 8 * ins;
  double W = f[0];
  double e[4] = {0, 0, 0, 0}; int fl = 1;
  if(lane < A.norient)
  {
    fl = 0;
    for(int a = 0; a < cs; a++)
    {
      const long long g = (ins * A.norient + lane) * cs + a;
      const double4 v = *reinterpret_cast<const double4*>(A.e4 + 4 * g);
      e[0] += v.x; e[1] += v.y; e[2] += v.z; e[3] += v.w; fl |= A.flag[g];
    }
  }
  double tot = e[0] + e[2]; if(P.vdw_real_bias) tot += e[1] + e[3];
  RosenResult r2 = rosenbluth_warp(-P.beta * tot, !fl, A.norient, A.uni[2 * ins + 1], true);
  int ok = r2.success && !(r2.R < 1e-150);
  const int so = r2.sel_lane;
  double ech[4];
#pragma unroll
  for(int k = 0; k < 4; k++) ech[k] = __shfl_sync(0xffffffffu, e[k], so);
  if(ok)
  {
    double W2 = r2.R / (double) A.norient;
    if(!P.vdw_real_bias) W2 *= exp(-P.beta * (ech[1] + ech[3]));
    W *= W2;
    if(W <= 1e-150) ok = 0;
  }
  double* rec = A.rec + (size_t) ins * rec_stride;
  if(!ok) { if(lane == 0) { A.stage[ins] = (r2.nsurv > 0) ? 2 : 3; rec[0] = 0.0; } return; }
  // positions of the selected orientation (recomputed: one quaternion)
  int blocked = 0;
  if(lane == so)
  {
    const double* r = A.pool3 + 3 * (or_off + so);
    if(A.C.npocket > 0 && blocked_pocket(P, A.C, f[5], f[6], f[7])) blocked = 1;
    for(int a = 0; a < cs; a++)
    {
      double vx = A.tx[1 + a] - A.tx[0], vy = A.ty[1 + a] - A.ty[0], vz = A.tz[1 + a] - A.tz[0];
      rotate_quaternion(vx, vy, vz, r[0], r[1], r[2]);
      const double cx = f[5] + vx, cy = f[6] + vy, cz = f[7] + vz;
      rec[8 + 3 * a] = cx; rec[9 + 3 * a] = cy; rec[10 + 3 * a] = cz;
      // after the growth EVERY atom of the molecule is tested against the block pockets (mc_swap_utilities.h:46-80)
      if(A.C.npocket > 0 && !blocked && blocked_pocket(P, A.C, cx, cy, cz)) blocked = 1;
    }
  }
  blocked = __shfl_sync(0xffffffffu, blocked, so);
  if(blocked) { if(lane == 0) { A.stage[ins] = 2; rec[0] = 0.0; } return; }
  if(lane == 0)
  {
    A.stage[ins] = 0;
    rec[0] = W;
    for(int k = 0; k < 4; k++) rec[1 + k] = f[1 + k] + ech[k];
    rec[5] = f[5]; rec[6] = f[6]; rec[7] = f[7];
  }
}
