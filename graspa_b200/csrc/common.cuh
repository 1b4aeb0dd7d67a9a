// graspa_b200 -- shared device/host definitions of the sm_100a energy engine.
// Product code: nothing here includes, links or calls oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "erfc_table.inc"
#include "erfc_table10.inc"

#define GBK_PI 3.14159265358979323846
#ifdef GBK_PHASE_TIMING
#include <cstdio>
__device__ long long g_marks[64];
__device__ int g_nmarks;
#define GBK_MARK() do { if(blockIdx.x == 0 && threadIdx.x == 0) g_marks[g_nmarks++] = clock64(); } while(0)
#else
#define GBK_MARK() do { } while(0)
#endif

#define GBK_MAX_SEG 16
#define GBK_QCAP 128           // per-warp in-cutoff queue entries (up to 31 left over + 3 x 32 pushed per iteration)
#define GBK_MAX_CS 32          // max atoms of one trial group handled by the warp pair loop
#define GBK_MAX_TRIALS 32      // one lane per trial in the Rosenbluth stage
#define GBK_ERFC_BYTES ((GBK_ERFC_DEG + 1) * GBK_ERFC_NINT * 8)
#define GBK_ERFC_BYTES_PAD ((GBK_ERFC_BYTES + 31) / 32 * 32)
#define GBK_SMEM_TABLES_OFF 32   // dynamic smem: [0,16) mbarrier, tables from byte 32 (32-byte aligned for double4 loads)

// Kernel-side view of Boxsize + ForceField scalars (data_struct.h:838-886).  Passed by value.
struct DevParams
{
  double cell[9];
  double inv[9];
  double cut_vdw2, cut_coul2, overlap;
  double alpha, prefactor, volume, recip_cutoff, beta;
  int ntypes, cubic, no_charges, vdw_real_bias, use1264, use_lammps;
  int kmax[3];
  int all_unit_scale;                 // every system atom has scale == scaleCoul == 1
  int cell_mode;                      // 0 general, 1 lower triangular (CIF cells, read_data.cpp:1545-1547), 2 orthorhombic
  int erfc_table_ok;                  // alpha*sqrt(cut_coul2) < GBK_ERFC_XMAX: every in-cutoff pair is inside the erfc table
  double cull_w[3];                   // tile culling: r_cut * |column i of the inverse cell| (reach of the cutoff sphere along fractional axis i)
  double cull_rcut;                   // sqrt(max(cut_vdw2, cut_coul2)) (cut_vdw2 alone without charges)
  const double4* __restrict__ ffA;    // LJ: {4*eps, sigma^2, shift, 1/sigma^2}; 12-6-4: {C12, C6, C4, shift}
  const double*  __restrict__ ffB;    // 12-6-4: C10
  const double*  __restrict__ erfc_tab;   // device copy of h_erfc_table
  const double*  __restrict__ erfc_tab10; // device copy of h_erfc_table10 (degree 10 on [0, GBK_ERFC10_XMAX): the short table of k_wc_energy_lt)
  int erfc10_ok;                      // alpha*sqrt(cut_coul2) < GBK_ERFC10_XMAX (EwaldPrecision 1e-6 gives alpha * r_cut = 3.1-3.2, read_data.cpp:693-697)
};

// system atoms, SoA over slots (fractional coordinates are derived from the Cartesian ones)
struct SysView
{
  const double* __restrict__ fx; const double* __restrict__ fy; const double* __restrict__ fz;  // fractional
  const double* __restrict__ q;                                                                   // charge
  const double* __restrict__ scale; const double* __restrict__ scoul;
  const int*    __restrict__ type;  const int* __restrict__ molid;
};

// the live atom ranges a pair loop runs over; kind 0 = HH, 1 = HG, 2 = GG
struct SegList
{
  int nseg;
  int start[GBK_MAX_SEG];
  int count[GBK_MAX_SEG];
  int comp[GBK_MAX_SEG];
  int kind[GBK_MAX_SEG];
  int staged[GBK_MAX_SEG];   // 1: the segment lives in the shared-memory staged pack at offset start
};

// %laneid through a volatile asm: evaluated once where it is called, never rematerialised as S2R+LOP3 inside hot loops
__device__ __forceinline__ unsigned lane_id() { unsigned l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// explicit shared-memory loads (the pair loop reads either the TMA-staged pack or global memory;
// keeping the state space in the instruction avoids generic LD and pointer selects in the hot loop)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
// address-space hints: pointers derived from the dynamic shared-memory block are declared shared so that loads
// through them compile to LDS (not generic LD) after inlining
#define GBK_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#define GBK_ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))

// d - nearest_integer(d) without leaving the FP64 pipe (|d| < 2^51).  The reference truncates
// static_cast<int>(s +- 0.5) (maths.cuh:442-444); both pick the same image except at exact half-integers.
__device__ __forceinline__ double frac_wrap(double d)
{
  const double M = 6755399441055744.0;   // 1.5 * 2^52
  double t = __dadd_rn(d, M);
  t = __dsub_rn(t, M);
  return d - t;
}

// fractional coordinate of a Cartesian position: the product of maths.cuh:438-440
__device__ __forceinline__ void to_frac(const DevParams& P, double x, double y, double z, double& sx, double& sy, double& sz)
{
  sx = P.inv[0] * x + P.inv[3] * y + P.inv[6] * z;
  sy = P.inv[1] * x + P.inv[4] * y + P.inv[7] * z;
  sz = P.inv[2] * x + P.inv[5] * y + P.inv[8] * z;
}

// cell matrix held in registers by the pair loops; MODE 1 skips the structural zeros of a lower-triangular
// cell (rows = lattice vectors: a = (c0,0,0), b = (c3,c4,0), c = (c6,c7,c8)), MODE 2 those of an orthorhombic one
template <int MODE>
struct CellRegs
{
  double c0, c1, c2, c3, c4, c5, c6, c7, c8;
  __device__ __forceinline__ void load(const DevParams& P)
  {
    c0 = P.cell[0]; c1 = P.cell[1]; c2 = P.cell[2]; c3 = P.cell[3]; c4 = P.cell[4]; c5 = P.cell[5]; c6 = P.cell[6]; c7 = P.cell[7]; c8 = P.cell[8];
  }
  // minimum-image squared distance from fractional differences (maths.cuh:442-448 + dot)
  __device__ __forceinline__ double r2(double dsx, double dsy, double dsz) const
  {
    dsx = frac_wrap(dsx); dsy = frac_wrap(dsy); dsz = frac_wrap(dsz);
    double dx, dy, dz;
    if(MODE == 2) { dx = c0 * dsx; dy = c4 * dsy; dz = c8 * dsz; }
    else if(MODE == 1)
    {
      dx = c0 * dsx + c3 * dsy + c6 * dsz;
      dy = c4 * dsy + c7 * dsz;
      dz = c8 * dsz;
    }
    else
    {
      dx = c0 * dsx + c3 * dsy + c6 * dsz;
      dy = c1 * dsx + c4 * dsy + c7 * dsz;
      dz = c2 * dsx + c5 * dsy + c8 * dsz;
    }
    return dx * dx + dy * dy + dz * dz;
  }
};

__device__ __forceinline__ double min_image_r2(const DevParams& P, double dsx, double dsy, double dsz)
{
  CellRegs<0> C; C.load(P);
  return C.r2(dsx, dsy, dsz);
}

// erfc(x) for x in [0, GBK_ERFC_XMAX): degree-12 piecewise polynomials on intervals of width 1/8 centred at k/8
// (tools/gen_erfc_table.py; worst relative error 5.6e-16 against 50-digit arithmetic).  Table in shared memory at
// byte address tab (layout coef[j][k]).  13 FP64 instructions + 13 LDS instead of libdevice's ~60 DFMA with
// immediate-constant moves.
__device__ __forceinline__ double erfc_table_eval(const double* __restrict__ tab, double x)
{
  GBK_ASSUME_SHARED(tab);
  const double M = 6755399441055744.0;
  const double y = x * GBK_ERFC_SCALE;
  double kd = __dadd_rn(y, M);
  const int k = __double2loint(kd);
  kd = __dsub_rn(kd, M);
  const double t = y - kd;
  const double* a = tab + k;
  double acc = a[GBK_ERFC_NINT * GBK_ERFC_DEG];
#pragma unroll
  for(int j = GBK_ERFC_DEG - 1; j >= 0; j--) acc = fma(acc, t, a[GBK_ERFC_NINT * j]);
  return acc;
}

// One in-cutoff pair: LJ 12-6 (+soft core, +shift) or 12-6-4 polynomial (maths.cuh:452-494) and the
// real-space Ewald term (maths.cuh:496-500).  One rsqrt feeds both.
// ffp: the LJ table (shared-memory copy or P.ffA); unit: warp-uniform "every scaling factor of this group is 1".
// libdevice's erfc for arguments beyond the table (alpha r > 6: never reached with the reference's Ewald set-ups); kept out
// of line so that its ~200 instructions are not replicated into every inlined copy of pair_energy
__device__ __noinline__ double erfc_beyond_table(double x) { return erfc(x); }

__device__ __forceinline__ void pair_energy(const DevParams& P, const double* __restrict__ etab, const double4* __restrict__ ffp, bool unit,
                                            double r2, int row, double scaling, double qq_scaled,
                                            double& e_vdw, double& e_real, int& flag)
{
  double rinv = rsqrt(r2);
  asm volatile("" : "+d"(rinv));                       // ONE reciprocal square root for both terms (the compiler otherwise sinks a copy into each branch)
  const double rinv2 = rinv * rinv;
  e_vdw = 0.0; e_real = 0.0; flag = 0;
  if(r2 < P.cut_vdw2)
  {
    const double4 f = ffp[row];
    double e;
    if(!P.use1264)
    {
      double rri3;
      if(unit) { const double x = f.y * rinv2; rri3 = x * x * x; e = f.x * (rri3 * (rri3 - 1.0)) - f.z; }
      else if(scaling == 1.0) { const double x = f.y * rinv2; rri3 = x * x * x; e = scaling * (f.x * (rri3 * (rri3 - 1.0)) - f.z); }
      else
      {
        const double t = r2 * f.w; const double t3 = t * t * t; const double om = 1.0 - scaling;
        rri3 = 1.0 / (t3 + 0.5 * om * om);
        e = scaling * (f.x * (rri3 * (rri3 - 1.0)) - f.z);
      }
    }
    else
    {
      const double ri4 = rinv2 * rinv2, ri6 = ri4 * rinv2, ri10 = ri4 * ri6, ri12 = ri6 * ri6;
      e = scaling * (f.x * ri12 - f.y * ri6 + __ldg(&P.ffB[row]) * ri10 + f.z * ri4 - f.w);
    }
    if(e > P.overlap) flag = 1;
    if(r2 < 0.01) flag = 1;
    e_vdw = e;
  }
  if(!P.no_charges && r2 < P.cut_coul2)
  {
    const double r = r2 * rinv;
    const double x = P.alpha * r;
    const double ec = (P.erfc_table_ok || x < GBK_ERFC_XMAX) ? erfc_table_eval(etab, x) : erfc_beyond_table(x);
    e_real = P.prefactor * qq_scaled * ec * rinv;
  }
}

// cooperative copy of the erfc table into shared memory (all threads of the CTA), returns its shared address
__device__ __forceinline__ void stage_erfc_table(const DevParams& P, double* dst)
{
  for(int i = threadIdx.x; i < (GBK_ERFC_DEG + 1) * GBK_ERFC_NINT; i += blockDim.x) dst[i] = __ldg(&P.erfc_tab[i]);
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
