// graspa_b200 -- shared device/host definitions of the sm_100a energy engine.
// Product code: nothing here includes, links or calls oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define GBK_PI 3.14159265358979323846
#define GBK_MAX_SEG 16
#define GBK_QCAP 64            // per-warp in-cutoff queue entries
#define GBK_MAX_CS 32          // max atoms of one trial group handled by the warp pair loop
#define GBK_MAX_TRIALS 32      // one lane per trial in the Rosenbluth stage

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
  const double4* __restrict__ ffA;    // LJ: {4*eps, sigma^2, shift, 1/sigma^2}; 12-6-4: {C12, C6, C4, shift}
  const double*  __restrict__ ffB;    // 12-6-4: C10
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

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

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

// d - nearest_integer(d) without leaving the FP64 pipe (|d| < 2^51).  The reference truncates
// static_cast<int>(s +- 0.5) (maths.cuh:442-444); both pick the same image except at exact half-integers.
__device__ __forceinline__ double frac_wrap(double d)
{
  const double M = 6755399441055744.0;   // 1.5 * 2^52
  double t = __dadd_rn(d, M);
  t = __dsub_rn(t, M);
  return d - t;
}

// fractional coordinate of a Cartesian position: s = InverseCell^T-style product of maths.cuh:438-440
__device__ __forceinline__ void to_frac(const DevParams& P, double x, double y, double z, double& sx, double& sy, double& sz)
{
  sx = P.inv[0] * x + P.inv[3] * y + P.inv[6] * z;
  sy = P.inv[1] * x + P.inv[4] * y + P.inv[7] * z;
  sz = P.inv[2] * x + P.inv[5] * y + P.inv[8] * z;
}

// minimum-image squared distance from fractional differences (maths.cuh:442-448 + dot)
__device__ __forceinline__ double min_image_r2(const DevParams& P, double dsx, double dsy, double dsz)
{
  dsx = frac_wrap(dsx); dsy = frac_wrap(dsy); dsz = frac_wrap(dsz);
  const double dx = P.cell[0] * dsx + P.cell[3] * dsy + P.cell[6] * dsz;
  const double dy = P.cell[1] * dsx + P.cell[4] * dsy + P.cell[7] * dsz;
  const double dz = P.cell[2] * dsx + P.cell[5] * dsy + P.cell[8] * dsz;
  return dx * dx + dy * dy + dz * dz;
}

// One in-cutoff pair: LJ 12-6 (+soft core, +shift) or 12-6-4 polynomial (maths.cuh:452-494) and the
// real-space Ewald term (maths.cuh:496-500).  One rsqrt feeds both.
__device__ __forceinline__ void pair_energy(const DevParams& P, double r2, int row, double scaling, double qq_scaled,
                                            double& e_vdw, double& e_real, int& flag)
{
  const double rinv = rsqrt(r2);
  const double rinv2 = rinv * rinv;
  if(r2 < P.cut_vdw2)
  {
    const double4 f = P.ffA[row];
    double e;
    if(!P.use1264)
    {
      double rri3;
      if(scaling == 1.0) { const double x = f.y * rinv2; rri3 = x * x * x; }
      else
      {
        const double t = r2 * f.w; const double t3 = t * t * t; const double om = 1.0 - scaling;
        rri3 = 1.0 / (t3 + 0.5 * om * om);
      }
      e = scaling * (f.x * (rri3 * (rri3 - 1.0)) - f.z);
    }
    else
    {
      const double ri4 = rinv2 * rinv2, ri6 = ri4 * rinv2, ri10 = ri4 * ri6, ri12 = ri6 * ri6;
      e = scaling * (f.x * ri12 - f.y * ri6 + P.ffB[row] * ri10 + f.z * ri4 - f.w);
    }
    if(e > P.overlap) flag = 1;
    if(r2 < 0.01) flag = 1;
    e_vdw += e;
  }
  if(!P.no_charges && r2 < P.cut_coul2)
  {
    const double r = r2 * rinv;
    e_real += P.prefactor * qq_scaled * erfc(P.alpha * r) * rinv;
  }
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

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
