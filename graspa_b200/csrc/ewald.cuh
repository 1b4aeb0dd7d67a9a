// graspa_b200 -- Ewald Fourier device code.
//
// Replaces Initialize_WaveVector_General (Ewald_Energy_Functions.h:162-185, eik recurrences :97-104) and
// Fourier_Ewald_Diff (:280-397).  The reference launches a 1-block table kernel and then 2*ceil(nvec/128)
// blocks in which every k is evaluated twice (same-type half, cross-type half) and inactive k idle.  Here
//   - only ACTIVE k-vectors exist: the host compacts (kx,ky,kz), temp(k) = factor*exp(-|k|^2/4a^2)/|k|^2 and
//     the slot in the stored structure-factor arrays into a table once per box;
//   - the eik tables of the <= 64 moved atoms are built in shared memory by the same kernel;
//   - same-type and cross-type energies come from one pass over k.
#pragma once
#include "common.cuh"

#define GBK_EW_MAX_ATOMS 64

struct KTable
{
  const int*    __restrict__ kpack;   // (kx << 16) | ((ky + 128) << 8) | (kz + 128)
  const double* __restrict__ temp;    // factor * exp(-rksq/(4 alpha^2)) / rksq, factor = prefactor*(1|2)
  const int*    __restrict__ slot;    // index kxyz in the full (kx,ky,kz) arrays
  int nact;
};

struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { cplx c; c.re = a.re * b.re - a.im * b.im; c.im = a.re * b.im + a.im * b.re; return c; }

// eik tables for n atoms: e?[k * n + i], k = 0..kmax.  Work item = (atom, axis); callable by any group of
// threads (tid in [0,nthreads)).
__device__ __forceinline__ void build_eik(const DevParams& P, const double* pos3 /* xyz interleaved */, int n,
                                          cplx* ex, cplx* ey, cplx* ez, int tid, int nthreads)
{
  for(int w = tid; w < 3 * n; w += nthreads)
  {
    const int i = w / 3, axis = w % 3;
    const double x = pos3[3 * i], y = pos3[3 * i + 1], z = pos3[3 * i + 2];
    double s = P.inv[axis] * x + P.inv[3 + axis] * y + P.inv[6 + axis] * z;   // matrix_multiply_by_vector, maths.cuh:133-138
    s *= 2 * GBK_PI;
    cplx* e = axis == 0 ? ex : (axis == 1 ? ey : ez);
    const int km = P.kmax[axis];
    cplx one; one.re = 1.0; one.im = 0.0;
    cplx e1; sincos(s, &e1.im, &e1.re);
    e[i] = one; e[n + i] = e1;
    cplx cur = e1;
    for(int k = 2; k <= km; k++) { cur = cmul(cur, e1); e[k * n + i] = cur; }
  }
}

// sum over atoms [a0, a1) of q * eik_x * eik_y * eik_z at one k (Ewald_Energy_Functions.h:336-357)
__device__ __forceinline__ cplx ck_sum(const cplx* ex, const cplx* ey, const cplx* ez, const double* qeff, int n,
                                       int a0, int a1, int kx, int ky, int kz)
{
  cplx s; s.re = 0.0; s.im = 0.0;
  const int aky = ky < 0 ? -ky : ky, akz = kz < 0 ? -kz : kz;
  for(int i = a0; i < a1; i++)
  {
    cplx t1 = ey[aky * n + i]; if(ky < 0) t1.im = -t1.im;
    const cplx exy = cmul(ex[kx * n + i], t1);
    cplx t2 = ez[akz * n + i]; if(kz < 0) t2.im = -t2.im;
    const cplx t = cmul(exy, t2);
    const double w = qeff[i];
    s.re += w * t.re; s.im += w * t.im;
  }
  return s;
}

__device__ __forceinline__ void unpack_k(int kp, int& kx, int& ky, int& kz)
{
  kx = kp >> 16; ky = ((kp >> 8) & 255) - 128; kz = (kp & 255) - 128;
}
