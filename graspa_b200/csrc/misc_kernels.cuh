// graspa_b200 -- state, Rosenbluth, Ewald and instrumentation kernels (included by engine.cu; the pair kernels live in pair_kernels.cuh).
#pragma once
#include "common.cuh"
#include "pair.cuh"
#include "ewald.cuh"

// ---------------------------------------------------------------------------------------------
// small state kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_frac_update(DevParams P, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                              double* fx, double* fy, double* fz, int start, int count)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < count) to_frac(P, x[start + i], y[start + i], z[start + i], fx[start + i], fy[start + i], fz[start + i]);
}

// ScalePositions (mc_box.h:18-64): a molecule follows its first atom, which scales with the box; the other atoms keep their
// minimum-image offset from it (old box).  One thread per molecule of the listed components.
struct ScaleArgs { int nseg; int start[GBK_MAX_SEG], ms[GBK_MAX_SEG], nmol[GBK_MAX_SEG]; double scale; };

__global__ void k_scale_molecules(DevParams P, ScaleArgs A, double* x, double* y, double* z)
{
  long long m = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  int seg = 0;
  while(seg < A.nseg && m >= A.nmol[seg]) { m -= A.nmol[seg]; seg++; }
  if(seg >= A.nseg) return;
  const int first = A.start[seg] + (int) m * A.ms[seg];
  const double cx = x[first], cy = y[first], cz = z[first];
  for(int a = 0; a < A.ms[seg]; a++)
  {
    double dx = x[first + a] - cx, dy = y[first + a] - cy, dz = z[first + a] - cz;
    if(P.cubic)
    {
      dx -= (double) static_cast<int>(dx * P.inv[0] + ((dx >= 0.0) ? 0.5 : -0.5)) * P.cell[0];
      dy -= (double) static_cast<int>(dy * P.inv[4] + ((dy >= 0.0) ? 0.5 : -0.5)) * P.cell[4];
      dz -= (double) static_cast<int>(dz * P.inv[8] + ((dz >= 0.0) ? 0.5 : -0.5)) * P.cell[8];
    }
    else
    {
      double sx = P.inv[0] * dx + P.inv[3] * dy + P.inv[6] * dz, sy = P.inv[1] * dx + P.inv[4] * dy + P.inv[7] * dz, sz = P.inv[2] * dx + P.inv[5] * dy + P.inv[8] * dz;
      sx -= (double) static_cast<int>(sx + ((sx >= 0.0) ? 0.5 : -0.5));
      sy -= (double) static_cast<int>(sy + ((sy >= 0.0) ? 0.5 : -0.5));
      sz -= (double) static_cast<int>(sz + ((sz >= 0.0) ? 0.5 : -0.5));
      dx = P.cell[0] * sx + P.cell[3] * sy + P.cell[6] * sz; dy = P.cell[1] * sx + P.cell[4] * sy + P.cell[7] * sz; dz = P.cell[2] * sx + P.cell[5] * sy + P.cell[8] * sz;
    }
    x[first + a] = cx * A.scale + dx; y[first + a] = cy * A.scale + dy; z[first + a] = cz * A.scale + dz;
  }
}

// staged pack of the framework (all host components' live atoms): [fx | fy | fz | q | type], each npad long
__global__ void k_build_pack(const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz,
                             const double* __restrict__ q, const double* __restrict__ scoul, const int* __restrict__ type,
                             SegList L, int npad, double* pack)
{
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= npad) return;
  int* ptype = reinterpret_cast<int*>(pack + 4 * (size_t) npad);
  int src = -1, acc = 0;
  for(int s = 0; s < L.nseg; s++) { if(g >= acc && g < acc + L.count[s]) src = L.start[s] + (g - acc); acc += L.count[s]; }
  if(src >= 0)
  {
    pack[g] = fx[src]; pack[npad + g] = fy[src]; pack[2 * (size_t) npad + g] = fz[src];
    pack[3 * (size_t) npad + g] = q[src] * scoul[src]; ptype[g] = type[src];
  }
  else { pack[g] = 0.0; pack[npad + g] = 0.0; pack[2 * (size_t) npad + g] = 0.0; pack[3 * (size_t) npad + g] = 0.0; ptype[g] = 0; }
}

// stage `bytes` of global memory into shared memory with TMA bulk copies (<= 32 KB each) on one mbarrier
__device__ __forceinline__ void stage_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  if(threadIdx.x == 0)
  {
    mbar_init(bar, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, bytes);
    uint32_t off = 0;
    while(off < bytes)
    {
      uint32_t chunk = bytes - off; if(chunk > 32768u) chunk = 32768u;
      tma_bulk_g2s(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(src) + off, chunk, bar);
      off += chunk;
    }
  }
  __syncthreads();
  mbar_wait(bar, 0);
}


// ---------------------------------------------------------------------------------------------
// Rosenbluth / Boltzmann stage inside one warp (mc_widom.h:14-39, 47-86, 305-383, 568-611).
// Lane t holds trial t.  Sums run sequentially in trial order like the host code they replace.
// ---------------------------------------------------------------------------------------------
struct RosenResult { int success; int sel_lane; int nsurv; double R; double R_minus_sel; };

__device__ __forceinline__ RosenResult rosenbluth_warp(double lb, bool surv, int ntr, double uniform, bool do_select)
{
  RosenResult r; r.success = 0; r.sel_lane = 0; r.nsurv = 0; r.R = 0.0; r.R_minus_sel = 0.0;
  const int lane = lane_id();
  const unsigned mask = __ballot_sync(0xffffffffu, surv && lane < ntr);
  r.nsurv = __popc(mask);
  if(mask == 0u) return r;
  // the largest surviving exponent: a maximum does not depend on the order it is taken in (log-step butterfly, not a trial loop)
  double largest = (surv && lane < ntr) ? lb : -INFINITY;
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) largest = fmax(largest, __shfl_xor_sync(0xffffffffu, largest, o));
  // every lane evaluates its own two exponentials once; the sums below only move them around, in trial order
  const double w_shift = exp(lb - largest), w_plain = exp(lb);
  int sel = __ffs(mask) - 1;
  // one pass for both running sums (the shifted weights of the selection and the plain Rosenbluth factor), each in trial order
  double sum = 0.0, R = 0.0;
  for(unsigned m = mask; m; m &= m - 1)
  {
    const int b = __ffs(m) - 1;
    sum += __shfl_sync(0xffffffffu, w_shift, b);
    R += __shfl_sync(0xffffffffu, w_plain, b);
  }
  if(do_select)
  {
    const double ws = uniform * sum;
    unsigned m = mask; int b = __ffs(m) - 1; m &= m - 1;
    double cumw = __shfl_sync(0xffffffffu, w_shift, b);
    while(cumw < ws && m) { b = __ffs(m) - 1; m &= m - 1; cumw += __shfl_sync(0xffffffffu, w_shift, b); }
    sel = b;
  }
  const double Rsel = __shfl_sync(0xffffffffu, w_plain, sel);
  r.sel_lane = sel; r.R = R; r.R_minus_sel = R - Rsel;
  r.success = 1;
  return r;
}

// quaternion rotation of one vector, mc_utilities.h:423-457
__device__ __forceinline__ void rotate_quaternion(double& vx, double& vy, double& vz, double u, double v, double w)
{
  const double pi = 3.14159265358979323846;
  double s1, c1, s2, c2;
  sincos(2 * pi * v, &s1, &c1); sincos(2 * pi * w, &s2, &c2);
  const double a = sqrt(1 - u), b = sqrt(u);
  const double q0 = a * s1, q1 = a * c1, q2 = b * s2, q3 = b * c2;
  const double a01 = q0 * q1, a02 = q0 * q2, a03 = q0 * q3, a11 = q1 * q1, a12 = q1 * q2, a13 = q1 * q3;
  const double a22 = q2 * q2, a23 = q2 * q3, a33 = q3 * q3;
  const double r0 = 1.0 - 2.0 * (a22 + a33), r1 = 2.0 * (a12 - a03), r2 = 2.0 * (a13 + a02);
  const double r3 = 2.0 * (a12 + a03), r4 = 1.0 - 2.0 * (a11 + a33), r5 = 2.0 * (a23 - a01);
  const double r6 = 2.0 * (a13 - a02), r7 = 2.0 * (a23 + a01), r8 = 1.0 - 2.0 * (a11 + a22);
  const double x = vx * r0 + vy * r1 + vz * r2;
  const double y = vx * r3 + vy * r4 + vz * r5;
  const double z = vx * r6 + vy * r7 + vz * r8;
  vx = x; vy = y; vz = z;
}


// ---------------------------------------------------------------------------------------------
// Batched Widom, stage B: Ewald Fourier delta of the grown molecule (GPU_EwaldDifference_General with INSERTION,
// Ewald_Energy_Functions.h:438-580), exclusion constant, tail correction, final Rosenbluth weight
// (mc_swap_utilities.h:96-108) and the block-average sums (RecordRosen data_struct.h:627-652, axpy.cu:177-185).
// One warp per insertion; the active-k table (k, temp, stored structure factors) is TMA-staged in shared memory.
// ---------------------------------------------------------------------------------------------
struct WidomB
{
  const double* rec; const int* stage; long long n; int ms;
  const double* __restrict__ tq; const double* __restrict__ tscoul;   // template charges
  // staged k table: [temp (nact) | sa (2 nact) | sf (2 nact) | kpack (nact int)]
  const double* __restrict__ ktab; int nact; int nact_pad; int stage_ktab;
  // row-ordered table (nrounds > 0: molecules of <= 4 atoms): a lane owns one (kx, ky) row of a round, the warp walks |kz| in lockstep
  const double* __restrict__ rtab; int npos; const int* __restrict__ rowmeta; const int* __restrict__ rounds; int nrounds;
  int do_ewald;
  double excl_const;       // (ExclusionIntra + ExclusionAtom) * scale^2, Ewald_Energy_Functions.h:553-556
  double tail;             // TailCorrectionDifference for this component (same for every ghost insertion)
  int nbins; long long gfirst, gn;   // bins on the global insertion index (sharded jobs)
  double* out8; int* out_stage;      // may be null
  double* partial;                   // [gridDim.x][nbins][12]
};

#ifndef GBK_EWALD_THREADS
#define GBK_EWALD_THREADS 384   // 12 warps per SM: 13.8 ms per 400 000 insertions against 15.8 ms with 8 and 16.3 ms with 16 (128-register cap)
#endif
#ifndef GBK_EWALD_CTAS
#define GBK_EWALD_CTAS 1
#endif
__global__ void __launch_bounds__(GBK_EWALD_THREADS, GBK_EWALD_CTAS)
k_widom_ewald(DevParams P, WidomB B)
{
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  double* ktab = reinterpret_cast<double*>(smem + 16);
  const size_t ktab_bytes = (B.do_ewald && B.stage_ktab) ? (size_t) B.nact_pad * 44 : 0;
  unsigned char* wbase = smem + 16 + (ktab_bytes + 15) / 16 * 16;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = lane_id();
  const int n = B.ms;
  const int kx1 = P.kmax[0] + 1, ky1 = P.kmax[1] + 1, kz1 = P.kmax[2] + 1;
  const size_t per_warp = ((size_t) n * (kx1 + ky1 + kz1) * sizeof(cplx) + (size_t) n * 4 * sizeof(double) + 15) / 16 * 16;
  cplx* ex = reinterpret_cast<cplx*>(wbase + warp * per_warp);
  cplx* ey = ex + (size_t) n * kx1; cplx* ez = ey + (size_t) n * ky1;
  double* qeff = reinterpret_cast<double*>(ez + (size_t) n * kz1);
  double* mpos = qeff + n;
  double* bins = reinterpret_cast<double*>(wbase + nwarps * per_warp);   // [nwarps][nbins][12]

  const double* g_temp = B.ktab; const double* g_sa = B.ktab + B.nact_pad; const double* g_sf = B.ktab + 3 * (size_t) B.nact_pad;
  const int* g_kp = reinterpret_cast<const int*>(B.ktab + 5 * (size_t) B.nact_pad);
  if(B.do_ewald && B.stage_ktab)
  {
    stage_bulk(ktab, B.ktab, (uint32_t) ktab_bytes, bar);
    g_temp = ktab; g_sa = ktab + B.nact_pad; g_sf = ktab + 3 * (size_t) B.nact_pad;
    g_kp = reinterpret_cast<const int*>(ktab + 5 * (size_t) B.nact_pad);
  }
  for(int k = lane; k < B.nbins * 12; k += 32) bins[(size_t) warp * B.nbins * 12 + k] = 0.0;
  __syncwarp();
  const long long gw = (long long) blockIdx.x * nwarps + warp, tw = (long long) gridDim.x * nwarps;
  const int rec_stride = 5 + 3 * B.ms;
  for(long long ins = gw; ins < B.n; ins += tw)
  {
    const int st = B.stage[ins];
    const int bin = (int)(((B.gfirst + ins) * B.nbins) / B.gn);
    double* mybin = bins + ((size_t) warp * B.nbins + bin) * 12;
    double o8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if(st == 0)
    {
      const double* rec = B.rec + (size_t) ins * rec_stride;
      double W = rec[0];
      double same = 0.0, cross = 0.0;
      if(B.do_ewald)
      {
        for(int k = lane; k < 3 * n; k += 32) mpos[k] = rec[5 + k];
        for(int k = lane; k < n; k += 32) qeff[k] = B.tscoul[k] * B.tq[k];
        __syncwarp();
        build_eik(P, mpos, n, ex, ey, ez, lane, 32);
        __syncwarp();
#if defined(GBK_EWALD_UNROLL) && GBK_EWALD_UNROLL == 2
#pragma unroll 2
#elif defined(GBK_EWALD_UNROLL) && GBK_EWALD_UNROLL == 4
#pragma unroll 4
#endif
        if(B.nrounds > 0)
        {
          // e^{i kx x} e^{i ky y} q is formed once per row and lane; e^{i kz z} is the same shared-memory word for the whole warp, and
          // +kz / -kz come out of one load.  Half the instructions per wave vector of the flat loop below.
          const double* __restrict__ Tt = B.rtab;
          const double* __restrict__ Tar = B.rtab + B.npos; const double* __restrict__ Tai = B.rtab + 2 * (size_t) B.npos;
          const double* __restrict__ Tfr = B.rtab + 3 * (size_t) B.npos; const double* __restrict__ Tfi = B.rtab + 4 * (size_t) B.npos;
          double same1 = 0.0, cross1 = 0.0;
          for(int g = 0; g < B.nrounds; g++)
          {
            const int off = B.rounds[2 * g] + lane, m = B.rounds[2 * g + 1];
            const int meta = B.rowmeta[g * 32 + lane];
            const int kx = meta >> 16, ky = ((meta >> 8) & 255) - 128, aky = ky < 0 ? -ky : ky;
            cplx a[4];
#pragma unroll
            for(int i = 0; i < 4; i++)
            {
              a[i].re = 0.0; a[i].im = 0.0;
              if(i < n)
              {
                cplx t1 = ey[aky * n + i]; if(ky < 0) t1.im = -t1.im;
                const cplx exy = cmul(ex[kx * n + i], t1);
                a[i].re = qeff[i] * exy.re; a[i].im = qeff[i] * exy.im;
              }
            }
            for(int j = 0; j <= m; j++)
            {
              double pr = 0.0, pi = 0.0, mr = 0.0, mi = 0.0;
#pragma unroll
              for(int i = 0; i < 4; i++)
                if(i < n)
                {
                  const cplx b = ez[j * n + i];
                  const double rr = a[i].re * b.re, ii = a[i].im * b.im, ri = a[i].re * b.im, ir = a[i].im * b.re;
                  pr += rr - ii; pi += ri + ir;          // a * b
                  mr += rr + ii; mi += ir - ri;          // a * conj(b)
                }
              const int p0 = off + j * 64, p1 = p0 + 32;
              // |S + d|^2 - |S|^2 = |d|^2 + 2 Re(conj(S) d): no cancellation of the stored term, and one accumulation per
              // wave vector on each of four independent accumulators
              {
                const double temp = Tt[p0], ore = Tar[p0], oim = Tai[p0];
                same += temp * ((pr * pr + pi * pi) + 2.0 * (ore * pr + oim * pi));
                cross += temp * (Tfr[p0] * pr + Tfi[p0] * pi);
              }
              {
                const double temp = Tt[p1], ore = Tar[p1], oim = Tai[p1];
                same1 += temp * ((mr * mr + mi * mi) + 2.0 * (ore * mr + oim * mi));
                cross1 += temp * (Tfr[p1] * mr + Tfi[p1] * mi);
              }
            }
          }
          same += same1; cross += cross1;
        }
        else
        for(int kk = lane; kk < B.nact; kk += 32)
        {
          int kx, ky, kz; unpack_k(g_kp[kk], kx, ky, kz);
          const cplx d = ck_sum(ex, ey, ez, qeff, n, 0, n, kx, ky, kz);
          const double temp = g_temp[kk];
          const double ore = g_sa[2 * kk], oim = g_sa[2 * kk + 1];
          const double nre = ore + d.re, nim = oim + d.im;
          same += temp * (nre * nre + nim * nim);
          same -= temp * (ore * ore + oim * oim);
          cross += temp * (g_sf[2 * kk] * d.re + g_sf[2 * kk + 1] * d.im);
        }
        same = warp_sum(same); cross = warp_sum(cross);
        same -= B.excl_const; cross *= 2.0;
        W *= exp(-P.beta * (same + cross));
        __syncwarp();
      }
      W *= exp(-P.beta * B.tail);
      o8[0] = W; o8[1] = rec[1]; o8[2] = rec[2]; o8[3] = rec[3]; o8[4] = rec[4]; o8[5] = same; o8[6] = cross; o8[7] = B.tail;
    }
    if(lane == 0)
    {
      if(B.out8) for(int k = 0; k < 8; k++) B.out8[(size_t) ins * 8 + k] = o8[k];
      if(B.out_stage) B.out_stage[ins] = st;
      mybin[0] += o8[0]; mybin[1] += o8[0] * o8[0]; mybin[2] += 1.0;
      for(int k = 0; k < 7; k++) mybin[3 + k] += o8[0] * o8[1 + k];
      if(st != 0) mybin[10] += 1.0;
    }
  }
  __syncthreads();
  // fixed-order CTA reduction of the per-warp bins
  for(int k = threadIdx.x; k < B.nbins * 12; k += blockDim.x)
  {
    double s = 0.0;
    for(int w = 0; w < nwarps; w++) s += bins[(size_t) w * B.nbins * 12 + k];
    B.partial[(size_t) blockIdx.x * B.nbins * 12 + k] = s;
  }
}

__global__ void k_reduce_partials(const double* __restrict__ partial, int nparts, int width, double* out)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= width) return;
  double s = 0.0;
  for(int p = 0; p < nparts; p++) s += partial[(size_t) p * width + k];
  out[k] = s;
}

// ---------------------------------------------------------------------------------------------
// Single-move Ewald delta: eik tables + per-k delta in one launch (replaces Initialize_WaveVector_General +
// Fourier_Ewald_Diff + the host block sum).  Threads run over ACTIVE k; writes tempEik; last CTA sums the
// CTA partials in fixed order and stores {same, 2*cross}.
// ---------------------------------------------------------------------------------------------
struct EwaldDeltaArgs
{
  const double* __restrict__ pos3;      // (nold + nnew) xyz
  const double* __restrict__ qeff;      // charge * scaleCoul
  int nold, nnew;
  KTable K;
  const double* __restrict__ same_sf;   // full arrays (nvec complex)
  const double* __restrict__ cross_sf;
  double* temp_sf;
  double* partial;                      // [gridDim.x][2]
  unsigned int* ticket;
  double* result;                       // {same, 2*cross}
  const double* dep;                    // fused moves: skip (result = 0) unless dep[0] != 0; nullptr: always run
};

__global__ void __launch_bounds__(128)
k_ewald_delta(DevParams P, EwaldDeltaArgs A)
{
  extern __shared__ __align__(16) unsigned char smem[];
  if(A.dep && A.dep[0] == 0.0)
  {
    if(blockIdx.x == 0 && threadIdx.x == 0) { A.result[0] = 0.0; A.result[1] = 0.0; }
    return;
  }
  const int n = A.nold + A.nnew;
  const int kx1 = P.kmax[0] + 1, ky1 = P.kmax[1] + 1, kz1 = P.kmax[2] + 1;
  cplx* ex = reinterpret_cast<cplx*>(smem);
  cplx* ey = ex + (size_t) n * kx1; cplx* ez = ey + (size_t) n * ky1;
  double* qeff = reinterpret_cast<double*>(ez + (size_t) n * kz1);
  __shared__ double red[2][4];
  __shared__ bool last;
  for(int k = threadIdx.x; k < n; k += blockDim.x) qeff[k] = A.qeff[k];
  build_eik(P, A.pos3, n, ex, ey, ez, threadIdx.x, blockDim.x);
  __syncthreads();
  const int kk = blockIdx.x * blockDim.x + threadIdx.x;
  double same = 0.0, cross = 0.0;
  if(kk < A.K.nact)
  {
    int kx, ky, kz; unpack_k(A.K.kpack[kk], kx, ky, kz);
    const cplx co = ck_sum(ex, ey, ez, qeff, n, 0, A.nold, kx, ky, kz);
    const cplx cn = ck_sum(ex, ey, ez, qeff, n, A.nold, n, kx, ky, kz);
    const double temp = A.K.temp[kk];
    const int slot = A.K.slot[kk];
    const double ore = A.same_sf[2 * slot], oim = A.same_sf[2 * slot + 1];
    const double nre = ore + cn.re - co.re, nim = oim + cn.im - co.im;
    same += temp * (nre * nre + nim * nim);
    same -= temp * (ore * ore + oim * oim);
    A.temp_sf[2 * slot] = nre; A.temp_sf[2 * slot + 1] = nim;
    cross += temp * (A.cross_sf[2 * slot] * (cn.re - co.re) + A.cross_sf[2 * slot + 1] * (cn.im - co.im));
  }
  same = warp_sum(same); cross = warp_sum(cross);
  if(lane_id() == 0) { red[0][threadIdx.x >> 5] = same; red[1][threadIdx.x >> 5] = cross; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    double s = 0.0, c = 0.0;
    for(int w = 0; w < (int)(blockDim.x >> 5); w++) { s += red[0][w]; c += red[1][w]; }
    A.partial[2 * blockIdx.x] = s; A.partial[2 * blockIdx.x + 1] = c;
    __threadfence();
    const unsigned int t = atomicAdd(A.ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if(last && threadIdx.x == 0)
  {
    __threadfence();
    double s = 0.0, c = 0.0;
    const volatile double* p = A.partial;
    for(unsigned int b = 0; b < gridDim.x; b++) { s += p[2 * b]; c += p[2 * b + 1]; }
    A.result[0] = s; A.result[1] = 2.0 * c;
    *A.ticket = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// Totals
// ---------------------------------------------------------------------------------------------
// total structure factors + Fourier energies (Ewald_Total ewald_preparation.h:84-171 / TotalFourierEwald
// Ewald_Energy_Functions.h:834-1147): one CTA per active k, threads stride over atoms.
struct EwaldTotalArgs
{
  const double* __restrict__ x; const double* __restrict__ y; const double* __restrict__ z;
  const double* __restrict__ q; const double* __restrict__ scoul;
  SegList L;                 // all live ranges; kind 0/1 = framework (comp < nhost), 2 = adsorbate
  KTable K;
  double* sf_ads; double* sf_fw;   // full arrays, written at active slots (others must be pre-zeroed)
  double* ek;                       // [nact][3] per-k GG, HH, HG contributions
  int has_fw;
};

__global__ void __launch_bounds__(128)
k_ewald_total(DevParams P, EwaldTotalArgs A)
{
  const int kk = blockIdx.x;
  int kx, ky, kz; unpack_k(A.K.kpack[kk], kx, ky, kz);
  double are = 0, aim = 0, fre = 0, fim = 0;
  for(int s = 0; s < A.L.nseg; s++)
  {
    const bool fw = A.L.kind[s] != 2;
    for(int i = A.L.start[s] + threadIdx.x; i < A.L.start[s] + A.L.count[s]; i += blockDim.x)
    {
      double sx, sy, sz; to_frac(P, A.x[i], A.y[i], A.z[i], sx, sy, sz);
      // phase of exp(i k.r) = 2 pi (kx sx + ky sy + kz sz); evaluated directly (no recurrence) with exact integer k
      const double ph = 2 * GBK_PI * ((double) kx * sx + (double) ky * sy + (double) kz * sz);
      double sn, cs; sincos(ph, &sn, &cs);
      const double w = A.scoul[i] * A.q[i];
      if(fw) { fre += w * cs; fim += w * sn; } else { are += w * cs; aim += w * sn; }
    }
  }
  __shared__ double red[4][4];
  are = warp_sum(are); aim = warp_sum(aim); fre = warp_sum(fre); fim = warp_sum(fim);
  const int w = threadIdx.x >> 5;
  if(lane_id() == 0) { red[w][0] = are; red[w][1] = aim; red[w][2] = fre; red[w][3] = fim; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    double v[4] = {0, 0, 0, 0};
    for(int i = 0; i < (int)(blockDim.x >> 5); i++) for(int k = 0; k < 4; k++) v[k] += red[i][k];
    if(!A.has_fw) { v[0] += v[2]; v[1] += v[3]; v[2] = 0.0; v[3] = 0.0; }   // no framework: everything is "adsorbate" (ewald_preparation.h:141-150)
    const double temp = A.K.temp[kk];
    const int slot = A.K.slot[kk];
    if(A.sf_ads) { A.sf_ads[2 * slot] = v[0]; A.sf_ads[2 * slot + 1] = v[1]; A.sf_fw[2 * slot] = v[2]; A.sf_fw[2 * slot + 1] = v[3]; }
    A.ek[3 * kk]     = temp * (v[0] * v[0] + v[1] * v[1]);
    A.ek[3 * kk + 1] = temp * (v[2] * v[2] + v[3] * v[3]);
    A.ek[3 * kk + 2] = temp * (v[2] * v[0] + v[3] * v[1]) * 2.0;
  }
}

// self + intra-molecular exclusion (ewald_preparation.h:176-227): one thread per live atom i, pairs (i, j>i) of its
// molecule.  Framework atoms of an even supercell sit at EXACT half-box separations, where the image the reference
// picks is decided by its own rounding; this (cold) kernel therefore follows the reference's Cartesian arithmetic
// operation by operation, unfused, with its truncating round (maths.cuh:437-448), instead of the hot path's
// fractional-space formulation.
__device__ __forceinline__ double ref_round_off(double s)
{
  return __dsub_rn(s, (double) static_cast<int>(__dadd_rn(s, (s >= 0.0) ? 0.5 : -0.5)));
}
__device__ __forceinline__ double dot3_unfused(double a0, double b0, double a1, double b1, double a2, double b2)
{
  return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}
__device__ __forceinline__ double min_image_r2_reference_order(const DevParams& P, double dx, double dy, double dz)
{
  if(P.cubic)
  {
    dx = __dsub_rn(dx, __dmul_rn((double) static_cast<int>(__dadd_rn(__dmul_rn(dx, P.inv[0]), (dx >= 0.0) ? 0.5 : -0.5)), P.cell[0]));
    dy = __dsub_rn(dy, __dmul_rn((double) static_cast<int>(__dadd_rn(__dmul_rn(dy, P.inv[4]), (dy >= 0.0) ? 0.5 : -0.5)), P.cell[4]));
    dz = __dsub_rn(dz, __dmul_rn((double) static_cast<int>(__dadd_rn(__dmul_rn(dz, P.inv[8]), (dz >= 0.0) ? 0.5 : -0.5)), P.cell[8]));
    return dot3_unfused(dx, dx, dy, dy, dz, dz);
  }
  double sx = dot3_unfused(P.inv[0], dx, P.inv[3], dy, P.inv[6], dz);
  double sy = dot3_unfused(P.inv[1], dx, P.inv[4], dy, P.inv[7], dz);
  double sz = dot3_unfused(P.inv[2], dx, P.inv[5], dy, P.inv[8], dz);
  sx = ref_round_off(sx); sy = ref_round_off(sy); sz = ref_round_off(sz);
  const double px = dot3_unfused(P.cell[0], sx, P.cell[3], sy, P.cell[6], sz);
  const double py = dot3_unfused(P.cell[1], sx, P.cell[4], sy, P.cell[7], sz);
  const double pz = dot3_unfused(P.cell[2], sx, P.cell[5], sy, P.cell[8], sz);
  return dot3_unfused(px, px, py, py, pz, pz);
}

struct ExclArgs
{
  const double* __restrict__ x; const double* __restrict__ y; const double* __restrict__ z;
  const double* __restrict__ q; const double* __restrict__ scoul;
  int start, natoms, ms;
  double* out;      // [natoms][2] self, intra (pairs i<j attributed to i)
};
__global__ void k_ewald_exclusion(DevParams P, ExclArgs A)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= A.natoms) return;
  const int i = A.start + t;
  const int mol_end = A.start + (t / A.ms + 1) * A.ms;
  const double pself = P.prefactor * P.alpha / sqrt(GBK_PI);
  const double fa = A.scoul[i] * A.q[i];
  const double self = pself * fa * fa;
  double intra = 0.0;
  const double xi = A.x[i], yi = A.y[i], zi = A.z[i];
  for(int j = i + 1; j < mol_end; j++)
  {
    const double fb = A.scoul[j] * A.q[j];
    const double r = sqrt(min_image_r2_reference_order(P, __dsub_rn(xi, A.x[j]), __dsub_rn(yi, A.y[j]), __dsub_rn(zi, A.z[j])));
    intra += P.prefactor * fa * fb * erf(P.alpha * r) / r;
  }
  A.out[2 * t] = self; A.out[2 * t + 1] = intra;
}

// total VDW + real: every live atom is a one-atom "trial group" against all live atoms with the own-molecule
// exclusion; 0.5 x the double-counted sum, exactly the reference's CPU loop structure (VDW_Coulomb.cu:94-206).

// ---------------------------------------------------------------------------------------------
// FP64 FMA peak microbenchmark (the roofline denominator bench.py reports for the pair kernel)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fp64_peak(double* out, int iters, double seed)
{
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.9999999, c = 1e-7;
  for(int i = 0; i < iters; i++)
  {
#pragma unroll
    for(int u = 0; u < 8; u++)
    {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
