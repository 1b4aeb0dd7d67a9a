/* TEST INFRASTRUCTURE ONLY -- see graspa_oracle.h.  Plain-C restatement of the gRASPA hot path.
 * File:line citations are relative to /root/reference/src_clean. */
#include "graspa_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846
#define ORC_THREADS 128 /* DEFAULTTHREAD, data_struct.h:17 */

/* ------------------------------------------------------------------ setup */

/* maths.cuh:28-56 */
void orc_inverse_cell(const double* x, double* r, double* det)
{
  double m11 = x[0], m21 = x[3], m31 = x[6];
  double m12 = x[1], m22 = x[4], m32 = x[7];
  double m13 = x[2], m23 = x[5], m33 = x[8];
  double d = +m11 * (m22 * m33 - m23 * m32) - m12 * (m21 * m33 - m23 * m31) + m13 * (m21 * m32 - m22 * m31);
  r[0] = +(m22 * m33 - m32 * m23) / d;
  r[3] = -(m21 * m33 - m31 * m23) / d;
  r[6] = +(m21 * m32 - m31 * m22) / d;
  r[1] = -(m12 * m33 - m32 * m13) / d;
  r[4] = +(m11 * m33 - m31 * m13) / d;
  r[7] = -(m11 * m32 - m31 * m12) / d;
  r[2] = +(m12 * m23 - m22 * m13) / d;
  r[5] = -(m11 * m23 - m21 * m13) / d;
  r[8] = +(m11 * m22 - m21 * m12) / d;
  *det = d;
}

/* read_data.cpp:1517-1547 */
void orc_cell_from_cif(double a, double b, double c, double al, double be, double ga, int nx, int ny, int nz, double* cell)
{
  double ax = al / (180.0 / 3.14159265358979323846);
  double ay = be / (180.0 / 3.14159265358979323846);
  double az = ga / (180.0 / 3.14159265358979323846);
  double dx = a;
  double dy = b * sin(az);
  double tempd = (cos(ax) - cos(az) * cos(ay)) / sin(az);
  double dz = c * sqrt(1 - pow(cos(ay), 2) - pow(tempd, 2));
  double bx = b * cos(az);
  double cx = c * cos(ay);
  double cy = c * tempd;
  cell[0] = nx * dx; cell[1] = 0.0;     cell[2] = 0.0;
  cell[3] = ny * bx; cell[4] = ny * dy; cell[5] = 0.0;
  cell[6] = nz * cx; cell[7] = nz * cy; cell[8] = nz * dz;
}

/* read_data.cpp:609-611, 691-702 (RASPA-2 heuristic) */
void orc_ewald_setup(double cutoff, double precision, orc_box* box)
{
  box->prefactor = 138935.483496;
  double tol = sqrt(fabs(log(precision * cutoff)));
  double alpha = sqrt(fabs(log(precision * cutoff * tol))) / cutoff;
  double tol1 = sqrt(-log(precision * cutoff * pow(2.0 * tol * alpha, 2)));
  box->alpha = alpha;
  box->kmax[0] = (int32_t) round(0.25 + box->cell[0] * alpha * tol1 / ORC_PI);
  box->kmax[1] = (int32_t) round(0.25 + box->cell[4] * alpha * tol1 / ORC_PI);
  box->kmax[2] = (int32_t) round(0.25 + box->cell[8] * alpha * tol1 / ORC_PI);
  int m = box->kmax[0]; if(box->kmax[1] > m) m = box->kmax[1]; if(box->kmax[2] > m) m = box->kmax[2];
  box->recip_cutoff = pow(1.05 * (double) m, 2);
  box->use_lammps_ewald = 0;
}

/* Get_Shifted_Value: the LJ energy at the cutoff through the same VDW routine (read_data.cpp, used at :1236-1239) */
static double shifted_value(double eps, double sig, double cutsq)
{
  double ffarg[5] = {eps, sig, 0.0, 0.0, 0.0}; double res[2];
  orc_vdw(ffarg, cutsq, 1.0, 0, res);
  return res[0];
}

/* read_data.cpp:833-846 (GetTailCorrectionValue) */
static double tail_value(double eps, double sig, double cutsq)
{
  double arg1 = eps, arg2 = sig * sig * sig, rr = sqrt(cutsq);
  double term1 = pow(arg2, 4) / (9.0 * pow(rr, 9));
  double term2 = pow(arg2, 2) / (3.0 * pow(rr, 3));
  return 16.0 * 3.14159265358979323846 / 2.0 * arg1 * (term1 - term2);
}

/* read_data.cpp:1179-1247 (ForceField_Processing, Use1264=false) */
void orc_ff_mix(int n, const double* eps_in, const double* sig_in, const int* shifted, const int* tail, double cutsq,
                double* eps, double* sigma, double* shift, int* use_tail, double* tail_e)
{
  for(int i = 0; i < n; i++)
    for(int j = 0; j < n; j++)
    {
      double e = sqrt(eps_in[i] * eps_in[j]) / 1.20272430057;
      double s = 0.5 * (sig_in[i] + sig_in[j]);
      eps[i*n+j] = e; sigma[i*n+j] = s;
      shift[i*n+j] = (shifted[i] && shifted[j]) ? shifted_value(e, s, cutsq) : 0.0;
      use_tail[i*n+j] = 0; tail_e[i*n+j] = 0.0;
    }
  for(int i = 0; i < n; i++)
    for(int j = 0; j < n; j++)
      if(tail[i] && tail[j]) { use_tail[i*n+j] = 1; tail_e[i*n+j] = tail_value(eps[i*n+j], sigma[i*n+j], cutsq); }
}

/* ------------------------------------------------------------------ pair primitives */

/* maths.cuh:427-450 */
void orc_pbc(double* p, const orc_box* b)
{
  const double* C = b->cell; const double* I = b->inv;
  if(b->cubic)
  {
    p[0] -= (int)(p[0] * I[0] + ((p[0] >= 0.0) ? 0.5 : -0.5)) * C[0];
    p[1] -= (int)(p[1] * I[4] + ((p[1] >= 0.0) ? 0.5 : -0.5)) * C[4];
    p[2] -= (int)(p[2] * I[8] + ((p[2] >= 0.0) ? 0.5 : -0.5)) * C[8];
  }
  else
  {
    double sx = I[0]*p[0] + I[3]*p[1] + I[6]*p[2];
    double sy = I[1]*p[0] + I[4]*p[1] + I[7]*p[2];
    double sz = I[2]*p[0] + I[5]*p[1] + I[8]*p[2];
    sx -= (int)(sx + ((sx >= 0.0) ? 0.5 : -0.5));
    sy -= (int)(sy + ((sy >= 0.0) ? 0.5 : -0.5));
    sz -= (int)(sz + ((sz >= 0.0) ? 0.5 : -0.5));
    p[0] = C[0]*sx + C[3]*sy + C[6]*sz;
    p[1] = C[1]*sx + C[4]*sy + C[7]*sz;
    p[2] = C[2]*sx + C[5]*sy + C[8]*sz;
  }
}

/* BlockedPocket, read_data.cpp:3466-3640: 1 when pos lies inside any of the n spheres {x, y, z, radius} (invert: when it
 * lies outside all of them).  The difference centre - pos takes the nearest image with RASPA-2's rule (apply_pbc_raspa2:
 * cubic boxes divide by the cell edge, :3519-3528; otherwise the fractional round of :3531-3547). */
int orc_blocked_pocket(const orc_box* b, const double* pockets, int n, int invert, const double* pos)
{
  const double* C = b->cell; const double* I = b->inv;
  if(n <= 0) return 0;
  for(int i = 0; i < n; i++)
  {
    double dx = pockets[4*i] - pos[0], dy = pockets[4*i+1] - pos[1], dz = pockets[4*i+2] - pos[2];
    if(b->cubic)
    {
      dx -= C[0] * (int)(dx / C[0] + ((dx >= 0.0) ? 0.5 : -0.5));
      dy -= C[4] * (int)(dy / C[4] + ((dy >= 0.0) ? 0.5 : -0.5));
      dz -= C[8] * (int)(dz / C[8] + ((dz >= 0.0) ? 0.5 : -0.5));
    }
    else
    {
      double sx = I[0]*dx + I[3]*dy + I[6]*dz, sy = I[1]*dx + I[4]*dy + I[7]*dz, sz = I[2]*dx + I[5]*dy + I[8]*dz;
      double tx = sx - (int)(sx + ((sx >= 0.0) ? 0.5 : -0.5));
      double ty = sy - (int)(sy + ((sy >= 0.0) ? 0.5 : -0.5));
      double tz = sz - (int)(sz + ((sz >= 0.0) ? 0.5 : -0.5));
      dx = C[0]*tx + C[3]*ty + C[6]*tz; dy = C[1]*tx + C[4]*ty + C[7]*tz; dz = C[2]*tx + C[5]*ty + C[8]*tz;
    }
    if(sqrt(dx*dx + dy*dy + dz*dz) < pockets[4*i+3]) return invert ? 0 : 1;
  }
  return invert ? 1 : 0;
}

/* maths.cuh:452-494 */
void orc_vdw(const double* F, double rr, double scaling, int use1264, double* result)
{
  if(use1264)
  {
    double C12 = F[0], C6 = F[1], C4 = F[2], shift = F[3], C10 = F[4];
    double ri2 = 1.0 / rr, ri4 = ri2 * ri2, ri6 = ri4 * ri2, ri10 = ri4 * ri6, ri12 = ri6 * ri6;
    double term = C12 * ri12 - C6 * ri6 + C10 * ri10 + C4 * ri4 - shift;
    result[0] = scaling * term; result[1] = 0.0;
  }
  else
  {
    double arg1 = 4.0 * F[0], arg2 = F[1] * F[1], arg3 = F[3];
    double temp = rr / arg2;
    double temp3 = temp * temp * temp;
    double rri3 = 1.0 / (temp3 + 0.5 * (1.0 - scaling) * (1.0 - scaling));
    double rri6 = rri3 * rri3;
    double term = arg1 * (rri3 * (rri3 - 1.0)) - arg3;
    double dl = scaling * arg1 * (rri6 * (2.0 * rri3 - 1.0));
    result[0] = scaling * term;
    result[1] = scaling < 1.0 ? term + (1.0 - scaling) * dl : 0.0;
  }
}

/* maths.cuh:496-500 */
double orc_coulomb_real(double qa, double qb, double r, double scaling, double prefactor, double alpha)
{
  double term = qa * qb * erfc(alpha * r);
  return prefactor * scaling * term / r;
}

/* component c owns slots [off[c], off[c]+alloc[c]); the first natoms[c] are live.  Slots beyond the live
 * range keep stale/template data exactly like the reference's device arrays (read_data.cpp:2122-2147). */
static void sys_offsets(const orc_system* s, int64_t* off)
{
  off[0] = 0;
  for(int c = 0; c < s->ncomp; c++) off[c+1] = off[c] + (s->alloc ? s->alloc[c] : s->natoms[c]);
}
static int64_t live_count(const orc_system* s, int c0, int c1)
{
  int64_t n = 0; for(int c = c0; c < c1; c++) n += s->natoms[c]; return n;
}

/* ------------------------------------------------------------------ trial generation */

/* mc_utilities.h:423-457 */
void orc_rotate_quaternions(double* V, const double* R)
{
  const double u = R[0], v = R[1], w = R[2];
  const double pi = 3.14159265358979323846;
  const double q0 = sqrt(1-u) * sin(2*pi*v);
  const double q1 = sqrt(1-u) * cos(2*pi*v);
  const double q2 = sqrt(u)   * sin(2*pi*w);
  const double q3 = sqrt(u)   * cos(2*pi*w);
  double rot[9];
  const double a01=q0*q1, a02=q0*q2, a03=q0*q3;
  const double a11=q1*q1, a12=q1*q2, a13=q1*q3;
  const double a22=q2*q2, a23=q2*q3, a33=q3*q3;
  rot[0]=1.0-2.0*(a22+a33); rot[1]=2.0*(a12-a03);     rot[2]=2.0*(a13+a02);
  rot[3]=2.0*(a12+a03);     rot[4]=1.0-2.0*(a11+a33); rot[5]=2.0*(a23-a01);
  rot[6]=2.0*(a13-a02);     rot[7]=2.0*(a23+a01);     rot[8]=1.0-2.0*(a11+a22);
  const double r = V[0]*rot[0] + V[1]*rot[1] + V[2]*rot[2];
  const double s = V[0]*rot[3] + V[1]*rot[4] + V[2]*rot[5];
  const double c = V[0]*rot[6] + V[1]*rot[7] + V[2]*rot[8];
  V[0] = r; V[1] = s; V[2] = c;
}

/* mc_widom.h:122-213.  scale/scale_coul = proposed_scale for insertion-type moves; for the
 * deletion/retrace types they are read from the existing molecule. */
void orc_trial_positions(const orc_box* box, const orc_system* sys, int movetype, int comp, int64_t start,
                         int ntrials, const double* rnd, double pscale, double pscale_coul,
                         double* tpos, double* tscale, double* tcharge, double* tscale_coul, int64_t* ttype)
{
  int64_t off[65]; sys_offsets(sys, off);
  const int64_t g = off[comp] + start;
  const double L[3] = {box->cell[0], box->cell[4], box->cell[8]};
  for(int i = 0; i < ntrials; i++)
  {
    double scale = 0.0, scoul = 0.0;
    int from_random = 0;
    switch(movetype)
    {
      case ORC_CBMC_INSERTION:        scale = pscale; scoul = pscale_coul; from_random = 1; break;
      case ORC_CBMC_DELETION: case ORC_REINSERTION_RETRACE:
        scale = sys->scale[g]; scoul = sys->scale_coul[g]; from_random = (i != 0); break;
      case ORC_REINSERTION_INSERTION: scale = sys->scale[g]; scoul = sys->scale_coul[g]; from_random = 1; break;
      case ORC_IDENTITY_SWAP_NEW:     scale = pscale; scoul = pscale_coul; from_random = 2; break; /* position preset by caller */
      case ORC_IDENTITY_SWAP_OLD:     scale = sys->scale[g]; scoul = sys->scale_coul[g]; from_random = (i == 0) ? 0 : 2; break;
    }
    if(from_random == 1) for(int d = 0; d < 3; d++) tpos[3*i+d] = L[d] * rnd[3*i+d];
    else if(from_random == 0) for(int d = 0; d < 3; d++) tpos[3*i+d] = sys->pos[3*g+d];
    tscale[i] = scale; tcharge[i] = sys->charge[g]; tscale_coul[i] = scoul; ttype[i] = sys->type[g];
  }
}

/* mc_widom.h:215-303.  start = start_position (index of the first chain atom's template: 1 for insertion). */
void orc_trial_orientations(const orc_system* sys, int movetype, int comp, int64_t start, int chainsize,
                            int norient, const double* rnd, const double* fb, double fb_scale, double fb_scale_coul,
                            double* tpos, double* tscale, double* tcharge, double* tscale_coul, int64_t* ttype)
{
  int64_t off[65]; sys_offsets(sys, off);
  const int64_t base = off[comp];
  for(int i = 0; i < norient * chainsize; i++)
  {
    int trial = i / chainsize, a = i % chainsize;
    double V[3];
    for(int d = 0; d < 3; d++) V[d] = sys->pos[3*(base+1+a)+d] - sys->pos[3*base+d]; /* template = molecule 0, :256 */
    int retrace_first = (movetype == ORC_CBMC_DELETION || movetype == ORC_REINSERTION_RETRACE || movetype == ORC_IDENTITY_SWAP_OLD) && trial == 0;
    if(retrace_first) for(int d = 0; d < 3; d++) tpos[3*i+d] = sys->pos[3*(base+start+a)+d];
    else
    {
      orc_rotate_quaternions(V, rnd + 3*trial);
      for(int d = 0; d < 3; d++) tpos[3*i+d] = fb[d] + V[d];
    }
    tscale[i] = fb_scale; tcharge[i] = sys->charge[base+start+a]; tscale_coul[i] = fb_scale_coul; ttype[i] = sys->type[base+start+a];
  }
}

/* ------------------------------------------------------------------ trial energies */

static void tree_reduce(double* s, int n) /* VDW_Coulomb.cu:1333-1343 */
{
  for(int r = n / 2; r != 0; r /= 2)
    for(int i = 0; i < r; i++) s[i] += s[i + r];
}

/* per-pair body, VDW_Coulomb.cu:1287-1322 */
static inline void pair_body(const orc_box* box, const orc_ff* ff, const double* pa, double scaleA, double chargeA, double scoulA, int64_t typeA,
                             const double* pb, double scaleB, double chargeB, double scoulB, int64_t typeB,
                             double* evdw, double* ecoul, int* flag, int64_t* counts)
{
  double v[3] = {pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2]};
  orc_pbc(v, box);
  const double rr = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
  int in = 0;
  if(counts) counts[0]++;
  if(rr < ff->cutoff_vdw_sq)
  {
    double res[2] = {0.0, 0.0};
    const int64_t row = typeA * ff->ntypes + typeB;
    const double F[5] = {ff->epsilon[row], ff->sigma[row], ff->z[row], ff->shift[row], ff->c10[row]};
    orc_vdw(F, rr, scaleA * scaleB, ff->use1264, res);
    if(res[0] > ff->overlap) *flag = 1;
    if(rr < 0.01) *flag = 1;
    *evdw += res[0];
    if(counts) counts[1]++;
    in = 1;
  }
  if(!ff->no_charges && rr < ff->cutoff_coul_sq)
  {
    const double r = sqrt(rr);
    *ecoul += orc_coulomb_real(chargeA, chargeB, r, scoulA * scoulB, box->prefactor, box->alpha);
    if(counts) counts[2]++;
    in = 1;
  }
  if(counts && in) counts[3]++;
}

/* VDW_Coulomb.cu:1183-1352 + mc_widom.h:42-119 */
void orc_trial_energies(const orc_box* box, const orc_ff* ff, const orc_system* sys, int ntrials, int chainsize,
                        const orc_atoms* T, int new_comp, int64_t new_molid, int excl_comp, int64_t excl_mol,
                        double* out, int32_t* out_flag, int64_t* counts)
{
  int64_t off[65]; sys_offsets(sys, off);
  int64_t nhost = live_count(sys, 0, sys->nhost), nguest = live_count(sys, sys->nhost, sys->ncomp);
  int64_t hg_blocks = (nhost * chainsize + ORC_THREADS - 1) / ORC_THREADS;
  int64_t gg_blocks = (nguest * chainsize + ORC_THREADS - 1) / ORC_THREADS;
  int64_t c4[4] = {0, 0, 0, 0};
  double sv[ORC_THREADS], sr[ORC_THREADS];
  for(int t = 0; t < ntrials; t++)
  {
    double e[4] = {0, 0, 0, 0}; int flag = 0;
    for(int seg = 0; seg < 2; seg++)
    {
      int64_t nblocks = seg == 0 ? hg_blocks : gg_blocks;
      int c0 = seg == 0 ? 0 : sys->nhost, c1 = seg == 0 ? sys->nhost : sys->ncomp;
      int64_t natom_seg = seg == 0 ? nhost : nguest;
      for(int64_t b = 0; b < nblocks; b++)
      {
        for(int k = 0; k < ORC_THREADS; k++)
        {
          sv[k] = 0.0; sr[k] = 0.0;
          int64_t ij = b * ORC_THREADS + k;
          int64_t i = ij / chainsize; int a = (int)(ij % chainsize);
          if(i >= natom_seg) continue;
          int comp = c0; int64_t posi = i;
          while(comp < c1 && posi >= sys->natoms[comp]) { posi -= sys->natoms[comp]; comp++; }
          if(comp >= c1) continue;
          int64_t g = off[comp] + posi;
          if(comp == excl_comp && sys->molid[g] == excl_mol) continue;           /* :1282 */
          if(sys->molid[g] == new_molid && comp == new_comp) continue;           /* :1283 */
          int j = t * chainsize + a;
          pair_body(box, ff, sys->pos + 3*g, sys->scale[g], sys->charge[g], sys->scale_coul[g], sys->type[g],
                    T->pos + 3*j, T->scale[j], T->charge[j], T->scale_coul[j], T->type[j], &sv[k], &sr[k], &flag, c4);
        }
        tree_reduce(sv, ORC_THREADS); tree_reduce(sr, ORC_THREADS);
        e[2*seg] += sv[0]; e[2*seg + 1] += sr[0];                                /* mc_widom.h:67-72 */
      }
    }
    for(int k = 0; k < 4; k++) out[4*t + k] = e[k];
    out_flag[t] = flag;
  }
  if(counts) for(int k = 0; k < 4; k++) counts[k] += c4[k];
}

/* ------------------------------------------------------------------ Rosenbluth */

/* mc_widom.h:14-39 */
int orc_select_trial(const double* lb, int n, double uniform)
{
  double largest = lb[0];
  for(int i = 1; i < n; i++) if(lb[i] > largest) largest = lb[i];
  double sum = 0.0; double sh[1024];
  for(int i = 0; i < n; i++) { sh[i] = exp(lb[i] - largest); sum += sh[i]; }
  int selected = 0; double cumw = sh[0]; double ws = uniform * sum;
  while(cumw < ws) cumw += sh[++selected];
  return selected;
}

/* mc_widom.h:305-383 (is_chain=0) and :568-611 (is_chain=1).  The VDWRealBias=false correction is applied by the caller. */
int orc_cbmc_finish(int movetype, int is_chain, double* rosen, int nsurv, int norm, double uniform,
                    double stored_in, double* stored_out, int* selected, double* rosenbluth)
{
  int good = 0; int sel = 0; double R = 0.0;
  int insertion_like = (movetype == ORC_CBMC_INSERTION || movetype == ORC_REINSERTION_INSERTION || (is_chain && movetype == ORC_IDENTITY_SWAP_NEW));
  if(insertion_like || movetype == ORC_IDENTITY_SWAP_NEW)
  {
    if(nsurv == 0) { *selected = 0; *rosenbluth = 0.0; return 0; }
    sel = (insertion_like) ? orc_select_trial(rosen, nsurv, uniform) : 0;
    for(int a = 0; a < nsurv; a++) { rosen[a] = exp(rosen[a]); }
    for(int a = 0; a < nsurv; a++) R += rosen[a];
    if(!(R < 1e-150)) good = 1;
  }
  else
  {
    sel = 0;
    for(int a = 0; a < nsurv; a++) { rosen[a] = exp(rosen[a]); }
    for(int a = 0; a < nsurv; a++) R += rosen[a];
    good = 1;
  }
  *selected = sel;
  if(!good) { *rosenbluth = 0.0; return 0; }
  if(!is_chain)
  {
    if(movetype == ORC_REINSERTION_INSERTION && stored_out) *stored_out = R - rosen[sel];
    if(movetype == ORC_REINSERTION_RETRACE) R += stored_in;
    if(movetype != ORC_IDENTITY_SWAP_OLD && movetype != ORC_IDENTITY_SWAP_NEW) R /= (double) norm;
  }
  else R = R / (double) norm;
  *rosenbluth = R;
  return 1;
}

/* ------------------------------------------------------------------ Ewald */

int64_t orc_nvec(const orc_box* b) { return (int64_t)(b->kmax[0] + 1) * (2 * b->kmax[1] + 1) * (2 * b->kmax[2] + 1); }

static double ksq_of(const orc_box* B, int kx, int ky, int kz) /* Ewald_Energy_Functions.h:303-322 */
{
  double ksqr = (double)(kx * kx + ky * ky + kz * kz);
  if(B->use_lammps_ewald)
  {
    const double lx = B->cell[0], ly = B->cell[4], lz = B->cell[8];
    const double xy = B->cell[3], xz = B->cell[6], yz = B->cell[7];
    const double ux = 2*ORC_PI/lx;
    const double uy = 2*ORC_PI*(-xy)/lx/ly;
    const double uz = 2*ORC_PI*(xy*yz - ly*xz)/lx/ly/lz;
    const double vy = 2*ORC_PI/ly;
    const double vz = 2*ORC_PI*(-yz)/ly/lz;
    const double wz = 2*ORC_PI/lz;
    const double kvx = kx*ux, kvy = kx*uy + ky*vy, kvz = kx*uz + ky*vz + kz*wz;
    ksqr = kvx*kvx + kvy*kvy + kvz*kvz;
  }
  return ksqr;
}

typedef struct { double re, im; } cplx;
static inline cplx cmul(cplx a, cplx b) { cplx c; c.re = a.re*b.re - a.im*b.im; c.im = a.re*b.im + a.im*b.re; return c; }

/* eik tables for n atoms: Ewald_Energy_Functions.h:97-104, 162-185.  ex[(k)*n + i] */
static void eik_tables(const orc_box* B, const double* pos, int n, cplx* ex, cplx* ey, cplx* ez)
{
  const double* I = B->inv;
  for(int i = 0; i < n; i++)
  {
    const double* p = pos + 3*i;
    double s[3];
    s[0] = I[0]*p[0] + I[3]*p[1] + I[6]*p[2];   /* matrix_multiply_by_vector, maths.cuh:133-138 */
    s[1] = I[1]*p[0] + I[4]*p[1] + I[7]*p[2];
    s[2] = I[2]*p[0] + I[5]*p[1] + I[8]*p[2];
    for(int d = 0; d < 3; d++) s[d] *= 2*ORC_PI;
    cplx one = {1.0, 0.0};
    ex[i] = one; ey[i] = one; ez[i] = one;
    ex[n + i].re = cos(s[0]); ex[n + i].im = sin(s[0]);
    ey[n + i].re = cos(s[1]); ey[n + i].im = sin(s[1]);
    ez[n + i].re = cos(s[2]); ez[n + i].im = sin(s[2]);
    for(int k = 2; k <= B->kmax[0]; k++) ex[k*n + i] = cmul(ex[(k-1)*n + i], ex[n + i]);
    for(int k = 2; k <= B->kmax[1]; k++) ey[k*n + i] = cmul(ey[(k-1)*n + i], ey[n + i]);
    for(int k = 2; k <= B->kmax[2]; k++) ez[k*n + i] = cmul(ez[(k-1)*n + i], ez[n + i]);
  }
}

/* Fourier_Ewald_Diff, Ewald_Energy_Functions.h:280-397, both halves of the grid, with the
 * block-of-128 tree and the serial host sum of GPU_EwaldDifference_General :542-543; result {same, 2*cross} (:579). */
void orc_ewald_delta(const orc_box* B, const double* pos, const double* charge, const double* scoul,
                     int nold, int nnew, const double* same_sf, const double* cross_sf, double* temp_sf,
                     double* out, int64_t* n_active)
{
  const int n = nold + nnew;
  const int kxm = B->kmax[0], kym = B->kmax[1], kzm = B->kmax[2];
  const int64_t nvec = orc_nvec(B);
  cplx* ex = (cplx*) malloc(sizeof(cplx) * (size_t)(n > 0 ? n : 1) * (kxm + 1));
  cplx* ey = (cplx*) malloc(sizeof(cplx) * (size_t)(n > 0 ? n : 1) * (kym + 1));
  cplx* ez = (cplx*) malloc(sizeof(cplx) * (size_t)(n > 0 ? n : 1) * (kzm + 1));
  eik_tables(B, pos, n, ex, ey, ez);
  const double alpha_sq = B->alpha * B->alpha;
  const double prefactor = B->prefactor * (2.0 * ORC_PI / B->volume);   /* :451 */
  const double* I = B->inv;
  const double ax[3] = {I[0], I[3], I[6]}, ay[3] = {I[1], I[4], I[7]}, az[3] = {I[2], I[5], I[8]};
  const int64_t nblock = (nvec + ORC_THREADS - 1) / ORC_THREADS;
  double same = 0.0, cross = 0.0; int64_t act = 0;
  double ss[ORC_THREADS], sc[ORC_THREADS];
  for(int64_t b = 0; b < nblock; b++)
  {
    for(int t = 0; t < ORC_THREADS; t++)
    {
      ss[t] = 0.0; sc[t] = 0.0;
      int64_t kxyz = b * ORC_THREADS + t;
      if(kxyz >= nvec) continue;
      int kz = (int)(kxyz % (2*kzm + 1)) - kzm;
      int kxy = (int)(kxyz / (2*kzm + 1));
      int kx = kxy / (2*kym + 1);
      int ky = kxy % (2*kym + 1) - kym;
      double ksqr = ksq_of(B, kx, ky, kz);
      if(!((ksqr > 1e-10) && (ksqr < B->recip_cutoff))) continue;
      act++;
      cplx ck_old = {0.0, 0.0}, ck_new = {0.0, 0.0};
      double kvx[3], kvy[3], kvz[3];
      for(int d = 0; d < 3; d++) { kvx[d] = ax[d] * 2.0 * ORC_PI * (double) kx; kvy[d] = ay[d] * 2.0 * ORC_PI * (double) ky; kvz[d] = az[d] * 2.0 * ORC_PI * (double) kz; }
      double factor = (kx == 0) ? (1.0 * prefactor) : (2.0 * prefactor);
      for(int i = 0; i < n; i++)
      {
        cplx t1 = ey[i + n * abs(ky)]; t1.im = ky >= 0 ? t1.im : -t1.im;
        cplx exy = cmul(ex[i + n * kx], t1);
        cplx t2 = ez[i + n * abs(kz)]; t2.im = kz >= 0 ? t2.im : -t2.im;
        cplx ti = cmul(exy, t2);
        double q = charge[i], sc_ = scoul[i];
        if(i < nold) { ck_old.re += sc_ * q * ti.re; ck_old.im += sc_ * q * ti.im; }
        else         { ck_new.re += sc_ * q * ti.re; ck_new.im += sc_ * q * ti.im; }
      }
      double kv[3] = {kvx[0] + kvy[0] + kvz[0], kvx[1] + kvy[1] + kvz[1], kvx[2] + kvy[2] + kvz[2]};
      double rksq = kv[0]*kv[0] + kv[1]*kv[1] + kv[2]*kv[2];
      double temp = factor * exp((-0.25 / alpha_sq) * rksq) / rksq;
      /* same type (first half of the grid) */
      {
        double ore = same_sf[2*kxyz], oim = same_sf[2*kxyz + 1];
        double nre = ore + ck_new.re - ck_old.re, nim = oim + ck_new.im - ck_old.im;
        double e = 0.0;
        e += temp * (nre*nre + nim*nim);
        e -= temp * (ore*ore + oim*oim);
        ss[t] = e;
        if(temp_sf) { temp_sf[2*kxyz] = nre; temp_sf[2*kxyz + 1] = nim; }
      }
      /* cross type (second half) */
      {
        double ore = cross_sf[2*kxyz], oim = cross_sf[2*kxyz + 1];
        double e = 0.0;
        e += temp * (ore * (ck_new.re - ck_old.re) + oim * (ck_new.im - ck_old.im));
        sc[t] = e;
      }
    }
    tree_reduce(ss, ORC_THREADS); tree_reduce(sc, ORC_THREADS);
    same += ss[0]; cross += sc[0];
  }
  out[0] = same; out[1] = 2.0 * cross;
  if(n_active) *n_active = act;
  free(ex); free(ey); free(ez);
}

/* ewald_preparation.h:5-259 */
void orc_ewald_total(const orc_box* B, const orc_system* S, int no_charges, double* outE, double* sf_ads, double* sf_fw)
{
  outE[0] = outE[1] = outE[2] = 0.0;
  const int64_t nvec = orc_nvec(B);
  if(sf_ads) memset(sf_ads, 0, sizeof(double) * 2 * nvec);
  if(sf_fw)  memset(sf_fw, 0, sizeof(double) * 2 * nvec);
  if(no_charges) return;
  int64_t off[65]; sys_offsets(S, off);
  const int64_t n = live_count(S, 0, S->ncomp);
  int64_t* gi = (int64_t*) malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1)); /* compact live index -> slot */
  { int64_t c_ = 0; for(int c = 0; c < S->ncomp; c++) for(int64_t p = 0; p < S->natoms[c]; p++) gi[c_++] = off[c] + p; }
  const int kxm = B->kmax[0], kym = B->kmax[1], kzm = B->kmax[2];
  const double alpha = B->alpha, alpha_sq = alpha * alpha;
  const double prefactor = B->prefactor * (2.0 * ORC_PI / B->volume);
  const double* I = B->inv;
  const double ax[3] = {I[0], I[3], I[6]}, ay[3] = {I[1], I[4], I[7]}, az[3] = {I[2], I[5], I[8]};
  const int has_fw = S->nhost > 0;
  size_t nn = (size_t)(n > 0 ? n : 1);
  cplx* ex = (cplx*) malloc(sizeof(cplx) * nn * (kxm + 1));
  cplx* ey = (cplx*) malloc(sizeof(cplx) * nn * (kym + 1));
  cplx* ez = (cplx*) malloc(sizeof(cplx) * nn * (kzm + 1));
  cplx* exy = (cplx*) malloc(sizeof(cplx) * nn);
  /* :43-80: k=0,1 for every atom, then recurrences k-major */
  for(int64_t i = 0; i < n; i++)
  {
    const double* p = S->pos + 3*gi[i]; double s[3];
    s[0] = I[0]*p[0] + I[3]*p[1] + I[6]*p[2];
    s[1] = I[1]*p[0] + I[4]*p[1] + I[7]*p[2];
    s[2] = I[2]*p[0] + I[5]*p[1] + I[8]*p[2];
    for(int d = 0; d < 3; d++) s[d] *= 2*ORC_PI;
    cplx one = {1.0, 0.0};
    ex[i] = one; ey[i] = one; ez[i] = one;
    ex[n + i].re = cos(s[0]); ex[n + i].im = sin(s[0]);
    ey[n + i].re = cos(s[1]); ey[n + i].im = sin(s[1]);
    ez[n + i].re = cos(s[2]); ez[n + i].im = sin(s[2]);
  }
  for(int k = 2; k <= kxm; k++) for(int64_t i = 0; i < n; i++) ex[k*n + i] = cmul(ex[(k-1)*n + i], ex[n + i]);
  for(int k = 2; k <= kym; k++) for(int64_t i = 0; i < n; i++) ey[k*n + i] = cmul(ey[(k-1)*n + i], ey[n + i]);
  for(int k = 2; k <= kzm; k++) for(int64_t i = 0; i < n; i++) ez[k*n + i] = cmul(ez[(k-1)*n + i], ez[n + i]);
  double GG = 0.0, HH = 0.0, HG = 0.0;
  int64_t iv = 0;
  for(int kx = 0; kx <= kxm; kx++)
  {
    double kvx[3]; for(int d = 0; d < 3; d++) kvx[d] = ax[d] * 2.0 * ORC_PI * (double) kx;
    double factor = (kx == 0) ? (1.0 * prefactor) : (2.0 * prefactor);
    for(int ky = -kym; ky <= kym; ky++)
    {
      double kvy[3]; for(int d = 0; d < 3; d++) kvy[d] = ay[d] * 2.0 * ORC_PI * (double) ky;
      for(int64_t i = 0; i < n; i++)
      {
        cplx t = ey[i + n * abs(ky)]; t.im = ky >= 0 ? t.im : -t.im;
        exy[i] = cmul(ex[i + n * kx], t);
      }
      for(int kz = -kzm; kz <= kzm; kz++)
      {
        double ksqr = ksq_of(B, kx, ky, kz);
        cplx ack = {0.0, 0.0}, fck = {0.0, 0.0};
        if((ksqr > 1e-10) && (ksqr < B->recip_cutoff))
        {
          double kvz[3]; for(int d = 0; d < 3; d++) kvz[d] = az[d] * 2.0 * ORC_PI * (double) kz;
          int64_t cnt = 0;
          for(int c = 0; c < S->ncomp; c++)
            for(int64_t p = 0; p < S->natoms[c]; p++, cnt++)
            {
              cplx t = ez[cnt + n * abs(kz)]; t.im = kz >= 0 ? t.im : -t.im;
              cplx v = cmul(exy[cnt], t);
              double w = S->scale_coul[gi[cnt]] * S->charge[gi[cnt]];
              ack.re += w * v.re; ack.im += w * v.im;
              if(c < S->nhost && has_fw) { fck.re += w * v.re; fck.im += w * v.im; }
            }
          double kv[3] = {kvx[0] + kvy[0] + kvz[0], kvx[1] + kvy[1] + kvz[1], kvx[2] + kvy[2] + kvz[2]};
          double rksq = kv[0]*kv[0] + kv[1]*kv[1] + kv[2]*kv[2];
          double temp = factor * exp((-0.25 / alpha_sq) * rksq) / rksq;
          if(has_fw) { ack.re -= fck.re; ack.im -= fck.im; }                  /* :149-150 */
          GG += temp * (ack.re * ack.re + ack.im * ack.im);
          HH += temp * (fck.re * fck.re + fck.im * fck.im);
          HG += temp * (fck.re * ack.re + fck.im * ack.im) * 2.0;
        }
        if(sf_fw)  { sf_fw[2*iv] = fck.re;  sf_fw[2*iv+1] = fck.im; }
        if(sf_ads) { sf_ads[2*iv] = ack.re; sf_ads[2*iv+1] = ack.im; }
        iv++;
      }
    }
  }
  GG += HH;                                                                    /* :174 */
  /* self, :177-191 */
  double pself = B->prefactor * alpha / sqrt(ORC_PI);
  {
    int64_t cnt = 0;
    for(int c = 0; c < S->ncomp; c++)
      for(int64_t p = 0; p < S->natoms[c]; p++, cnt++)
      {
        double q = S->charge[gi[cnt]], sc_ = S->scale_coul[gi[cnt]];
        GG -= pself * sc_ * q * sc_ * q;
        if(c < S->nhost && has_fw) HH -= pself * sc_ * q * sc_ * q;
      }
  }
  /* intra-molecular exclusion, :194-227 */
  for(int c = 0; c < S->ncomp; c++)
  {
    if(S->natoms[c] == 0 || S->molsize[c] == 0) continue;
    int64_t ms = S->molsize[c], nmol = S->natoms[c] / ms;
    for(int64_t m = 0; m < nmol; m++)
    {
      int64_t a0 = off[c] + m * ms;
      for(int64_t i = a0; i < a0 + ms - 1; i++)
      {
        double fa = S->scale_coul[i] * S->charge[i];
        for(int64_t j = i + 1; j < a0 + ms; j++)
        {
          double fb = S->scale_coul[j] * S->charge[j];
          double v[3] = {S->pos[3*i] - S->pos[3*j], S->pos[3*i+1] - S->pos[3*j+1], S->pos[3*i+2] - S->pos[3*j+2]};
          orc_pbc(v, B);
          double r = sqrt(v[0]*v[0] + v[1]*v[1] + v[2]*v[2]);
          double e = B->prefactor * fa * fb * erf(alpha * r) / r;
          GG -= e;
          if(c < S->nhost && has_fw) HH -= e;
        }
      }
    }
  }
  outE[0] = GG; outE[1] = HH; outE[2] = HG;
  free(ex); free(ey); free(ez); free(exy); free(gi);
}

/* ewald_preparation.h:261-298 */
void orc_exclusion_rigid(const orc_box* B, int ms, const double* pos, const double* charge, const double* scoul, double* intra, double* self)
{
  double E = 0.0;
  for(int i = 0; i + 1 < ms; i++)
  {
    double fa = scoul[i] * charge[i];
    for(int j = i + 1; j < ms; j++)
    {
      double fb = scoul[j] * charge[j];
      double v[3] = {pos[3*i] - pos[3*j], pos[3*i+1] - pos[3*j+1], pos[3*i+2] - pos[3*j+2]};
      orc_pbc(v, B);
      double r = sqrt(v[0]*v[0] + v[1]*v[1] + v[2]*v[2]);
      E += B->prefactor * fa * fb * erf(B->alpha * r) / r;
    }
  }
  *intra = ms > 0 ? E : 0.0;
  double S = 0.0, ps = B->prefactor * B->alpha / sqrt(ORC_PI);
  for(int i = 0; i < ms; i++) S += ps * scoul[i] * charge[i] * scoul[i] * charge[i];
  *self = S;
}

/* ------------------------------------------------------------------ tail */

/* TailCorrection_Energy_Functions.h:3-21 */
double orc_tail_total(int n, const int64_t* np, const int32_t* use, const double* e, double V)
{
  double T = 0.0;
  for(int i = 0; i < n; i++)
    for(int j = i; j < n; j++)
      if(use[i*n+j])
      {
        double pe = e[i*n+j] * (double)((size_t) np[i] * (size_t) np[j]); if(i != j) pe *= 2.0;
        T += pe;
      }
  return T / V;
}

/* TailCorrection_Energy_Functions.h:36-82; sign = +1 insertion/Widom, -1 deletion */
double orc_tail_difference(int n, const int64_t* np, const int32_t* use, const double* e, double V, const int32_t* cnt, int sign)
{
  double T = 0.0;
  for(int i = 0; i < n; i++)
  {
    int di = cnt[i] * sign;
    for(int j = i; j < n; j++)
    {
      int dj = cnt[j] * sign;
      if(use[i*n+j])
      {
        int Ni = (int) np[i], Nj = (int) np[j];
        int dN = Ni * dj + Nj * di + di * dj;
        double dE = e[i*n+j] * (double) dN; if(i != j) dE *= 2.0;
        T += dE;
      }
    }
  }
  return T / V;
}

/* TailCorrection_Energy_Functions.h:85-113 */
double orc_tail_identity_swap(int n, const int64_t* np, const int32_t* use, const double* e, double V, const int32_t* cn, const int32_t* co)
{
  double T = 0.0;
  for(int i = 0; i < n; i++)
  {
    int di = cn[i] - co[i];
    for(int j = i; j < n; j++)
    {
      int dj = cn[j] - co[j];
      if(use[i*n+j])
      {
        int Ni = (int) np[i], Nj = (int) np[j];
        int dN = Ni * dj + Nj * di + di * dj;
        double dE = e[i*n+j] * (double) dN; if(i != j) dE *= 2.0;
        T += dE;
      }
    }
  }
  return T / V;
}

/* ------------------------------------------------------------------ single body */

/* VDW_Coulomb.cu:626-841 + host sum mc_single_particle.h:183-200 */
void orc_single_body_delta(const orc_box* box, const orc_ff* ff, const orc_system* sys, int comp_id, int64_t molid,
                           const orc_atoms* O, const orc_atoms* N, int do_new, int do_old, double* out, int32_t* flag)
{
  int64_t off[65]; sys_offsets(sys, off);
  const int chainsize = (int)(do_new ? N->n : O->n);
  int64_t nhost = live_count(sys, 0, sys->nhost), nguest = live_count(sys, sys->nhost, sys->ncomp);
  const int moved_is_host = comp_id < sys->nhost;
  int64_t ncross = moved_is_host ? nguest : nhost;
  int64_t hh_b = moved_is_host ? (nhost * chainsize + ORC_THREADS - 1) / ORC_THREADS : 0;
  int64_t hg_b = (ncross * chainsize + ORC_THREADS - 1) / ORC_THREADS;
  int64_t gg_b = moved_is_host ? 0 : (nguest * chainsize + ORC_THREADS - 1) / ORC_THREADS;
  for(int k = 0; k < 6; k++) out[k] = 0.0;
  int fl = 0;
  double sv[ORC_THREADS], sr[ORC_THREADS];
  for(int seg = 0; seg < 3; seg++)
  {
    int64_t nb = seg == 0 ? hh_b : (seg == 1 ? hg_b : gg_b);
    int c0, c1;
    if(seg == 0) { c0 = 0; c1 = sys->nhost; }
    else if(seg == 2) { c0 = sys->nhost; c1 = sys->ncomp; }
    else if(moved_is_host) { c0 = sys->nhost; c1 = sys->ncomp; } else { c0 = 0; c1 = sys->nhost; }
    for(int64_t b = 0; b < nb; b++)
    {
      for(int k = 0; k < ORC_THREADS; k++)
      {
        sv[k] = 0.0; sr[k] = 0.0;
        int64_t ij = b * ORC_THREADS + k;
        int64_t i = ij / chainsize; int j = (int)(ij % chainsize);
        int comp = c0; int64_t posi = i;
        while(comp < c1 && posi >= sys->natoms[comp]) { posi -= sys->natoms[comp]; comp++; }
        if(comp >= c1) continue;
        int64_t g = off[comp] + posi;
        if(sys->molid[g] == molid && comp == comp_id) continue;
        if(do_new)
        {
          double ev = 0.0, ec = 0.0; int f = 0;
          pair_body(box, ff, sys->pos + 3*g, sys->scale[g], sys->charge[g], sys->scale_coul[g], sys->type[g],
                    N->pos + 3*j, N->scale[j], N->charge[j], N->scale_coul[j], N->type[j], &ev, &ec, &f, NULL);
          if(f) fl = 1;
          sv[k] += ev; sr[k] += ec;
        }
        if(do_old)
        {
          double ev = 0.0, ec = 0.0; int f = 0;
          pair_body(box, ff, sys->pos + 3*g, sys->scale[g], sys->charge[g], sys->scale_coul[g], sys->type[g],
                    O->pos + 3*j, O->scale[j], O->charge[j], O->scale_coul[j], O->type[j], &ev, &ec, &f, NULL);
          sv[k] -= ev; sr[k] -= ec;
        }
      }
      tree_reduce(sv, ORC_THREADS); tree_reduce(sr, ORC_THREADS);
      out[2*seg] += sv[0]; out[2*seg + 1] += sr[0];
    }
  }
  *flag = fl;
}

/* ------------------------------------------------------------------ totals */

/* VDW_Coulomb.cu:94-206 */
void orc_total_vdw_real(const orc_box* box, const orc_ff* ff, const orc_system* sys, double* out)
{
  int64_t off[65]; sys_offsets(sys, off);
  double tv[3] = {0, 0, 0}, tr[3] = {0, 0, 0};
  for(int ci = 0; ci < sys->ncomp; ci++)
    for(int64_t i = off[ci]; i < off[ci] + sys->natoms[ci]; i++)
      for(int cj = 0; cj < sys->ncomp; cj++)
      {
        int kind;
        if(ci < sys->nhost || cj < sys->nhost) kind = (!(ci < sys->nhost) || !(cj < sys->nhost)) ? 1 : 0; else kind = 2;
        for(int64_t j = off[cj]; j < off[cj] + sys->natoms[cj]; j++)
        {
          if((sys->molid[i] == sys->molid[j]) && (ci == cj)) continue;
          double v[3] = {sys->pos[3*i] - sys->pos[3*j], sys->pos[3*i+1] - sys->pos[3*j+1], sys->pos[3*i+2] - sys->pos[3*j+2]};
          orc_pbc(v, box);
          const double rr = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
          if(rr < ff->cutoff_vdw_sq)
          {
            double res[2];
            const int64_t row = sys->type[i] * ff->ntypes + sys->type[j];
            const double F[5] = {ff->epsilon[row], ff->sigma[row], ff->z[row], ff->shift[row], ff->c10[row]};
            orc_vdw(F, rr, sys->scale[i] * sys->scale[j], ff->use1264, res);
            tv[kind] += 0.5 * res[0];
          }
          if(!ff->no_charges && rr < ff->cutoff_coul_sq)
            tr[kind] += 0.5 * orc_coulomb_real(sys->charge[i], sys->charge[j], sqrt(rr), sys->scale_coul[i] * sys->scale_coul[j], box->prefactor, box->alpha);
        }
      }
  out[0] = tv[0]; out[1] = tr[0]; out[2] = tv[1]; out[3] = tr[1]; out[4] = tv[2]; out[5] = tr[2];
}

/* ------------------------------------------------------------------ Widom insertion */

/* Insertion_Body, mc_swap_utilities.h:3-133, with Widom_Move_FirstBead_PARTIAL (mc_widom.h:385-507),
 * Widom_Move_Chain_PARTIAL (:509-614) and GPU_EwaldDifference_General(INSERTION) inlined. */
void orc_widom_insertion(const orc_box* box, const orc_ff* ff, const orc_system* sys, const orc_widom_cfg* cfg,
                         const double* rnd_fb, double u_fb, const double* rnd_or, double u_or,
                         double* out, int32_t* out_stage, int32_t* out_sel, double* out_pos, int64_t* counts)
{
  enum { MAXT = 64, MAXA = 64 };
  const int comp = cfg->comp;
  const int ms = (int) sys->molsize[comp];
  const int NT = cfg->ntrials, NO = cfg->norient;
  int64_t off[65]; sys_offsets(sys, off);
  const int64_t new_molid = sys->natoms[comp] / (ms > 0 ? ms : 1);   /* NumberOfMolecule_for_Component, mc_widom.h:414 */
  for(int k = 0; k < 8; k++) out[k] = 0.0;
  out_sel[0] = out_sel[1] = 0; *out_stage = 0;

  double tpos[3*MAXT*MAXA], tsc[MAXT*MAXA], tq[MAXT*MAXA], tscc[MAXT*MAXA]; int64_t tty[MAXT*MAXA];
  double e[4*MAXT]; int32_t fl[MAXT];
  orc_atoms T = {tpos, tsc, tq, tscc, tty, 0};

  /* ---- first bead ---- */
  orc_trial_positions(box, sys, ORC_CBMC_INSERTION, comp, 0, NT, rnd_fb, 1.0, 1.0, tpos, tsc, tq, tscc, tty);
  T.n = NT;
  orc_trial_energies(box, ff, sys, NT, 1, &T, comp, new_molid, -1, -1, e, fl, counts);
  double rosen[MAXT]; int idx[MAXT]; int ns = 0;
  for(int t = 0; t < NT; t++) if(!fl[t])                                 /* mc_widom.h:47-86 */
  {
    double tot = e[4*t] + e[4*t+2];
    if(ff->vdw_real_bias) tot += e[4*t+1] + e[4*t+3];
    rosen[ns] = -cfg->beta * tot; idx[ns] = t; ns++;
  }
  int sel = 0; double W = 0.0;
  int ok = orc_cbmc_finish(ORC_CBMC_INSERTION, 0, rosen, ns, NT, u_fb, 0.0, NULL, &sel, &W);
  if(ok && !ff->vdw_real_bias) W *= exp(-cfg->beta * (e[4*idx[sel]+1] + e[4*idx[sel]+3]));
  if(W <= 1e-150) ok = 0;                                               /* mc_swap_utilities.h:21 */
  if(!ok) { *out_stage = 1; return; }
  const int fbt = idx[sel];
  out_sel[0] = fbt;
  double efb[4]; for(int k = 0; k < 4; k++) efb[k] = e[4*fbt + k];
  double fbpos[3] = {tpos[3*fbt], tpos[3*fbt+1], tpos[3*fbt+2]};
  double fbq = tq[fbt]; double fbs = tsc[fbt], fbsc = tscc[fbt];
  double ech[4] = {0, 0, 0, 0};
  double mpos[3*MAXA], mq[MAXA], mscc[MAXA];
  mpos[0] = fbpos[0]; mpos[1] = fbpos[1]; mpos[2] = fbpos[2]; mq[0] = fbq; mscc[0] = fbsc;

  /* ---- chain ---- */
  if(ms > 1)
  {
    const int cs = ms - 1;
    orc_trial_orientations(sys, ORC_CBMC_INSERTION, comp, 1, cs, NO, rnd_or, fbpos, fbs, fbsc, tpos, tsc, tq, tscc, tty);
    T.n = (int64_t) NO * cs;
    orc_trial_energies(box, ff, sys, NO, cs, &T, comp, new_molid, -1, -1, e, fl, counts);
    ns = 0;
    for(int t = 0; t < NO; t++) if(!fl[t])
    {
      double tot = e[4*t] + e[4*t+2];
      if(ff->vdw_real_bias) tot += e[4*t+1] + e[4*t+3];
      rosen[ns] = -cfg->beta * tot; idx[ns] = t; ns++;
    }
    double W2 = 0.0;
    ok = orc_cbmc_finish(ORC_CBMC_INSERTION, 1, rosen, ns, NO, u_or, 0.0, NULL, &sel, &W2);
    if(!ok) { *out_stage = 2; return; }                                   /* mc_widom.h:594-598 */
    if(!ff->vdw_real_bias) W2 *= exp(-cfg->beta * (e[4*idx[sel]+1] + e[4*idx[sel]+3]));
    W *= W2;
    if(W <= 1e-150) { *out_stage = 2; return; }                          /* mc_swap_utilities.h:32 */
    const int ot = idx[sel];
    out_sel[1] = ot;
    for(int k = 0; k < 4; k++) ech[k] = e[4*ot + k];
    for(int a = 0; a < cs; a++)
    {
      for(int d = 0; d < 3; d++) mpos[3*(a+1)+d] = tpos[3*(ot*cs + a)+d];
      mq[a+1] = tq[ot*cs + a]; mscc[a+1] = tscc[ot*cs + a];
    }
  }
  /* ---- Ewald, mc_swap_utilities.h:96-104 + Ewald_Energy_Functions.h:547-557 ---- */
  double ew[2] = {0.0, 0.0};
  if(!ff->no_charges && cfg->has_charge)
  {
    int64_t act = 0;
    orc_ewald_delta(box, mpos, mq, mscc, 0, ms, cfg->sf_ads, cfg->sf_fw, NULL, ew, &act);
    ew[0] -= (cfg->excl_intra + cfg->excl_self) * (pow(1.0, 2) - 0.0);
    if(counts) counts[4] += act * ms;
    W *= exp(-cfg->beta * (ew[0] + ew[1]));
  }
  /* ---- tail, mc_swap_utilities.h:105-108 ---- */
  double tail = 0.0;
  if(cfg->has_tail) tail = orc_tail_difference(cfg->ntypes, cfg->npseudo, cfg->use_tail, cfg->tail_e, box->volume, cfg->species_counts, +1);
  W *= exp(-cfg->beta * tail);
  out[0] = W;
  out[1] = efb[0] + ech[0]; out[2] = efb[1] + ech[1]; out[3] = efb[2] + ech[2]; out[4] = efb[3] + ech[3];
  out[5] = ew[0]; out[6] = ew[1]; out[7] = tail;
  if(out_pos) for(int k = 0; k < 3*ms; k++) out_pos[k] = mpos[k];
}

int orc_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_widom_batch(const orc_box* box, const orc_ff* ff, const orc_system* sys, const orc_widom_cfg* cfg,
                     int64_t n, const double* rnd, const double* uni, int nthreads,
                     double* out, int32_t* stage, int64_t* counts)
{
  const int per = cfg->ntrials + cfg->norient;
  int64_t tot[5] = {0, 0, 0, 0, 0};
#ifdef _OPENMP
  if(nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+:tot[:5])
#endif
  for(int64_t i = 0; i < n; i++)
  {
    int64_t c[5] = {0, 0, 0, 0, 0}; int32_t sel[2];
    const double* r = rnd + 3 * per * i;
    orc_widom_insertion(box, ff, sys, cfg, r, uni[2*i], r + 3 * cfg->ntrials, uni[2*i+1], out + 8*i, stage + i, sel, NULL, c);
    for(int k = 0; k < 5; k++) tot[k] += c[k];
  }
  if(counts) for(int k = 0; k < 5; k++) counts[k] = tot[k];
}

/* data_struct.cpp:6-11 with std::srand(RANDOMSEED), data_struct.h:1340 */
void orc_uniform_stream(int seed, int64_t n, double* out)
{
  srand((unsigned) seed);
  for(int64_t i = 0; i < n; i++) out[i] = (double) rand() / RAND_MAX;
}
