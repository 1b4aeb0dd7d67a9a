// graspa_b200 -- host side of the C ABI (include/graspa_b200.h): device state, uploads, launches.
// Product code.  There is no CPU fallback anywhere in this file: every energy comes from a kernel in misc_kernels.cuh / pair_kernels.cuh.
#include "../../include/graspa_b200.h"
#include "misc_kernels.cuh"
#include "pair_kernels.cuh"
#include "move_kernels.cuh"
#include "fused_kernel.cuh"
#include "move_server.cuh"
#include "widom_cells.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <map>
#include <vector>
#include <atomic>
#include <chrono>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges show up in nsys / ncu --nvtx, cost nothing without a tool attached

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CUDA_TRY(call)                                                                             \
  do {                                                                                             \
    cudaError_t err__ = (call);                                                                    \
    if(err__ != cudaSuccess)                                                                       \
      return fail(GB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));             \
  } while(0)

// NVTX range for the scope of one ABI call / stage (the reference has only omp_get_wtime stopwatches)
struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

template <typename T>
struct DevBuf
{
  T* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t n)
  {
    if(n <= cap) return cudaSuccess;
    if(p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 4 + 16;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if(e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if(p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Comp
{
  int offset = 0, alloc = 0, molsize = 0, natoms = 0;
  bool uploaded = false;
  double excl_intra = 0.0, excl_atom = 0.0; int rigid = 1, has_charge = 0;
  double* d_pocket = nullptr; int npocket = 0, pocket_invert = 0;      // block pockets: npocket x {x, y, z, radius} on the device
};

} // namespace

struct gb_engine
{
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop{};
  size_t smem_optin = 0;
  DevParams P{};
  bool have_ff = false, have_box = false;
  int ntypes = 0;
  DevBuf<double4> d_ffA; DevBuf<double> d_ffB; DevBuf<double> d_erfc, d_erfc10;
  std::vector<int> tail_use; std::vector<double> tail_e; bool has_tail = false;
  DevBuf<int> d_tail_use; DevBuf<double> d_tail_e;

  int ncomp = 0, nhost = 0;
  std::vector<Comp> comps;
  // tail-correction deltas already evaluated on the device, keyed by (count delta per type, N_pseudo per type): the delta is a
  // function of integer occupation numbers only, and a GCMC run revisits the same few hundred states over and over
  std::map<std::vector<long long>, double> tail_memo; std::vector<long long> tail_key;
  int nslots = 0;
  // host staging of the slot arrays (authoritative until the first device-side commit)
  std::vector<double> hx, hy, hz, hq, hscale, hscoul; std::vector<int> htype, hmolid;
  DevBuf<double> dx, dy, dz, dfx, dfy, dfz, dq, dscale, dscoul; DevBuf<int> dtype, dmolid;
  bool device_stale = true;
  std::vector<long long> npseudo;          // Components::NumberOfPseudoAtoms

  // framework pack
  DevBuf<double> d_pack; int pack_n = 0, pack_npad = 0, pack_ntp = 0; bool pack_dirty = true;

  // Ewald
  long long nvec = 0;
  std::vector<int> h_kpack, h_kslot; std::vector<double> h_ktemp;
  // what the INDEX part of the k table (active list, slots, row-ordered walk) was built for: without the LAMMPS-style set-up the active set is
  // decided by integer kx^2 + ky^2 + kz^2 against the reciprocal cutoff, i.e. by kmax alone -- a volume move that keeps kmax reuses it
  bool kindex_valid = false; int kindex_kmax[3] = {-1, -1, -1}; double kindex_rcut = 0.0;
  DevBuf<int> d_kpack, d_kslot; DevBuf<double> d_ktemp;
  int nact = 0, nact_pad = 0;
  DevBuf<double> d_sf[3];                 // ads, fw, temp (full arrays)
  int i_ads = 0, i_fw = 1, i_tmp = 2;     // pointer swap = index swap (Update_Vector_Ewald)
  bool have_sf = false;
  DevBuf<double> d_ktab; bool ktab_dirty = true;
  cudaStream_t copy_stream = nullptr; cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // host-input Widom batches: H2D of chunk c+1 under the kernels of chunk c
  // row-ordered k table of k_widom_ewald: a lane owns one (kx, ky) row and walks kz with its warp in lockstep
  std::vector<int> h_rowidx, h_rowmeta, h_round; DevBuf<int> d_rowidx, d_rowmeta, d_round; DevBuf<double> d_rtab; int nrounds = 0, npos = 0;
  // volume move (gb_volume_move_trial / _finish): the state to fall back to on rejection
  gb_box cur_box{}, vol_old_box{}; bool vol_pending = false, vol_had_sf = false; long long vol_old_nvec = 0;
  DevBuf<double> d_vol_xyz, d_vol_sf;

  // cell-sorted Widom pair stage (widom_cells.cuh)
  DevBuf<double> wc_udelta, wc_e4, wc_fbres, wc_afx, wc_afy, wc_afz, wc_aq; DevBuf<double4> wc_srec;
  DevBuf<int> wc_ucell, wc_flag, wc_count, wc_off, wc_cursor, wc_items, wc_ctl, wc_atk;
  // first-bead trial energies kept by gb_widom_first_bead_success for gb_widom_batch(resume_first_bead): indexed by pool row
  DevBuf<double> fbk_e4; DevBuf<int> fbk_flag; bool fbk_valid = false; long long fbk_row0 = 0, fbk_rows = 0, fbk_serial = -1, pool_gen = 0, fbk_pool_gen = -1; int fbk_comp = -1;
  long long call_serial = 0;             // entry points that went through ready() (Widom calls take themselves out again)
  bool wc_overflowed = false;            // a candidate list did not fit even with one CTA per SM: this engine keeps to k_widom_pair
  int wc_ctas_cap = 8, wc_last_ctas = 1; // CTAs per SM of the energy kernel: lowered (larger lists per CTA) when a list overflowed
  // CBMC
  int ntrials = 10, norient = 10; bool have_cbmc = false;
  DevBuf<double> d_pool; long long n_pool = 0;

  // scratch
  DevBuf<double> d_rec, d_out8, d_partial, d_sums, d_uni, d_scratch, d_result;
  DevBuf<int> d_stage, d_iscratch;
  DevBuf<long long> d_idx0, d_idx1;
  DevBuf<unsigned int> d_ticket;
  double* h_pinned = nullptr;            // 4 KB pinned result slot
  double* h_results = nullptr;           // 2 KB of h_pinned: tagged result records of k_move
  unsigned long long move_seq = 0;       // sequence number of the last k_move launch (published to h_pinned + 256)
  bool move_cooperative = true;          // cudaLaunchCooperativeKernel: co-residency of the move kernel's CTAs guaranteed by the driver
  // resident move server (move_server.cuh): one cooperative launch that executes move after move; the fused move calls post
  // commands to it, the accept calls queue their commits for the next command, every other call stops it first (ready())
  bool counted = false;
  double growth_scale[2] = {1.0, 1.0};   // scaling factors of the molecule the last CBMC insertion grew (a fractional molecule clears P.all_unit_scale when it is committed)
  bool srv_enabled = true, srv_running = false; int srv_compat = 0, srv_grid = 0; size_t srv_smem = 0;
  cudaStream_t srv_stream = nullptr;
  unsigned long long* srv_hcmd = nullptr;        // 4 KB pinned: [0, 2 KB) command records, [2 KB] status word
  DevBuf<unsigned long long> srv_dcmd, srv_done;
  DevParams srv_P{}; SysView srv_S{};            // what the running server was launched with
  unsigned long long srv_next = 0;               // sequence number the server expects next
  std::vector<CommitOp> pending_commits;         // accepted moves whose state change the next command carries
  long long srv_starts = 0, srv_commands = 0;

  // single-move path
  DevBuf<double> d_mv, d_ewpos; DevBuf<int> d_mvi;
  int last_fb_selected = 0, last_cbmc_comp = -1; long long last_cbmc_selected = 0;
  int sb_comp = -1, sb_type = -1; long long sb_molecule = 0;
  double lambda_scale[2] = {1.0, 1.0};            // new scaling factors of the last gb_lambda_change_delta
  bool committed = false;                // device slots changed by accept calls: the device is authoritative

  long long launches = 0;
  bool timing = false;
  double ms_pair = 0.0, ms_ewald = 0.0, ms_wc = 0.0; long long n_pair = 0, n_ewald = 0, n_wc = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

// Every copy and memset of the engine goes through ITS stream.  The stream is non-blocking, so the legacy default stream is not
// ordered against it: a plain cudaMemcpy from pageable memory returns once the data is staged, cudaMemset does not wait at all, and
// a kernel on the engine's stream could still see the old contents (the k tables of the previous box in a volume move).
cudaError_t copy_on_stream(gb_engine* e, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind)
{
  cudaError_t err = cudaMemcpyAsync(dst, src, bytes, kind, e->stream);
  if(err != cudaSuccess) return err;
  return cudaStreamSynchronize(e->stream);
}

SysView sys_view(gb_engine* e)
{
  SysView S; S.fx = e->dfx.p; S.fy = e->dfy.p; S.fz = e->dfz.p; S.q = e->dq.p; S.scale = e->dscale.p; S.scoul = e->dscoul.p;
  S.type = e->dtype.p; S.molid = e->dmolid.p;
  return S;
}

// live ranges; which = 0: all, 1: host only, 2: adsorbate only.  kind: host 1 (HG), adsorbate 2 (GG) -- trial-move convention.
SegList seg_list(gb_engine* e, int which)
{
  SegList L; memset(&L, 0, sizeof(L));
  for(int c = 0; c < e->ncomp && L.nseg < GBK_MAX_SEG; c++)
  {
    const bool host = c < e->nhost;
    if((which == 1 && !host) || (which == 2 && host)) continue;
    if(e->comps[c].natoms == 0) continue;
    L.start[L.nseg] = e->comps[c].offset; L.count[L.nseg] = e->comps[c].natoms; L.comp[L.nseg] = c;
    L.kind[L.nseg] = host ? 1 : 2; L.staged[L.nseg] = 0; L.nseg++;
  }
  return L;
}

int sync_slots_to_device(gb_engine* e)
{
  if(!e->device_stale) return GB_OK;
  const size_t n = (size_t) e->nslots;
  if(n == 0) return GB_OK;
  CUDA_TRY(e->dx.reserve(n)); CUDA_TRY(e->dy.reserve(n)); CUDA_TRY(e->dz.reserve(n));
  CUDA_TRY(e->dfx.reserve(n)); CUDA_TRY(e->dfy.reserve(n)); CUDA_TRY(e->dfz.reserve(n));
  CUDA_TRY(e->dq.reserve(n)); CUDA_TRY(e->dscale.reserve(n)); CUDA_TRY(e->dscoul.reserve(n));
  CUDA_TRY(e->dtype.reserve(n)); CUDA_TRY(e->dmolid.reserve(n));
  const size_t b = n * sizeof(double);
  CUDA_TRY(cudaMemcpyAsync(e->dx.p, e->hx.data(), b, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->dy.p, e->hy.data(), b, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->dz.p, e->hz.data(), b, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->dq.p, e->hq.data(), b, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->dscale.p, e->hscale.data(), b, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->dscoul.p, e->hscoul.data(), b, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->dtype.p, e->htype.data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->dmolid.p, e->hmolid.data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  if(e->have_box)
  {
    k_frac_update<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->P, e->dx.p, e->dy.p, e->dz.p, e->dfx.p, e->dfy.p, e->dfz.p, 0, (int) n);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->device_stale = false; e->pack_dirty = true;
  bool unit = true;
  for(int c = 0; c < e->ncomp; c++)
    for(int i = 0; i < e->comps[c].alloc; i++)
    {
      const size_t g = (size_t) e->comps[c].offset + i;
      if(e->hscale[g] != 1.0 || e->hscoul[g] != 1.0) unit = false;
    }
  e->P.all_unit_scale = unit ? 1 : 0;
  return GB_OK;
}

// resident move server (move_server.cuh; host side in fused_moves.inc).  Calls that may run while the server is resident -- they post
// commands, queue commits, or launch only tiny kernels / copies on the engine's stream -- hold a SrvCompat; every other entry point
// stops the server (and brings the device slots up to date) in ready().
struct SrvCompat { gb_engine* e; explicit SrvCompat(gb_engine* e_) : e(e_) { if(e) e->srv_compat++; } ~SrvCompat() { if(e) e->srv_compat--; } };
int server_stop(gb_engine* e);
int commit_op(gb_engine* e, const CommitOp& c);
std::atomic<int> g_engines_alive{0};

int ready(gb_engine* e)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  // a server-compatible call while the move server is resident: everything below was checked when the server was started, and the call
  // touches the device through the mailbox only (the few that launch something set the device themselves)
  if(e->srv_running && e->srv_compat > 0) { e->call_serial++; return GB_OK; }
  if(!e->have_ff) return fail(GB_ERR_STATE, "gb_upload_forcefield has not been called");
  if(!e->have_box) return fail(GB_ERR_STATE, "gb_upload_box has not been called");
  if(e->ncomp == 0) return fail(GB_ERR_STATE, "gb_set_components has not been called");
  for(int c = 0; c < e->ncomp; c++) if(!e->comps[c].uploaded) return fail(GB_ERR_STATE, "component " + std::to_string(c) + " has not been uploaded");
  CUDA_TRY(cudaSetDevice(e->device));
  e->call_serial++;
  if((e->srv_running || !e->pending_commits.empty()) && e->srv_compat == 0) { int rc = server_stop(e); if(rc) return rc; }
  e->P.erfc_table_ok = (e->P.alpha * std::sqrt(e->P.cut_coul2) < GBK_ERFC_XMAX) ? 1 : 0;
  e->P.erfc10_ok = (e->P.alpha * std::sqrt(e->P.cut_coul2) < GBK_ERFC10_XMAX) ? 1 : 0;
  {
    // reach of the cutoff sphere along each fractional axis: |s_i| = |r . inv[:,i]| <= |r| |inv[:,i]|  (tile culling)
    const double rc = std::sqrt(e->P.no_charges ? e->P.cut_vdw2 : std::max(e->P.cut_vdw2, e->P.cut_coul2));
    e->P.cull_rcut = rc;
    for(int i = 0; i < 3; i++)
    {
      const double nrm = std::sqrt(e->P.inv[i] * e->P.inv[i] + e->P.inv[3 + i] * e->P.inv[3 + i] + e->P.inv[6 + i] * e->P.inv[6 + i]);
      e->P.cull_w[i] = rc * nrm * (1.0 + 1e-12) + 1e-12;
    }
  }
  return sync_slots_to_device(e);
}

// bytes of the staged pack: 5 arrays of npad (the type array is int, stored in a double-wide slot) + 10 tile arrays of ntp
static size_t pack_bytes_for(int npad, int ntp) { return gbk_pack_bytes(npad, ntp); }

// framework pack for TMA staging (tile-sorted, see pair.cuh).  Only valid for moves whose exclusions do not touch host components.
int ensure_pack(gb_engine* e, bool& usable, size_t extra_smem)
{
  usable = false;
  SegList L = seg_list(e, 1);
  int n = 0; for(int s = 0; s < L.nseg; s++) n += L.count[s];
  if(n == 0 || !e->P.all_unit_scale) return GB_OK;
  const int npad = (n + 31) / 32 * 32, ntiles = npad / 32, ntp = (ntiles + 31) / 32 * 32;
  if(pack_bytes_for(npad, ntp) + 64 + extra_smem > e->smem_optin) return GB_OK;
  if(e->pack_dirty || e->pack_n != n)
  {
    // ---- fetch the host components' live atoms (fractional coordinates, charge, type)
    std::vector<double> f[3], q(n), sc(n); std::vector<int> ty(n);
    for(int k = 0; k < 3; k++) f[k].resize(n);
    int acc = 0;
    for(int s = 0; s < L.nseg; s++)
    {
      const size_t b8 = (size_t) L.count[s] * sizeof(double);
      CUDA_TRY(cudaMemcpyAsync(f[0].data() + acc, e->dfx.p + L.start[s], b8, cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(cudaMemcpyAsync(f[1].data() + acc, e->dfy.p + L.start[s], b8, cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(cudaMemcpyAsync(f[2].data() + acc, e->dfz.p + L.start[s], b8, cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(cudaMemcpyAsync(q.data() + acc, e->dq.p + L.start[s], b8, cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(cudaMemcpyAsync(sc.data() + acc, e->dscoul.p + L.start[s], b8, cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(cudaMemcpyAsync(ty.data() + acc, e->dtype.p + L.start[s], (size_t) L.count[s] * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
      acc += L.count[s];
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for(int k = 0; k < 3; k++) for(int i = 0; i < n; i++) { f[k][i] -= std::floor(f[k][i]); if(f[k][i] >= 1.0) f[k][i] = 0.0; }
    // ---- k-d split into tiles of 32 atoms: full tiles go left, so only the last tile can be partial
    const double* H = e->P.cell;
    const double len[3] = { std::sqrt(H[0] * H[0] + H[1] * H[1] + H[2] * H[2]), std::sqrt(H[3] * H[3] + H[4] * H[4] + H[5] * H[5]),
                            std::sqrt(H[6] * H[6] + H[7] * H[7] + H[8] * H[8]) };
    std::vector<int> idx(n);
    for(int i = 0; i < n; i++) idx[i] = i;
    std::vector<std::pair<int, int>> todo; todo.push_back({0, n});
    while(!todo.empty())
    {
      const int lo = todo.back().first, hi = todo.back().second; todo.pop_back();
      const int m = (hi - lo + 31) / 32;
      if(m <= 1) continue;
      const int nl = (m / 2) * 32;
      int axis = 0; double best = -1.0;
      for(int k = 0; k < 3; k++)
      {
        double mn = 1e300, mx = -1e300;
        for(int i = lo; i < hi; i++) { mn = std::min(mn, f[k][idx[i]]); mx = std::max(mx, f[k][idx[i]]); }
        if((mx - mn) * len[k] > best) { best = (mx - mn) * len[k]; axis = k; }
      }
      const std::vector<double>& fa = f[axis];
      std::nth_element(idx.begin() + lo, idx.begin() + lo + nl, idx.begin() + hi, [&](int a, int b) { return fa[a] < fa[b] || (fa[a] == fa[b] && a < b); });
      todo.push_back({lo, lo + nl}); todo.push_back({lo + nl, hi});
    }
    // ---- pack [x | y | z | q*scoul | type] + tile table [lo xyz | hi xyz | centre xyz | radius]
    std::vector<double> pk(gbk_pack_bytes(npad, ntp) / sizeof(double), 0.0);
    double* X = pk.data(); double* Y = X + npad; double* Z = Y + npad; double* Q = Z + npad; int* T = reinterpret_cast<int*>(Q + npad);
    double* tile = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(pk.data()) + (size_t) npad * 36);
    for(int i = 0; i < npad; i++)
    {
      if(i < n)
      {
        const int j = idx[i];
        X[i] = H[0] * f[0][j] + H[3] * f[1][j] + H[6] * f[2][j];
        Y[i] = H[1] * f[0][j] + H[4] * f[1][j] + H[7] * f[2][j];
        Z[i] = H[2] * f[0][j] + H[5] * f[1][j] + H[8] * f[2][j];
        Q[i] = q[j] * sc[j]; T[i] = ty[j];
      }
      else { X[i] = Y[i] = Z[i] = 1e30; Q[i] = 0.0; T[i] = 0; }
    }
    for(int t = 0; t < ntp; t++)
    {
      double lo3[3] = {1e30, 1e30, 1e30}, hi3[3] = {1e30, 1e30, 1e30}, c3[3] = {0, 0, 0}, rad = 0.0;
      if(t < ntiles)
      {
        const int i0 = t * 32, i1 = std::min(n, i0 + 32);
        for(int k = 0; k < 3; k++) { lo3[k] = 1e300; hi3[k] = -1e300; }
        for(int i = i0; i < i1; i++)
        {
          const int j = idx[i];
          for(int k = 0; k < 3; k++) { lo3[k] = std::min(lo3[k], f[k][j]); hi3[k] = std::max(hi3[k], f[k][j]); }
          c3[0] += X[i]; c3[1] += Y[i]; c3[2] += Z[i];
        }
        for(int k = 0; k < 3; k++) c3[k] /= (double) (i1 - i0);
        for(int i = i0; i < i1; i++)
          rad = std::max(rad, std::sqrt((X[i] - c3[0]) * (X[i] - c3[0]) + (Y[i] - c3[1]) * (Y[i] - c3[1]) + (Z[i] - c3[2]) * (Z[i] - c3[2])));
        rad = rad * (1.0 + 1e-12) + 1e-9;
        // the fractional coordinates the kernel reconstructs differ from f by rounding: widen the box by a few ulps
        for(int k = 0; k < 3; k++) { lo3[k] -= 1e-12; hi3[k] += 1e-12; }
      }
      for(int k = 0; k < 3; k++) { tile[(size_t) k * ntp + t] = lo3[k]; tile[(size_t)(3 + k) * ntp + t] = hi3[k]; tile[(size_t)(6 + k) * ntp + t] = c3[k]; }
      tile[(size_t) 9 * ntp + t] = rad;
    }
    CUDA_TRY(e->d_pack.reserve(pk.size()));
    CUDA_TRY(cudaMemcpyAsync(e->d_pack.p, pk.data(), pk.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));          // pk is a local
    e->pack_n = n; e->pack_npad = npad; e->pack_ntp = ntp; e->pack_dirty = false;
  }
  usable = true;
  return GB_OK;
}

// stored structure factors gathered at the active k, in the staged layout of k_widom_ewald
__global__ void k_build_ktab(const double* __restrict__ temp, const int* __restrict__ kpack, const int* __restrict__ slot,
                             const double* __restrict__ sa, const double* __restrict__ sf, int nact, int npad, double* ktab)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= npad) return;
  int* kp = reinterpret_cast<int*>(ktab + 5 * (size_t) npad);
  if(k < nact)
  {
    const int s = slot[k];
    ktab[k] = temp[k];
    ktab[npad + 2 * k] = sa[2 * s]; ktab[npad + 2 * k + 1] = sa[2 * s + 1];
    ktab[3 * (size_t) npad + 2 * k] = sf[2 * s]; ktab[3 * (size_t) npad + 2 * k + 1] = sf[2 * s + 1];
    kp[k] = kpack[k];
  }
  else
  {
    ktab[k] = 0.0; ktab[npad + 2 * k] = 0.0; ktab[npad + 2 * k + 1] = 0.0;
    ktab[3 * (size_t) npad + 2 * k] = 0.0; ktab[3 * (size_t) npad + 2 * k + 1] = 0.0; kp[k] = (128 << 8) | 128;
  }
}

__global__ void k_add_sums(const double* __restrict__ src, double* dst, int n)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k < n) dst[k] += src[k];
}

// the same gathered in row order: 5 arrays of npos [temp | sa.re | sa.im | sf.re | sf.im]; unused positions carry temp = 0
__global__ void k_build_rtab(const double* __restrict__ temp, const int* __restrict__ slot, const int* __restrict__ rowidx,
                             const double* __restrict__ sa, const double* __restrict__ sf, int npos, double* rtab)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p >= npos) return;
  const int k = rowidx[p];
  double t = 0.0, ar = 0.0, ai = 0.0, fr = 0.0, fi = 0.0;
  if(k >= 0) { const int s = slot[k]; t = temp[k]; ar = sa[2 * s]; ai = sa[2 * s + 1]; fr = sf[2 * s]; fi = sf[2 * s + 1]; }
  rtab[p] = t; rtab[(size_t) npos + p] = ar; rtab[2 * (size_t) npos + p] = ai; rtab[3 * (size_t) npos + p] = fr; rtab[4 * (size_t) npos + p] = fi;
}

int ensure_ktab(gb_engine* e)
{
  if(!e->ktab_dirty) return GB_OK;
  if(e->nact == 0) { e->ktab_dirty = false; return GB_OK; }
  CUDA_TRY(e->d_ktab.reserve((size_t) e->nact_pad * 6));
  k_build_ktab<<<(e->nact_pad + 255) / 256, 256, 0, e->stream>>>(e->d_ktemp.p, e->d_kpack.p, e->d_kslot.p, e->d_sf[e->i_ads].p, e->d_sf[e->i_fw].p,
                                                                   e->nact, e->nact_pad, e->d_ktab.p);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  if(e->npos > 0)
  {
    CUDA_TRY(e->d_rtab.reserve((size_t) e->npos * 5));
    k_build_rtab<<<(e->npos + 255) / 256, 256, 0, e->stream>>>(e->d_ktemp.p, e->d_kslot.p, e->d_rowidx.p, e->d_sf[e->i_ads].p, e->d_sf[e->i_fw].p, e->npos, e->d_rtab.p);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  e->ktab_dirty = false;
  return GB_OK;
}

// tail corrections on the device (TailCorrection_Energy_Functions.h:3-113): tiny, one thread
__global__ void k_tail(int n, const long long* __restrict__ np, const int* __restrict__ use, const double* __restrict__ te,
                       const int* __restrict__ dcount /* null: total */, double volume, double* out)
{
  if(threadIdx.x != 0 || blockIdx.x != 0) return;
  double T = 0.0;
  for(int i = 0; i < n; i++)
    for(int j = i; j < n; j++)
      if(use[i * n + j])
      {
        double v;
        if(dcount)
        {
          const int Ni = (int) np[i], Nj = (int) np[j], di = dcount[i], dj = dcount[j];
          const int dN = Ni * dj + Nj * di + di * dj;
          v = te[i * n + j] * (double) dN;
        }
        else v = te[i * n + j] * (double)((unsigned long long) np[i] * (unsigned long long) np[j]);
        if(i != j) v *= 2.0;
        T += v;
      }
  *out = T / volume;
}

int tail_device(gb_engine* e, const std::vector<int>* dcount, double* d_out)
{
  const int n = e->ntypes;
  CUDA_TRY(cudaSetDevice(e->device));
  if(!e->has_tail) { CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(double), e->stream)); return GB_OK; }
  CUDA_TRY(e->d_iscratch.reserve((size_t) n + 16));
  CUDA_TRY(e->d_idx0.reserve((size_t) n + 16));
  CUDA_TRY(cudaMemcpyAsync(e->d_idx0.p, e->npseudo.data(), n * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
  if(dcount) CUDA_TRY(cudaMemcpyAsync(e->d_iscratch.p, dcount->data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  k_tail<<<1, 32, 0, e->stream>>>(n, e->d_idx0.p, e->d_tail_use.p, e->d_tail_e.p, dcount ? e->d_iscratch.p : nullptr, e->P.volume, d_out);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return GB_OK;
}

// tail delta for a change of `d` pseudo-atoms per type at the current occupation, evaluated by k_tail once per distinct state
int tail_delta_memo(gb_engine* e, const std::vector<int>& d, double* out)
{
  // the key is built in a buffer the engine keeps (a hit allocates nothing): moves of a GCMC run ask for the same few hundred states
  std::vector<long long>& key = e->tail_key;
  key.resize(2 * (size_t) e->ntypes);
  for(int i = 0; i < e->ntypes; i++) { key[i] = d[i]; key[e->ntypes + i] = e->npseudo[i]; }
  auto it = e->tail_memo.find(key);
  if(it != e->tail_memo.end()) { *out = it->second; return GB_OK; }
  int rc = tail_device(e, &d, e->d_result.p + 8); if(rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->h_pinned + 8, e->d_result.p + 8, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  *out = e->h_pinned[8];
  if(e->tail_memo.size() > 200000) e->tail_memo.clear();
  e->tail_memo.emplace(key, *out);
  return GB_OK;
}

std::vector<int> species_counts(gb_engine* e, int comp)
{
  std::vector<int> c(e->ntypes, 0);
  const Comp& C = e->comps[comp];
  for(int a = 0; a < C.molsize; a++) { const int t = e->htype[(size_t) C.offset + a]; if(t >= 0 && t < e->ntypes) c[t]++; }
  return c;
}

// CUDA-event timing of a group of launches on the engine's stream (only when gb_timing_enable is on).  Families: 0 the pair stage of
// the batched Widom path and the stage-call pair kernels, 1 the Fourier kernels, 3 the k_wc_energy launches alone (nested inside 0).
// Each timer owns its events, so timers may nest.
struct Timer
{
  gb_engine* e; int fam; bool on; cudaEvent_t a = nullptr, b = nullptr;
  Timer(gb_engine* e_, int fam_) : e(e_), fam(fam_), on(e_->timing)
  {
    if(on) { on = cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess; if(on) cudaEventRecord(a, e->stream); }
  }
  void stop(long long launches)
  {
    if(!on) return;
    cudaEventRecord(b, e->stream); cudaEventSynchronize(b);
    float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
    if(fam == 0) { e->ms_pair += ms; e->n_pair += launches; } else if(fam == 1) { e->ms_ewald += ms; e->n_ewald += launches; } else { e->ms_wc += ms; e->n_wc += launches; }
    on = false;
  }
  ~Timer() { if(a) cudaEventDestroy(a); if(b) cudaEventDestroy(b); }
};

} // namespace

extern "C" {

static int engine_init(gb_engine* e, int device);
int gb_abi_version(void) { return GB_ABI_VERSION; }
const char* gb_last_error(void) { return g_err.c_str(); }

int gb_engine_create(gb_engine** out, int device)
{
  if(!out) return fail(GB_ERR_ARG, "out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if(err != cudaSuccess || ndev == 0)
    return fail(GB_ERR_CUDA, std::string("no CUDA device: graspa_b200 has no CPU path (") + cudaGetErrorString(err) + ")");
  if(device < 0) CUDA_TRY(cudaGetDevice(&device));
  if(device >= ndev) return fail(GB_ERR_ARG, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  gb_engine* e = new gb_engine();
  e->device = device;
  const int rc_init = engine_init(e, device);
  if(rc_init != GB_OK) { const std::string msg = g_err; gb_engine_destroy(e); return fail(rc_init, msg); }     // nothing of a half-built engine is leaked
  e->counted = true; g_engines_alive++;       // the resident move server is used only while ONE engine lives in the process (header: one engine per GPU)
  *out = e;
  return GB_OK;
}

static int engine_init(gb_engine* e, int device)
{
  CUDA_TRY(cudaGetDeviceProperties(&e->prop, device));
  int optin = 0; CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  e->smem_optin = (size_t) optin;
  CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&e->ev0)); CUDA_TRY(cudaEventCreate(&e->ev1));
  CUDA_TRY(cudaMallocHost(&e->h_pinned, 8192)); memset(e->h_pinned, 0, 8192); e->h_results = e->h_pinned + 512;
  CUDA_TRY(e->d_ticket.reserve(16)); CUDA_TRY(cudaMemsetAsync(e->d_ticket.p, 0, 16 * sizeof(unsigned int), e->stream));
  CUDA_TRY(e->d_result.reserve(512));
  {
    // dynamic + static shared memory must stay within the opt-in limit
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_widom_pair<0>));
    CUDA_TRY(cudaFuncSetAttribute(k_widom_pair<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int) fa.sharedSizeBytes));
    CUDA_TRY(cudaFuncSetAttribute(k_widom_pair<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int) fa.sharedSizeBytes));
    CUDA_TRY(cudaFuncSetAttribute(k_widom_pair<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int) fa.sharedSizeBytes));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_widom_ewald));
    CUDA_TRY(cudaFuncSetAttribute(k_widom_ewald, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int) fa.sharedSizeBytes));
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
#define GBK_WC_ATTR(F) \
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy_lt<0, false, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024)); \
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy_lt<1, false, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024)); \
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy_lt<2, false, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024)); \
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy_lt<0, true, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024)); \
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy_lt<1, true, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024)); \
    CUDA_TRY(cudaFuncSetAttribute(k_wc_energy_lt<2, true, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
    GBK_WC_ATTR(0) GBK_WC_ATTR(1) GBK_WC_ATTR(2) GBK_WC_ATTR(3) GBK_WC_ATTR(4) GBK_WC_ATTR(5)
#undef GBK_WC_ATTR
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_move));
    CUDA_TRY(cudaFuncSetAttribute(k_move, cudaFuncAttributeMaxDynamicSharedMemorySize, GBF_MAX_DYN_SMEM));
    {
      int coop = 0; CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
      e->move_cooperative = coop != 0 && !(std::getenv("GB_MOVE_COOP") && std::atoi(std::getenv("GB_MOVE_COOP")) == 0);     // GB_MOVE_COOP=0: A/B timing only (1 % on the GCMC decks)
    }
    {
      // resident move server: same scratch capacity as the largest k_move launch; needs a cooperative launch (co-resident CTAs)
      e->srv_smem = std::min((size_t) optin - 1024, srv_smem_fixed() + (size_t) GBF_MAX_DYN_SMEM - ((sizeof(FusedSmem) + 127) / 128) * 128);
      CUDA_TRY(cudaFuncSetAttribute(k_move_server, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) e->srv_smem));
      e->srv_grid = std::min(e->prop.multiProcessorCount, 256);
      e->srv_enabled = e->move_cooperative && !(std::getenv("GB_MOVE_SERVER") && std::atoi(std::getenv("GB_MOVE_SERVER")) == 0);
      CUDA_TRY(cudaStreamCreateWithFlags(&e->srv_stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaMallocHost(&e->srv_hcmd, 4096)); memset(e->srv_hcmd, 0, 4096);
      CUDA_TRY(e->srv_dcmd.reserve(2 * (GBS_NREC + GBS_CREC))); CUDA_TRY(e->srv_done.reserve(256));
    }
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_ewald_delta));
    CUDA_TRY(cudaFuncSetAttribute(k_ewald_delta, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int) fa.sharedSizeBytes));
    e->smem_optin -= 1024;   // head room for static shared memory of the kernels
  }
  CUDA_TRY(e->d_erfc.reserve((GBK_ERFC_DEG + 1) * GBK_ERFC_NINT));
  CUDA_TRY(copy_on_stream(e, e->d_erfc.p, h_erfc_table, sizeof(h_erfc_table), cudaMemcpyHostToDevice));
  e->P.erfc_tab = e->d_erfc.p;
  CUDA_TRY(e->d_erfc10.reserve((GBK_ERFC10_DEG + 1) * GBK_ERFC10_NINT));
  CUDA_TRY(copy_on_stream(e, e->d_erfc10.p, h_erfc_table10, sizeof(h_erfc_table10), cudaMemcpyHostToDevice));
  e->P.erfc_tab10 = e->d_erfc10.p;
  return GB_OK;
}

int gb_engine_destroy(gb_engine* e)
{
  if(!e) return GB_OK;
  cudaSetDevice(e->device);
  if(e->srv_running || !e->pending_commits.empty()) server_stop(e);
  if(e->counted) { g_engines_alive--; e->counted = false; }
  if(e->srv_stream) cudaStreamDestroy(e->srv_stream);
  if(e->srv_hcmd) cudaFreeHost(e->srv_hcmd);
  e->srv_dcmd.release(); e->srv_done.release();
  if(e->stream) cudaStreamSynchronize(e->stream);
  e->d_ffA.release(); e->d_ffB.release(); e->d_erfc.release(); e->d_erfc10.release(); e->d_tail_use.release(); e->d_tail_e.release();
  e->dx.release(); e->dy.release(); e->dz.release(); e->dfx.release(); e->dfy.release(); e->dfz.release();
  e->dq.release(); e->dscale.release(); e->dscoul.release(); e->dtype.release(); e->dmolid.release();
  e->d_pack.release(); e->d_kpack.release(); e->d_kslot.release(); e->d_ktemp.release();
  for(int i = 0; i < 3; i++) e->d_sf[i].release();
  if(e->copy_stream) { cudaStreamDestroy(e->copy_stream); for(int c = 0; c < 8; c++) if(e->ev_chunk[c]) cudaEventDestroy(e->ev_chunk[c]); }
  e->d_ktab.release(); e->d_pool.release(); e->d_rec.release(); e->d_out8.release(); e->d_partial.release(); e->d_sums.release();
  e->d_uni.release(); e->d_scratch.release(); e->d_result.release(); e->d_stage.release(); e->d_iscratch.release();
  e->d_idx0.release(); e->d_idx1.release(); e->d_ticket.release(); e->d_mv.release(); e->d_mvi.release(); e->d_ewpos.release();
  e->fbk_e4.release(); e->fbk_flag.release();
  e->wc_udelta.release(); e->wc_e4.release(); e->wc_fbres.release(); e->wc_afx.release(); e->wc_afy.release(); e->wc_afz.release(); e->wc_aq.release(); e->wc_srec.release();
  e->wc_ucell.release(); e->wc_flag.release(); e->wc_count.release(); e->wc_off.release(); e->wc_cursor.release(); e->wc_items.release(); e->wc_ctl.release(); e->wc_atk.release();
  e->d_vol_xyz.release(); e->d_vol_sf.release(); e->d_rowidx.release(); e->d_rowmeta.release(); e->d_round.release(); e->d_rtab.release();
  for(auto& C : e->comps) if(C.d_pocket) cudaFree(C.d_pocket);
  if(e->h_pinned) cudaFreeHost(e->h_pinned);
  if(e->ev0) cudaEventDestroy(e->ev0);
  if(e->ev1) cudaEventDestroy(e->ev1);
  if(e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return GB_OK;
}

int gb_device_info(gb_engine* e, int* sm_count, int* cc_major, int* cc_minor, int64_t* smem)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  if(sm_count) *sm_count = e->prop.multiProcessorCount;
  if(cc_major) *cc_major = e->prop.major;
  if(cc_minor) *cc_minor = e->prop.minor;
  if(smem) *smem = (int64_t) e->smem_optin;
  return GB_OK;
}

int gb_synchronize(gb_engine* e) { if(!e) return fail(GB_ERR_ARG, "null engine"); CUDA_TRY(cudaStreamSynchronize(e->stream)); return GB_OK; }
void* gb_stream(gb_engine* e) { return e ? (void*) e->stream : nullptr; }

int gb_upload_forcefield(gb_engine* e, const gb_forcefield* ff, const gb_tail_table* tail)
{
  if(!e || !ff) return fail(GB_ERR_ARG, "null argument");
  if(ff->size <= 0 || ff->size > 1024) return fail(GB_ERR_ARG, "bad force-field size");
  CUDA_TRY(cudaSetDevice(e->device));
  const int n = ff->size; const size_t n2 = (size_t) n * n;
  std::vector<double4> A(n2); std::vector<double> B(n2);
  for(size_t i = 0; i < n2; i++)
  {
    if(!ff->use1264) { const double s = ff->sigma[i]; A[i] = make_double4(4.0 * ff->epsilon[i], s * s, ff->shift[i], 1.0 / (s * s)); B[i] = 0.0; }
    else { A[i] = make_double4(ff->epsilon[i], ff->sigma[i], ff->z[i], ff->shift[i]); B[i] = ff->c10 ? ff->c10[i] : 0.0; }
  }
  CUDA_TRY(e->d_ffA.reserve(n2)); CUDA_TRY(e->d_ffB.reserve(n2));
  CUDA_TRY(copy_on_stream(e, e->d_ffA.p, A.data(), n2 * sizeof(double4), cudaMemcpyHostToDevice));
  CUDA_TRY(copy_on_stream(e, e->d_ffB.p, B.data(), n2 * sizeof(double), cudaMemcpyHostToDevice));
  e->ntypes = n;
  e->P.ntypes = n; e->P.cut_vdw2 = ff->cutoff_vdw_sq; e->P.cut_coul2 = ff->cutoff_coul_sq; e->P.overlap = ff->overlap_criteria;
  e->P.no_charges = ff->no_charges; e->P.vdw_real_bias = ff->vdw_real_bias; e->P.use1264 = ff->use1264;
  e->P.ffA = e->d_ffA.p; e->P.ffB = e->d_ffB.p;
  e->tail_use.assign(n2, 0); e->tail_e.assign(n2, 0.0); e->has_tail = false; e->tail_memo.clear();
  if(tail && tail->use_tail && tail->energy)
  {
    if(tail->size != n) return fail(GB_ERR_ARG, "tail table size differs from force-field size");
    for(size_t i = 0; i < n2; i++) { e->tail_use[i] = tail->use_tail[i]; e->tail_e[i] = tail->energy[i]; if(tail->use_tail[i]) e->has_tail = true; }
  }
  CUDA_TRY(e->d_tail_use.reserve(n2)); CUDA_TRY(e->d_tail_e.reserve(n2));
  CUDA_TRY(copy_on_stream(e, e->d_tail_use.p, e->tail_use.data(), n2 * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(copy_on_stream(e, e->d_tail_e.p, e->tail_e.data(), n2 * sizeof(double), cudaMemcpyHostToDevice));
  e->npseudo.assign(n, 0);
  e->have_ff = true;
  return GB_OK;
}

namespace {
// Box scalars, active k table and structure-factor storage for `box`.  Does not touch the atom slots.
int apply_box(gb_engine* e, const gb_box* box)
{
  if(box->kmax[0] > 127 || box->kmax[1] > 127 || box->kmax[2] > 127 || box->kmax[0] < 0 || box->kmax[1] < 0 || box->kmax[2] < 0)
    return fail(GB_ERR_ARG, "kmax out of range [0,127]");      // before any engine state changes
  e->tail_memo.clear();                                           // the tail deltas carry 1 / volume
  CUDA_TRY(cudaSetDevice(e->device));
  e->cur_box = *box;
  for(int i = 0; i < 9; i++) { e->P.cell[i] = box->cell[i]; e->P.inv[i] = box->inverse_cell[i]; }
  e->P.volume = box->volume; e->P.alpha = box->alpha; e->P.prefactor = box->prefactor; e->P.recip_cutoff = box->reciprocal_cutoff;
  e->P.cubic = box->cubic; e->P.use_lammps = box->use_lammps_ewald;
  {
    const double* c = box->cell;
    const bool upper_zero = c[1] == 0.0 && c[2] == 0.0 && c[5] == 0.0;
    const bool lower_zero = c[3] == 0.0 && c[6] == 0.0 && c[7] == 0.0;
    e->P.cell_mode = (upper_zero && lower_zero) ? 2 : (upper_zero ? 1 : 0);
  }
  for(int i = 0; i < 3; i++) e->P.kmax[i] = box->kmax[i];
  // active k table: Ewald_Energy_Functions.h:299-334, 358-360
  const int kxm = box->kmax[0], kym = box->kmax[1], kzm = box->kmax[2];
  e->nvec = (long long)(kxm + 1) * (2 * kym + 1) * (2 * kzm + 1);
  const bool reuse_index = e->kindex_valid && !box->use_lammps_ewald && box->alpha > 0.0 && e->kindex_kmax[0] == kxm && e->kindex_kmax[1] == kym &&
                           e->kindex_kmax[2] == kzm && e->kindex_rcut == box->reciprocal_cutoff && !std::getenv("GB_NO_KINDEX_CACHE");
  e->h_ktemp.clear();
  if(!reuse_index) { e->h_kpack.clear(); e->h_kslot.clear(); e->kindex_valid = false; }      // valid again only once every table of the new index is queued
  const double* I = box->inverse_cell;
  const double ax[3] = {I[0], I[3], I[6]}, ay[3] = {I[1], I[4], I[7]}, az[3] = {I[2], I[5], I[8]};
  const double alpha_sq = box->alpha * box->alpha;
  const double prefactor = box->prefactor * (2.0 * GBK_PI / box->volume);
  if(box->alpha > 0.0 && !reuse_index)
    for(long long kxyz = 0; kxyz < e->nvec; kxyz++)
    {
      const int kz = (int)(kxyz % (2 * kzm + 1)) - kzm;
      const int kxy = (int)(kxyz / (2 * kzm + 1));
      const int kx = kxy / (2 * kym + 1);
      const int ky = kxy % (2 * kym + 1) - kym;
      double ksqr = (double)(kx * kx + ky * ky + kz * kz);
      if(box->use_lammps_ewald)
      {
        const double lx = box->cell[0], ly = box->cell[4], lz = box->cell[8], xy = box->cell[3], xz = box->cell[6], yz = box->cell[7];
        const double ux = 2 * GBK_PI / lx, uy = 2 * GBK_PI * (-xy) / lx / ly, uz = 2 * GBK_PI * (xy * yz - ly * xz) / lx / ly / lz;
        const double vy = 2 * GBK_PI / ly, vz = 2 * GBK_PI * (-yz) / ly / lz, wz = 2 * GBK_PI / lz;
        const double kvx = kx * ux, kvy = kx * uy + ky * vy, kvz = kx * uz + ky * vz + kz * wz;
        ksqr = kvx * kvx + kvy * kvy + kvz * kvz;
      }
      if(!((ksqr > 1e-10) && (ksqr < box->reciprocal_cutoff))) continue;
      e->h_kpack.push_back((kx << 16) | ((ky + 128) << 8) | (kz + 128));
      e->h_kslot.push_back((int) kxyz);
    }
  if(box->alpha <= 0.0) { e->h_kpack.clear(); e->h_kslot.clear(); }
  e->nact = (int) e->h_kpack.size(); e->nact_pad = (e->nact + 31) / 32 * 32;
  // the weights of the active wave vectors: the only part of the table that follows the cell (one code path, rebuilt and reused index alike)
  e->h_ktemp.reserve((size_t) e->nact);
  for(int k = 0; k < e->nact; k++)
  {
    const int kx = e->h_kpack[k] >> 16, ky = ((e->h_kpack[k] >> 8) & 255) - 128, kz = (e->h_kpack[k] & 255) - 128;
    double kv[3];
    for(int d = 0; d < 3; d++) kv[d] = ax[d] * 2.0 * GBK_PI * (double) kx + ay[d] * 2.0 * GBK_PI * (double) ky + az[d] * 2.0 * GBK_PI * (double) kz;
    const double rksq = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2];
    const double factor = (kx == 0) ? (1.0 * prefactor) : (2.0 * prefactor);
    e->h_ktemp.push_back(factor * std::exp((-0.25 / alpha_sq) * rksq) / rksq);
  }
  if(!reuse_index)
  {
    // rows (kx, ky) of the active list, longest kz reach first; 32 rows per round, positions [round][j = |kz|][sign][lane]
    std::map<int, std::vector<int>> rows;                  // key = kpack >> 8 (kx, ky), values = active indices
    std::vector<int> order;
    for(int k = 0; k < e->nact; k++) { const int key = e->h_kpack[k] >> 8; if(!rows.count(key)) order.push_back(key); rows[key].push_back(k); }
    auto reach = [&](int key) { int m = 0; for(int k : rows[key]) m = std::max(m, std::abs((e->h_kpack[k] & 255) - 128)); return m; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return reach(a) > reach(b); });
    e->nrounds = ((int) order.size() + 31) / 32;
    e->h_round.assign(2 * (size_t) std::max(e->nrounds, 1), 0); e->h_rowmeta.assign(32 * (size_t) std::max(e->nrounds, 1), (128 << 8));
    e->h_rowidx.clear();
    for(int g = 0; g < e->nrounds; g++)
    {
      const int m = reach(order[(size_t) g * 32]);
      const int off = (int) e->h_rowidx.size();
      e->h_round[2 * g] = off; e->h_round[2 * g + 1] = m;
      e->h_rowidx.resize((size_t) off + (size_t) (m + 1) * 64, -1);
      for(int l = 0; l < 32 && (size_t) g * 32 + l < order.size(); l++)
      {
        const int key = order[(size_t) g * 32 + l];
        e->h_rowmeta[(size_t) g * 32 + l] = key << 8;
        for(int k : rows[key])
        {
          const int kz = (e->h_kpack[k] & 255) - 128, j = std::abs(kz), sgn = kz < 0 ? 1 : 0;
          e->h_rowidx[(size_t) off + ((size_t) j * 2 + sgn) * 32 + l] = k;
        }
      }
    }
    e->npos = (int) e->h_rowidx.size();
    if(e->npos > 0)
    {
      CUDA_TRY(e->d_rowidx.reserve(e->npos)); CUDA_TRY(e->d_rowmeta.reserve(e->h_rowmeta.size())); CUDA_TRY(e->d_round.reserve(e->h_round.size()));
      CUDA_TRY(cudaMemcpyAsync(e->d_rowidx.p, e->h_rowidx.data(), e->npos * sizeof(int), cudaMemcpyHostToDevice, e->stream));
      CUDA_TRY(cudaMemcpyAsync(e->d_rowmeta.p, e->h_rowmeta.data(), e->h_rowmeta.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
      CUDA_TRY(cudaMemcpyAsync(e->d_round.p, e->h_round.data(), e->h_round.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    }
  }
  if(e->nact > 0)
  {
    CUDA_TRY(e->d_kpack.reserve(e->nact)); CUDA_TRY(e->d_kslot.reserve(e->nact)); CUDA_TRY(e->d_ktemp.reserve(e->nact));
    if(!reuse_index)
    {
      CUDA_TRY(cudaMemcpyAsync(e->d_kpack.p, e->h_kpack.data(), e->nact * sizeof(int), cudaMemcpyHostToDevice, e->stream));
      CUDA_TRY(cudaMemcpyAsync(e->d_kslot.p, e->h_kslot.data(), e->nact * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(e->d_ktemp.p, e->h_ktemp.data(), e->nact * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  }
  e->kindex_valid = !box->use_lammps_ewald && box->alpha > 0.0;
  e->kindex_kmax[0] = kxm; e->kindex_kmax[1] = kym; e->kindex_kmax[2] = kzm; e->kindex_rcut = box->reciprocal_cutoff;
  for(int i = 0; i < 3; i++)
  {
    CUDA_TRY(e->d_sf[i].reserve((size_t) std::max<long long>(2 * e->nvec, 2)));
    CUDA_TRY(cudaMemsetAsync(e->d_sf[i].p, 0, (size_t) std::max<long long>(2 * e->nvec, 2) * sizeof(double), e->stream));
  }
  e->i_ads = 0; e->i_fw = 1; e->i_tmp = 2; e->have_sf = false; e->ktab_dirty = true;
  e->have_box = true;
  CUDA_TRY(cudaStreamSynchronize(e->stream));       // one wait for all the table uploads above (their host vectors are rebuilt by the next call)
  return GB_OK;
}
} // namespace

int gb_upload_box(gb_engine* e, const gb_box* box)
{
  if(!e || !box) return fail(GB_ERR_ARG, "null argument");
  if(e->vol_pending) return fail(GB_ERR_STATE, "a volume move is pending: call gb_volume_move_finish first");
  int rc = apply_box(e, box); if(rc) return rc;
  if(e->committed && !e->device_stale && e->nslots > 0)
  {
    // accept calls or a volume move made the device copy authoritative: the host staging arrays are stale and must not be uploaded
    // over it.  Only the fractional mirror depends on the box; refresh it in place (as gb_volume_move_trial does).
    const size_t n = (size_t) e->nslots;
    k_frac_update<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->P, e->dx.p, e->dy.p, e->dz.p, e->dfx.p, e->dfy.p, e->dfz.p, 0, (int) n);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->pack_dirty = true;
    return GB_OK;
  }
  e->device_stale = true;
  return GB_OK;
}

int gb_set_components(gb_engine* e, int32_t n_total, int32_t n_host)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  if(n_total <= 0 || n_total > GBK_MAX_SEG || n_host < 0 || n_host > n_total) return fail(GB_ERR_ARG, "bad component counts");
  for(auto& C : e->comps) if(C.d_pocket) { cudaFree(C.d_pocket); C.d_pocket = nullptr; }     // block-pocket lists belong to the old layout
  e->ncomp = n_total; e->nhost = n_host; e->comps.assign(n_total, Comp()); e->tail_memo.clear();
  e->nslots = 0; e->hx.clear(); e->hy.clear(); e->hz.clear(); e->hq.clear(); e->hscale.clear(); e->hscoul.clear(); e->htype.clear(); e->hmolid.clear();
  e->device_stale = true;
  return GB_OK;
}

int gb_upload_atoms(gb_engine* e, int32_t c, const gb_atoms* a)
{
  if(!e || !a) return fail(GB_ERR_ARG, "null argument");
  if(c < 0 || c >= e->ncomp) return fail(GB_ERR_ARG, "component out of range");
  if(a->n_live > a->n_upload || a->n_upload > a->n_alloc || (a->molsize <= 0 && a->n_alloc > 0)) return fail(GB_ERR_ARG, "inconsistent atom counts");   // an empty component (a box without framework atoms) may have molsize 0
  if(!e->have_ff) return fail(GB_ERR_STATE, "upload the force field before atoms");
  if(a->n_upload > 0 && !a->pos) return fail(GB_ERR_ARG, "null positions");
  if(a->type) for(int64_t i = 0; i < a->n_upload; i++) if(a->type[i] >= (uint64_t) e->ntypes) return fail(GB_ERR_ARG, "atom type outside the force-field table");   // checked before any state changes
  if(e->comps[c].uploaded && e->comps[c].alloc != (int) a->n_alloc) return fail(GB_ERR_ARG, "component re-uploaded with a different n_alloc");
  if(!e->comps[c].uploaded) for(int k = 0; k < c; k++) if(!e->comps[k].uploaded) return fail(GB_ERR_STATE, "upload components in order");
  if(e->committed && !e->device_stale && e->nslots > 0)
  {
    // accept calls changed the device copy: bring the host staging arrays up to date before they are re-uploaded
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    const size_t n = (size_t) e->nslots, b = n * sizeof(double);
    CUDA_TRY(copy_on_stream(e, e->hx.data(), e->dx.p, b, cudaMemcpyDeviceToHost)); CUDA_TRY(copy_on_stream(e, e->hy.data(), e->dy.p, b, cudaMemcpyDeviceToHost));
    CUDA_TRY(copy_on_stream(e, e->hz.data(), e->dz.p, b, cudaMemcpyDeviceToHost)); CUDA_TRY(copy_on_stream(e, e->hq.data(), e->dq.p, b, cudaMemcpyDeviceToHost));
    CUDA_TRY(copy_on_stream(e, e->hscale.data(), e->dscale.p, b, cudaMemcpyDeviceToHost)); CUDA_TRY(copy_on_stream(e, e->hscoul.data(), e->dscoul.p, b, cudaMemcpyDeviceToHost));
    CUDA_TRY(copy_on_stream(e, e->htype.data(), e->dtype.p, n * sizeof(int), cudaMemcpyDeviceToHost)); CUDA_TRY(copy_on_stream(e, e->hmolid.data(), e->dmolid.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    e->committed = false;
  }
  Comp& C = e->comps[c];
  if(!C.uploaded)
  {
    C.offset = e->nslots; C.alloc = (int) a->n_alloc; e->nslots += C.alloc;
    const size_t n = (size_t) e->nslots;
    e->hx.resize(n, 0.0); e->hy.resize(n, 0.0); e->hz.resize(n, 0.0); e->hq.resize(n, 0.0); e->hscale.resize(n, 1.0); e->hscoul.resize(n, 1.0);
    e->htype.resize(n, 0); e->hmolid.resize(n, 0);
  }
  else
  {
    for(int i = 0; i < C.natoms; i++) { const int t = e->htype[(size_t) C.offset + i]; if(t >= 0 && t < e->ntypes) e->npseudo[t]--; }
  }
  C.molsize = std::max(1, (int) a->molsize); C.natoms = (int) a->n_live; C.uploaded = true;
  for(int64_t i = 0; i < a->n_upload; i++)
  {
    const size_t g = (size_t) C.offset + i;
    e->hx[g] = a->pos[3 * i]; e->hy[g] = a->pos[3 * i + 1]; e->hz[g] = a->pos[3 * i + 2];
    e->hq[g] = a->charge ? a->charge[i] : 0.0; e->hscale[g] = a->scale ? a->scale[i] : 1.0; e->hscoul[g] = a->scale_coul ? a->scale_coul[i] : 1.0;
    e->htype[g] = a->type ? (int) a->type[i] : 0; e->hmolid[g] = a->molid ? (int) a->molid[i] : 0;
  }
  for(int i = 0; i < C.natoms; i++) e->npseudo[e->htype[(size_t) C.offset + i]]++;
  bool hc = false;
  for(int i = 0; i < C.molsize && i < (int) a->n_upload; i++) if(std::fabs(e->hq[(size_t) C.offset + i]) > 1e-10) hc = true;
  C.has_charge = hc ? 1 : 0;
  e->device_stale = true; e->pack_dirty = true;
  return GB_OK;
}

int gb_download_atoms(gb_engine* e, int32_t c, double* pos, double* scale, double* charge, double* scale_coul,
                      uint64_t* type, uint64_t* molid, int64_t* n_live)
{
  int rc = ready(e); if(rc) return rc;
  if(c < 0 || c >= e->ncomp) return fail(GB_ERR_ARG, "component out of range");
  const Comp& C = e->comps[c];
  const size_t n = (size_t) C.alloc;
  std::vector<double> t(n * 6); std::vector<int> ti(n * 2);
  const size_t b = n * sizeof(double);
  CUDA_TRY(cudaMemcpyAsync(t.data(), e->dx.p + C.offset, b, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaMemcpyAsync(t.data() + n, e->dy.p + C.offset, b, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaMemcpyAsync(t.data() + 2 * n, e->dz.p + C.offset, b, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaMemcpyAsync(t.data() + 3 * n, e->dscale.p + C.offset, b, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaMemcpyAsync(t.data() + 4 * n, e->dq.p + C.offset, b, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaMemcpyAsync(t.data() + 5 * n, e->dscoul.p + C.offset, b, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaMemcpyAsync(ti.data(), e->dtype.p + C.offset, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaMemcpyAsync(ti.data() + n, e->dmolid.p + C.offset, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  for(size_t i = 0; i < n; i++)
  {
    if(pos) { pos[3 * i] = t[i]; pos[3 * i + 1] = t[n + i]; pos[3 * i + 2] = t[2 * n + i]; }
    if(scale) scale[i] = t[3 * n + i];
    if(charge) charge[i] = t[4 * n + i];
    if(scale_coul) scale_coul[i] = t[5 * n + i];
    if(type) type[i] = (uint64_t) ti[i];
    if(molid) molid[i] = (uint64_t) ti[n + i];
  }
  if(n_live) *n_live = C.natoms;
  return GB_OK;
}

// Snapshot of LIVE molecules only (restart / movie writers, write_data.h:109-263, axpy.cu:25-69): the reference copies every
// component's Allocate_size slots back (10 240 per adsorbate by default); a writer needs size = N_molecules x Molsize.
int gb_snapshot_molecules(gb_engine* e, int32_t c, int64_t first, int64_t count, double* pos, double* charge, double* scale, double* scale_coul)
{
  int rc = ready(e); if(rc) return rc;
  if(c < 0 || c >= e->ncomp) return fail(GB_ERR_ARG, "component out of range");
  const Comp& C = e->comps[c];
  const int64_t nmol = C.natoms / C.molsize;
  if(first < 0 || count < 0 || first + count > nmol) return fail(GB_ERR_ARG, "molecule range outside the live molecules");
  const size_t n = (size_t) count * C.molsize, o = (size_t) C.offset + (size_t) first * C.molsize;
  if(n == 0) return GB_OK;
  std::vector<double> t(n * 3);
  const size_t b = n * sizeof(double);
  if(pos)
  {
    CUDA_TRY(cudaMemcpyAsync(t.data(), e->dx.p + o, b, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaMemcpyAsync(t.data() + n, e->dy.p + o, b, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaMemcpyAsync(t.data() + 2 * n, e->dz.p + o, b, cudaMemcpyDeviceToHost, e->stream));
  }
  if(charge) CUDA_TRY(cudaMemcpyAsync(charge, e->dq.p + o, b, cudaMemcpyDeviceToHost, e->stream));
  if(scale) CUDA_TRY(cudaMemcpyAsync(scale, e->dscale.p + o, b, cudaMemcpyDeviceToHost, e->stream));
  if(scale_coul) CUDA_TRY(cudaMemcpyAsync(scale_coul, e->dscoul.p + o, b, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if(pos) for(size_t i = 0; i < n; i++) { pos[3 * i] = t[i]; pos[3 * i + 1] = t[n + i]; pos[3 * i + 2] = t[2 * n + i]; }
  return GB_OK;
}

int gb_upload_structure_factors(gb_engine* e, const double* ads, const double* fw)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  if(!e->have_box) return fail(GB_ERR_STATE, "upload the box first");
  CUDA_TRY(cudaSetDevice(e->device));
  const size_t b = (size_t) e->nvec * 2 * sizeof(double);
  if(ads) CUDA_TRY(copy_on_stream(e, e->d_sf[e->i_ads].p, ads, b, cudaMemcpyHostToDevice));
  if(fw)  CUDA_TRY(copy_on_stream(e, e->d_sf[e->i_fw].p, fw, b, cudaMemcpyHostToDevice));
  e->have_sf = true; e->ktab_dirty = true;
  return GB_OK;
}

int gb_download_structure_factors(gb_engine* e, double* ads, double* fw, double* tmp)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  CUDA_TRY(cudaSetDevice(e->device));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  const size_t b = (size_t) e->nvec * 2 * sizeof(double);
  if(ads) CUDA_TRY(copy_on_stream(e, ads, e->d_sf[e->i_ads].p, b, cudaMemcpyDeviceToHost));
  if(fw)  CUDA_TRY(copy_on_stream(e, fw, e->d_sf[e->i_fw].p, b, cudaMemcpyDeviceToHost));
  if(tmp) CUDA_TRY(copy_on_stream(e, tmp, e->d_sf[e->i_tmp].p, b, cudaMemcpyDeviceToHost));
  return GB_OK;
}

int gb_set_exclusion_constants(gb_engine* e, int32_t c, double intra, double atom, int32_t rigid, int32_t has_charge)
{
  if(!e || c < 0 || c >= e->ncomp) return fail(GB_ERR_ARG, "bad component");
  e->comps[c].excl_intra = intra; e->comps[c].excl_atom = atom; e->comps[c].rigid = rigid; e->comps[c].has_charge = has_charge;
  return GB_OK;
}

// ReadBlockingPockets / ReplicateBlockPockets (read_data.cpp:3290-3454) leave Cartesian centres and radii per component;
// BlockedPocket (:3466-3640) tests trial positions against them on the host.  Here the list lives on the device and the
// move kernels run the test themselves (first-bead trials, grown molecules, translation / rotation proposals).
int gb_set_block_pockets(gb_engine* e, int32_t c, int32_t n, const double* centers, const double* radii, int32_t invert)
{
  if(!e || c < 0 || c >= e->ncomp) return fail(GB_ERR_ARG, "bad component");
  if(n < 0 || (n > 0 && (!centers || !radii))) return fail(GB_ERR_ARG, "bad block-pocket list");
  CUDA_TRY(cudaSetDevice(e->device));
  Comp& C = e->comps[c];
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if(C.d_pocket) { cudaFree(C.d_pocket); C.d_pocket = nullptr; }
  C.npocket = 0; C.pocket_invert = invert ? 1 : 0;
  if(n == 0) return GB_OK;
  std::vector<double> h((size_t) 4 * n);
  for(int i = 0; i < n; i++) { h[4 * i] = centers[3 * i]; h[4 * i + 1] = centers[3 * i + 1]; h[4 * i + 2] = centers[3 * i + 2]; h[4 * i + 3] = radii[i]; }
  CUDA_TRY(cudaMalloc(&C.d_pocket, h.size() * sizeof(double)));
  CUDA_TRY(copy_on_stream(e, C.d_pocket, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  C.npocket = n;
  return GB_OK;
}

int gb_upload_random_pool(gb_engine* e, const double* r3, int64_t n)
{
  if(!e || !r3 || n <= 0) return fail(GB_ERR_ARG, "bad pool");
  CUDA_TRY(cudaSetDevice(e->device));
  // a refill of the same size runs beside the resident move server (a copy on the engine's stream); a LARGER pool frees the old
  // buffer, and cudaFree waits for every kernel on the device: stop the server first
  if(e->srv_running && (size_t) n * 3 > e->d_pool.cap) { int rc = server_stop(e); if(rc) return rc; }
  CUDA_TRY(e->d_pool.reserve((size_t) n * 3));
  CUDA_TRY(cudaMemcpyAsync(e->d_pool.p, r3, (size_t) n * 3 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->n_pool = n; e->pool_gen++;
  return GB_OK;
}

int gb_set_cbmc(gb_engine* e, int32_t ntp, int32_t nto, double beta)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  if(ntp <= 0 || ntp > GBK_MAX_TRIALS || nto <= 0 || nto > GBK_MAX_TRIALS) return fail(GB_ERR_ARG, "trial counts must be in [1,32]");
  e->ntrials = ntp; e->norient = nto; e->P.beta = beta; e->have_cbmc = true;
  return GB_OK;
}

int gb_get_pseudo_atom_counts(gb_engine* e, int64_t* counts)
{
  if(!e || !counts) return fail(GB_ERR_ARG, "null argument");
  for(int i = 0; i < e->ntypes; i++) counts[i] = e->npseudo[i];
  return GB_OK;
}

int gb_number_of_molecules(gb_engine* e, int32_t c, int64_t* n)
{
  if(!e || !n || c < 0 || c >= e->ncomp) return fail(GB_ERR_ARG, "bad argument");
  *n = e->comps[c].molsize > 0 ? e->comps[c].natoms / e->comps[c].molsize : 0;
  return GB_OK;
}

// ---------------------------------------------------------------------------------------------- trial energies
static int widom_cells_grid(gb_engine* e, WcGrid& G, long long n_trials);
static int trial_energies_cells(gb_engine* e, long long ntr, int cs, const TrialBuf& B, const double* d_tq, const int* d_ttype, double* d_out6, int* d_flag, bool* done);

int gb_trial_energies(gb_engine* e, int32_t ntr, int32_t cs, const double* pos, const double* scale, const double* charge,
                      const double* scale_coul, const uint64_t* type, int32_t new_comp, int64_t new_molid,
                      int32_t excl_comp, int64_t excl_mol, double* out_energy, int32_t* out_flag)
{
  int rc = ready(e); if(rc) return rc;
  if(ntr <= 0 || cs <= 0 || cs > GBK_MAX_CS || !pos || !type || !out_energy || !out_flag) return fail(GB_ERR_ARG, "bad trial arguments");
  const size_t n = (size_t) ntr * cs;
  std::vector<double> h(n * 5); std::vector<int> ht(n);
  // fractional coordinates on the host is set-up arithmetic only (same product as to_frac)
  for(size_t i = 0; i < n; i++)
  {
    const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    h[i] = e->P.inv[0] * x + e->P.inv[3] * y + e->P.inv[6] * z;
    h[n + i] = e->P.inv[1] * x + e->P.inv[4] * y + e->P.inv[7] * z;
    h[2 * n + i] = e->P.inv[2] * x + e->P.inv[5] * y + e->P.inv[8] * z;
    h[3 * n + i] = (charge ? charge[i] : 0.0) * (scale_coul ? scale_coul[i] : 1.0);
    h[4 * n + i] = scale ? scale[i] : 1.0;
    ht[i] = (int) type[i];
    if(ht[i] < 0 || ht[i] >= e->ntypes) return fail(GB_ERR_ARG, "trial atom type outside the force-field table");
  }
  CUDA_TRY(e->d_scratch.reserve(n * 5 + (size_t) ntr * 6)); CUDA_TRY(e->d_iscratch.reserve(n + ntr));
  CUDA_TRY(cudaMemcpyAsync(e->d_scratch.p, h.data(), n * 5 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->d_iscratch.p, ht.data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  TrialBuf B; B.fx = e->d_scratch.p; B.fy = B.fx + n; B.fz = B.fy + n; B.q = B.fz + n; B.scale = B.q + n; B.type = e->d_iscratch.p;
  double* d_out = e->d_scratch.p + n * 5; int* d_flag = e->d_iscratch.p + n;
  SegList L = seg_list(e, 0);
  // a large batch of equal groups of an adsorbate molecule that excludes nothing of the system (trial insertions): the cell-sorted
  // energy kernel of the Widom stage (from 64 trial atoms per 2 A cell on, like gb_widom_batch); otherwise one CTA per group
  bool cells = false;
  {
    bool eligible = e->P.all_unit_scale && !e->wc_overflowed && cs <= 33 && excl_comp < 0 && (new_comp < 0 || new_comp >= e->nhost) && !std::getenv("GB_TRIAL_NO_CELLS");
    if(eligible && new_comp >= 0 && new_comp < e->ncomp) eligible = new_molid < 0 || new_molid >= e->comps[new_comp].natoms / std::max(1, e->comps[new_comp].molsize);
    if(eligible) { WcGrid G; widom_cells_grid(e, G, 0); eligible = (long long) n >= 64LL * G.ncells; }
    for(size_t i = 0; i < n && eligible; i++) eligible = h[4 * n + i] == 1.0 && h[3 * n + i] == h[3 * n + i % cs] && ht[i] == ht[i % cs];
    // a candidate list that overflows its shared-memory capacity: again with fewer CTAs per SM (larger lists), like gb_widom_batch
    while(eligible && !cells && !e->wc_overflowed)
    {
      Timer tm(e, 0);
      rc = trial_energies_cells(e, ntr, cs, B, B.q, B.type, d_out, d_flag, &cells); if(rc) return rc;
      if(cells) tm.stop(6);
    }
  }
  if(!cells)
  {
    Timer tm(e, 0);
    k_trial_energies<<<ntr, 256, 0, e->stream>>>(e->P, sys_view(e), L, B, cs, new_comp, (int) new_molid, excl_comp, (int) excl_mol, d_out, d_flag);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    tm.stop(1);
  }
  std::vector<double> o((size_t) ntr * 6);
  CUDA_TRY(cudaMemcpyAsync(o.data(), d_out, o.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaMemcpyAsync(out_flag, d_flag, ntr * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  for(int t = 0; t < ntr; t++)
  {
    // HH contributions (a framework component inserted against other framework components) are reported with HG,
    // as the reference's HG blocks cover every host component (VDW_Coulomb.cu:1232-1235)
    out_energy[4 * t] = o[6 * t + 2] + o[6 * t]; out_energy[4 * t + 1] = o[6 * t + 3] + o[6 * t + 1];
    out_energy[4 * t + 2] = o[6 * t + 4]; out_energy[4 * t + 3] = o[6 * t + 5];
  }
  return GB_OK;
}

// ---------------------------------------------------------------------------------------------- Ewald delta
// enqueue the Ewald delta kernel (no synchronisation): result {same, 2*cross} goes to d_result2, skipped when *d_dep == 0
static int ewald_delta_enqueue(gb_engine* e, bool framework_moved, int nold, int nnew, const double* d_pos3, const double* d_qeff,
                               double* d_result2, const double* d_dep, bool same_is_temp = false)
{
  const int n = nold + nnew;
  if(n <= 0 || n > GBK_EW_MAX_ATOMS) return fail(GB_ERR_ARG, "Ewald delta supports 1..64 moved atoms");
  EwaldDeltaArgs A;
  A.pos3 = d_pos3; A.qeff = d_qeff; A.nold = nold; A.nnew = nnew;
  A.K.kpack = e->d_kpack.p; A.K.temp = e->d_ktemp.p; A.K.slot = e->d_kslot.p; A.K.nact = e->nact;
  A.same_sf = framework_moved ? e->d_sf[e->i_fw].p : e->d_sf[e->i_ads].p;
  if(same_is_temp) A.same_sf = e->d_sf[e->i_tmp].p;   // UseTempVector (second step of a CBCF deletion, Ewald_Energy_Functions.h:713-716): read and written per k in place
  A.cross_sf = framework_moved ? e->d_sf[e->i_ads].p : e->d_sf[e->i_fw].p;
  A.temp_sf = e->d_sf[e->i_tmp].p;     // only active k are written; inactive entries of all three arrays stay zero
  const int nblk = (e->nact + 127) / 128;
  CUDA_TRY(e->d_partial.reserve((size_t) std::max(nblk * 2, 64)));
  A.partial = e->d_partial.p; A.ticket = e->d_ticket.p; A.result = d_result2; A.dep = d_dep;
  const size_t smem = (size_t) n * (e->P.kmax[0] + e->P.kmax[1] + e->P.kmax[2] + 3) * sizeof(cplx) + (size_t) n * sizeof(double) + 16;
  if(smem > e->smem_optin) return fail(GB_ERR_ARG, "eik tables exceed shared memory");
  Timer tm(e, 1);
  k_ewald_delta<<<nblk, 128, smem, e->stream>>>(e->P, A);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  tm.stop(1);
  return GB_OK;
}

static int ewald_delta_launch(gb_engine* e, bool framework_moved, int nold, int nnew, const double* d_pos3, const double* d_qeff, double out[2],
                              bool same_is_temp = false)
{
  if(e->nact == 0) { out[0] = 0.0; out[1] = 0.0; return GB_OK; }
  int rc = ewald_delta_enqueue(e, framework_moved, nold, nnew, d_pos3, d_qeff, e->d_result.p, nullptr, same_is_temp); if(rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->h_pinned, e->d_result.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  out[0] = e->h_pinned[0]; out[1] = e->h_pinned[1];
  return GB_OK;
}

int gb_ewald_delta_explicit(gb_engine* e, int32_t fw_moved, int32_t nold, int32_t nnew, const double* pos, const double* charge,
                            const double* scale_coul, double out[2])
{
  int rc = ready(e); if(rc) return rc;
  if(!pos || !charge || !out) return fail(GB_ERR_ARG, "null argument");
  const int n = nold + nnew;
  if(n <= 0 || n > GBK_EW_MAX_ATOMS) return fail(GB_ERR_ARG, "Ewald delta supports 1..64 moved atoms");
  std::vector<double> h((size_t) n * 4);
  for(int i = 0; i < 3 * n; i++) h[i] = pos[i];
  for(int i = 0; i < n; i++) h[3 * n + i] = (scale_coul ? scale_coul[i] : 1.0) * charge[i];
  CUDA_TRY(e->d_scratch.reserve((size_t) n * 4));
  CUDA_TRY(cudaMemcpyAsync(e->d_scratch.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  return ewald_delta_launch(e, fw_moved != 0, nold, nnew, e->d_scratch.p, e->d_scratch.p + 3 * n, out);
}

int gb_ewald_commit(gb_engine* e, int32_t c)
{
  if(!e || c < 0 || c >= e->ncomp) return fail(GB_ERR_ARG, "bad component");
  if(c < e->nhost) std::swap(e->i_fw, e->i_tmp); else std::swap(e->i_ads, e->i_tmp);   // Ewald_Energy_Functions.h:423-433
  e->ktab_dirty = true;
  return GB_OK;
}

// ---------------------------------------------------------------------------------------------- tail
int gb_tail_total(gb_engine* e, double* out)
{
  int rc = ready(e); if(rc) return rc;
  rc = tail_device(e, nullptr, e->d_result.p + 8); if(rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->h_pinned + 8, e->d_result.p + 8, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  *out = e->h_pinned[8];
  return GB_OK;
}

int gb_tail_difference(gb_engine* e, int32_t c, int32_t move_type, double* out)
{
  if(e && e->have_ff && !e->has_tail && out) { *out = 0.0; return GB_OK; }     // HasTailCorrection == false (:39)
  SrvCompat srv_ok(e);
  int rc = ready(e); if(rc) return rc;
  if(c < 0 || c >= e->ncomp) return fail(GB_ERR_ARG, "bad component");
  if(move_type == GB_CBCF_INSERTION || move_type == GB_CBCF_DELETION) return fail(GB_ERR_UNIMPLEMENTED, "tail corrections are not defined for CBCF moves (TailCorrection_Energy_Functions.h:48-55)");
  std::vector<int> d = species_counts(e, c);
  const int sign = (move_type == GB_DELETION) ? -1 : 1;     // every other type keeps sign = +1 (:40-61)
  for(auto& v : d) v *= sign;
  return tail_delta_memo(e, d, out);
}

int gb_tail_identity_swap(gb_engine* e, int32_t newc, int32_t oldc, double* out)
{
  SrvCompat srv_ok(e);
  int rc = ready(e); if(rc) return rc;
  if(newc < 0 || newc >= e->ncomp || oldc < 0 || oldc >= e->ncomp) return fail(GB_ERR_ARG, "bad component");
  std::vector<int> dn = species_counts(e, newc), dold = species_counts(e, oldc);
  for(int i = 0; i < e->ntypes; i++) dn[i] -= dold[i];
  return tail_delta_memo(e, dn, out);
}

// ---------------------------------------------------------------------------------------------- totals
namespace {
int total_vdw_real_impl(gb_engine* e, gb_move_energy* out, int32_t* overlap);
int total_ewald_impl(gb_engine* e, int32_t store, bool device_convention, gb_move_energy* out);
}

int gb_total_vdw_real(gb_engine* e, gb_move_energy* out) { return total_vdw_real_impl(e, out, nullptr); }
int gb_total_ewald(gb_engine* e, int32_t store, gb_move_energy* out) { return total_ewald_impl(e, store, false, out); }

namespace {
// overlap != NULL: also report whether any pair trips OverlapCriteria or r^2 < 0.01 (VDWCoulEnergy_Total, VDW_Coulomb.cu:1385-1386)
int total_vdw_real_impl(gb_engine* e, gb_move_energy* out, int32_t* overlap)
{
  NvtxRange nvtx_call("gb_total_vdw_real");
  int rc = ready(e); if(rc) return rc;
  if(!out) return fail(GB_ERR_ARG, "null out");
  memset(out, 0, sizeof(*out));
  if(overlap) *overlap = 0;
  TotalArgs A; A.L = seg_list(e, 0); A.nhost = e->nhost; A.flag = nullptr;
  int nlive = 0; for(int s = 0; s < A.L.nseg; s++) nlive += A.L.count[s];
  if(nlive == 0) return GB_OK;
  CUDA_TRY(e->d_scratch.reserve((size_t) nlive * 6 + 8));
  A.out = e->d_scratch.p;
  if(overlap)
  {
    CUDA_TRY(e->d_iscratch.reserve(16));
    CUDA_TRY(cudaMemsetAsync(e->d_iscratch.p, 0, sizeof(int), e->stream));
    A.flag = e->d_iscratch.p;
  }
  Timer tm(e, 0);
  k_total_vdw_real<<<nlive, 128, 0, e->stream>>>(e->P, sys_view(e), A);
  k_reduce_partials<<<1, 32, 0, e->stream>>>(e->d_scratch.p, nlive, 6, e->d_result.p + 16);
  e->launches += 2;
  CUDA_TRY(cudaGetLastError());
  tm.stop(2);
  CUDA_TRY(cudaMemcpyAsync(e->h_pinned + 16, e->d_result.p + 16, 6 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  const double* r = e->h_pinned + 16;
  out->HHVDW = r[0]; out->HHReal = r[1]; out->HGVDW = r[2]; out->HGReal = r[3]; out->GGVDW = r[4]; out->GGReal = r[5];
  if(overlap)
  {
    int f = 0;
    CUDA_TRY(copy_on_stream(e, &f, e->d_iscratch.p, sizeof(int), cudaMemcpyDeviceToHost));
    *overlap = f ? 1 : 0;
  }
  return GB_OK;
}

// device_convention = false: the terms as the CPU Ewald_Total reports them (GG includes HH, ewald_preparation.h:174);
// true: as the device Ewald_TotalEnergy does (HH, HG, GG separate, each minus its own exclusions; Ewald_Energy_Functions.h:1366-1428)
int total_ewald_impl(gb_engine* e, int32_t store, bool device_convention, gb_move_energy* out)
{
  NvtxRange nvtx_call("gb_total_ewald");
  int rc = ready(e); if(rc) return rc;
  if(!out) return fail(GB_ERR_ARG, "null out");
  memset(out, 0, sizeof(*out));
  if(e->P.no_charges || e->nact == 0) return GB_OK;
  EwaldTotalArgs A;
  A.x = e->dx.p; A.y = e->dy.p; A.z = e->dz.p; A.q = e->dq.p; A.scoul = e->dscoul.p;
  A.L = seg_list(e, 0);
  A.K.kpack = e->d_kpack.p; A.K.temp = e->d_ktemp.p; A.K.slot = e->d_kslot.p; A.K.nact = e->nact;
  int nhost_atoms = 0; for(int c = 0; c < e->nhost; c++) nhost_atoms += e->comps[c].natoms;
  A.has_fw = (e->nhost > 0 && nhost_atoms > 0) ? 1 : 0;
  if(store)
  {
    CUDA_TRY(cudaMemsetAsync(e->d_sf[e->i_ads].p, 0, (size_t) e->nvec * 2 * sizeof(double), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->d_sf[e->i_fw].p, 0, (size_t) e->nvec * 2 * sizeof(double), e->stream));
    A.sf_ads = e->d_sf[e->i_ads].p; A.sf_fw = e->d_sf[e->i_fw].p;
  }
  else { A.sf_ads = nullptr; A.sf_fw = nullptr; }
  // exclusion records: one {self, intra} pair per live atom, then a fixed-order column sum per component
  int nlive_total = 0; for(int c = 0; c < e->ncomp; c++) nlive_total += e->comps[c].natoms;
  CUDA_TRY(e->d_scratch.reserve((size_t) e->nact * 3 + (size_t) nlive_total * 2 + 64));
  A.ek = e->d_scratch.p;
  Timer tm(e, 1);
  k_ewald_total<<<e->nact, 128, 0, e->stream>>>(e->P, A);
  k_reduce_partials<<<1, 32, 0, e->stream>>>(A.ek, e->nact, 3, e->d_result.p + 32);
  e->launches += 2;
  CUDA_TRY(cudaGetLastError());
  double* d_ex = e->d_scratch.p + (size_t) e->nact * 3;
  int aoff = 0;
  for(int c = 0; c < e->ncomp; c++)
  {
    const Comp& C = e->comps[c];
    if(C.natoms > 0)
    {
      ExclArgs X; X.x = e->dx.p; X.y = e->dy.p; X.z = e->dz.p; X.q = e->dq.p; X.scoul = e->dscoul.p; X.start = C.offset; X.natoms = C.natoms; X.ms = C.molsize;
      X.out = d_ex + 2 * (size_t) aoff;
      k_ewald_exclusion<<<(C.natoms + 127) / 128, 128, 0, e->stream>>>(e->P, X);
      k_reduce_partials<<<1, 32, 0, e->stream>>>(X.out, C.natoms, 2, e->d_result.p + 40 + 2 * c);
      e->launches += 2;
      CUDA_TRY(cudaGetLastError());
    }
    else CUDA_TRY(cudaMemsetAsync(e->d_result.p + 40 + 2 * c, 0, 2 * sizeof(double), e->stream));
    aoff += C.natoms;
  }
  tm.stop(2 + 2 * e->ncomp);
  CUDA_TRY(cudaMemcpyAsync(e->h_pinned + 32, e->d_result.p + 32, (8 + 2 * GBK_MAX_SEG) * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  double GG = e->h_pinned[32], HH = e->h_pinned[33], HG = e->h_pinned[34];
  if(!device_convention) GG += HH;                                     // ewald_preparation.h:174
  for(int c = 0; c < e->ncomp; c++)
  {
    const double self = e->h_pinned[40 + 2 * c], intra = e->h_pinned[40 + 2 * c + 1];
    if(device_convention)
    {
      if(c < e->nhost) { HH -= self; HH -= intra; } else { GG -= self; GG -= intra; }
      continue;
    }
    GG -= self; GG -= intra;
    if(c < e->nhost && A.has_fw) { HH -= self; HH -= intra; }
  }
  out->GGEwaldE = GG; out->HHEwaldE = HH; out->HGEwaldE = HG;
  if(store) { e->have_sf = true; e->ktab_dirty = true; }
  return GB_OK;
}
} // namespace

// ------------------------------------------------------------------------------------------------
// NPT volume move (VolumeMove, mc_box.h:196-320)
// ------------------------------------------------------------------------------------------------
int gb_volume_move_trial(gb_engine* e, const gb_box* new_box, double scale, gb_move_energy* out, int32_t* overlap)
{
  int rc = ready(e); if(rc) return rc;
  if(!new_box || !out || !overlap) return fail(GB_ERR_ARG, "null argument");
  if(e->vol_pending) return fail(GB_ERR_STATE, "a volume move is already pending");
  if(!(scale > 0.0)) return fail(GB_ERR_ARG, "scale must be positive");
  memset(out, 0, sizeof(*out)); *overlap = 0;
  const size_t n = (size_t) e->nslots;
  // 1. what a rejection falls back to: positions, box, stored structure factors
  CUDA_TRY(e->d_vol_xyz.reserve(3 * n + 8));
  CUDA_TRY(cudaMemcpyAsync(e->d_vol_xyz.p, e->dx.p, n * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->d_vol_xyz.p + n, e->dy.p, n * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->d_vol_xyz.p + 2 * n, e->dz.p, n * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  e->vol_old_box = e->cur_box; e->vol_had_sf = e->have_sf; e->vol_old_nvec = e->nvec;
  const size_t sfn = (size_t) std::max<long long>(2 * e->nvec, 2);
  CUDA_TRY(e->d_vol_sf.reserve(2 * sfn));
  CUDA_TRY(cudaMemcpyAsync(e->d_vol_sf.p, e->d_sf[e->i_ads].p, sfn * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->d_vol_sf.p + sfn, e->d_sf[e->i_fw].p, sfn * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  // 2. ScalePositions (mc_box.h:41-64): every molecule of components >= 1 moves with its first atom, rigid about it
  ScaleArgs SA; memset(&SA, 0, sizeof(SA)); SA.scale = scale;
  long long nmol = 0;
  for(int c = 1; c < e->ncomp && SA.nseg < GBK_MAX_SEG; c++)
  {
    const Comp& C = e->comps[c];
    if(C.natoms == 0) continue;
    SA.start[SA.nseg] = C.offset; SA.ms[SA.nseg] = C.molsize; SA.nmol[SA.nseg] = C.natoms / C.molsize; nmol += SA.nmol[SA.nseg]; SA.nseg++;
  }
  if(nmol > 0)
  {
    k_scale_molecules<<<(unsigned)((nmol + 127) / 128), 128, 0, e->stream>>>(e->P, SA, e->dx.p, e->dy.p, e->dz.p);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->vol_pending = true; e->committed = true;
  // 3. the new box: scalars, k table, empty structure factors; fractional coordinates of every slot.
  //    From here on a failure puts the old state back before it is reported, so that the engine is never left half-scaled.
  auto undo = [&](int code) { const std::string msg = gb_last_error(); gb_volume_move_finish(e, 0); return fail(code, msg); };
  rc = apply_box(e, new_box); if(rc) return undo(rc);
  rc = ready(e); if(rc) return undo(rc);
  if(n > 0)
  {
    k_frac_update<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->P, e->dx.p, e->dy.p, e->dz.p, e->dfx.p, e->dfy.p, e->dfz.p, 0, (int) n);
    e->launches++;
    { const cudaError_t le = cudaGetLastError(); if(le != cudaSuccess) { fail(GB_ERR_CUDA, std::string("k_frac_update: ") + cudaGetErrorString(le)); return undo(GB_ERR_CUDA); } }
  }
  e->pack_dirty = true;
  // 4. total energies of the scaled system; the structure factors of the new state are stored on the way
  gb_move_energy v, w;
  rc = total_vdw_real_impl(e, &v, overlap); if(rc) return undo(rc);
  rc = total_ewald_impl(e, 1, true, &w); if(rc) return undo(rc);
  *out = v; out->HHEwaldE = w.HHEwaldE; out->HGEwaldE = w.HGEwaldE; out->GGEwaldE = w.GGEwaldE;
  return GB_OK;
}

int gb_volume_move_finish(gb_engine* e, int32_t accept)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  if(!e->vol_pending) return fail(GB_ERR_STATE, "no volume move is pending");
  CUDA_TRY(cudaSetDevice(e->device));
  e->vol_pending = false;
  if(accept) return GB_OK;                                            // CopyScaledPositions + swap of the structure factors: already in place
  // Revert_Boxsize (mc_box.h:129-190) + the old positions and structure factors
  const size_t n = (size_t) e->nslots;
  int rc = apply_box(e, &e->vol_old_box); if(rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->dx.p, e->d_vol_xyz.p, n * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->dy.p, e->d_vol_xyz.p + n, n * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->dz.p, e->d_vol_xyz.p + 2 * n, n * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  const size_t sfn = (size_t) std::max<long long>(2 * e->vol_old_nvec, 2);
  CUDA_TRY(cudaMemcpyAsync(e->d_sf[e->i_ads].p, e->d_vol_sf.p, sfn * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->d_sf[e->i_fw].p, e->d_vol_sf.p + sfn, sfn * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  e->have_sf = e->vol_had_sf; e->ktab_dirty = true; e->pack_dirty = true;
  rc = ready(e); if(rc) return rc;
  if(n > 0)
  {
    k_frac_update<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->P, e->dx.p, e->dy.p, e->dz.p, e->dfx.p, e->dfy.p, e->dfz.p, 0, (int) n);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return GB_OK;
}

// ---------------------------------------------------------------------------------------------- batched Widom
// stage A launch shared by gb_widom_batch and gb_widom_first_bead_success
// ins0 / n_total: this launch covers insertions [ins0, ins0 + n) of a batch of n_total (its inputs already point at ins0)
static bool widom_cells_wanted(gb_engine* e, int comp, long long n_total);
static int widom_stage_a_cells(gb_engine* e, int comp, long long n, const double* d_pool, const long long* d_fb, const long long* d_or, const double* d_uni,
                               int first_bead_only, long long ins0, long long n_total, bool resume = false, bool keep = false);

static int widom_stage_a(gb_engine* e, int comp, long long n, const double* d_pool, const long long* d_fb, const long long* d_or, const double* d_uni,
                         int first_bead_only, long long ins0 = 0, long long n_total = 0)
{
  int rc;
  // large batches (and components with block pockets) take the cell-sorted pair stage; the choice is made on the size of the WHOLE
  // batch so that the chunks of a pipelined upload all go the same way
  if(widom_cells_wanted(e, comp, n_total > 0 ? n_total : ins0 + n)) return widom_stage_a_cells(e, comp, n, d_pool, d_fb, d_or, d_uni, first_bead_only, ins0, n_total);
  const Comp& C = e->comps[comp];
  const int ms = C.molsize, cs = ms - 1;
  const int rec_stride = 5 + 3 * ms;
  if(n_total < ins0 + n) n_total = ins0 + n;
  if(ins0 == 0) { CUDA_TRY(e->d_rec.reserve((size_t) n_total * rec_stride)); CUDA_TRY(e->d_stage.reserve((size_t) n_total)); }
  const int warpsA = 16;
  const size_t per_warpA = widom_per_warp_bytes(e->norient, cs);
  bool use_pack = false;
  const bool stage_ff = e->ntypes <= 24;
  const size_t ff_bytes = stage_ff ? (size_t) e->ntypes * e->ntypes * sizeof(double4) : 0;
  rc = ensure_pack(e, use_pack, warpsA * per_warpA + GBK_ERFC_BYTES + ff_bytes + sizeof(SegList) + 128); if(rc) return rc;
  WidomA A;
  A.pool3 = d_pool; A.fb_index = d_fb; A.or_index = d_or; A.uni = d_uni; A.n = n;
  A.ntrials = e->ntrials; A.norient = e->norient; A.ms = ms; A.comp = comp; A.new_molid = C.natoms / ms;
  A.tx = e->dx.p + C.offset; A.ty = e->dy.p + C.offset; A.tz = e->dz.p + C.offset; A.tq = e->dq.p + C.offset;
  A.tscoul = e->dscoul.p + C.offset; A.ttype = e->dtype.p + C.offset;
  A.pack = e->d_pack.p; A.npad = e->pack_npad; A.ntp = e->pack_ntp; A.pack_n = e->pack_n; A.use_pack = use_pack ? 1 : 0; A.stage_ff = stage_ff ? 1 : 0;
  A.first_bead_only = first_bead_only;
  A.rec = e->d_rec.p + (size_t) ins0 * rec_stride; A.stage = e->d_stage.p + ins0;
  SegList L = seg_list(e, 0);
  if(use_pack)
  {
    int acc = 0;
    for(int s = 0; s < L.nseg; s++) if(L.comp[s] < e->nhost) { L.staged[s] = 1; L.start[s] = acc; acc += L.count[s]; }
  }
  const size_t headA = (GBK_SMEM_TABLES_OFF + GBK_ERFC_BYTES_PAD + ff_bytes + sizeof(SegList) + 15) / 16 * 16;
  const size_t smemA = headA + (use_pack ? pack_bytes_for(e->pack_npad, e->pack_ntp) : 0) + warpsA * per_warpA;
  if(smemA > e->smem_optin) return fail(GB_ERR_ARG, "Widom stage A shared memory exceeds the device limit");
  const int gridA = (int) std::min<long long>((n + warpsA - 1) / warpsA, e->prop.multiProcessorCount);
  Timer tm(e, 0);
  if(e->P.cell_mode == 2)      k_widom_pair<2><<<gridA, warpsA * 32, smemA, e->stream>>>(e->P, sys_view(e), L, A);
  else if(e->P.cell_mode == 1) k_widom_pair<1><<<gridA, warpsA * 32, smemA, e->stream>>>(e->P, sys_view(e), L, A);
  else                         k_widom_pair<0><<<gridA, warpsA * 32, smemA, e->stream>>>(e->P, sys_view(e), L, A);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  tm.stop(1);
  return GB_OK;
}

// ---- cell-sorted pair stage (widom_cells.cuh): same inputs and the same rec / stage outputs as widom_stage_a
static int widom_cells_grid(gb_engine* e, WcGrid& G, long long n_trials = 0);
static bool widom_cells_wanted(gb_engine* e, int comp, long long n_total)
{
  const Comp& C = e->comps[comp];
  if(!e->P.all_unit_scale || C.molsize > 33 || e->wc_overflowed) return false;
  const char* env = std::getenv("GB_WIDOM_PATH");             // A/B measurements and tests: "cells" / "warp"
  if(env && std::strcmp(env, "cells") == 0) return true;
  if(env && std::strcmp(env, "warp") == 0) return C.npocket > 0;
  if(C.npocket > 0) return true;
  // the per-cell list build is amortised, and the lanes of the lane-per-trial kernel are filled, from ~64 first-bead trials per cell on
  WcGrid G; widom_cells_grid(e, G);
  return n_total * (long long) e->ntrials >= 64LL * G.ncells;
}

// Cells are 2 A wide (the measured optimum for large batches; fixed, so that an insertion's result does not depend on the size of the
// batch it is in).  n_trials > 0 (the RNG-exact replay: the rows of the pool it walks through): a pool with fewer than 64 rows per
// such cell gets wider cells, up to 4 A, so that a candidate list is still built for ~64 trial atoms
static int widom_cells_grid(gb_engine* e, WcGrid& G, long long n_trials)
{
  const double* H = e->P.cell;
  double h = 2.0;
  if(const char* env = std::getenv("GB_WC_H")) h = std::max(0.5, std::atof(env));
  for(;;)
  {
    long long nc = 1;
    for(int k = 0; k < 3; k++)
    {
      const double len = std::sqrt(H[3 * k] * H[3 * k] + H[3 * k + 1] * H[3 * k + 1] + H[3 * k + 2] * H[3 * k + 2]);
      G.n[k] = std::max(1, (int) std::lround(len / h));
      nc *= G.n[k];
    }
    const bool sparse = n_trials > 0 && n_trials < 64 * nc && h < 4.0 && !std::getenv("GB_WC_H");
    if(nc <= 131072 && !sparse) { G.ncells = (int) nc; break; }
    h *= 1.1;
  }
  G.rcell = 0.0;
  for(int k = 0; k < 3; k++) { G.inv_n[k] = 1.0 / (double) G.n[k]; G.margin[k] = 0.5 * G.inv_n[k] + 1e-9; }
  for(int sx = -1; sx <= 1; sx += 2) for(int sy = -1; sy <= 1; sy += 2)
  {
    const double f[3] = {0.5 * sx * G.inv_n[0], 0.5 * sy * G.inv_n[1], 0.5 * G.inv_n[2]};
    const double x = H[0] * f[0] + H[3] * f[1] + H[6] * f[2], y = H[1] * f[0] + H[4] * f[1] + H[7] * f[2], z = H[2] * f[0] + H[5] * f[1] + H[8] * f[2];
    G.rcell = std::max(G.rcell, std::sqrt(x * x + y * y + z * z));
  }
  G.rcell = G.rcell * (1.0 + 1e-9) + 1e-9;
  G.chunk = 512;
  if(const char* env = std::getenv("GB_WC_CHUNK")) G.chunk = std::max(32, std::atoi(env));
  return GB_OK;
}

// what one pass of the cell-sorted energy kernel needs: the grid, the live atoms gathered contiguously, the launch shape with the
// capacities of the candidate lists, and buffers for nmax trial atoms.  The caller fills in the trial atoms' template (E.tq, E.tscoul,
// E.ttype, E.ms), bins its trial atoms (wc_ucell / wc_udelta / wc_count) and calls wc_sort_and_energy.
struct WcPlan { WcGrid G; WcEnergy E; int mode = 1, ctas = 4, thrE = 192, gridE = 0, nads = 0, fast = 0; size_t smemE = 0; };

static int wc_plan(gb_engine* e, long long grid_basis, long long nmax, WcPlan& W)
{
  WcGrid& G = W.G;
  int rc = widom_cells_grid(e, G, grid_basis); if(rc) return rc;
  // ---- live atoms of every component, contiguous
  SegList L = seg_list(e, 0);
  int ntot = 0, nads = 0;
  for(int s = 0; s < L.nseg; s++) { ntot += L.count[s]; if(L.kind[s] == 2) nads += L.count[s]; }
  W.nads = nads;
  const size_t na = (size_t) std::max(ntot, 1);
  CUDA_TRY(e->wc_afx.reserve(na)); CUDA_TRY(e->wc_afy.reserve(na)); CUDA_TRY(e->wc_afz.reserve(na)); CUDA_TRY(e->wc_aq.reserve(na)); CUDA_TRY(e->wc_atk.reserve(na));
  if(ntot > 0)
  {
    k_wc_pack<<<(ntot + 255) / 256, 256, 0, e->stream>>>(sys_view(e), L, ntot, e->wc_afx.p, e->wc_afy.p, e->wc_afz.p, e->wc_aq.p, e->wc_atk.p);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  // ---- launch shape, capacities and shared memory of the energy kernel
  // mode 1 (default): a lane per trial atom, several small CTAs per SM, each with the lists of its own cell;
  // mode 0: a warp per trial atom, one large CTA per SM
  int& mode = W.mode; int& ctas = W.ctas; int& thrE = W.thrE;
  mode = 1; ctas = 4; thrE = 192;
  if(const char* env = std::getenv("GB_WC_MODE")) mode = std::atoi(env) ? 1 : 0;
  if(mode == 0) { ctas = 1; thrE = 768; }
  // the pair body without run-time switches (wc_sort_and_energy) needs 59 registers: four CTAs of 256 threads fit an SM (32 warps);
  // the general body (66 registers) four of 192
  const bool stage_ff = e->ntypes <= 24;
  W.fast = (stage_ff && !e->P.use1264 && !std::getenv("GB_WC_GENERAL")) ? (e->P.no_charges ? 2 : (e->P.erfc_table_ok ? 1 : 0)) : 0;
  if(W.fast == 1 && e->P.cut_vdw2 == e->P.cut_coul2 && !std::getenv("GB_WC_NO_SAMECUT")) W.fast = 3;
  if((W.fast == 1 || W.fast == 3) && e->P.erfc10_ok && !std::getenv("GB_WC_NO_SHORT_ERFC")) W.fast = (W.fast == 1) ? 4 : 5;      // 4 / 5: the short erfc table
  if(mode == 1 && (e->wc_ctas_cap < 4 || W.fast == 1 || W.fast >= 3)) thrE = 256;
  if(const char* env = std::getenv("GB_WC_CTAS")) ctas = std::max(1, std::min(mode ? 6 : 1, std::atoi(env)));
  ctas = std::max(1, std::min(ctas, e->wc_ctas_cap));
  e->wc_last_ctas = ctas;
  if(const char* env = std::getenv("GB_WC_THREADS")) thrE = std::min(mode ? 256 : 768, std::max(64, std::atoi(env) / 32 * 32));
  const size_t budget = std::min(e->smem_optin, (size_t) (e->prop.sharedMemPerMultiprocessor / ctas) - 1024 - 128);
  G.cap_fast = std::min((std::max(ntot, 32) + 1) & ~1, 2304); G.cap_slow = std::min((std::max(ntot, 32) + 1) & ~1, 768);
  while(wc_energy_smem(e->ntypes, stage_ff, G.cap_fast, G.cap_slow) > budget && G.cap_fast > 256) { G.cap_fast -= 64; G.cap_slow = std::max(128, G.cap_slow - 16); }
  W.smemE = wc_energy_smem(e->ntypes, stage_ff, G.cap_fast, G.cap_slow);
  if(W.smemE > e->smem_optin) return fail(GB_ERR_ARG, "cell-sorted Widom stage: shared memory exceeds the device limit");
  if(mode == 1 && !std::getenv("GB_WC_CHUNK")) G.chunk = 32 * (thrE / 32) * 4;
  // ---- buffers
  CUDA_TRY(e->wc_ucell.reserve((size_t) nmax)); CUDA_TRY(e->wc_udelta.reserve((size_t) nmax * 3)); CUDA_TRY(e->wc_srec.reserve((size_t) nmax));
  CUDA_TRY(e->wc_e4.reserve((size_t) nmax * 4)); CUDA_TRY(e->wc_flag.reserve((size_t) nmax));
  CUDA_TRY(e->wc_count.reserve((size_t) G.ncells + 1)); CUDA_TRY(e->wc_off.reserve((size_t) G.ncells + 1)); CUDA_TRY(e->wc_cursor.reserve((size_t) G.ncells + 1));
  CUDA_TRY(e->wc_items.reserve(3 * ((size_t) G.ncells + (size_t) (nmax / G.chunk) + 2))); CUDA_TRY(e->wc_ctl.reserve(8));
  WcEnergy& E = W.E;
  memset(&E, 0, sizeof(E));
  E.atoms.fx = e->wc_afx.p; E.atoms.fy = e->wc_afy.p; E.atoms.fz = e->wc_afz.p; E.atoms.q = e->wc_aq.p; E.atoms.tk = e->wc_atk.p; E.atoms.n = ntot;
  E.srec = e->wc_srec.p; E.items = e->wc_items.p; E.ctl = e->wc_ctl.p;
  E.stage_ff = stage_ff ? 1 : 0; E.e4 = e->wc_e4.p; E.flag = e->wc_flag.p; E.overflow = e->wc_ctl.p + 2;
  W.gridE = e->prop.multiProcessorCount * ctas;
  return GB_OK;
}

// counting sort of the binned trial atoms (wc_ucell / wc_udelta / wc_count) and the energy kernel over the cells
static int wc_sort_and_energy(gb_engine* e, WcPlan& W, long long nitems_src, int amod, int abase)
{
  WcGrid& G = W.G; WcEnergy& E = W.E;
  const int gridE = W.gridE, thrE = W.thrE; const size_t smemE = W.smemE;
  k_wc_scan<<<1, 1024, 0, e->stream>>>(e->wc_count.p, G.ncells, G.chunk, e->wc_off.p, e->wc_cursor.p, e->wc_items.p, e->wc_ctl.p);
  k_wc_scatter<<<(unsigned) ((nitems_src + 255) / 256), 256, 0, e->stream>>>(e->wc_ucell.p, e->wc_udelta.p, nitems_src, e->wc_off.p, e->wc_cursor.p, e->wc_srec.p);
  E.amod = amod; E.abase = abase;
  const bool gg = W.nads > 0;
#define GBK_WC_LAUNCH(K, ...) do { \
    if(e->P.cell_mode == 2)      { if(gg) K<2, true __VA_ARGS__><<<gridE, thrE, smemE, e->stream>>>(e->P, G, E); else K<2, false __VA_ARGS__><<<gridE, thrE, smemE, e->stream>>>(e->P, G, E); } \
    else if(e->P.cell_mode == 1) { if(gg) K<1, true __VA_ARGS__><<<gridE, thrE, smemE, e->stream>>>(e->P, G, E); else K<1, false __VA_ARGS__><<<gridE, thrE, smemE, e->stream>>>(e->P, G, E); } \
    else                         { if(gg) K<0, true __VA_ARGS__><<<gridE, thrE, smemE, e->stream>>>(e->P, G, E); else K<0, false __VA_ARGS__><<<gridE, thrE, smemE, e->stream>>>(e->P, G, E); } } while(0)
  {
    Timer te(e, 3);
    // the common case -- plain 12-6 LJ, LJ table staged in shared memory, every in-cutoff erfc argument inside the table -- runs a pair body
    // without run-time switches (1: with real-space Coulomb, 2: no charges); everything else the general one
    const int fast = W.fast;
    if(W.mode != 1) GBK_WC_LAUNCH(k_wc_energy);
    else if(fast == 1) GBK_WC_LAUNCH(k_wc_energy_lt, , 1);
    else if(fast == 2) GBK_WC_LAUNCH(k_wc_energy_lt, , 2);
    else if(fast == 3) GBK_WC_LAUNCH(k_wc_energy_lt, , 3);
    else if(fast == 4) GBK_WC_LAUNCH(k_wc_energy_lt, , 4);
    else if(fast == 5) GBK_WC_LAUNCH(k_wc_energy_lt, , 5);
    else GBK_WC_LAUNCH(k_wc_energy_lt, , 0);
    te.stop(1);
  }
#undef GBK_WC_LAUNCH
  e->launches += 3;
  CUDA_TRY(cudaGetLastError());
  return GB_OK;
}

// gb_trial_energies for a large batch of equal trial groups with nothing of the system excluded: the trial atoms (fractional coordinates
// in B) are binned into the cells and evaluated by the energy kernel of the Widom stage; d_out6 / d_flag as k_trial_energies leaves them.
// *done = false (nothing launched) when the route does not apply or a candidate list overflowed.
static int trial_energies_cells(gb_engine* e, long long ntr, int cs, const TrialBuf& B, const double* d_tq, const int* d_ttype, double* d_out6, int* d_flag, bool* done)
{
  *done = false;
  const long long n = ntr * cs;
  WcPlan W; int rc = wc_plan(e, 0, n, W); if(rc) return rc;
  CUDA_TRY(e->d_uni.reserve(64));                                   // 64 ones: the scaleCoul column of the trial template
  {
    std::vector<double> ones(64, 1.0);
    CUDA_TRY(cudaMemcpyAsync(e->d_uni.p, ones.data(), 64 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  }
  W.E.tq = d_tq; W.E.tscoul = e->d_uni.p; W.E.ttype = d_ttype; W.E.ms = cs;
  CUDA_TRY(cudaMemsetAsync(e->wc_count.p, 0, ((size_t) W.G.ncells + 1) * sizeof(int), e->stream));
  CUDA_TRY(cudaMemsetAsync(e->wc_ctl.p + 2, 0, sizeof(int), e->stream));
  k_wc_gen_explicit<<<(unsigned) ((n + 255) / 256), 256, 0, e->stream>>>(e->P, W.G, B.fx, B.fy, B.fz, n, e->wc_ucell.p, e->wc_udelta.p, e->wc_count.p);
  e->launches++;
  rc = wc_sort_and_energy(e, W, n, cs, 0); if(rc) return rc;
  k_wc_sum_groups<<<(unsigned) ((ntr + 255) / 256), 256, 0, e->stream>>>(e->wc_e4.p, e->wc_flag.p, ntr, cs, d_out6, d_flag);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<int*>(e->h_pinned + 12), e->wc_ctl.p + 2, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if(*reinterpret_cast<int*>(e->h_pinned + 12) != 0)
  {
    if(e->wc_last_ctas > 1) e->wc_ctas_cap = e->wc_last_ctas - 1; else e->wc_overflowed = true;
    return GB_OK;                                                   // the caller runs k_trial_energies instead
  }
  *done = true;
  return GB_OK;
}

static int widom_stage_a_cells(gb_engine* e, int comp, long long n, const double* d_pool, const long long* d_fb, const long long* d_or, const double* d_uni,
                               int first_bead_only, long long ins0, long long n_total, bool resume, bool keep)
{
  NvtxRange nvtx_stage("widom pair stage (cell-sorted)");
  const Comp& C = e->comps[comp];
  const int ms = C.molsize, cs = ms - 1;
  const int rec_stride = 5 + 3 * ms;
  if(n_total < ins0 + n) n_total = ins0 + n;
  if(ins0 == 0) { CUDA_TRY(e->d_rec.reserve((size_t) n_total * rec_stride)); CUDA_TRY(e->d_stage.reserve((size_t) n_total)); }
  if(n_total < ins0 + n) n_total = ins0 + n;
  // RNG-exact replay (kept / resumed first-bead energies): every row of the engine's pool is a trial position of the replay, whatever
  // share of it this call or this GPU evaluates -- the grid follows the pool, so results do not depend on how the pool was cut
  const long long nfb = n * e->ntrials, nch = n * (long long) e->norient * cs, nmax = std::max(nfb, nch);
  WcPlan W; int rc = wc_plan(e, (resume || keep) ? e->n_pool : 0, nmax, W); if(rc) return rc;
  CUDA_TRY(e->wc_fbres.reserve((size_t) n * 8));
  WcGrid& G = W.G;
  W.E.tq = e->dq.p + C.offset; W.E.tscoul = e->dscoul.p + C.offset; W.E.ttype = e->dtype.p + C.offset; W.E.ms = ms;
  auto sort_and_energy = [&](long long nitems_src, int amod, int abase) -> int { return wc_sort_and_energy(e, W, nitems_src, amod, abase); };
  Timer tm(e, 0);
  // ---- first beads (resume: their trial energies were kept, by pool row, by gb_widom_first_bead_success)
  if(ins0 == 0) CUDA_TRY(cudaMemsetAsync(e->wc_ctl.p + 2, 0, sizeof(int), e->stream));
  if(!resume)
  {
    CUDA_TRY(cudaMemsetAsync(e->wc_count.p, 0, ((size_t) G.ncells + 1) * sizeof(int), e->stream));
    WcGen Gn; Gn.pool3 = d_pool; Gn.fb_index = d_fb; Gn.n = n; Gn.ntrials = e->ntrials; Gn.norient = e->norient;
    Gn.ucell = e->wc_ucell.p; Gn.udelta = e->wc_udelta.p; Gn.count = e->wc_count.p;
    k_wc_gen_fb<<<(unsigned) ((nfb + 255) / 256), 256, 0, e->stream>>>(e->P, G, Gn);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    rc = sort_and_energy(nfb, 1, 0); if(rc) return rc;
    if(keep)
    {
      CUDA_TRY(e->fbk_e4.reserve((size_t) nfb * 4)); CUDA_TRY(e->fbk_flag.reserve((size_t) nfb));
      CUDA_TRY(cudaMemcpyAsync(e->fbk_e4.p, e->wc_e4.p, (size_t) nfb * 4 * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
      CUDA_TRY(cudaMemcpyAsync(e->fbk_flag.p, e->wc_flag.p, (size_t) nfb * sizeof(int), cudaMemcpyDeviceToDevice, e->stream));
    }
  }
  // ---- first-bead selection + chain trial atoms
  WcSel S;
  S.pool3 = d_pool; S.fb_index = d_fb; S.or_index = d_or; S.uni = d_uni; S.n = n; S.ntrials = e->ntrials; S.norient = e->norient; S.ms = ms;
  S.tx = e->dx.p + C.offset; S.ty = e->dy.p + C.offset; S.tz = e->dz.p + C.offset;
  memset(&S.C, 0, sizeof(S.C)); S.C.pocket = C.d_pocket; S.C.npocket = C.npocket; S.C.pocket_invert = C.pocket_invert;
  S.e4 = resume ? e->fbk_e4.p - 4 * e->fbk_row0 : e->wc_e4.p; S.flag = resume ? e->fbk_flag.p - e->fbk_row0 : e->wc_flag.p; S.e4_by_row = resume ? 1 : 0; S.fbres = e->wc_fbres.p;
  S.rec = e->d_rec.p + (size_t) ins0 * rec_stride; S.stage = e->d_stage.p + ins0; S.first_bead_only = first_bead_only;
  S.ucell = e->wc_ucell.p; S.udelta = e->wc_udelta.p; S.count = e->wc_count.p;
  if(cs > 0 && !first_bead_only)
  {
    CUDA_TRY(cudaMemsetAsync(e->wc_count.p, 0, ((size_t) G.ncells + 1) * sizeof(int), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->wc_ucell.p, 0xFF, (size_t) nch * sizeof(int), e->stream));           // -1: insertions whose first bead failed leave no chain atoms
  }
  k_wc_select_fb<<<(unsigned) ((n * 32 + 255) / 256), 256, 0, e->stream>>>(e->P, G, S);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  if(cs > 0 && !first_bead_only)
  {
    rc = sort_and_energy(nch, cs, 1); if(rc) return rc;
    S.e4 = e->wc_e4.p; S.flag = e->wc_flag.p; S.e4_by_row = 0;
    k_wc_select_chain<<<(unsigned) ((n * 32 + 255) / 256), 256, 0, e->stream>>>(e->P, S);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  tm.stop(resume ? 5 : (first_bead_only || cs == 0 ? 6 : 10));
  return GB_OK;
}

int gb_widom_first_bead_success(gb_engine* e, int32_t comp, int64_t n, const double* pool3, int64_t n_pool, const int64_t* fb_index, int32_t* code)
{
  int rc = ready(e); if(rc) return rc;
  e->call_serial--;                                            // a Widom call changes nothing of the system
  if(n <= 0 || !fb_index || !code) return fail(GB_ERR_ARG, "bad arguments");
  if(comp < e->nhost || comp >= e->ncomp) return fail(GB_ERR_ARG, "Widom component must be an adsorbate component");
  if(!e->have_cbmc) return fail(GB_ERR_STATE, "gb_set_cbmc has not been called");
  if(!pool3 && e->n_pool <= 0) return fail(GB_ERR_STATE, "pool3 is NULL and gb_upload_random_pool has not been called");
  // consecutive blocks in pool order (fb_index[k] = fb_index[0] + k * ntrials; a shard of the pool in a multi-GPU replay) on the engine's own pool: the per-trial energies can be kept by pool
  // row for gb_widom_batch(resume_first_bead); that needs the cell-sorted stage, which is taken from 16 pool rows per 2 A cell on here
  bool in_order = !pool3 && fb_index[0] % e->ntrials == 0;
  for(int64_t k = 0; k < n && in_order; k++) in_order = fb_index[k] == fb_index[0] + k * (int64_t) e->ntrials;
  WcGrid G; widom_cells_grid(e, G);
  const Comp& C = e->comps[comp];
  const bool keep = in_order && e->P.all_unit_scale && C.molsize <= 33 && !e->wc_overflowed && !std::getenv("GB_WIDOM_NO_RESUME") && e->n_pool >= 16LL * G.ncells;
  const bool cells = keep || widom_cells_wanted(e, comp, n);
  if(C.npocket > 0 && !cells) return fail(GB_ERR_UNIMPLEMENTED, "block pockets need the cell-sorted Widom stage (unit scaling factors, molecules of <= 33 atoms)");
  e->fbk_valid = false;
  if(pool3)
  {
    CUDA_TRY(e->d_pool.reserve((size_t) n_pool * 3));
    CUDA_TRY(cudaMemcpyAsync(e->d_pool.p, pool3, (size_t) n_pool * 3 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    e->n_pool = n_pool; e->pool_gen++;
  }
  for(int64_t k = 0; k < n; k++) if(fb_index[k] < 0 || fb_index[k] + e->ntrials > e->n_pool) return fail(GB_ERR_ARG, "first-bead block outside the pool");
  CUDA_TRY(e->d_idx0.reserve((size_t) n));
  CUDA_TRY(cudaMemcpyAsync(e->d_idx0.p, fb_index, (size_t) n * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
  if(cells) rc = widom_stage_a_cells(e, comp, n, e->d_pool.p, e->d_idx0.p, e->d_idx0.p, nullptr, 1, 0, n, false, keep);
  else      rc = widom_stage_a(e, comp, n, e->d_pool.p, e->d_idx0.p, e->d_idx0.p, nullptr, 1);
  if(rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(code, e->d_stage.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  if(cells) CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<int*>(e->h_pinned + 12), e->wc_ctl.p + 2, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if(cells && *reinterpret_cast<int*>(e->h_pinned + 12) != 0)
  {
    // a candidate list overflowed its shared-memory capacity: again with larger lists (fewer CTAs per SM), in the end with the warp kernel
    if(e->wc_last_ctas > 1) e->wc_ctas_cap = e->wc_last_ctas - 1; else e->wc_overflowed = true;
    return gb_widom_first_bead_success(e, comp, n, pool3, n_pool, fb_index, code);
  }
  if(keep) { e->fbk_valid = true; e->fbk_row0 = fb_index[0]; e->fbk_rows = n * (long long) e->ntrials; e->fbk_comp = comp; e->fbk_pool_gen = e->pool_gen; e->fbk_serial = e->call_serial; }
  return GB_OK;
}

int gb_widom_batch(gb_engine* e, int32_t comp, int64_t n, const gb_widom_inputs* in, double* out8, int32_t* stage,
                   int32_t outputs_on_device, double* sums)
{
  NvtxRange nvtx_call("gb_widom_batch");
  int rc = ready(e); if(rc) return rc;
  e->call_serial--;                                            // a Widom call changes nothing of the system
  if(!in || n <= 0 || !in->uniforms) return fail(GB_ERR_ARG, "bad Widom inputs");
  const bool engine_pool = !in->pool3;
  if(engine_pool && (in->inputs_on_device || e->n_pool <= 0)) return fail(GB_ERR_STATE, "pool3 is NULL: needs host inputs and a pool left by gb_upload_random_pool");
  const bool resume = in->resume_first_bead != 0;
  if(resume)
  {
    if(!engine_pool || !in->fb_index || !in->or_index) return fail(GB_ERR_ARG, "resume_first_bead needs pool3 == NULL and fb_index / or_index");
    if(e->wc_overflowed) return fail(GB_ERR_STATE, "resume_first_bead: the cell-sorted stage is not available for this system (candidate lists overflow)");
    if(!e->fbk_valid || e->fbk_comp != comp || e->fbk_pool_gen != e->pool_gen || e->fbk_serial != e->call_serial)
      return fail(GB_ERR_STATE, "resume_first_bead: no valid first-bead energies (gb_widom_first_bead_success on this pool, nothing but Widom calls since)");
    for(int64_t i = 0; i < n; i++)
      if(in->fb_index[i] < e->fbk_row0 || in->fb_index[i] % e->ntrials != 0 || in->fb_index[i] + e->ntrials > e->fbk_row0 + e->fbk_rows) return fail(GB_ERR_ARG, "resume_first_bead: fb_index is not the start of an evaluated block");
  }
  if(comp < e->nhost || comp >= e->ncomp) return fail(GB_ERR_ARG, "Widom component must be an adsorbate component");
  if(!e->have_cbmc) return fail(GB_ERR_STATE, "gb_set_cbmc has not been called");
  const Comp& C = e->comps[comp];
  const bool cells = resume || widom_cells_wanted(e, comp, n);
  if(C.npocket > 0 && !cells) return fail(GB_ERR_UNIMPLEMENTED, "block pockets need the cell-sorted Widom stage (unit scaling factors, molecules of <= 33 atoms)");
  const int ms = C.molsize, cs = ms - 1;
  if(cs > GBK_MAX_CS) return fail(GB_ERR_ARG, "molecule too large for the CBMC chain stage");
  const bool do_ewald = !e->P.no_charges && C.has_charge && e->nact > 0;
  if(do_ewald && !e->have_sf) return fail(GB_ERR_STATE, "structure factors have not been uploaded or computed (gb_total_ewald(store=1))");
  if(ms > GBK_EW_MAX_ATOMS) return fail(GB_ERR_ARG, "molecule too large for the Ewald stage");
  const int nbins = in->n_blocks > 0 ? in->n_blocks : 1;
  const int per = e->ntrials + e->norient;
  // tail-correction difference of one more molecule: constant over the batch, a function of the occupation numbers only, evaluated
  // once per state (memoised) and BEFORE anything of the batch is queued, so that no synchronisation falls between the stages
  double tail_value = 0.0;
  if(e->has_tail) { std::vector<int> dc = species_counts(e, comp); rc = tail_delta_memo(e, dc, &tail_value); if(rc) return rc; }

  // ---- inputs on the device
  const double* d_pool = in->pool3; const long long* d_fb = (const long long*) in->fb_index; const long long* d_or = (const long long*) in->or_index;
  const double* d_uni = in->uniforms;
  // host inputs in the packed layout and a batch worth pipelining: the randoms go up in chunks on a second stream while the pair
  // kernel already works on the chunks that have arrived (measured: 2-4 chunks are equivalent, 8 lose more in kernel tails than
  // they hide; default = a first chunk of n/8 and the rest, so that 7/8 of the 0.48 KB per insertion travel under compute)
  const int nchunk = (!in->inputs_on_device && !engine_pool && !in->fb_index && n >= 65536 && in->n_pool >= n * per && !std::getenv("GB_WIDOM_NO_OVERLAP")) ? (std::getenv("GB_WIDOM_CHUNKS") ? std::max(2, std::min(8, std::atoi(std::getenv("GB_WIDOM_CHUNKS")))) : 2) : 1;
  if(nchunk > 1)
  {
    if(!e->copy_stream)
    {
      CUDA_TRY(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
      for(int c = 0; c < 8; c++) CUDA_TRY(cudaEventCreateWithFlags(&e->ev_chunk[c], cudaEventDisableTiming));
    }
    CUDA_TRY(e->d_pool.reserve((size_t) in->n_pool * 3)); CUDA_TRY(e->d_uni.reserve((size_t) n * 2));
    d_pool = e->d_pool.p; d_uni = e->d_uni.p; e->n_pool = in->n_pool; e->pool_gen++;
    CUDA_TRY(cudaStreamSynchronize(e->stream));                      // nothing of an earlier call may still read the buffers
    long long c0[9];
    for(int c = 0; c <= nchunk; c++) c0[c] = (n * c / nchunk) / 32 * 32;
    if(nchunk == 2) c0[1] = (n / 8) / 32 * 32;          // a short first chunk gets the kernels going, the rest travels under it
    c0[nchunk] = n;
    for(int c = 0; c < nchunk; c++)
    {
      const long long a = c0[c], b = c0[c + 1];
      const size_t po = (size_t) a * per * 3, pn = (size_t) ((c + 1 == nchunk ? (long long) in->n_pool : b * per) - a * per) * 3;
      CUDA_TRY(cudaMemcpyAsync(e->d_pool.p + po, in->pool3 + po, pn * sizeof(double), cudaMemcpyHostToDevice, e->copy_stream));
      CUDA_TRY(cudaMemcpyAsync(e->d_uni.p + 2 * a, in->uniforms + 2 * a, (size_t) (b - a) * 2 * sizeof(double), cudaMemcpyHostToDevice, e->copy_stream));
      CUDA_TRY(cudaEventRecord(e->ev_chunk[c], e->copy_stream));
    }
    for(int c = 0; c < nchunk; c++)
    {
      const long long a = c0[c], b = c0[c + 1];
      CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_chunk[c], 0));
      rc = widom_stage_a(e, comp, b - a, d_pool + (size_t) a * per * 3, nullptr, nullptr, d_uni + 2 * a, 0, a, n); if(rc) return rc;
    }
  }
  else if(!in->inputs_on_device)
  {
    if(!engine_pool)
    {
      CUDA_TRY(e->d_pool.reserve((size_t) in->n_pool * 3));
      CUDA_TRY(cudaMemcpyAsync(e->d_pool.p, in->pool3, (size_t) in->n_pool * 3 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
      e->n_pool = in->n_pool; e->pool_gen++;
    }
    d_pool = e->d_pool.p;
    CUDA_TRY(e->d_uni.reserve((size_t) n * 2));
    CUDA_TRY(cudaMemcpyAsync(e->d_uni.p, in->uniforms, (size_t) n * 2 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    d_uni = e->d_uni.p;
    if(in->fb_index)
    {
      CUDA_TRY(e->d_idx0.reserve((size_t) n)); CUDA_TRY(e->d_idx1.reserve((size_t) n));
      CUDA_TRY(cudaMemcpyAsync(e->d_idx0.p, in->fb_index, (size_t) n * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
      CUDA_TRY(cudaMemcpyAsync(e->d_idx1.p, in->or_index, (size_t) n * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
      d_fb = e->d_idx0.p; d_or = e->d_idx1.p;
    }
  }
  if(!in->fb_index && (engine_pool ? e->n_pool : in->n_pool) < n * per) return fail(GB_ERR_ARG, "random pool smaller than n*(trial positions+orientations)");

  // ---- stage A
  if(resume) { rc = widom_stage_a_cells(e, comp, n, d_pool, d_fb, d_or, d_uni, 0, 0, n, true, false); if(rc) return rc; }
  else if(nchunk == 1) { rc = widom_stage_a(e, comp, n, d_pool, d_fb, d_or, d_uni, 0); if(rc) return rc; }
  // ---- stage B
  rc = ensure_ktab(e); if(rc) return rc;
  const int warpsB = GBK_EWALD_THREADS / 32;
  WidomB B;
  B.rec = e->d_rec.p; B.stage = e->d_stage.p; B.n = n; B.ms = ms; B.tq = e->dq.p + C.offset; B.tscoul = e->dscoul.p + C.offset;
  B.ktab = e->d_ktab.p; B.nact = e->nact; B.nact_pad = e->nact_pad; B.do_ewald = do_ewald ? 1 : 0;
  B.rtab = e->d_rtab.p; B.npos = e->npos; B.rowmeta = e->d_rowmeta.p; B.rounds = e->d_round.p; B.nrounds = (e->npos > 0 && C.molsize <= 4 && !std::getenv("GB_EWALD_FLAT")) ? e->nrounds : 0;      // GB_EWALD_FLAT: the one-k-per-lane loop, for A/B timing
  B.excl_const = C.rigid ? (C.excl_intra + C.excl_atom) * 1.0 : 0.0;
  B.tail = tail_value; B.nbins = nbins;
  B.gn = in->global_n > 0 ? in->global_n : n; B.gfirst = in->global_n > 0 ? in->global_first : 0;
  if(B.gfirst < 0 || B.gfirst + n > B.gn) return fail(GB_ERR_ARG, "shard [global_first, global_first + n) outside the job");
  const size_t per_warpB = ((size_t) ms * (e->P.kmax[0] + e->P.kmax[1] + e->P.kmax[2] + 3) * sizeof(cplx) + (size_t) ms * 4 * sizeof(double) + 15) / 16 * 16;
  const size_t fixedB = 16 + warpsB * per_warpB + (size_t) warpsB * nbins * 12 * sizeof(double) + 64;
  const size_t ktab_bytes = ((size_t) e->nact_pad * 44 + 15) / 16 * 16;
#ifdef GBK_EWALD_NO_STAGE
  B.stage_ktab = 0;
#else
  B.stage_ktab = (do_ewald && fixedB + ktab_bytes <= e->smem_optin) ? 1 : 0;
#endif
  const size_t smemB = fixedB + (B.stage_ktab ? ktab_bytes : 0);
  if(smemB > e->smem_optin) return fail(GB_ERR_ARG, "Widom stage B shared memory exceeds the device limit");
#ifndef GBK_EWALD_CTAS
#define GBK_EWALD_CTAS 1
#endif
  const int gridB = (int) std::min<long long>((n + warpsB - 1) / warpsB, (long long) e->prop.multiProcessorCount * GBK_EWALD_CTAS);
  if(out8 && !outputs_on_device) { CUDA_TRY(e->d_out8.reserve((size_t) n * 8)); B.out8 = e->d_out8.p; } else B.out8 = out8;
  B.out_stage = nullptr;   // stage already lives in d_stage
  CUDA_TRY(e->d_partial.reserve((size_t) gridB * nbins * 12)); CUDA_TRY(e->d_sums.reserve((size_t) nbins * 12));
  B.partial = e->d_partial.p;
  {
    NvtxRange nvtx_stage("widom fourier stage");
    Timer tm(e, 1);
    k_widom_ewald<<<gridB, warpsB * 32, smemB, e->stream>>>(e->P, B);
    k_reduce_partials<<<(nbins * 12 + 63) / 64, 64, 0, e->stream>>>(e->d_partial.p, gridB, nbins * 12, e->d_sums.p);
    e->launches += 2;
    if(in->sums_device) { k_add_sums<<<(nbins * 12 + 63) / 64, 64, 0, e->stream>>>(e->d_sums.p, in->sums_device, nbins * 12); e->launches++; }
    CUDA_TRY(cudaGetLastError());
    tm.stop(2);
  }
  // ---- outputs
  // the block sums come back through the engine's pinned block (h_pinned + 64 .. : 5 x 12 doubles fit; more bins go through a vector)
  std::vector<double> hs_big; double* hs = e->h_pinned + 64;
  if((size_t) nbins * 12 > 400) { hs_big.resize((size_t) nbins * 12); hs = hs_big.data(); }
  if(sums) CUDA_TRY(cudaMemcpyAsync(hs, e->d_sums.p, (size_t) nbins * 12 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if(out8 && !outputs_on_device) CUDA_TRY(cudaMemcpyAsync(out8, e->d_out8.p, (size_t) n * 8 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if(stage)
  {
    if(outputs_on_device) CUDA_TRY(cudaMemcpyAsync(stage, e->d_stage.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToDevice, e->stream));
    else CUDA_TRY(cudaMemcpyAsync(stage, e->d_stage.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  }
  if(cells) CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<int*>(e->h_pinned + 12), e->wc_ctl.p + 2, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if(cells && *reinterpret_cast<int*>(e->h_pinned + 12) != 0)
  {
    // a candidate list of the cell-sorted stage did not fit its shared-memory capacity (dense system, long cutoff): nothing of this
    // call is kept; the batch is evaluated again with fewer CTAs per SM (larger lists), in the end by the warp-per-insertion kernel
    if(e->wc_last_ctas > 1) e->wc_ctas_cap = e->wc_last_ctas - 1;          // fewer CTAs per SM = more shared memory for the lists of each
    else e->wc_overflowed = true;
    if(std::getenv("GB_DEBUG")) std::fprintf(stderr, "graspa_b200: cell-sorted Widom stage overflowed its candidate lists (n = %lld, %d CTAs per SM); %s\n", (long long) n,
                                             e->wc_last_ctas, e->wc_overflowed ? "falling back to k_widom_pair" : "retrying with larger lists");
    return gb_widom_batch(e, comp, n, in, out8, stage, outputs_on_device, sums);
  }
  if(sums) for(size_t i = 0; i < (size_t) nbins * 12; i++) sums[i] = hs[i];
  return GB_OK;
}

// ---------------------------------------------------------------------------------------------- instrumentation
int gb_launch_count(gb_engine* e, int64_t* n, int32_t reset)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  if(n) *n = e->launches;
  if(reset) e->launches = 0;
  return GB_OK;
}

int gb_timing_enable(gb_engine* e, int32_t on) { if(!e) return fail(GB_ERR_ARG, "null engine"); e->timing = on != 0; return GB_OK; }

int gb_move_server(gb_engine* e, int32_t on, int64_t* starts, int64_t* commands)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  if(on == 0)
  {
    CUDA_TRY(cudaSetDevice(e->device));
    if(e->srv_running || !e->pending_commits.empty()) { int rc = server_stop(e); if(rc) return rc; }
    e->srv_enabled = false;
  }
  else if(on > 0) e->srv_enabled = e->move_cooperative;
  if(starts) *starts = e->srv_starts;
  if(commands) *commands = e->srv_commands;
  return GB_OK;
}

int gb_timing_read(gb_engine* e, int32_t family, double* ms, int64_t* launches, int32_t reset)
{
  if(!e) return fail(GB_ERR_ARG, "null engine");
  double m = family == 0 ? e->ms_pair : (family == 1 ? e->ms_ewald : (family == 3 ? e->ms_wc : e->ms_pair + e->ms_ewald));
  long long l = family == 0 ? e->n_pair : (family == 1 ? e->n_ewald : (family == 3 ? e->n_wc : e->n_pair + e->n_ewald));
  if(ms) *ms = m;
  if(launches) *launches = l;
  if(reset) { e->ms_pair = e->ms_ewald = e->ms_wc = 0.0; e->n_pair = e->n_ewald = e->n_wc = 0; }
  return GB_OK;
}

int gb_measure_fp64_peak(gb_engine* e, double* tflops)
{
  if(!e || !tflops) return fail(GB_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(e->device));
  const int blocks = e->prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  CUDA_TRY(e->d_scratch.reserve((size_t) blocks * threads));
  float best = 1e30f;
  for(int rep = 0; rep < 4; rep++)
  {
    cudaEventRecord(e->ev0, e->stream);
    k_fp64_peak<<<blocks, threads, 0, e->stream>>>(e->d_scratch.p, iters, 1.0 + rep);
    cudaEventRecord(e->ev1, e->stream);
    CUDA_TRY(cudaEventSynchronize(e->ev1));
    float ms = 0.f; cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    if(rep > 0 && ms < best) best = ms;
    e->launches++;
  }
  const double flop = 2.0 * 64.0 * (double) iters * blocks * threads;
  *tflops = flop / (best * 1e-3) / 1e12;
  return GB_OK;
}

} // extern "C"
#include "moves.inc"
#include "fused_moves.inc"
