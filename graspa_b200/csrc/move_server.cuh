// graspa_b200 -- the resident move server: k_move without the launch.
//
// A GCMC move is ~700 warp instructions per warp of straight-line code; launched as its own kernel it spends its time fetching
// those instructions into cold instruction caches and ramping a grid up and down (profiles/r2_move_kernel.md: stall_no_instruction
// 16 per issue, 14 us per launch of which ~7 us is work).  k_move_server is ONE cooperative launch of one CTA per SM that stays
// resident while a Monte Carlo loop runs and executes move after move (move_body<true>, the same code as k_move):
//
//   host                              CTA 0                                   CTAs 1 .. G-1
//   ----                              -----                                   -------------
//   writes the FusedArgs of the move  polls the mailbox in pinned host        poll the relay copy of the command in device
//   as tagged 16-byte records into    memory (one PCIe read round trip),      memory (L2), acquire, run their share of the
//   the mailbox (pinned memory)       checks that every CTA has finished      move, then store "done = seq"
//                                     the previous command, applies the
//                                     queued state commits, releases, relays
//                                     the command, runs the move, stores the
//   polls the tagged result records   result records to pinned host memory
//
// Every value that crosses between the host and the GPU or between CTAs is a self-validating record {lo, tag, hi, tag} (tag = low
// 32 bits of the command's sequence number), so no flag has to be ordered against its data.  What a kernel boundary used to
// guarantee is restated explicitly: (1) a command is relayed only after every CTA has stored `done` for the previous one, so the
// hand-over records of a move are never overwritten while a CTA still polls them; (2) the commit of an accepted move (the accept calls
// queue it, the NEXT command carries it) is written by CTA 0, followed by a gpu-scope fence, before the relay, and every CTA runs an
// acquire fence after it has seen the command -- that also drops the L1 lines a CTA may still hold of the slot arrays; the move body
// reads F and the slot-array pointers from shared memory, so no load of mutable data is routed through the non-coherent path.
// The server never waits without a clock: CTA 0 leaves after `idle_ns` without a command (the host restarts it on demand), any
// other wait that exceeds `stuck_ns` traps (the host then reports GB_ERR_CUDA instead of hanging).
#pragma once
#include "fused_kernel.cuh"

#define GBS_NREC 128                     // 16-byte records of one command: 8 payload bytes each
static_assert(sizeof(FusedArgs) <= 8 * GBS_NREC, "FusedArgs does not fit the command mailbox");
static_assert(sizeof(FusedArgs) % 8 == 0, "FusedArgs is copied in 8-byte words");

enum { GBS_RUNNING = 1, GBS_EXITED = 2, GBS_IDLE_EXIT = 3 };

struct ServerCtl
{
  const unsigned long long* host_cmd;    // pinned host memory (UVA): GBS_NREC tagged records, written by the host
  unsigned long long* dev_cmd;           // device memory: the relay copy CTA 0 publishes for the other CTAs
  unsigned long long* done;              // device memory: per CTA, the sequence number of the last command it finished
  unsigned long long* host_status;       // pinned host memory: [0] = GBS_* state of the server
  unsigned long long first_seq;          // sequence number of the first command
  unsigned long long idle_ns, stuck_ns;
  SlotArrays slots;                      // where the commits go
  MoveBufs B;
};

struct SrvShared
{
  unsigned long long raw[GBS_NREC];      // the command: a FusedArgs
  DevParams P;                           // shared-memory copies: the out-of-line stage routines take them by reference
  SysView S;
  int abort;
};

__device__ __forceinline__ bool ll_load_u64(const unsigned long long* rec, int j, unsigned int tag, unsigned long long& v)
{
  unsigned int lo, t0, hi, t1;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(rec + 2 * j) : "memory");
  v = ((unsigned long long) hi << 32) | lo;
  return t0 == tag && t1 == tag;
}
__device__ __forceinline__ void ll_store_u64(unsigned long long* rec, int j, unsigned long long v, unsigned int tag)
{
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(rec + 2 * j), "r"((unsigned int) v), "r"(tag), "r"((unsigned int) (v >> 32)), "r"(tag) : "memory");
}
__device__ __forceinline__ unsigned long long srv_timer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__host__ __device__ inline unsigned long long srv_next_seq(unsigned long long s) { s++; if((s & 0xffffffffULL) == 0) s++; return s; }

// the commit kernels' bodies (k_commit_from_buffer / k_commit_delete, move_kernels.cuh) cut in two: field f of atom i of the source
// (0-8: x y z fx fy fz q scale scoul as bits, 9: type), and the store of that value into the slot arrays
__device__ __forceinline__ unsigned long long srv_commit_source(const SlotArrays& S, const MoveBufs& B, const CommitOp& c, int i, int f)
{
  if(c.op == 1) return f < 9 ? (unsigned long long) __double_as_longlong(B.mol(c.buf, f)[i]) : (unsigned long long) (unsigned int) B.mol_type(c.buf)[i];
  const int j = c.src + i;
  const double* a = f == 0 ? S.x : f == 1 ? S.y : f == 2 ? S.z : f == 3 ? S.fx : f == 4 ? S.fy : f == 5 ? S.fz : f == 6 ? S.q : f == 7 ? S.scale : S.scoul;
  return f < 9 ? (unsigned long long) __double_as_longlong(a[j]) : (unsigned long long) (unsigned int) S.type[j];
}
__device__ __forceinline__ void srv_commit_store(const SlotArrays& S, const CommitOp& c, int i, int f, unsigned long long v)
{
  const int j = c.dst + i;
  // op 1: what 0 = positions, 1 = + charge and scaling factors, 2 = + type and MolID; op 2 (deletion): everything but the MolID
  const int lim = c.op == 2 ? 10 : (c.what == 0 ? 6 : (c.what == 1 ? 9 : 10));
  if(f >= lim) return;
  if(f < 9)
  {
    double* a = f == 0 ? S.x : f == 1 ? S.y : f == 2 ? S.z : f == 3 ? S.fx : f == 4 ? S.fy : f == 5 ? S.fz : f == 6 ? S.q : f == 7 ? S.scale : S.scoul;
    a[j] = __longlong_as_double((long long) v);
  }
  else { S.type[j] = (int) (unsigned int) v; if(c.op == 1) S.molid[j] = c.molid; }
}

#define GBS_CREC (2 * 10 * GBK_MV_MOL_SLOTS)     // relay records of the commit payload: two operations of up to 64 atoms, 10 fields each

__host__ __device__ inline size_t srv_smem_fixed() { return ((sizeof(FusedSmem) + 127) / 128) * 128 + ((sizeof(SrvShared) + 127) / 128) * 128; }

__global__ void __launch_bounds__(256, 1)
k_move_server(DevParams P_in, SysView S_in, ServerCtl C)
{
  extern __shared__ __align__(128) unsigned char dyn_all[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(dyn_all);
  SrvShared& sh = *reinterpret_cast<SrvShared*>(dyn_all + ((sizeof(FusedSmem) + 127) / 128) * 128);
  unsigned char* dyn = dyn_all + srv_smem_fixed();
  const FusedArgs& F = *reinterpret_cast<const FusedArgs*>(sh.raw);
  if(!P_in.no_charges) stage_erfc_table(P_in, sm.etab);
  if(threadIdx.x == 0) { sh.P = P_in; sh.S = S_in; sh.abort = 0; }
  __syncthreads();
  const DevParams& P = sh.P;
  volatile int* abort_flag = &sh.abort;
  unsigned long long seq = C.first_seq, prev = 0;
  const bool lead = blockIdx.x == 0;
  for(;;)
  {
    const unsigned int tag = (unsigned int) seq;
    const unsigned long long t0 = srv_timer();
    if(threadIdx.x < GBS_NREC)
    {
      // ---- the command: thread t waits for record t (CTA 0: from the host's mailbox, over PCIe; the others: from the relay copy)
      const unsigned long long* src = lead ? C.host_cmd : C.dev_cmd;
      unsigned long long v = 0; unsigned int spins = 0;
      while(!ll_load_u64(src, threadIdx.x, tag, v))
      {
        if(*abort_flag) break;
        if(!lead) __nanosleep(40);
        if(threadIdx.x == 0 && (++spins & 63u) == 0)
        {
          const unsigned long long dt = srv_timer() - t0;
          if(lead ? dt > C.idle_ns : dt > C.stuck_ns) *abort_flag = 1;
        }
      }
      sh.raw[threadIdx.x] = v;
    }
    else if(lead && prev != 0)
    {
      // ---- meanwhile: every CTA has finished the previous command (its polls of the hand-over records and its reads of the slot arrays)
      for(int b = threadIdx.x - GBS_NREC + 1; b < (int) gridDim.x; b += (int) blockDim.x - GBS_NREC)
      {
        unsigned int spins = 0;
        while(*reinterpret_cast<volatile unsigned long long*>(C.done + b) != prev)
          if((++spins & 1023u) == 0 && srv_timer() - t0 > C.stuck_ns) __trap();
      }
    }
    __syncthreads();
#ifdef GBK_PHASE_TIMING
    if(lead && threadIdx.x == 0) g_nmarks = 0;
#endif
    GBK_MARK();
    if(*abort_flag)
    {
      if(!lead) __trap();                                   // the relay never came: CTA 0 is gone
      // idle: tell the other CTAs to leave (an EXIT command under the expected tag), then leave
      if(threadIdx.x < GBS_NREC) ll_store_u64(C.dev_cmd, threadIdx.x, threadIdx.x == 0 ? (unsigned long long) GBF_EXIT : 0ULL, tag);
      if(threadIdx.x == 0) { *reinterpret_cast<volatile unsigned long long*>(C.host_status) = GBS_IDLE_EXIT; }
      return;
    }
    // relay first: the other CTAs start on the command while CTA 0 is still busy below
    if(lead && threadIdx.x < GBS_NREC) ll_store_u64(C.dev_cmd, threadIdx.x, sh.raw[threadIdx.x], tag);
    // acquire: what the host copied to the device before this command (a refilled random pool: the copy had completed before the command
    // was posted) and what other CTAs wrote during earlier moves (structure factors of an accepted move) is visible to the loads below;
    // stale L1 lines are dropped (CCTL.IVALL)
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    if(F.kind != GBF_EXIT && F.ncommit > 0)
    {
      // the state change of the previous accepted move.  No CTA reads it from another CTA's stores: CTA 0 reads the source (the
      // molecule buffer it exported itself, or the slots of the last molecule), relays every value as a tagged record, and EVERY CTA
      // writes the same values into the slot arrays before it reads them -- identical stores from all CTAs, no fence on the path
      int base = 0;
      for(int k = 0; k < F.ncommit && k < 2; k++)
      {
        const CommitOp c = F.commit[k];
        if(c.op == 2 && c.dst == c.src) continue;
        for(int idx = threadIdx.x; idx < 10 * c.n; idx += blockDim.x)
        {
          const int i = idx / 10, f = idx - 10 * i;
          unsigned long long v;
          if(lead) { v = srv_commit_source(C.slots, C.B, c, i, f); ll_store_u64(C.dev_cmd, GBS_NREC + base + idx, v, tag); }
          else
          {
            unsigned int spins = 0;
            while(!ll_load_u64(C.dev_cmd, GBS_NREC + base + idx, tag, v)) if((++spins & 1023u) == 0 && srv_timer() - t0 > C.stuck_ns) __trap();
          }
          srv_commit_store(C.slots, c, i, f, v);
        }
        base += 10 * c.n;
      }
      __syncthreads();
    }
    GBK_MARK();
    if(F.kind == GBF_EXIT)
    {
      if(lead && threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(C.host_status) = GBS_EXITED;
      return;
    }
    if((int) blockIdx.x < F.ngrid) move_body<true>(P, sh.S, F, sm, dyn);
    __syncthreads();
    if(!lead && threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(C.done + blockIdx.x) = seq;
    prev = seq; seq = srv_next_seq(seq);
  }
}
