// steal_board.h - the ticket-board protocol of the tile-stealing lane kernels (conv_umma.cu, conv_steal_kernel).
//
// One board per lane (microbatch in flight).  The protocol is written once, over a tiny atomics shim, so the SAME code
// runs in the CUDA kernel (PTX acquire / release, atomicCAS) and in a host model driven by std::threads
// (tests/steal_model.cpp, tests/test_steal_protocol.py), where it is stress-tested: every tile claimed exactly once,
// no tile of op o+1 claimed before every tile of op o is complete, re-arming while other lanes poll.
//
// ticket = [epoch:16 | op:16 | next tile:32].  States of a lane:
//   all-zero board            never armed (n_ops == 0 rejects it)
//   (e, STEAL_OP_DONE, 0)     run e complete, or the lane is being re-armed
//   (e, o, k), k <  tiles(o)  tile k of op o is claimable
//   (e, o, k), k == tiles(o)  op o fully issued; the finisher of its last tile will publish (e, o+1, 0) / DONE
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define STEAL_HD __host__ __device__ __forceinline__
#else
#define STEAL_HD inline
#endif

namespace defer {

constexpr unsigned STEAL_OP_DONE = 0xffffu;

template <class OpT>
struct alignas(64) LaneBoardT {
  unsigned long long ticket;     // [epoch:16 | op:16 | next tile:32]
  unsigned int done;             // tiles of the current op that are complete
  unsigned int target_epoch;     // epoch of the lane's current run (set by arm)
  const OpT* ops;                // op descriptors of the current run; OpT has int m_tiles, n_tiles
  int n_ops;
  int pad_[9];
};

STEAL_HD unsigned long long pack_ticket(unsigned epoch, unsigned op, unsigned tile) {
  return ((unsigned long long)(epoch & 0xffffu) << 48) | ((unsigned long long)(op & 0xffffu) << 32) | tile;
}
STEAL_HD unsigned ticket_epoch(unsigned long long t) { return (unsigned)(t >> 48); }
STEAL_HD unsigned ticket_op(unsigned long long t) { return (unsigned)((t >> 32) & 0xffffu); }
STEAL_HD unsigned ticket_tile(unsigned long long t) { return (unsigned)(t & 0xffffffffu); }

// ---- atomics shim
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long sb_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sb_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool sb_cas(unsigned long long* p, unsigned long long expect, unsigned long long desired) {
  return atomicCAS(p, expect, desired) == expect;
}
__device__ __forceinline__ unsigned long long sb_xchg(unsigned long long* p, unsigned long long v) { return atomicExch(p, v); }
__device__ __forceinline__ unsigned sb_add(unsigned int* p, unsigned v) { return atomicAdd(p, v); }
__device__ __forceinline__ void sb_fence() { __threadfence(); }
template <class T>
__device__ __forceinline__ T sb_ld_relaxed(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
template <class T>
__device__ __forceinline__ void sb_st_relaxed(T* p, T v) { *reinterpret_cast<volatile T*>(p) = v; }
#else
inline unsigned long long sb_ld_acquire(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void sb_st_release(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline bool sb_cas(unsigned long long* p, unsigned long long expect, unsigned long long desired) {
  return __atomic_compare_exchange_n(p, &expect, desired, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
}
inline unsigned long long sb_xchg(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline unsigned sb_add(unsigned int* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline void sb_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T>
inline T sb_ld_relaxed(const T* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
template <class T>
inline void sb_st_relaxed(T* p, T v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
#endif

// (Re-)arm a lane for a new run.  Single caller, ordered after the lane's previous run (stream order on the GPU).
template <class OpT>
STEAL_HD void steal_arm(LaneBoardT<OpT>* b, const OpT* ops, int n_ops) {
  // Close the lane first: from here on a poller reads STEAL_OP_DONE, and a CAS that still expects an older ticket value
  // (e.g. the all-zero ticket of a never-armed board read just before) fails.
  const unsigned e = (sb_ld_relaxed(&b->target_epoch) + 1u) & 0xffffu;
  sb_xchg(&b->ticket, pack_ticket(e, STEAL_OP_DONE, 0));
  sb_fence();
  sb_st_relaxed(&b->target_epoch, e);
  sb_st_relaxed(&b->ops, ops);
  sb_st_relaxed(&b->n_ops, n_ops);
  sb_st_relaxed(&b->done, 0u);
  sb_fence();
  sb_st_release(&b->ticket, pack_ticket(e, 0, 0));
}

// Is run `epoch` of this lane complete?
template <class OpT>
STEAL_HD bool steal_lane_done(const LaneBoardT<OpT>* b, unsigned epoch) {
  const unsigned long long t = sb_ld_acquire(&b->ticket);
  return ticket_epoch(t) == epoch && ticket_op(t) == STEAL_OP_DONE;
}

struct StealClaim {
  int lane, op, tile;
};

// Try to claim one tile: own lane first, then the others round-robin.  Returns false if nothing is claimable right now.
// A successful claim was made with acquire semantics on the ticket: everything the previous op of that lane stored is
// visible to the claiming thread.
template <class OpT>
STEAL_HD bool steal_try_claim(LaneBoardT<OpT>* boards, int n_lanes, int my_lane, StealClaim* out, const OpT** op_out) {
  for (int i = 0; i < n_lanes; ++i) {
    int l = my_lane + i;
    if (l >= n_lanes) l -= n_lanes;
    LaneBoardT<OpT>* b = boards + l;
    const unsigned long long t = sb_ld_acquire(&b->ticket);
    const unsigned o = ticket_op(t);
    if (o == STEAL_OP_DONE) continue;                                   // finished, or being re-armed
    if (ticket_epoch(t) != sb_ld_relaxed(&b->target_epoch)) continue;   // stale
    if ((int)o >= sb_ld_relaxed(&b->n_ops)) continue;                   // never armed (zeroed board)
    const OpT* cand = sb_ld_relaxed(&b->ops) + o;
    const unsigned k = ticket_tile(t);
    if (k >= (unsigned)(cand->m_tiles * cand->n_tiles)) continue;       // fully issued, completion pending
    if (sb_cas(&b->ticket, t, t + 1ull)) {
      out->lane = l; out->op = (int)o; out->tile = (int)k;
      *op_out = cand;
      return true;
    }
  }
  return false;
}

// A claimed tile is complete (its stores are performed and fenced by the caller).  The completer of the LAST tile of an
// op publishes the lane's next op - or the finished state.
template <class OpT>
STEAL_HD void steal_complete(LaneBoardT<OpT>* b, int op_index, unsigned n_tiles) {
  sb_fence();                                    // release: this tile's stores before the counter
  const unsigned prev = sb_add(&b->done, 1u);
  if (prev + 1u == n_tiles) {
    sb_fence();                                  // acquire side of the counter's release sequence
    sb_st_relaxed(&b->done, 0u);
    const unsigned e = sb_ld_relaxed(&b->target_epoch);
    const int lane_ops = sb_ld_relaxed(&b->n_ops);
    const unsigned next = (op_index + 1 >= lane_ops) ? STEAL_OP_DONE : (unsigned)op_index + 1u;
    sb_fence();
    sb_st_release(&b->ticket, pack_ticket(e, next, 0));
  }
}

}  // namespace defer
