// Flag-in-data transport between row slabs over peer memory (NVLink): every double travels as ONE 16-byte store
// (lo, seq, hi, seq) -- each 8-byte half carries its own copy of the sequence number, the protocol of NCCL's LL
// primitives -- so a transfer needs no fence, no separate flag and no ticket: the receiver polls the slot until both
// halves show the expected number.  Slots are double-buffered by the parity of the sequence number; a stale slot holds
// an OLDER number, so equality is the only test.  Measured on 8 B200s: an exchange of interface rows costs 4.5-6 us
// against 13-19 us for "payload, fence.sys, release-store of a flag / acquire-poll, read payload".
#pragma once
#include "topo_internal.cuh"

namespace {

__device__ __forceinline__ void ll_store(uint4 *dst, double v, unsigned seq) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(seq), "r"(hi), "r"(seq) : "memory");
}
// bounded spin (~ seconds): a missing peer must not hang the GPU
__device__ __forceinline__ bool ll_load(const uint4 *src, unsigned seq, double &v, unsigned int *error) {
  for (long long it = 0; it < (1ll << 24); ++it) {
    unsigned a, b, c, d;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(src) : "memory");
    if (b == seq && d == seq) {
      v = __hiloint2double((int)c, (int)a);
      return true;
    }
    if (it > 2000) {
      if (*reinterpret_cast<volatile unsigned int *>(error)) return false;   // somebody already gave up
      __nanosleep(100);
    }
  }
  *error = 1u;
  return false;
}

// What a kernel needs to swap nodes of its first / last node row with the neighbouring slabs.  c == nullptr: no exchange
// inside this kernel (single GPU, or the host runs separate exchange kernels).
struct RowXchg {
  char *region, *peer_lo, *peer_hi;        // peer_*: nullptr = no neighbour on that side
  size_t off_lo, off_hi, stride;           // row slots "from below" / "from above": [parity][n] x 2 doubles x 16 B
  P2PCounters *c;
};
__device__ __forceinline__ unsigned long long ll_rows_seq(const RowXchg &x) { return x.c->halo_push + 1; }
// node i of my first (lo) / last row: push my value into the neighbour's slot, return the neighbour's value of the node
__device__ __forceinline__ bool ll_row_swap(const RowXchg &x, bool lo, int i, unsigned long long seq, double2 mine, double2 &theirs) {
  const unsigned s32 = (unsigned)seq;
  const size_t buf = (size_t)(seq & 1ull) * x.stride;
  uint4 *dst = reinterpret_cast<uint4 *>((lo ? x.peer_lo : x.peer_hi) + (lo ? x.off_hi : x.off_lo) + buf) + 2 * (size_t)i;
  ll_store(dst, mine.x, s32);
  ll_store(dst + 1, mine.y, s32);
  const uint4 *rcv = reinterpret_cast<const uint4 *>(x.region + (lo ? x.off_lo : x.off_hi) + buf) + 2 * (size_t)i;
  return ll_load(rcv, s32, theirs.x, &x.c->error) && ll_load(rcv + 1, s32, theirs.y, &x.c->error);
}
// thread 0 of EVERY block of the kernel, after a block barrier: the last block advances the sequence number
__device__ __forceinline__ void ll_rows_done(const RowXchg &x, unsigned long long seq, unsigned total_blocks) {
  __threadfence();
  if (atomicAdd(&x.c->fticket[2], 1u) == total_blocks - 1u) {
    x.c->fticket[2] = 0u;
    x.c->halo_push = seq;
    x.c->halo_wait = seq;
  }
}

}  // namespace

struct PeerTable {
  char *p[TOPO_P2P_MAXW];                  // mapped regions by rank (own region at [rank])
};

RowXchg topo_dist_row_xchg(const topo_handle *h);
PeerTable topo_dist_peer_table(const topo_handle *h);   // dist.cu: the handle's slots if in-kernel exchanges are on, else zeros
