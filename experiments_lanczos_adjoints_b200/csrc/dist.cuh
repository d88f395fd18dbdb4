// Peer-memory communicator of the row-sharded path (one process per GPU of one NVSwitch box).
//
// Every rank owns a small "mailbox" in its own HBM, mapped into the address space of every other
// rank (CUDA IPC).  A reduction of the Krylov loop is ONE single-block kernel per rank: push the
// local partial sums into the peers' mailboxes with plain stores over NVLink, publish a sequence
// number (st.release.sys), spin on the own mailbox until every peer's number has arrived
// (ld.acquire.sys), add the `world` contributions in rank order (bit-identical on every rank)
// and run the epilogue that turns the sums into the coefficients of the next kernel -- what the
// NCCL route does with an all-reduce launch plus an epilogue launch.  The wave stencil's halo rows
// travel the same way.  Two parity buffers per slot make the protocol safe without
// acknowledgements: a rank can only start exchange e+2 (same parity as e) after it finished e+1,
// which needed the peer's e+1 flag, which the peer published after it had consumed e.
#pragma once

#include "common.cuh"

namespace bl {
namespace dist {

constexpr int kMaxRanks = 8;
constexpr int kRedSlots = 16384;           // doubles per reduction (K*K of the dense-cotangent set-up, K <= 128)
constexpr size_t kHaloRowBytes = 512 * 1024;  // one halo row (<= 65536 doubles)
constexpr int kHaloSlots = 4;              // field 0 from above / below, field 1 from above / below
constexpr size_t kFlagBytes = 1024;        // [0..8) reduce flags by source, [8..16) halo, [16..32) gather slots 0/1, [64] error
constexpr size_t kRedOff = kFlagBytes;
constexpr size_t kHaloOff = kRedOff + 2ull * kMaxRanks * kRedSlots * 8;
constexpr size_t kMailboxBytes = kHaloOff + 2ull * kHaloSlots * kHaloRowBytes;
constexpr long long kSpinTimeout = 20000000000ll;  // ~10 s of SM clocks: a dead peer must not hang the GPU

struct PeerView {
  int rank = 0, world = 1;
  unsigned long long seq = 0;
  unsigned char* mail[kMaxRanks] = {};  // mailbox of every rank in THIS rank's address space
};

__device__ __forceinline__ unsigned long long* red_flag(unsigned char* mail, int src) {
  return reinterpret_cast<unsigned long long*>(mail) + src;
}
__device__ __forceinline__ unsigned long long* halo_flag(unsigned char* mail, int src) {
  return reinterpret_cast<unsigned long long*>(mail) + 8 + src;
}
__device__ __forceinline__ unsigned long long* gather_flag(unsigned char* mail, int slot, int src) {
  return reinterpret_cast<unsigned long long*>(mail) + 16 + slot * 8 + src;
}
// header of the own mailbox (written at connect time): rank, world and the mapped mailbox of every rank, so
// that kernels which only carry the own mailbox pointer and a sequence number can rebuild the PeerView
constexpr size_t kHeaderOff = 768;
struct MailHeader {
  int rank, world;
  unsigned char* mail[kMaxRanks];
};
__device__ __forceinline__ unsigned long long* error_flag(unsigned char* mail) {
  return reinterpret_cast<unsigned long long*>(mail) + 64;
}
__host__ __device__ __forceinline__ double* red_slot(unsigned char* mail, int parity, int src) {
  return reinterpret_cast<double*>(mail + kRedOff) + ((size_t)parity * kMaxRanks + src) * kRedSlots;
}
__host__ __device__ __forceinline__ unsigned char* halo_slot(unsigned char* mail, int parity, int slot) {
  return mail + kHaloOff + ((size_t)parity * kHaloSlots + slot) * kHaloRowBytes;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// spin until *flag >= seq; false after the timeout (and the mailbox's error word is set)
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long seq,
                                          unsigned char* own_mail) {
  // once a wait has timed out the communicator is dead: later exchanges fail at once instead of
  // stalling ~10 s each
  volatile unsigned long long* err = error_flag(own_mail);
  if (*err != 0ull) return ld_acquire_sys(flag) >= seq;
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < seq) {
    if (clock64() - t0 > kSpinTimeout) {
      *err = 1ull;
      return false;
    }
  }
  return true;
}

// Block-wide: red[0..count) <- sum over ranks of red[0..count), same bits on every rank.
// Call with all threads of a single block (blockDim.x >= world); ends with a __syncthreads().
__device__ __forceinline__ void peer_allreduce_block(const PeerView& pv, double* red, int count) {
  if (pv.world <= 1) {
    __syncthreads();
    return;
  }
  const int t = threadIdx.x, nt = blockDim.x, parity = (int)(pv.seq & 1ull);
  for (int j = t; j < count; j += nt) {
    const double v = red[j];
    for (int p = 0; p < pv.world; ++p)
      if (p != pv.rank) red_slot(pv.mail[p], parity, pv.rank)[j] = v;  // NVLink store
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int ok;
  if (t == 0) ok = 1;
  __syncthreads();
  if (t < pv.world && t != pv.rank) {
    st_release_sys(red_flag(pv.mail[t], pv.rank), pv.seq);
    if (!wait_flag(red_flag(pv.mail[pv.rank], t), pv.seq, pv.mail[pv.rank])) ok = 0;
  }
  __syncthreads();
  const double poison = ok ? 0.0 : __longlong_as_double(0x7ff8000000000000ll);
  for (int j = t; j < count; j += nt) {
    double s = poison;
    for (int r = 0; r < pv.world; ++r)
      s += r == pv.rank ? red[j] : ld_volatile(red_slot(pv.mail[pv.rank], parity, r) + j);
    red[j] = s;
  }
  __syncthreads();
}

// All-gather window (optional, bl_dist_comm_window_*): every rank owns [2 parities][2 slots][slot_bytes] in
// its HBM, mapped by the peers; a rank pushes its chunk of a vector into every peer's window.
struct WindowView {
  size_t slot_bytes = 0;
  unsigned char* win[kMaxRanks] = {};
  __host__ __device__ __forceinline__ unsigned char* slot(int rank, int parity, int slot_index) const {
    return win[rank] + ((size_t)parity * 2 + slot_index) * slot_bytes;
  }
};

__device__ __forceinline__ PeerView view_from_mailbox(unsigned char* own_mail, unsigned long long seq) {
  const MailHeader* h = reinterpret_cast<const MailHeader*>(own_mail + kHeaderOff);
  PeerView pv;
  pv.rank = h->rank;
  pv.world = h->world;
  pv.seq = seq;
#pragma unroll
  for (int p = 0; p < kMaxRanks; ++p) pv.mail[p] = h->mail[p];
  return pv;
}

// host side (dist.cu)
bool active();
int world();
int next_reduce(PeerView* pv);  // the active communicator's view with the next reduction number
int next_halo(PeerView* pv);    // ... next halo-exchange number

}  // namespace dist
}  // namespace bl

struct bl_comm;  // opaque (include/b200_lanczos.h: bl_comm_t)
namespace bl {
namespace dist {
int view_of(bl_comm* comm, bool halo, PeerView* pv);  // advances the sequence number of `comm`
int gather_view_of(bl_comm* comm, PeerView* pv, WindowView* wv);  // advances the gather sequence number
}
}  // namespace bl
