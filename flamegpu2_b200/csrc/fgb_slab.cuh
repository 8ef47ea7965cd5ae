// fgb_slab.cuh -- device side of the multi-GPU z-slab exchange (no reference counterpart: a
// reference simulation never spans GPUs).  One process per GPU.  Halo messages and migrating agents are packed
// by the ordinary compaction kernel (fgb_compact_limited) whose OUTPUT pointers address the neighbour's
// staging buffer directly (peer memory over NVLink, mapped with cudaIpcOpenMemHandle), so packing IS the transfer.
// What this file adds is the synchronisation around it:
//   k_slab_signal : after the pack kernels of one exchange, publish "epoch e is complete" in the neighbour's flag word
//   k_slab_wait   : spin (one thread) until the neighbours' epoch-e flags arrive, check the received counts
//   k_slab_allreduce : 8-byte all-reduce over per-rank mailboxes (HostAgentAPI reductions under slabs)
// Epochs increase by one per simulation step and live in a device word, so the kernels are CUDA-graph replayable.
// A wait gives up after `timeout_ns` and raises an error word instead of hanging the GPU (a dead peer must not
// wedge the box).
#pragma once
#include "fgb_common.cuh"

namespace fgb {

#ifdef __CUDACC__

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

enum : unsigned int { kSlabErrTimeout = 1u, kSlabErrOverflow = 2u, kSlabErrBound = 4u };

// flags[s] (peer memory, NULL = no neighbour on that side) <- *d_epoch + 1.  The data the flag covers was written by
// earlier kernels of this stream; the system-scope fence + release orders it before the flag for the peer's acquire.
__global__ void k_slab_signal(unsigned long long *flag_lo, unsigned long long *flag_hi, const unsigned int *d_epoch) {
  __threadfence_system();
  const unsigned long long e = static_cast<unsigned long long>(*d_epoch) + 1ull;
  if (threadIdx.x == 0 && flag_lo) st_release_sys_u64(flag_lo, e);
  if (threadIdx.x == 1 && flag_hi) st_release_sys_u64(flag_hi, e);
}

// Wait until the local flag words (written by the neighbours) reach *d_epoch + 1; then validate the received counts
// against the staging capacity (the sender clamps its writes, the count reports every selected item).
__global__ void k_slab_wait(const unsigned long long *flag_lo, const unsigned long long *flag_hi, const unsigned int *count_lo,
                            const unsigned int *count_hi, unsigned int capacity, const unsigned int *d_epoch,
                            unsigned int *d_err, unsigned long long timeout_ns) {
  const unsigned long long *flag = threadIdx.x == 0 ? flag_lo : flag_hi;
  const unsigned int *count = threadIdx.x == 0 ? count_lo : count_hi;
  if (threadIdx.x > 1 || !flag) return;
  const unsigned long long e = static_cast<unsigned long long>(*d_epoch) + 1ull;
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys_u64(flag) < e) {
    if (global_timer_ns() - t0 > timeout_ns) {
      atomicOr(d_err, kSlabErrTimeout);
      return;
    }
    __nanosleep(100);
  }
  if (*reinterpret_cast<const volatile unsigned int *>(count) > capacity) atomicOr(d_err, kSlabErrOverflow);
}

// *d_count > bound: agents beyond a list's launch bound would silently stop executing -> raise the error word.
__global__ void k_slab_check_bound(const unsigned int *d_count, unsigned int bound, unsigned int *d_err) {
  if (*d_count > bound) atomicOr(d_err, kSlabErrBound);
}

// All-reduce of one 8-byte value.  mailbox[r] is rank r's mailbox array (peer memory for r != rank):
// 2 (epoch parity) x world slots of {value bits, epoch}.  Lane r < world writes this rank's value into rank r's
// slot [parity][rank], then waits for rank r's value in the local slot [parity][r]; lane 0 folds the world values
// in RANK ORDER (every rank computes the identical result).  op: 0 sum, 1 min, 2 max; dtype as fgb_dtype.
struct SlabMailSlot {
  unsigned long long value;
  unsigned long long epoch;
};
constexpr int kSlabMaxWorld = 32;
struct SlabMailboxes {
  SlabMailSlot *box[kSlabMaxWorld];
};

template <typename A>
__device__ __forceinline__ A slab_combine(A a, A b, int op) {
  return op == 0 ? a + b : (op == 1 ? (b < a ? b : a) : (b > a ? b : a));
}

template <typename A>
__global__ void k_slab_allreduce(A *value_inout, const __grid_constant__ SlabMailboxes mb, int rank, int world, int op,
                                 unsigned long long epoch, unsigned int *d_err, unsigned long long timeout_ns) {
  __shared__ A s_val[kSlabMaxWorld];
  __shared__ int s_ok;
  const int r = threadIdx.x;
  const int par = static_cast<int>(epoch & 1ull);
  if (r == 0) s_ok = 1;
  __syncwarp();
  if (r < world) {
    const A mine = *value_inout;
    unsigned long long bits = 0;
    memcpy(&bits, &mine, sizeof(A));
    SlabMailSlot *dst = mb.box[r] + par * world + rank;
    *reinterpret_cast<volatile unsigned long long *>(&dst->value) = bits;
    __threadfence_system();
    st_release_sys_u64(&dst->epoch, epoch);
    const SlabMailSlot *src = mb.box[rank] + par * world + r;
    const unsigned long long t0 = global_timer_ns();
    bool ok = true;
    while (ld_acquire_sys_u64(&src->epoch) != epoch) {
      if (global_timer_ns() - t0 > timeout_ns) {
        ok = false;
        break;
      }
      __nanosleep(100);
    }
    if (!ok) {
      atomicOr(d_err, kSlabErrTimeout);
      s_ok = 0;
    }
    const unsigned long long got = *reinterpret_cast<const volatile unsigned long long *>(&src->value);
    A v;
    memcpy(&v, &got, sizeof(A));
    s_val[r] = v;
  }
  __syncwarp();
  if (r == 0 && s_ok) {
    A acc = s_val[0];
    for (int i = 1; i < world; ++i) acc = slab_combine(acc, s_val[i], op);
    *value_inout = acc;
  }
}

#endif  // __CUDACC__
}  // namespace fgb
