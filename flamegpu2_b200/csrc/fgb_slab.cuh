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
// Consecutive calls alternate the parity slot, so a rank can never overwrite a value its neighbour has not consumed
// (it needs that neighbour's next contribution before it can get two calls ahead).
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
                                 unsigned long long *d_epoch, unsigned int *d_err, unsigned long long timeout_ns) {
  __shared__ A s_val[kSlabMaxWorld];
  __shared__ int s_ok;
  const int r = threadIdx.x;
  // the epoch lives in a device word that every call advances by one (identically on every rank): the launch carries no
  // host-side counter and can be replayed from a CUDA graph
  const unsigned long long epoch = *d_epoch + 1ull;
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
  if (r == 0) *d_epoch = epoch;
  if (r == 0 && s_ok) {
    A acc = s_val[0];
    for (int i = 1; i < world; ++i) acc = slab_combine(acc, s_val[i], op);
    *value_inout = acc;
  }
}

#endif  // __CUDACC__
}  // namespace fgb

// ------------------------------------------------------------------------------------------------------------------
// Migration without rewriting the list.  Only a few agents per step leave a slab (those that crossed one of its two
// boundary planes), so instead of a stable compaction of EVERY variable of EVERY agent (read + write of the whole list)
//   k_slab_select : one read of the position along the decomposed axis; the indices of the leavers are collected in
//                   two small lists (warp-aggregated atomics; order is arrival order)
//   k_slab_pack   : gathers the leavers' variables straight into the neighbours' staging buffers (peer memory) + counts
//   k_slab_holes  : ONE block pairs every hole below the new list end with a surviving agent of the tail
//   k_slab_fill   : moves those tail agents into the holes; the list's count becomes n - leavers
// The list stays dense; the relative order of the agents that stay changes only for the few moved tail agents (the
// order of a state list is observable, but a slab's list has no single-GPU counterpart to match; the order inside a
// PBM bin is unspecified in the reference).
// ------------------------------------------------------------------------------------------------------------------
namespace fgb {
#ifdef __CUDACC__

struct SlabSel {           // device-resident control block of one exchange (zeroed by k_slab_fill for the next one)
  unsigned int cnt[2];     // leavers towards rank-1 / rank+1 (may exceed the list capacity: overflow is reported)
  unsigned int pairs;      // holes below the new end == tail agents to move
  unsigned int new_n;
};

__global__ void __launch_bounds__(256) k_slab_select(const float *__restrict__ p, uint32_t n_max, const unsigned int *d_n, float mn, float radius,
                                                     int dim, int lo, int hi, uint32_t cap, SlabSel *sel, uint32_t *idx_lo, uint32_t *idx_hi) {
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  int side = -1;
  if (i < n) {
    int c = static_cast<int>(floorf(__fdiv_rn(__ldg(p + i) - mn, radius)));
    c = c < 0 ? 0 : (c >= dim ? dim - 1 : c);
    side = c < lo ? 0 : (c >= hi ? 1 : -1);
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const unsigned int m = __ballot_sync(0xFFFFFFFFu, side == s);
    if (!m) continue;
    const unsigned int lane = threadIdx.x & 31u;
    unsigned int base = 0;
    if (lane == static_cast<unsigned int>(__ffs(m) - 1)) base = atomicAdd(&sel->cnt[s], static_cast<unsigned int>(__popc(m)));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
    if (side == s) {
      const unsigned int at = base + __popc(m & ((1u << lane) - 1u));
      if (at < cap) (s == 0 ? idx_lo : idx_hi)[at] = i;
    }
  }
}

// blockIdx.y = side.  vt_lo / vt_hi: in = the list's columns, out = the neighbour's staging columns (NULL table n == 0:
// no neighbour on that side; its leavers are still removed -- they left the global domain's decomposition range never
// happens with clamped planes, so the lists are empty there).
// reset_done != NULL (halo pack: nothing is removed, so no k_slab_fill follows): the last block to finish clears the
// control block for the next exchange.
__global__ void __launch_bounds__(256) k_slab_pack(SlabSel *sel, const uint32_t *__restrict__ idx_lo, const uint32_t *__restrict__ idx_hi,
                                                   uint32_t cap, const __grid_constant__ VarTable vt_lo, const __grid_constant__ VarTable vt_hi,
                                                   unsigned int *peer_count_lo, unsigned int *peer_count_hi, unsigned int *reset_done) {
  const int s = blockIdx.y;
  const VarTable &vt = s == 0 ? vt_lo : vt_hi;
  const uint32_t *idx = s == 0 ? idx_lo : idx_hi;
  unsigned int *peer_count = s == 0 ? peer_count_lo : peer_count_hi;
  const unsigned int total = sel->cnt[s];
  if (blockIdx.x == 0 && threadIdx.x == 0 && peer_count) *peer_count = total;  // the receiver flags total > cap as an overflow
  const unsigned int m = vt.n == 0 ? 0u : (total < cap ? total : cap);
  for (uint32_t j = blockIdx.x * 256 + threadIdx.x; j < m; j += gridDim.x * 256) {
    const uint32_t src = idx[j];
    for (uint32_t v = 0; v < vt.n; ++v) copy_item(vt, v, src, j);
  }
  if (reset_done) {
    __shared__ unsigned int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(reset_done, 1u) == gridDim.x * gridDim.y - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
      sel->cnt[0] = sel->cnt[1] = 0u;
      *reset_done = 0u;
    }
  }
}

// One block.  holes = idx_lo[0..c0) U idx_hi[0..c1); new_n = n - (c0 + c1).  to[k] = k-th hole below new_n,
// from[k] = k-th surviving index in [new_n, n).  Tail membership is a bitmap in shared memory (tail length == #holes).
constexpr int kSlabHoleBits = 1 << 18;  // up to 262144 leavers per step and rank (32 KB bitmap)
__global__ void __launch_bounds__(1024) k_slab_holes(SlabSel *sel, const uint32_t *__restrict__ idx_lo, const uint32_t *__restrict__ idx_hi,
                                                     uint32_t cap, uint32_t n_max, const unsigned int *d_n, uint32_t *to, uint32_t *from,
                                                     unsigned int *d_err) {
  __shared__ uint32_t s_bits[kSlabHoleBits / 32];
  __shared__ uint32_t s_scan[33];
  __shared__ uint32_t s_cursor;
  const uint32_t n = load_count_coherent(d_n, n_max);
  uint32_t c0 = sel->cnt[0], c1 = sel->cnt[1];
  if (c0 > cap || c1 > cap) {
    if (threadIdx.x == 0) atomicOr(d_err, kSlabErrOverflow);
    c0 = c0 < cap ? c0 : cap;
    c1 = c1 < cap ? c1 : cap;
  }
  uint32_t h = c0 + c1;
  if (h > static_cast<uint32_t>(kSlabHoleBits)) {  // cannot happen with capacities below the bitmap size; keep memory safe regardless
    if (threadIdx.x == 0) atomicOr(d_err, kSlabErrOverflow);
    h = kSlabHoleBits;
    c1 = h > c0 ? h - c0 : 0u;
    c0 = h - c1;
  }
  const uint32_t new_n = n - h;
  for (uint32_t w = threadIdx.x; w < (h + 31u) / 32u; w += 1024) s_bits[w] = 0u;
  if (threadIdx.x == 0) s_cursor = 0u;
  __syncthreads();
  // holes below the new end need filling; holes in the tail are simply cut off (marked in the bitmap)
  for (uint32_t k = threadIdx.x; k < h; k += 1024) {
    const uint32_t i = k < c0 ? idx_lo[k] : idx_hi[k - c0];
    if (i >= new_n) atomicOr(&s_bits[(i - new_n) >> 5], 1u << ((i - new_n) & 31u));
    else to[atomicAdd(&s_cursor, 1u)] = i;
  }
  __syncthreads();
  const uint32_t pairs = s_cursor;
  // survivors of the tail, in index order: thread t owns bitmap words [t*W, (t+1)*W)
  const uint32_t words = (h + 31u) / 32u;
  const uint32_t W = (words + 1023u) / 1024u;
  uint32_t mine = 0;
  for (uint32_t w = threadIdx.x * W; w < (threadIdx.x + 1) * W && w < words; ++w) {
    const uint32_t valid = (w + 1u) * 32u <= h ? 0xFFFFFFFFu : ((1u << (h - w * 32u)) - 1u);
    mine += __popc(~s_bits[w] & valid);
  }
  uint32_t total;
  uint32_t at = block_exclusive_scan(mine, s_scan, &total);
  for (uint32_t w = threadIdx.x * W; w < (threadIdx.x + 1) * W && w < words; ++w) {
    const uint32_t valid = (w + 1u) * 32u <= h ? 0xFFFFFFFFu : ((1u << (h - w * 32u)) - 1u);
    uint32_t live = ~s_bits[w] & valid;
    while (live) {
      const int b = __ffs(static_cast<int>(live)) - 1;
      live &= live - 1u;
      from[at++] = new_n + w * 32u + static_cast<uint32_t>(b);
    }
  }
  if (threadIdx.x == 0) {
    sel->pairs = pairs;  // == total by construction
    sel->new_n = new_n;
  }
}

// moves from[k] -> to[k] for every variable, publishes the new count and clears the control block for the next exchange
__global__ void __launch_bounds__(256) k_slab_fill(SlabSel *sel, const uint32_t *__restrict__ to, const uint32_t *__restrict__ from,
                                                   const __grid_constant__ VarTable vt, unsigned int *d_n, unsigned int *done) {
  __shared__ unsigned int s_last;
  const unsigned int pairs = sel->pairs;
  for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < pairs; k += gridDim.x * 256) {
    const uint32_t a = to[k], b = from[k];
    for (uint32_t v = 0; v < vt.n; ++v) copy_item(vt, v, b, a);  // in == out == the list's columns
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(done, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    *d_n = sel->new_n;
    sel->cnt[0] = sel->cnt[1] = sel->pairs = 0u;
    *done = 0u;
  }
}

#endif  // __CUDACC__
}  // namespace fgb
