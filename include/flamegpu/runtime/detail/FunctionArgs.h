// flamegpu/runtime/detail/FunctionArgs.h -- what an agent-function kernel receives.
//
// One POD passed BY VALUE (kernel parameter space == constant bank) instead of the reference's
// 12 separate arguments plus a 8-12 KB cuRVE hash table copied into shared memory by every block
// (runtime/AgentFunction.cuh:40-76, runtime/detail/SharedBlock.h:15-31) and a 12 KB host-to-device
// table upload before every launch (runtime/detail/curve/HostCurve.cu:159-163).
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_DETAIL_FUNCTIONARGS_H_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_DETAIL_FUNCTIONARGS_H_

#include <cstdint>
#include <limits>
#include <type_traits>

#include "flamegpu/defines.h"
#include "flamegpu/detail/hash.h"

namespace flamegpu {
namespace detail {

// name-hash -> SoA base pointer table of one list (agent state list, message list, new-agent list).
// A per-table PERFECT hash: the host picks `salt` so that (hash * salt) >> (32 - kSlotBits) is
// collision free for the table's names, so a lookup is one multiply-shift and one indexed
// constant-bank load, with no loop (and is hoisted out of message loops as loop-invariant code).
constexpr int kSlotBits = 6;
constexpr int kSlots = 1 << kSlotBits;
struct DevVars {
  uint32_t n;
  uint32_t salt;
  uint32_t hash[kSlots];  // 0 == empty slot
  char *ptr[kSlots];
};

// By-value copy of MessageSpatial2D/3D::MetaData (reference MessageSpatial3D.h:38-68); the
// reference passes a pointer and re-reads it from global memory in every iterator step.
struct SpatialMeta {
  float min[3];
  float max[3];
  float radius;
  float env_width[3];
  int grid_dim[3];      // global grid (clamp of getGridPosition)
  int win_begin;        // slab window on the slowest axis: first plane stored locally
  int win_count;        // ... and how many (== grid_dim[slowest] on a single GPU)
  int wrap_compatible;
  // b200 iterator modes: 0 = reference order, every message of the Moore neighbourhood in strip order;
  //  1 = radius-filtered: only messages within the radius of the search origin (a superset of what
  //      `sqrtf(d2) < radius` accepts), in the same relative order; the lanes of a warp walk in lock-step
  //      and queue their accepted messages in shared memory (MessageSpatial3D.cuh advance_filtered).
  int iter_mode;
  float radius2_eps;   // radius^2 * (1 + 1e-5): conservative in-radius test of the iterator
  unsigned int pad_index;  // index of the list's padding message (one past its capacity; location far outside any environment)
  const unsigned int *pbm;
};

struct DevEnv {
  uint32_t n;
  uint32_t salt;
  uint32_t hash[kSlots];
  uint32_t offset[kSlots];
  const char *buffer;  // device copy of the packed property values
};

struct FunctionArgs {
  const unsigned int *d_count;  // device-resident agent count (<= bound)
  const unsigned int *exec_perm; // thread t runs agent exec_perm[t] (bin order); NULL: agent t
  const unsigned int *d_perm_limit;  // non-NULL: exec_perm is valid for threads below this device word only, the rest run agent t
  const unsigned int *d_agent_offset;  // function condition: agents [0, *d_agent_offset) are disabled (NULL: 0)
  unsigned int bound;           // launch bound
  unsigned int first_thread;    // this launch covers threads [first_thread, last_thread) of the function (a function may be
  unsigned int last_thread;     // launched in several thread-range chunks, CUDASimulation::streamPopulationDataSoA); 0,0 = all
  DevVars agent;                // variables of the executing agent's state list
  DevVars msg_in;               // input message list (bin-sorted for spatial messages)
  DevVars msg_out;              // output message list
  DevVars agent_out;            // new-agent scratch list (one slot per parent thread)
  // new-agent variables in declaration order, for completing a child with its defaults
  uint32_t agent_out_nvars;
  uint32_t agent_out_slot[b200::kMaxVars];      // slot of variable v in agent_out
  uint32_t agent_out_len[b200::kMaxVars];       // bytes per item
  const char *agent_out_defaults[b200::kMaxVars];  // device pointer to the default value
  SpatialMeta in_meta;
  // Fused index build (mandatory spatial output into an emptied list): the kernel publishes the bin key of every
  // message it writes and counts it in the list's histogram, so buildIndex starts at the scan and never re-reads the
  // locations (fgb_spatial_writer_args / fgb_build_index_ex in flamegpu2_b200.h).  NULL: off.
  unsigned int *out_keys;
  unsigned int *out_hist;
  unsigned int *d_keyed;         // receives the number of messages keyed here (== the list's new count)
  float out_min[3];
  float out_radius;
  int out_grid_dim[3];
  int out_win_begin, out_win_count;
  unsigned int slot_by_thread;   // 1: message / new-agent slot = thread index (bin-ordered output), 0: = agent index
  const unsigned int *d_msg_in_count;   // brute-force style inputs
  const unsigned int *d_msg_out_offset; // append offset of the output list (NULL -> 0)
  unsigned int *d_msg_out_count;        // non-NULL: mandatory output into a truncated list, the kernel itself publishes
                                        // the list's new count (= number of executing agents)
  unsigned int *death_flag;      // scan flags, one per thread (NULL when the feature is off)
  unsigned int *msg_out_flag;
  unsigned int *agent_out_flag;
  id_t *next_id;                 // per-agent-type id counter for births (atomic)
  const unsigned int *d_step;    // device-resident step counter
  DevEnv env;
};

// Slot of a hashed name in a table.  Like the reference, an unknown name is only diagnosed when
// FLAMEGPU_SEATBELTS is on (reference DeviceCurve.cuh:357-371); with seatbelts off (the benchmark
// configuration of both code bases) the lookup is a multiply-shift and nothing else.
#ifndef FLAMEGPU_SEATBELTS
#define FLAMEGPU_SEATBELTS 0
#endif
FGB_HD int find_slot(const DevVars &t, uint32_t h) {
  const uint32_t i = (h * t.salt) >> (32 - kSlotBits);
#if FLAMEGPU_SEATBELTS
  return t.hash[i] == h ? static_cast<int>(i) : -1;
#else
  return static_cast<int>(i);
#endif
}
FGB_HD int find_slot(const DevEnv &t, uint32_t h) {
  const uint32_t i = (h * t.salt) >> (32 - kSlotBits);
#if FLAMEGPU_SEATBELTS
  return t.hash[i] == h ? static_cast<int>(i) : -1;
#else
  return static_cast<int>(i);
#endif
}
// message location variables have fixed names: their base pointers are resolved once per thread
constexpr uint32_t kHashX = name_hash("x");
constexpr uint32_t kHashY = name_hash("y");
constexpr uint32_t kHashZ = name_hash("z");
struct LocPtrs {
  const char *x, *y, *z;
};
FGB_HD LocPtrs make_loc(const FunctionArgs &a) {
  LocPtrs l;
  l.x = a.msg_in.ptr[(kHashX * a.msg_in.salt) >> (32 - kSlotBits)];
  l.y = a.msg_in.ptr[(kHashY * a.msg_in.salt) >> (32 - kSlotBits)];
  l.z = a.msg_in.ptr[(kHashZ * a.msg_in.salt) >> (32 - kSlotBits)];  // 2D lists: an empty slot, never read
#if defined(__CUDA_ARCH__)
  // keep the three pointers in registers: without this the compiler re-derives them (hash * salt, shift, indexed
  // constant load) inside the message loop, ~10 instructions per message
  asm("" : "+l"(l.x), "+l"(l.y), "+l"(l.z));
#endif
  return l;
}

// Radius-filtered iterator (MessageSpatial2D/3D::In::Filter, ITER_MODE 1; StripWalk.cuh).  Each lane walks its strips
// in chunks of up to 32 consecutive messages and keeps, per chunk with at least one message within the radius, the
// pair {first message index - 1, accepted bit mask} in dynamic shared memory: entry k of thread t is the 8-byte word
// [k * blockDim.x + t] (conflict free).  The scheduler launches with kFilterQueueWords * blockDim.x words.
// 8 chunks = 64 B of shared memory per thread.  Shared memory is carved out of the L1: with 16 chunks per lane the
// L1 hit rate of the message loads fell from 74 % to 9 % and the walk became latency bound (stress model 2x
// slower than the unfiltered iterator); with 4 the queues fill too often.  Measured: 16 / 8 / 4 chunks ->
// Circles 16.8 M move 9.2 / 8.0 / 9.9 ms, stress 4 M update 1.90 / 0.96 / 1.09 ms.
#ifndef FGB_FILTER_CHUNKS
#define FGB_FILTER_CHUNKS 10
#endif
constexpr unsigned int kFilterChunks = FGB_FILTER_CHUNKS;     // queued chunks per lane (each up to 32 accepted messages)
constexpr unsigned int kFilterQueueWords = 2 * kFilterChunks;
// location of the padding message a lane is shown while other lanes of its warp still have accepted messages:
// far outside any environment, finite so that distance arithmetic on it stays on the fast paths
template <typename T>
FGB_HD T pad_location() {
  if constexpr (std::is_floating_point<T>::value) {
    return static_cast<T>(1.0e18f);
  } else {
    return T{};
  }
}
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t *filter_queue() {
  extern __shared__ uint32_t fgb_filter_queue_smem[];
  return fgb_filter_queue_smem;
}
// packed fp32x2 arithmetic of sm_100 (two messages per instruction)
__device__ __forceinline__ unsigned long long f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long f32x2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f32x2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f32x2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
#endif

// host: choose a salt that makes the table collision free and fill hash[]; returns false if none found
template <typename Table>
inline bool build_perfect_table(Table &t, const uint32_t *hashes, uint32_t n, uint32_t *slot_of) {
  t.n = n;
  for (uint32_t salt = 0x9E3779B1u, tries = 0; tries < (1u << 22); ++tries, salt = salt * 0x01000193u + 0x7F4A7C15u) {
    const uint32_t s = salt | 1u;
    uint64_t used = 0;
    bool ok = true;
    for (uint32_t k = 0; k < n && ok; ++k) {
      const uint32_t i = (hashes[k] * s) >> (32 - kSlotBits);
      if (used & (1ull << i)) ok = false;
      used |= 1ull << i;
    }
    if (!ok) continue;
    t.salt = s;
    for (int i = 0; i < kSlots; ++i) t.hash[i] = 0u;
    for (uint32_t k = 0; k < n; ++k) {
      const uint32_t i = (hashes[k] * s) >> (32 - kSlotBits);
      t.hash[i] = hashes[k];
      slot_of[k] = i;
    }
    return true;
  }
  return false;
}

}  // namespace detail
}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_RUNTIME_DETAIL_FUNCTIONARGS_H_
