// flamegpu/runtime/detail/FunctionArgs.h -- what an agent-function kernel receives.
//
// One POD passed BY VALUE (kernel parameter space == constant bank) instead of the reference's
// 12 separate arguments plus a 8-12 KB cuRVE hash table copied into shared memory by every block
// (runtime/AgentFunction.cuh:40-76, runtime/detail/SharedBlock.h:15-31) and a 12 KB host-to-device
// table upload before every launch (runtime/detail/curve/HostCurve.cu:159-163).
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_DETAIL_FUNCTIONARGS_H_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_DETAIL_FUNCTIONARGS_H_

#include <cstdint>

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
  // b200 iterator modes (0 = reference order, every message of the Moore neighbourhood in strip order):
  //  1 = radius-first: the same messages, each exactly once, but those within the radius of the search
  //      origin first (in strip order), then the rest (in strip order);
  //  2 = radius-only: only messages within the radius (a superset of what `sqrtf(d2) < radius` accepts).
  int iter_mode;
  float radius2_eps;   // radius^2 * (1 + 1e-5): conservative in-radius test of the iterator
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
  const unsigned int *d_agent_offset;  // function condition: agents [0, *d_agent_offset) are disabled (NULL: 0)
  unsigned int bound;           // launch bound
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
  const unsigned int *d_msg_in_count;   // brute-force style inputs
  const unsigned int *d_msg_out_offset; // append offset of the output list (NULL -> 0)
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
  return l;
}

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
