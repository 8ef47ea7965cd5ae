// flamegpu/simulation/CUDASimulation.h -- the per-step scheduler of the hot path.
//
// Replaces CUDASimulation::{step, stepLayer, spatialSortAgent_async} and the device-state managers
// they drive (reference src/flamegpu/simulation/CUDASimulation.cu:463-1109, CUDAAgent*.cu,
// CUDAFatAgent*.cu, CUDAMessage*.cu) for models built from agents + MessageSpatial2D/3D (and
// brute-force) lists.  The design is B200-first rather than a translation:
//
//  * every count (agents per state, messages per list, next ids, step counter) lives in a small
//    device control block; kernels are launched for a host-known upper BOUND and clamp to the
//    device count, so nothing returns to the host inside a step.  The reference reads a count back
//    and synchronises after every compaction (CUDAScatter.cu:175-178) - >= 6 host syncs per step;
//  * a whole step (all layers, the agent sorts, PBM builds, agent functions, death / birth /
//    optional-message compactions) is recorded once into a CUDA graph, with one captured stream
//    per concurrently running function of a layer, and replayed; graphs are cached per
//    (buffer-pointer parity, bounds) because double-buffered lists swap pointers every step;
//  * all data movement goes through the sm_100a library behind the C ABI (flamegpu2_b200.h).
//
// Order of operations inside a layer follows the reference exactly: auto agent
// sort -> PBM build for the input list -> agent function -> optional-message compaction ->
// death compaction -> state transition -> birth append.
#ifndef FGB_INCLUDE_FLAMEGPU_SIMULATION_CUDASIMULATION_H_
#define FGB_INCLUDE_FLAMEGPU_SIMULATION_CUDASIMULATION_H_

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "flamegpu/model/ModelDescription.h"
#include "flamegpu/simulation/AgentVector.h"
#include "flamegpu2_b200.h"

namespace flamegpu {

#define FGB_CUDA_THROW(expr)                                                                                     \
  do {                                                                                                           \
    cudaError_t _e = (expr);                                                                                     \
    if (_e != cudaSuccess)                                                                                       \
      throw flamegpu::exception::CUDAError(std::string(#expr) + " failed: " + cudaGetErrorString(_e));          \
  } while (0)
#define FGB_ABI_THROW(expr)                                                                                      \
  do {                                                                                                           \
    fgb_status _s = (expr);                                                                                      \
    if (_s != 0) throw flamegpu::exception::CUDAError(std::string(#expr) + " failed: " + fgb_error_string(_s)); \
  } while (0)

namespace detail {

#if defined(__CUDACC__)
// HostAgentAPI::reduce / transformReduce with user functors: grid-stride accumulation in thread order, block partials
// in shared memory folded by thread 0, the last block to arrive folds the partials in block order (reproducible)
constexpr unsigned int kUserRedBlocks = 592;  // 4 blocks per SM
template <typename InT, typename OutT, typename Transform, typename Reduce>
__global__ void __launch_bounds__(256) k_user_transform_reduce(const InT *in, unsigned int bound, const unsigned int *d_n, OutT init, OutT *partial,
                                                               unsigned int *done, OutT *out) {
  __shared__ OutT s_val[256];
  __shared__ unsigned int s_last;
  unsigned int n = bound;
  if (d_n) {
    const unsigned int c = *d_n;
    n = c < n ? c : n;
  }
  Transform tr;
  Reduce rd;
  OutT acc = init;
  bool any = false;
  for (unsigned int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const OutT v = tr(in[i]);
    acc = any ? rd(acc, v) : v;
    any = true;
  }
  // fold the threads that saw data, in thread order (init is applied once, at the very end)
  s_val[threadIdx.x] = acc;
  __shared__ unsigned char s_any[256];
  s_any[threadIdx.x] = any ? 1 : 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    bool have = false;
    OutT b = init;
    for (int t = 0; t < 256; ++t)
      if (s_any[t]) {
        b = have ? rd(b, s_val[t]) : s_val[t];
        have = true;
      }
    partial[blockIdx.x] = b;
    reinterpret_cast<unsigned int *>(partial + kUserRedBlocks)[blockIdx.x] = have ? 1u : 0u;
    __threadfence();
    s_last = atomicAdd(done, 1u) == gridDim.x - 1 ? 1u : 0u;
    __threadfence();
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    OutT all = init;
    const volatile unsigned int *have = reinterpret_cast<const volatile unsigned int *>(partial + kUserRedBlocks);
    for (unsigned int b = 0; b < gridDim.x; ++b)
      if (have[b]) {
        OutT p;
        const volatile unsigned char *src = reinterpret_cast<const volatile unsigned char *>(partial + b);
        unsigned char *dst = reinterpret_cast<unsigned char *>(&p);
        for (unsigned int k = 0; k < sizeof(OutT); ++k) dst[k] = src[k];
        all = rd(all, p);
      }
    *out = all;
    *done = 0u;
  }
}
// end-of-step bookkeeping on the device: ++step, reset the counts of non-persistent message lists
// (reference CUDASimulation.cu:619-625 does this on the host)
__global__ void k_end_of_step(unsigned int *ctrl, unsigned int step_slot, const unsigned int *zero_slots, unsigned int n_zero,
                              unsigned int epoch_slot) {
  if (threadIdx.x == 0) ctrl[step_slot] += 1u;
  if (threadIdx.x == 1 && epoch_slot) ctrl[epoch_slot] += 1u;  // slab exchange epoch (one halo + one migration per step)
  for (unsigned int i = threadIdx.x; i < n_zero; i += blockDim.x) ctrl[zero_slots[i]] = 0u;
}
__global__ void k_copy_word(unsigned int *dst, const unsigned int *src) { *dst = *src; }
// results of the reductions recorded behind a step, then (system fence) the step's epoch = the step counter AFTER the
// step, to mapped host memory: mapped[0] epoch, mapped[1 + k] result k
__global__ void k_publish_reductions(unsigned long long *mapped, const unsigned long long *vals, unsigned int count, const unsigned int *ctrl,
                                     unsigned int step_slot) {
  if (threadIdx.x < count) *reinterpret_cast<volatile unsigned long long *>(mapped + 1 + threadIdx.x) = vals[threadIdx.x];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(mapped) = ctrl[step_slot];
}
// limit of a partially valid tile-local permutation (DevList::perm_partial): mode 0 before the list changes
// (limit = count, or min(limit, count) if it was partial already), mode 1 afterwards (min with the new count, rounded
// down to whole 2048-item tiles)
__global__ void k_perm_limit(unsigned int *limit, const unsigned int *count, int was_partial, int after) {
  unsigned int l = (was_partial || after) ? (*limit < *count ? *limit : *count) : *count;
  if (after) l &= ~2047u;
  *limit = l;
}
// function condition helpers: words[active] = words[count] - words[failed]; death flags of the
// disabled front are forced to "alive"; move flags select the executing survivors for a state change
__global__ void k_sub_word(unsigned int *dst, const unsigned int *a, const unsigned int *b) { *dst = *a - *b; }
__global__ void k_fill_front(unsigned int *flags, const unsigned int *d_front, unsigned int bound, unsigned int value) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < bound && i < *d_front) flags[i] = value;
}
__global__ void k_move_flags(unsigned int *move, const unsigned int *death, const unsigned int *d_front, const unsigned int *d_n,
                             unsigned int bound) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bound || i >= *d_n) return;
  move[i] = (i >= *d_front && (!death || death[i] == 1u)) ? 1u : 0u;
}
__global__ void k_iota(unsigned int *dst, unsigned int n, unsigned int first) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = first + i;
}
#endif

// One SoA list on the device: an agent state list, a message list or a new-agent scratch list.
struct DevList {
  std::vector<std::string> names;  // alphabetical (std::map order), like the reference
  std::vector<Variable> meta;
  std::vector<char *> data, swap;
  unsigned int capacity = 0;
  unsigned int bound = 0;       // host-known upper bound of the device count
  unsigned int count_slot = 0;  // control-block slot of the device count
  unsigned int appended_this_step = 0;  // upper bound of what listAppend added since the last endStep
  bool double_buffered = true;
  // Spatial message lists keep a PADDING item at index `capacity` of every buffer (floating-point variables 1e18, others
  // zero): the radius-filtered iterator shows it to lanes that wait for the other lanes of their warp (StripWalk.cuh).
  // Nothing ever writes at or beyond `capacity`, so the item is written once per allocation.
  bool pad_slot = false;
  // the owning simulation's allocation generation, bumped whenever a buffer of this list moves: launches captured
  // into a CUDA graph hold raw pointers, CUDASimulation::step() drops its cached graphs when the generation moved
  unsigned long long *gen = nullptr;
  // Bin-order permutation of this list's items computed by its most recent spatial reader (thread t ran item
  // cached_perm[t]); valid while the list's membership and order are unchanged (order_version).  A mandatory spatial
  // OUTPUT function of the same list reuses it to write its messages bin-grouped (run_function step 3).
  unsigned long long order_version = 1, cached_perm_version = 0;
  const unsigned int *cached_perm = nullptr;
  void touch() { ++order_version; }
  bool perm_valid() const { return cached_perm != nullptr && cached_perm_version == order_version; }
  // Slab migration changes a few items of the list in place (holes filled from the tail, arrivals appended): the
  // tile-local permutation stays a bijection on every 2048-item tile below min(count before, count after), so it is
  // kept, valid up to the device word at control slot perm_limit_slot (threads beyond it run the identity).
  bool perm_partial = false;
  unsigned int perm_limit_slot = 0;
  bool cached_perm_global = false;  // the cached permutation orders the WHOLE list by bin (not inside 2048-item tiles)

  void init(const VariableMap &vars, bool dbl) {
    double_buffered = dbl;
    for (const auto &v : vars) {
      names.push_back(v.first);
      meta.push_back(v.second);
      data.push_back(nullptr);
      swap.push_back(nullptr);
    }
  }
  int index_of(const std::string &n) const {
    for (size_t i = 0; i < names.size(); ++i)
      if (names[i] == n) return static_cast<int>(i);
    return -1;
  }
  // grow to at least `need` items, keeping the first `keep` items of data[]
  void reserve(unsigned int need, unsigned int keep) {
    if (need <= capacity) return;
    if (gen) ++*gen;
    unsigned int cap = std::max(capacity, 256u);
    while (cap < need) cap = static_cast<unsigned int>(std::min<unsigned long long>(0xFFFFFFF0ull, static_cast<unsigned long long>(cap) * 5 / 4 + 64));
    cap = (cap + 63u) & ~63u;
    const size_t items = static_cast<size_t>(cap) + (pad_slot ? 8u : 0u);
    for (size_t v = 0; v < names.size(); ++v) {
      const size_t b = meta[v].bytes();
      char *nd = nullptr;
      FGB_CUDA_THROW(cudaMalloc(&nd, items * b));
      if (data[v] && keep) FGB_CUDA_THROW(cudaMemcpy(nd, data[v], static_cast<size_t>(std::min(keep, capacity)) * b, cudaMemcpyDeviceToDevice));
      if (data[v]) cudaFree(data[v]);
      data[v] = nd;
      if (double_buffered) {
        if (swap[v]) cudaFree(swap[v]);
        FGB_CUDA_THROW(cudaMalloc(&swap[v], items * b));
      }
      if (pad_slot) {
        std::vector<char> pad(b, 0);
        if (meta[v].type == std::type_index(typeid(float)))
          for (unsigned int e = 0; e < meta[v].elements; ++e) reinterpret_cast<float *>(pad.data())[e] = detail::pad_location<float>();
        else if (meta[v].type == std::type_index(typeid(double)))
          for (unsigned int e = 0; e < meta[v].elements; ++e) reinterpret_cast<double *>(pad.data())[e] = detail::pad_location<double>();
        FGB_CUDA_THROW(cudaMemcpy(data[v] + static_cast<size_t>(cap) * b, pad.data(), b, cudaMemcpyHostToDevice));
        if (double_buffered) FGB_CUDA_THROW(cudaMemcpy(swap[v] + static_cast<size_t>(cap) * b, pad.data(), b, cudaMemcpyHostToDevice));
      }
    }
    capacity = cap;
  }
  void swap_buffers() {  // every caller swaps because it just reordered or compacted the list
    for (size_t v = 0; v < names.size(); ++v) std::swap(data[v], swap[v]);
    touch();
  }
  void release() {
    for (auto &p : data)
      if (p) cudaFree(p), p = nullptr;
    for (auto &p : swap)
      if (p) cudaFree(p), p = nullptr;
    capacity = 0;
  }
  // fgb_var table data -> swap (or the reverse)
  std::vector<fgb_var> vars(bool from_data = true) const {
    std::vector<fgb_var> out(names.size());
    for (size_t v = 0; v < names.size(); ++v) {
      out[v].type_len = meta[v].bytes();
      out[v].in = from_data ? data[v] : swap[v];
      out[v].out = from_data ? swap[v] : data[v];
    }
    return out;
  }
  // perfect-hash slots of the variable names, computed once per list
  mutable std::vector<uint32_t> slots;
  mutable DevVars proto{};
  void fill_table(DevVars &t, bool use_swap = false) const {
    if (names.size() > static_cast<size_t>(b200::kMaxVars)) throw exception::UnsupportedFeature("more than b200::kMaxVars variables in one list");
    if (slots.size() != names.size()) {
      std::vector<uint32_t> h(names.size());
      for (size_t v = 0; v < names.size(); ++v) h[v] = name_hash_rt(names[v]);
      slots.assign(names.size(), 0u);
      std::memset(&proto, 0, sizeof(proto));
      if (!build_perfect_table(proto, h.data(), static_cast<uint32_t>(h.size()), slots.data()))
        throw exception::UnsupportedFeature("could not build a collision-free variable table");
    }
    t = proto;
    for (size_t v = 0; v < names.size(); ++v) t.ptr[slots[v]] = use_swap ? swap[v] : data[v];
  }
};

struct DevFlags {  // grow-only u32 scan-flag array (one per thread of a function)
  unsigned int *p = nullptr;
  unsigned int cap = 0;
  unsigned long long *gen = nullptr;  // see DevList::gen
  void reserve(unsigned int n) {
    if (n <= cap) return;
    if (gen) ++*gen;
    if (p) cudaFree(p);
    cap = (n + n / 4 + 255u) & ~255u;
    FGB_CUDA_THROW(cudaMalloc(&p, static_cast<size_t>(cap) * 4));
    FGB_CUDA_THROW(cudaMemset(p, 0, static_cast<size_t>(cap) * 4));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct CUDAMessage {
  std::shared_ptr<MessageData> desc;
  DevList list;
  fgb_spatial *spatial = nullptr;   // index handler: spatial 2D/3D lists and bucket lists (bucket == true)
  bool bucket = false;
  fgb_spatial_metadata md{};
  bool pbm_dirty = true;
  bool truncate = true;
  // fused index build: the function that wrote the list published every message's bin key and histogram count
  bool keyed_by_writer = false;      // keys + histogram exist for the first *keyed_slot items
  bool appended_after_keyed = false; // ... and more items were appended since (they are keyed by the build)
  bool hist_dirty = false;           // the histogram holds counts that no build has consumed yet
  bool written_bin_ordered = false;  // the writer ran in bin order (slot == thread): the build may expect grouped tiles
  unsigned int keyed_slot = 0;       // control word: number of items keyed by the writer
  int win_begin = 0, win_count = -1;  // slab window (planes of the slowest axis held by this process)
};

struct CUDAAgent {
  std::shared_ptr<AgentData> desc;
  std::map<std::string, DevList> states;
  unsigned int next_id_slot = 0;
  id_t host_next_id = 1;
  // Host-known upper bound of the agents of this type over ALL states.  State transitions move agents between
  // lists but cannot create any, so a list bound never needs to exceed it: without this clamp a conditional
  // transition A -> B adds the source bound to B's bound every step (the device counts are not read back).
  unsigned int pop_bound = 0;
  void recompute_pop_bound() {
    unsigned long long t = 0;
    for (const auto &s : states) t += s.second.bound;
    pop_bound = static_cast<unsigned int>(std::min<unsigned long long>(t, 0xFFFFFFF0ull));
  }
};

struct FunctionRT {
  std::shared_ptr<AgentFunctionData> fn;
  CUDAAgent *agent = nullptr;
  CUDAMessage *msg_in = nullptr, *msg_out = nullptr;
  CUDAAgent *out_agent = nullptr;
  DevList scratch_new;  // new-agent slots, one per parent thread
  char *d_defaults = nullptr;
  std::vector<size_t> default_offsets;
  DevFlags death_flag, msg_flag, birth_flag, cond_flag, move_flag;
  unsigned int failed_slot = 0, active_slot = 0;  // function condition: device words
  bool sortable = false;
  int sort_dims = 0;
  fgb_spatial *exec_binner = nullptr;  // bins the executing agents on the input list's grid
  DevFlags exec_perm;
  unsigned int tmp_slot = 0;  // control word scratch (survivor count)
  int block_size = 128;
};

}  // namespace detail

class CUDASimulation;

// Host agent creation (reference include/flamegpu/runtime/agent/HostNewAgentAPI.h): an agent made by a host function lives
// in a host-side array of structs (defaults pre-filled) until the host function returns; CUDASimulation then uploads the
// structs and fgb_scatter_new_agents transposes them behind the state list (reference CUDAAgent::scatterHostCreation,
// CUDAScatter.cu:348-395).
namespace detail {
struct HostNewAgents {
  std::vector<std::string> names;
  std::vector<Variable> meta;
  std::vector<size_t> offset;  // of every variable inside one struct
  size_t agent_size = 0;
  std::vector<char> defaults;  // one struct holding the default values
  std::vector<char> data;      // count structs
  unsigned int count = 0;
};
}  // namespace detail
class HostNewAgentAPI {
 public:
  HostNewAgentAPI(detail::HostNewAgents *b, unsigned int i) : buf(b), index(i) {}
  template <typename T>
  void setVariable(const std::string &name, T value) {
    *reinterpret_cast<T *>(slot<T>(name, 1)) = value;
  }
  template <typename T, flamegpu::size_type N>
  void setVariable(const std::string &name, unsigned int i, T value) {
    if (i >= N) throw exception::OutOfBoundsException("array index out of bounds in HostNewAgentAPI::setVariable()");
    reinterpret_cast<T *>(slot<T>(name, N))[i] = value;
  }
  template <typename T>
  T getVariable(const std::string &name) const {
    return *reinterpret_cast<const T *>(const_cast<HostNewAgentAPI *>(this)->slot<T>(name, 1));
  }

 private:
  template <typename T>
  char *slot(const std::string &name, unsigned int elements) {
    if (name == ID_VARIABLE_NAME) throw exception::ReservedName("agent ids are assigned by the simulation");
    for (size_t v = 0; v < buf->names.size(); ++v)
      if (buf->names[v] == name) {
        if (buf->meta[v].type != std::type_index(typeid(T)) || buf->meta[v].elements != elements)
          throw exception::InvalidVarType("wrong type for variable '" + name + "' in HostNewAgentAPI");
        return buf->data.data() + static_cast<size_t>(index) * buf->agent_size + buf->offset[v];
      }
    throw exception::InvalidAgentVar("new agent has no variable '" + name + "'");
  }
  detail::HostNewAgents *buf;
  unsigned int index;
};

// Host-side API handed to init/step/exit functions (subset of the reference's HostAPI).
// sum / min / max run on the device (fgb_reduce: one kernel, 8 bytes come back) instead of the reference's
// cub::DeviceReduce + copy (HostAgentAPI.cuh:540-700).
class HostAgentAPI {
 public:
  HostAgentAPI(CUDASimulation *s, std::string a, std::string st) : sim(s), agent(std::move(a)), state(std::move(st)) {}
  inline unsigned int count();
  inline HostNewAgentAPI newAgent();  // reference HostAgentAPI.cuh:138-142
  template <typename T>
  inline T sum(const std::string &variable);
  template <typename T>
  inline T min(const std::string &variable);
  template <typename T>
  inline T max(const std::string &variable);
  // reference HostAgentAPI.cuh:700-718 (thrust::count) and :561-604 (mean, POPULATION standard deviation)
  template <typename T>
  inline unsigned int count(const std::string &variable, T value);
  template <typename T>
  inline std::pair<double, double> meanStandardDeviation(const std::string &variable);
  // reference HostAgentAPI.cuh:241,417,720-760 (cub::DeviceHistogram::HistogramEven): counts of lower <= v < upper in even bins
  template <typename InT, typename OutT = unsigned int>
  inline std::vector<OutT> histogramEven(const std::string &variable, unsigned int histogramBins, InT lowerBound, InT upperBound);
#if defined(__CUDACC__)
  // reference HostAgentAPI.cuh:255-268,776-850 (thrust::reduce / thrust::transform_reduce with user functors declared by
  // FLAMEGPU_CUSTOM_REDUCTION / FLAMEGPU_CUSTOM_TRANSFORM): one kernel, block partials folded by the last block in block order
  template <typename InT, typename reductionOperatorT>
  inline InT reduce(const std::string &variable, reductionOperatorT reductionOperator, InT init);
  template <typename InT, typename OutT, typename transformOperatorT, typename reductionOperatorT>
  inline OutT transformReduce(const std::string &variable, transformOperatorT transformOperator, reductionOperatorT reductionOperator, OutT init);
  template <typename InT, typename transformOperatorT, typename reductionOperatorT>
  inline InT transformReduce(const std::string &variable, transformOperatorT t, reductionOperatorT r, InT init) {
    return transformReduce<InT, InT, transformOperatorT, reductionOperatorT>(variable, t, r, init);
  }
#endif

 private:
  template <typename T, typename R>
  inline R reduce(const std::string &variable, int op);
  template <typename T, typename R>
  inline R transform_reduce(const std::string &variable, int transform, const void *param);
  CUDASimulation *sim;
  std::string agent, state;
};
class HostEnvironment {
 public:
  explicit HostEnvironment(CUDASimulation *s) : sim(s) {}
  template <typename T>
  inline T getProperty(const std::string &name);
  template <typename T>
  inline T setProperty(const std::string &name, T value);

 private:
  CUDASimulation *sim;
};
class HostAPI {
 public:
  explicit HostAPI(CUDASimulation *s) : environment(s), sim(s) {}
  HostAgentAPI agent(const std::string &agent_name, const std::string &state = DEFAULT_STATE) { return HostAgentAPI(sim, agent_name, state); }
  inline unsigned int getStepCounter() const;
  HostEnvironment environment;

 private:
  CUDASimulation *sim;
};

enum CONDITION_RESULT { CONTINUE, EXIT };

class CUDASimulation {
 public:
  struct Config {  // reference include/flamegpu/simulation/Simulation.h:37-71 (subset)
    std::string input_file;
    uint64_t random_seed = 0;
    unsigned int steps = 1;
    bool timing = false;
    bool truncate_log_files = true;
    int verbosity = 1;
  };
  struct CUDAConfig_t {  // reference CUDASimulation.h:100-129 (+ b200 extensions)
    int device_id = 0;
    bool inLayerConcurrency = true;
    bool useCUDAGraphs = true;        // b200: capture each step as a CUDA graph
    bool stableMessageOrder = false;  // b200: deterministic (source) order inside PBM bins
    int spatialIterationMode = -1;    // b200: -1 per function (radius-filtered where the function declared the radius
                                      // contract, setMessageInputRadiusFiltered), 0 reference order everywhere,
                                      // 1 radius-filtered lock-step walk for every spatial reader (FunctionArgs.h)
    bool binOrderExecution = true;    // b200: run functions that read spatial messages in bin order
    int agentFunctionBlockSize = 128;  // b200: threads per block of the agent function kernels
    bool tileLocalExecOrder = true;   // b200: bin-order execution groups inside 2048-agent tiles when the list was just sorted
    bool overlapIndexBuild = true;    // b200: build the input list's PBM on a second stream while the agents are sorted
    bool profile = false;             // b200: eager execution with CUDA events around every phase (getProfile())
    unsigned int slabRefreshPeriod = 16;  // b200 slabs: steps between host re-reads of the list counts (launch bounds)
    unsigned int slabTimeoutMs = 20000;   // b200 slabs: a wait for a neighbour gives up (error word) after this long
    bool fusedIndexBuild = true;      // b200: mandatory spatial output publishes bin keys + histogram (buildIndex starts at the scan)
    bool binOrderedOutput = true;     // b200: mandatory spatial output writes its messages in the bin order of the last reader
    bool trueSpatialSortKey = false;  // b200: sort 3D agents by the intended x,y,z key (the reference's
                                      // key collapses z, CUDASimulation.cu:487; see sort_geometry())
  };

  explicit CUDASimulation(const ModelDescription &model_desc, int argc = 0, const char **argv = nullptr)
      : model(model_desc.model), host_api(this) {
    parse_args(argc, argv);
  }
  ~CUDASimulation() { destroy(); }
  CUDASimulation(const CUDASimulation &) = delete;
  CUDASimulation &operator=(const CUDASimulation &) = delete;

  Config &SimulationConfig() { return config; }
  const Config &getSimulationConfig() const { return config; }
  CUDAConfig_t &CUDAConfig() { return cuda_config; }

  void setPopulationData(AgentVector &pop, const std::string &state = DEFAULT_STATE);
  void getPopulationData(AgentVector &pop, const std::string &state = DEFAULT_STATE);
  bool step();
  void simulate();
  unsigned int getStepCounter() const { return step_count; }
  void resetStepCounter() {
    step_count = 0;
    last_step_prefetched = 0;
    if (initialised) FGB_CUDA_THROW(cudaMemset(d_ctrl + kStepSlot, 0, 4));
  }
  // seconds per step (reference CUDASimulation.h:362); measured with CUDA events on the step stream
  std::vector<double> getElapsedTimeSteps();
  double getElapsedTimeSimulation() const { return elapsed_simulation; }
  template <typename T>
  T getEnvironmentProperty(const std::string &name) { return EnvironmentDescription(model).getProperty<T>(name); }
  template <typename T>
  T setEnvironmentProperty(const std::string &name, T value) {
    T old = EnvironmentDescription(model).setProperty<T>(name, value);
    env_dirty = true;
    return old;
  }
  unsigned int getAgentCount(const std::string &agent_name, const std::string &state = DEFAULT_STATE);

  // ---- b200 multi-GPU extension: z-slab decomposition (no reference counterpart) -------
  // Must be called before the first step / setPopulationData: this process stores only planes
  // [plane_begin, plane_begin+plane_count) of the message list's slowest grid axis (own planes + ghosts).
  void setMessageWindow(const std::string &message_name, int plane_begin, int plane_count) {
    if (initialised) throw exception::InvalidArgument("setMessageWindow must be called before the simulation is initialised");
    windows[message_name] = std::make_pair(plane_begin, plane_count);
  }
  // Slab decomposition BEHIND step(): this process is rank `rank` of `world` (one process per GPU) and owns a contiguous
  // range of bin planes of `message`'s slowest grid axis.  Call before the first step / setPopulationData, upload only
  // the agents whose position lies in slabPlanes(), exchange the staging handles (slabExportHandle on every rank, an
  // all-gather by the caller -- torch.distributed / MPI: plumbing -- then slabConnect) and call step() as usual:
  //   * before `message` is indexed for its first reader, the messages of the two boundary planes are packed straight
  //     into the neighbours' staging buffers (peer memory over NVLink), the neighbours' halos are appended;
  //   * at the end of the step agents whose position left the slab are packed to the neighbour the same way, removed
  //     here and the arriving ones appended;
  //   * HostAgentAPI reductions (step functions) are all-reduced over per-rank mailboxes.
  // Everything is kernel launches on the step's streams: the slab step is captured in the same CUDA graphs, nothing
  // returns to the host (list bounds are re-read every CUDAConfig().slabRefreshPeriod steps).
  void configureSlabs(int rank, int world, const std::string &message, unsigned int halo_capacity, unsigned int migrate_capacity);
  size_t slabHandleBytes() const { return sizeof(cudaIpcMemHandle_t); }
  void slabExportHandle(void *out);
  void slabConnect(const void *all_handles);
  void slabPlanes(int *z0, int *z1) const {
    if (z0) *z0 = slab.z0;
    if (z1) *z1 = slab.z1;
  }
  unsigned int slabError();  // FGB_SLAB_ERR_* bits raised on the device so far (synchronises)
  // run layers [first, last) of the model eagerly (no end-of-step bookkeeping); endStep() finishes the step
  void runLayers(unsigned int first, unsigned int last);
  void endStep();
  // Select the items of a list whose position along the slowest axis lies in planes < lo / >= hi
  // (fgb_plane_flags + fgb_compact), packed into caller-provided device buffers (one per variable, list
  // order); returns the two counts.  remove != 0 additionally drops them from the list (agent migration).
  // Nothing returns to the host: the two counts go to the device words *d_count_lo and *d_count_hi.
  void slabPack(bool is_message, const std::string &name, const std::string &state, const std::string &geometry_message,
                int lo, int hi, void *const *dst_lo, void *const *dst_hi, unsigned int capacity, bool remove,
                unsigned int *d_count_lo, unsigned int *d_count_hi);
  // append up to n_max items (actual count in the device word d_n) from device buffers (one per variable,
  // list order) to a list
  void listAppend(bool is_message, const std::string &name, const std::string &state, unsigned int n_max,
                  const unsigned int *d_n, const void *const *src);
  // re-read every list count from the device (one sync) and tighten the host-side launch bounds
  void refreshBounds() { initialise(); refresh_bounds(); }
  // Pipelined variant for the slab driver: the counts of step t are copied to pinned memory asynchronously and
  // consumed one step later, while the GPU is already running step t+1, so the host never drains the device.
  void endStepPipelined();
  // Message exchange on its own stream (multi-GPU slab driver): between beginMessageExchange() and
  // endMessageExchange() slabPack / listAppend run on exchangeStream() -- and so does the caller's NCCL traffic --
  // while the main stream already sorts the agents of the reading layer; endMessageExchange() builds the PBM of
  // `message` on that stream, and the reading functions wait for it right before their kernels.
  void beginMessageExchange();
  void endMessageExchange(const std::string &message);
  cudaStream_t exchangeStream() const { return index_stream; }
  std::vector<std::pair<std::string, size_t>> listLayout(bool is_message, const std::string &name);

  // ---- b200 extensions used by the parity harness / bench (no reference counterpart) -------------
  // Bulk SoA population exchange straight between caller buffers (ideally pinned) and the device
  // lists, asynchronous on the simulation stream: the AgentVector path above stages every variable
  // through pageable host vectors.  Variables not listed are reset to their defaults; ids restart at 1.
  void setPopulationDataSoA(const std::string &agent_name, const std::string &state, unsigned int n, unsigned int nvars,
                            const char *const *names, const void *const *host_ptrs);
  // Streamed download: arms the NEXT step() to copy the listed variables of a state list into the caller's (pinned) host
  // buffers WHILE the step's last function on that list is still running -- the function is launched in `chunks`
  // thread-range chunks and every finished chunk is copied on a second stream, so only the last chunk's copy is exposed
  // (the reference's getPopulationData copies after the step, CUDASimulation.cu:1417-1440).  Falls back to one copy at
  // the end of the step when the list's last function can die / birth / change state, has a condition, or runs in a
  // global (not tile-local) bin order.  finishStreamedPopulation() waits for the copies and returns the agent count.
  void streamPopulationDataSoA(const std::string &agent_name, const std::string &state, unsigned int nvars, const char *const *names,
                               void *const *host_ptrs, unsigned int chunks = 8);
  unsigned int finishStreamedPopulation();
  // Returns the agent count; copies the listed variables (device order) into the caller's buffers.
  unsigned int getPopulationDataSoA(const std::string &agent_name, const std::string &state, unsigned int nvars,
                                    const char *const *names, void *const *host_ptrs, unsigned int capacity);
  // device pointers of a state list variable / message variable (current read buffers)
  void *getAgentVariableDevicePtr(const std::string &agent_name, const std::string &state, const std::string &var);
  void *getMessageVariableDevicePtr(const std::string &message_name, const std::string &var);
  unsigned int getMessageCount(const std::string &message_name);
  fgb_spatial *getSpatialHandler(const std::string &message_name);
  HostAPI &hostAPI() { return host_api; }
  unsigned long long getLaunchCount() const { return (ctx ? fgb_launch_count(ctx) : 0ull) + own_launches; }
  unsigned int getGraphCount() const { return static_cast<unsigned int>(graphs.size()); }
  unsigned int getGraphWidth() const { return last_graph_width; }
  void getListBound(const std::string &agent_name, const std::string &state, unsigned int *bound, unsigned int *capacity) {
    initialise();
    const detail::DevList &l = state_list(agent_name, state);
    if (bound) *bound = l.bound;
    if (capacity) *capacity = l.capacity;
  }
  cudaStream_t getStream() const { return main_stream; }
  // phase name -> (total milliseconds, calls) since the last call; only filled when CUDAConfig().profile
  std::map<std::string, std::pair<double, unsigned int>> getProfile();
  void synchronize() { if (initialised) FGB_CUDA_THROW(cudaStreamSynchronize(main_stream)); }

 private:
  friend class HostAgentAPI;
  friend class HostEnvironment;
  friend class HostAPI;
  static constexpr unsigned int kCtrlWords = 1024;
  static constexpr unsigned int kStepSlot = 0;

  void parse_args(int argc, const char **argv);
  void initialise();
  void destroy();
  detail::CUDAAgent &agent_rt(const std::string &name) {
    auto it = agents.find(name);
    if (it == agents.end()) throw exception::InvalidCudaAgent("agent '" + name + "' was not found");
    return it->second;
  }
  detail::DevList &state_list(const std::string &agent_name, const std::string &state) {
    auto &a = agent_rt(agent_name);
    auto it = a.states.find(state);
    if (it == a.states.end()) throw exception::InvalidStateName("agent '" + agent_name + "' has no state '" + state + "'");
    return it->second;
  }
  unsigned int alloc_slot() {
    if (next_slot >= kCtrlWords) throw exception::UnsupportedFeature("control block exhausted");
    return next_slot++;
  }
  unsigned int *slot_ptr(unsigned int s) const { return d_ctrl + s; }
  unsigned int read_slot(unsigned int s) {
    unsigned int v = 0;
    FGB_CUDA_THROW(cudaMemcpyAsync(&v, d_ctrl + s, 4, cudaMemcpyDeviceToHost, main_stream));
    FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
    return v;
  }
  void write_slot(unsigned int s, unsigned int v) { FGB_CUDA_THROW(cudaMemcpy(d_ctrl + s, &v, 4, cudaMemcpyHostToDevice)); }

  // ---- slab decomposition state --------------------------------------------------------------------------------
  struct SlabStaging {  // receive buffer for one (list, side) inside the arena: variables back to back, count, flag
    std::vector<size_t> var_off;
    size_t count_off = 0, flag_off = 0;
  };
  struct SlabList {
    bool is_message = false;
    detail::DevList *list = nullptr;
    detail::CUDAAgent *agent = nullptr;
    int pos_var = -1;           // index of the position variable along the decomposed axis
    unsigned int capacity = 0;  // items per staging buffer
    SlabStaging st[2];          // [0]: data arriving from rank-1, [1]: from rank+1
  };
  struct Slab {
    bool enabled = false, connected = false;
    int rank = 0, world = 1;
    std::string message;
    int planes = 0, z0 = 0, z1 = 0;
    unsigned int halo_cap = 0, mig_cap = 0;
    char *arena = nullptr;  // every staging buffer + the reduction mailboxes of THIS rank, one IPC-exported allocation
    size_t arena_bytes = 0, mail_off = 0;
    std::vector<char *> peer;  // arena of every rank as mapped here (peer[rank] == arena)
    std::vector<SlabList> lists;  // [0] = the halo message list, then every agent state list that carries the position
    unsigned int epoch_slot = 0, err_slot = 0;
  } slab;
  struct Streamed {
    bool armed = false, chunked = false;
    detail::DevList *list = nullptr;
    const detail::FunctionRT *function = nullptr;  // the step's last function on the list (chunked launch), or NULL
    std::vector<int> vars;
    std::vector<void *> host;
    unsigned int chunks = 8, n = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t chunk_done = nullptr;
  } streamed;
  std::map<std::pair<std::string, std::string>, detail::HostNewAgents> host_new_agents;  // (agent, state) -> agents made by host functions
  detail::HostNewAgents &host_new_buffer(const std::string &agent_name, const std::string &state);
  void flush_host_agents();  // uploads + transposes them behind their state lists (after every host function phase)
  char *d_new_aos = nullptr;
  size_t new_aos_bytes = 0;
  unsigned int *h_words = nullptr;  // pinned staging for small host -> device control words
  unsigned int soa_uploads = 0;
  cudaStream_t stream_copy = nullptr;        // streamed downloads
  cudaEvent_t stream_chunk_done = nullptr;
  static constexpr unsigned int kSlabScratchSlot = 127;  // fgb_ctx scratch slot of the exchange (FGB_MAX_STREAMS - 1)
  void slab_setup();  // arena layout + allocation (initialise)
  void slab_exchange(SlabList &S, int lo_plane, int hi_plane, bool remove, cudaStream_t st);
  void slab_refresh_bounds();
  void slab_allreduce(void *d_value, int dtype, int op);
  void upload_environment();
  void plan_step();                         // reserve capacities for the coming step (may allocate)
  std::vector<unsigned int> last_plan_sig;  // list bounds the last plan was made for (plan_step is skipped while they hold)
  void record_step(cudaStream_t main);      // enqueue one whole step
  void record_layers(cudaStream_t main, size_t first, size_t last);
  void record_end_of_step(cudaStream_t main);
  void run_function(detail::FunctionRT &f, cudaStream_t st, unsigned int stream_id);
  void build_input_index(detail::CUDAMessage &M, cudaStream_t st);
  bool filtered_iteration(const detail::FunctionRT &f) const;
  void refresh_bounds();                    // births only: read the counts back once per step
  int sort_geometry(const detail::FunctionRT &f, float mn[3], float width[3], unsigned int gd[3]) const;
  std::vector<unsigned long long> graph_key() const;
  static unsigned int graph_width(cudaGraph_t graph);
  static unsigned int quantise(unsigned int n) {
    if (n <= 4096u) return (n + 255u) & ~255u;
    unsigned int p = 1u;
    while (p < n) p <<= 1;
    const unsigned int q = p >> 5;  // 32 steps per octave
    return (n + q - 1) / q * q;
  }

  std::shared_ptr<ModelData> model;
  Config config;
  CUDAConfig_t cuda_config;
  bool initialised = false;
  fgb_ctx *ctx = nullptr;
  cudaStream_t main_stream = nullptr;
  std::vector<cudaStream_t> side_streams;
  std::vector<cudaEvent_t> join_events;
  cudaEvent_t fork_event = nullptr;
  void *d_reduce_out = nullptr;          // 8-byte result word of HostAgentAPI reductions
  // Reductions a step function asked for are LEARNED and from the next step on recorded behind the step itself (inside
  // its CUDA graph): their results and the step's epoch land in mapped pinned host memory, so the step function finds the
  // value with a spin on one host word -- no kernel launch behind a drained stream, no D2H copy, no stream synchronise.
  struct PrefetchedReduction {
    std::string agent, state, variable;
    int op = 0, dtype = 0;
  };
  std::vector<PrefetchedReduction> prefetch;
  static constexpr unsigned int kMaxPrefetch = 32;
  unsigned long long *h_prefetch = nullptr;  // [0] epoch = step counter after the step, [1 + k] result of prefetch[k]
  unsigned long long *d_prefetch = nullptr;  // device alias of h_prefetch
  unsigned long long *d_prefetch_vals = nullptr;  // device-resident results (all-reduced in place under slabs) before they are published
  unsigned long long *d_reduce_epoch = nullptr;   // device word: epoch of the slab all-reduce mailboxes (fgb_slab_allreduce)
  unsigned int last_step_prefetched = 0;     // how many reductions the step that just ran has recorded
  bool in_step_function = false;             // a prefetched value describes the lists as the step left them: step functions only
  void record_prefetched_reductions(cudaStream_t st);
  bool prefetched_result(const std::string &agent, const std::string &state, const std::string &variable, int op, int dtype, void *out, size_t bytes);
  void *d_user_reduce = nullptr;         // block partials + arrival counter of the user-functor reductions
  unsigned int *d_hist_out = nullptr;    // histogramEven counts
  unsigned int hist_cap = 0;
  cudaStream_t index_stream = nullptr;   // PBM builds of a layer's input lists (overlapIndexBuild)
  cudaEvent_t index_fork = nullptr, index_done = nullptr;
  bool index_pending = false;
  bool exchange_active = false;   // slabPack / listAppend target the exchange stream (scratch slot 1)
  unsigned int *d_ctrl = nullptr;
  unsigned int next_slot = 1;
  unsigned int slab_tmp_slot = 0;
  unsigned int *d_zero_slots = nullptr;
  unsigned int n_zero_slots = 0;
  char *d_env = nullptr;
  detail::DevEnv env_table{};
  std::vector<char> env_host;
  std::vector<size_t> env_offsets;
  bool env_dirty = true;
  std::map<std::string, detail::CUDAAgent> agents;
  std::map<std::string, detail::CUDAMessage> messages;
  std::map<std::string, std::pair<int, int>> windows;
  detail::DevFlags slab_flags[3];
  unsigned int *h_ctrl_pinned[2] = {nullptr, nullptr};
  cudaEvent_t ctrl_events[2] = {nullptr, nullptr};
  unsigned long long pipelined_steps = 0;
  std::vector<std::vector<detail::FunctionRT>> layers;  // [layer][function]
  std::vector<bool> layer_serial;                        // layer whose members must not run concurrently
  bool model_has_births = false;
  bool model_has_host_layers = false;
  unsigned int step_count = 0;
  unsigned long long own_launches = 0;
  HostAPI host_api;
  struct GraphEntry {
    std::vector<unsigned long long> key;
    cudaGraphExec_t exec = nullptr;
    unsigned long long launches = 0;
    unsigned int prefetched = 0;  // reductions recorded behind the step (prefetch[0 .. prefetched))
    // host-side state after the step (pointer parity, bounds, flags) to restore on replay
    std::vector<unsigned long long> post_state;
  };
  std::vector<GraphEntry> graphs;
  unsigned int last_graph_width = 0;
  unsigned long long alloc_gen = 0;          // bumped by every reallocation of a list / flag array of this simulation
  unsigned long long graphs_generation = 0;  // allocation generation the cached graphs were captured under
  struct ProfRec {
    std::string name;
    cudaEvent_t e0, e1;
  };
  std::vector<ProfRec> prof;
  void prof_begin(const std::string &name, cudaStream_t st) {
    if (!cuda_config.profile) return;
    ProfRec r;
    r.name = name;
    FGB_CUDA_THROW(cudaEventCreate(&r.e0));
    FGB_CUDA_THROW(cudaEventCreate(&r.e1));
    FGB_CUDA_THROW(cudaEventRecord(r.e0, st));
    prof.push_back(r);
  }
  void prof_end(cudaStream_t st) {
    if (!cuda_config.profile) return;
    FGB_CUDA_THROW(cudaEventRecord(prof.back().e1, st));
  }
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> step_events;
  std::vector<cudaEvent_t> event_pool;  // recycled timing events (take_event / harvest_step_events)
  cudaEvent_t take_event();
  void harvest_step_events(bool only_finished);
  std::vector<double> step_seconds;
  double elapsed_simulation = 0.0;

  std::vector<unsigned long long> snapshot_host_state() const;
  void restore_host_state(const std::vector<unsigned long long> &s);
};

}  // namespace flamegpu

// reference include/flamegpu/runtime/agent/HostAgentAPI.cuh:69-79,104-114: user functors for HostAgentAPI::reduce / transformReduce
#define FLAMEGPU_CUSTOM_REDUCTION(funcName, a, b)                                                              \
  struct funcName##_impl {                                                                                     \
   public:                                                                                                     \
    template <typename OutT>                                                                                   \
    struct binary_function {                                                                                   \
      __host__ __device__ __forceinline__ OutT operator()(const OutT &a, const OutT &b) const;                 \
    };                                                                                                         \
  };                                                                                                           \
  funcName##_impl funcName;                                                                                    \
  template <typename OutT>                                                                                     \
  __host__ __device__ __forceinline__ OutT funcName##_impl::binary_function<OutT>::operator()(const OutT &a, const OutT &b) const
#define FLAMEGPU_CUSTOM_TRANSFORM(funcName, a)                                                                 \
  struct funcName##_impl {                                                                                     \
   public:                                                                                                     \
    template <typename InT, typename OutT>                                                                     \
    struct unary_function {                                                                                    \
      __host__ __device__ OutT operator()(const InT &a) const;                                                 \
    };                                                                                                         \
  };                                                                                                           \
  funcName##_impl funcName;                                                                                    \
  template <typename InT, typename OutT>                                                                       \
  __host__ __device__ __forceinline__ OutT funcName##_impl::unary_function<InT, OutT>::operator()(const InT &a) const

#include "flamegpu/simulation/CUDASimulation_impl.h"

#endif  // FGB_INCLUDE_FLAMEGPU_SIMULATION_CUDASIMULATION_H_
