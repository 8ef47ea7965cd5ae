// flamegpu/runtime/DeviceAPI.cuh -- the object an agent function receives as `FLAMEGPU`.
// API compatible with the reference's include/flamegpu/runtime/DeviceAPI.cuh:163-401 for the hot
// path: agent variable access, message_in / message_out, agent_out (births), environment
// properties, getID, getStepCounter.  (RNG, macro properties and directed graphs are outside the
// hot-path scope)
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_DEVICEAPI_CUH_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_DEVICEAPI_CUH_

#include "flamegpu/defines.h"
#include "flamegpu/runtime/detail/FunctionArgs.h"

namespace flamegpu {

#if defined(__CUDACC__)

// Read-only environment properties (reference runtime/environment/DeviceEnvironment.cuh)
class DeviceEnvironment {
 public:
  __device__ __forceinline__ explicit DeviceEnvironment(const detail::FunctionArgs &args) : a(args) {}
  template <typename T, unsigned int N>
  __device__ __forceinline__ T getProperty(const char (&name)[N]) const {
    const int slot = detail::find_slot(a.env, detail::name_hash(name));
    if (slot < 0) return T{};
    return __ldg(reinterpret_cast<const T *>(a.env.buffer + a.env.offset[slot]));
  }
  template <typename T, flamegpu::size_type N, unsigned int M>
  __device__ __forceinline__ T getProperty(const char (&name)[M], unsigned int index) const {
    const int slot = detail::find_slot(a.env, detail::name_hash(name));
    if (slot < 0 || index >= N) return T{};
    return __ldg(reinterpret_cast<const T *>(a.env.buffer + a.env.offset[slot]) + index);
  }

 private:
  const detail::FunctionArgs &a;
};

template <typename MessageIn, typename MessageOut>
class DeviceAPI {
 public:
  // Births (reference DeviceAPI.cuh:486-557).  The child lives in slot `parent thread index` of a
  // scratch SoA list; which variables the parent set is tracked in a register so that only real
  // births get the remaining defaults (the reference pre-fills ALL N scratch slots with defaults
  // before every launch: broadcastInit, CUDAAgent.cu:378-383).
  class AgentOut {
   public:
    __device__ __forceinline__ AgentOut(const detail::FunctionArgs &args, unsigned int index)
        : a(args), slot(index), set_mask(0ull), id(ID_NOT_SET) {}
    template <typename T, unsigned int N>
    __device__ __forceinline__ void setVariable(const char (&name)[N], T value) const {
      if (!a.agent_out_flag) return;
      genID();
      const int s = detail::find_slot(a.agent_out, detail::name_hash(name));
      if (s >= 0) {
        reinterpret_cast<T *>(a.agent_out.ptr[s])[slot] = value;
        set_mask |= 1ull << s;
      }
    }
    template <typename T, flamegpu::size_type N, unsigned int M>
    __device__ __forceinline__ void setVariable(const char (&name)[M], unsigned int index, T value) const {
      if (!a.agent_out_flag) return;
      genID();
      const int s = detail::find_slot(a.agent_out, detail::name_hash(name));
      if (s >= 0 && index < N) {
        T *base = reinterpret_cast<T *>(a.agent_out.ptr[s]) + static_cast<size_t>(slot) * N;
        if (!(set_mask & (1ull << s))) {  // first touch of an array variable: start from its default
          const T *def = nullptr;
          for (uint32_t v = 0; v < a.agent_out_nvars; ++v)
            if (a.agent_out_slot[v] == static_cast<uint32_t>(s)) def = reinterpret_cast<const T *>(a.agent_out_defaults[v]);
          for (unsigned int e = 0; e < N; ++e) base[e] = def ? def[e] : T{};
          set_mask |= 1ull << s;
        }
        base[index] = value;
      }
    }
    __device__ __forceinline__ id_t getID() const {
      if (!a.agent_out_flag) return ID_NOT_SET;
      genID();
      return id;
    }
    __device__ __forceinline__ bool born() const { return id != ID_NOT_SET; }
    // Called once by the kernel wrapper: publish the birth flag and complete the child.
    __device__ __forceinline__ void finalise() const {
      if (!a.agent_out_flag) return;
      a.agent_out_flag[slot] = born() ? 1u : 0u;
      if (!born()) return;
      for (uint32_t v = 0; v < a.agent_out_nvars; ++v) {
        const uint32_t s = a.agent_out_slot[v];
        if (set_mask & (1ull << s)) continue;
        const uint32_t len = a.agent_out_len[v];
        char *dst = a.agent_out.ptr[s] + static_cast<size_t>(slot) * len;
        const char *def = a.agent_out_defaults[v];
        if ((len & 3u) == 0) {
          for (uint32_t w = 0; w < len; w += 4) *reinterpret_cast<uint32_t *>(dst + w) = *reinterpret_cast<const uint32_t *>(def + w);
        } else {
          for (uint32_t b = 0; b < len; ++b) dst[b] = def[b];
        }
      }
    }

   private:
    // reference DeviceAPI.cuh:548-557: first touch draws a fresh id and marks the slot
    __device__ __forceinline__ void genID() const {
      if (id == ID_NOT_SET) {
        id = atomicAdd(a.next_id, 1u);
        const int s = detail::find_slot(a.agent_out, detail::name_hash("_id"));
        if (s >= 0) {
          reinterpret_cast<id_t *>(a.agent_out.ptr[s])[slot] = id;
          set_mask |= 1ull << s;
        }
      }
    }
    const detail::FunctionArgs &a;
    unsigned int slot;
    mutable unsigned long long set_mask;
    mutable id_t id;
  };

  __device__ __forceinline__ DeviceAPI(const detail::FunctionArgs &args, unsigned int idx, unsigned int slot, int iter_mode = 0)
      : message_in(args, iter_mode), message_out(args, slot), agent_out(args, slot), environment(args), a(args), index(idx) {}

  template <typename T, unsigned int N>
  __device__ __forceinline__ T getVariable(const char (&name)[N]) const {
    const int s = detail::find_slot(a.agent, detail::name_hash(name));
    if (s < 0) return T{};
    return reinterpret_cast<const T *>(a.agent.ptr[s])[index];
  }
  template <typename T, flamegpu::size_type N, unsigned int M>
  __device__ __forceinline__ T getVariable(const char (&name)[M], unsigned int i) const {
    const int s = detail::find_slot(a.agent, detail::name_hash(name));
    if (s < 0 || i >= N) return T{};
    return reinterpret_cast<const T *>(a.agent.ptr[s])[static_cast<size_t>(index) * N + i];
  }
  template <typename T, unsigned int N>
  __device__ __forceinline__ void setVariable(const char (&name)[N], T value) const {
    const int s = detail::find_slot(a.agent, detail::name_hash(name));
    if (s >= 0) reinterpret_cast<T *>(a.agent.ptr[s])[index] = value;
  }
  template <typename T, flamegpu::size_type N, unsigned int M>
  __device__ __forceinline__ void setVariable(const char (&name)[M], unsigned int i, T value) const {
    const int s = detail::find_slot(a.agent, detail::name_hash(name));
    if (s >= 0 && i < N) reinterpret_cast<T *>(a.agent.ptr[s])[static_cast<size_t>(index) * N + i] = value;
  }
  __device__ __forceinline__ id_t getID() const { return getVariable<id_t>("_id"); }
  __device__ __forceinline__ unsigned int getStepCounter() const { return a.d_step ? __ldg(a.d_step) : 0u; }
  __device__ __forceinline__ unsigned int getIndex() const { return index; }
  __device__ __forceinline__ unsigned int getThreadIndex() const { return index; }

  const typename MessageIn::In message_in;
  const typename MessageOut::Out message_out;
  const AgentOut agent_out;
  const DeviceEnvironment environment;

 private:
  const detail::FunctionArgs &a;
  const unsigned int index;
};

// The object a function condition receives (reference runtime/DeviceAPI.cuh:40-160 ReadOnlyDeviceAPI)
class ReadOnlyDeviceAPI {
 public:
  __device__ __forceinline__ ReadOnlyDeviceAPI(const detail::FunctionArgs &args, unsigned int idx)
      : environment(args), a(args), index(idx) {}
  template <typename T, unsigned int N>
  __device__ __forceinline__ T getVariable(const char (&name)[N]) const {
    const int s = detail::find_slot(a.agent, detail::name_hash(name));
    if (s < 0) return T{};
    return reinterpret_cast<const T *>(a.agent.ptr[s])[index];
  }
  template <typename T, flamegpu::size_type N, unsigned int M>
  __device__ __forceinline__ T getVariable(const char (&name)[M], unsigned int i) const {
    const int s = detail::find_slot(a.agent, detail::name_hash(name));
    if (s < 0 || i >= N) return T{};
    return reinterpret_cast<const T *>(a.agent.ptr[s])[static_cast<size_t>(index) * N + i];
  }
  __device__ __forceinline__ id_t getID() const { return getVariable<id_t>("_id"); }
  __device__ __forceinline__ unsigned int getStepCounter() const { return a.d_step ? __ldg(a.d_step) : 0u; }
  const DeviceEnvironment environment;

 private:
  const detail::FunctionArgs &a;
  const unsigned int index;
};

#endif  // __CUDACC__

}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_RUNTIME_DEVICEAPI_CUH_
