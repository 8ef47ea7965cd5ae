// flamegpu/runtime/AgentFunction.cuh -- FLAMEGPU_AGENT_FUNCTION and the kernel that runs it.
// Same user-facing macro as the reference (runtime/AgentFunction_shim.cuh:32-40); the kernel
// takes ONE by-value argument block and needs no shared memory (the reference's wrapper fills a
// cuRVE table into shared memory and __syncthreads() before any agent code, AgentFunction.cuh:77-93).
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_AGENTFUNCTION_CUH_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_AGENTFUNCTION_CUH_

#include <typeindex>

#include "flamegpu/defines.h"
#include "flamegpu/runtime/DeviceAPI.cuh"
#include "flamegpu/runtime/detail/FunctionArgs.h"
#include "flamegpu/runtime/messaging/MessageBruteForce.cuh"
#include "flamegpu/runtime/messaging/MessageBucket.cuh"
#include "flamegpu/runtime/messaging/MessageNone.h"
#include "flamegpu/runtime/messaging/MessageSpatial2D.cuh"
#include "flamegpu/runtime/messaging/MessageSpatial3D.cuh"

namespace flamegpu {

typedef void(AgentFunctionWrapper)(const detail::FunctionArgs);
typedef void(AgentFunctionConditionWrapper)(const detail::FunctionArgs);

#if defined(__CUDACC__)
// ITER_MODE: spatial iterator variant compiled into this instance (0 reference visit order, 1 radius-filtered
// lock-step walk); it reaches the iterator as a constant, so each instance contains one variant only
template <typename AgentFunction, typename MessageIn, typename MessageOut, int ITER_MODE = 0>
__global__ void agent_function_wrapper(const __grid_constant__ detail::FunctionArgs args) {
  const unsigned int index = args.first_thread + blockIdx.x * blockDim.x + threadIdx.x;
  if (args.last_thread && index >= args.last_thread) return;
  unsigned int n = args.bound;
  if (args.d_count) {
    const unsigned int c = __ldg(args.d_count);
    n = c < n ? c : n;
  }
  // function condition: the agents that failed it sit at the front of the list and do not execute
  // (reference CUDAFatAgent::processFunctionCondition, CUDAFatAgent.cu:186-236)
  const unsigned int offset = args.d_agent_offset ? __ldg(args.d_agent_offset) : 0u;
  if (MessageOut::HAS_OUTPUT) {
    // mandatory output into an emptied list: every executing agent writes exactly one message, so the count is
    // known here (the reference derives it on the host, CUDAMessage.cu:196-205; a separate 1-thread kernel before)
    if (index == 0 && args.d_msg_out_count) {
      *args.d_msg_out_count = n > offset ? n - offset : 0u;
      if (args.d_keyed) *args.d_keyed = n > offset ? n - offset : 0u;
    }
  }
  if (index + offset >= n) return;
  // Run agents in the bin order of the input list when the scheduler provides it: lanes of a warp then
  // walk the same message strips (coalesced / broadcast loads).  Every per-agent slot (variables, scan
  // flags, message and new-agent slots) is addressed by the AGENT index, so results do not change.
  const bool permuted = args.exec_perm && (!args.d_perm_limit || index < __ldg(args.d_perm_limit));
  const unsigned int agent = permuted ? __ldg(args.exec_perm + index) : offset + index;
  if (agent >= n) return;  // a stale permutation must never address outside the list (the scheduler versions it)
  // message / new-agent slot: index among the executing agents, or -- for a mandatory output that the scheduler runs in
  // bin order -- the thread index, so that the list is WRITTEN bin-grouped (any bijection agent <-> slot is a legal
  // instance: the reference's order inside a PBM bin is atomic-arrival order)
  const unsigned int slot = args.slot_by_thread ? index : agent - offset;
  DeviceAPI<MessageIn, MessageOut> api(args, agent, slot, ITER_MODE);
  const AGENT_STATUS status = AgentFunction()(&api);
  // one flag store per thread, no memset beforehand (reference AgentFunction.cuh:111-119 +
  // CUDAScanCompaction::zero_async)
  if (args.death_flag) args.death_flag[agent] = static_cast<unsigned int>(status);
  if (MessageOut::HAS_OUTPUT) {
    if (args.msg_out_flag) args.msg_out_flag[slot] = api.message_out.written() ? 1u : 0u;
    if constexpr (MessageOut::SPATIAL) api.message_out.template publish_index<MessageOut::DIMS>();
  }
  api.agent_out.finalise();
}

// reference runtime/AgentFunctionCondition.cuh:45: the boolean goes into the AGENT_DEATH scan flags
template <typename AgentFunctionCondition>
__global__ void agent_function_condition_wrapper(const __grid_constant__ detail::FunctionArgs args) {
  const unsigned int index = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int n = args.bound;
  if (args.d_count) {
    const unsigned int c = __ldg(args.d_count);
    n = c < n ? c : n;
  }
  if (index >= n) return;
  ReadOnlyDeviceAPI api(args, index);
  args.death_flag[index] = AgentFunctionCondition()(&api) ? 1u : 0u;
}
#endif

}  // namespace flamegpu

#define FLAMEGPU_AGENT_FUNCTION(funcName, message_in, message_out)                                                      \
  struct funcName##_impl {                                                                                              \
    __device__ __forceinline__ flamegpu::AGENT_STATUS operator()(                                                       \
        flamegpu::DeviceAPI<message_in, message_out> *FLAMEGPU) const;                                                  \
    static constexpr flamegpu::AgentFunctionWrapper *fnPtr() {                                                          \
      return &flamegpu::agent_function_wrapper<funcName##_impl, message_in, message_out>;                               \
    }                                                                                                                   \
    static constexpr flamegpu::AgentFunctionWrapper *fnPtrFiltered() {                                                  \
      return &flamegpu::agent_function_wrapper<funcName##_impl, message_in, message_out, message_in::SPATIAL ? 1 : 0>;  \
    }                                                                                                                   \
    static std::type_index inType() { return std::type_index(typeid(message_in)); }                                     \
    static std::type_index outType() { return std::type_index(typeid(message_out)); }                                   \
  };                                                                                                                    \
  funcName##_impl funcName;                                                                                             \
  __device__ __forceinline__ flamegpu::AGENT_STATUS funcName##_impl::operator()(                                        \
      flamegpu::DeviceAPI<message_in, message_out> *FLAMEGPU) const

#define FLAMEGPU_AGENT_FUNCTION_CONDITION(funcName)                                                                     \
  struct funcName##_cdn_impl {                                                                                          \
    __device__ __forceinline__ bool operator()(flamegpu::ReadOnlyDeviceAPI *FLAMEGPU) const;                            \
    static constexpr flamegpu::AgentFunctionConditionWrapper *fnPtr() {                                                 \
      return &flamegpu::agent_function_condition_wrapper<funcName##_cdn_impl>;                                          \
    }                                                                                                                   \
  };                                                                                                                    \
  funcName##_cdn_impl funcName;                                                                                         \
  __device__ __forceinline__ bool funcName##_cdn_impl::operator()(flamegpu::ReadOnlyDeviceAPI *FLAMEGPU) const

#define FLAMEGPU_DEVICE_FUNCTION __device__ __forceinline__
#define FLAMEGPU_HOST_DEVICE_FUNCTION __host__ __device__ __forceinline__

#endif  // FGB_INCLUDE_FLAMEGPU_RUNTIME_AGENTFUNCTION_CUH_
