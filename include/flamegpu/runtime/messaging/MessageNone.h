// flamegpu/runtime/messaging/MessageNone.h -- "no messages" specialisation
// (reference include/flamegpu/runtime/messaging/MessageNone.h + MessageNone/MessageNoneDevice.cuh).
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGENONE_H_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGENONE_H_

#include "flamegpu/runtime/detail/FunctionArgs.h"

namespace flamegpu {

class MessageNone {
 public:
  static constexpr int DIMS = 0;
  static constexpr bool SPATIAL = false;
  static constexpr bool HAS_OUTPUT = false;
#if defined(__CUDACC__)
  class In {
   public:
    __device__ __forceinline__ explicit In(const detail::FunctionArgs &, int = 0) {}
  };
  class Out {
   public:
    __device__ __forceinline__ Out(const detail::FunctionArgs &, unsigned int) {}
    __device__ __forceinline__ bool written() const { return false; }
  };
#endif
};

}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGENONE_H_
