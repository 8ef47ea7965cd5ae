// flamegpu/runtime/messaging/MessageBruteForce.cuh -- the generic message writer every spatial
// Out derives from (reference MessageBruteForce/MessageBruteForceDevice.cuh:260-276) and the
// all-to-all reader.  Only Out is on the spatial hot path; In is provided so that models mixing
// spatial and brute-force lists still compile.
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGEBRUTEFORCE_CUH_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGEBRUTEFORCE_CUH_

#include "flamegpu/runtime/detail/FunctionArgs.h"

namespace flamegpu {

class MessageBruteForce {
 public:
  class Description;  // host side
  static constexpr int DIMS = 0;
  static constexpr bool SPATIAL = false;
  static constexpr bool HAS_OUTPUT = true;
#if defined(__CUDACC__)
  class In {
   public:
    class Message {
      const detail::FunctionArgs &a;
      unsigned int idx;

     public:
      __device__ __forceinline__ Message(const detail::FunctionArgs &args, unsigned int i) : a(args), idx(i) {}
      __device__ __forceinline__ bool operator!=(const Message &rhs) const { return idx != rhs.idx; }
      __device__ __forceinline__ Message &operator++() {
        ++idx;
        return *this;
      }
      __device__ __forceinline__ Message &operator*() { return *this; }
      __device__ __forceinline__ unsigned int getIndex() const { return idx; }
      template <typename T, unsigned int N>
      __device__ __forceinline__ T getVariable(const char (&name)[N]) const {
        const int s = detail::find_slot(a.msg_in, detail::name_hash(name));
        if (s < 0) return T{};
        return __ldg(reinterpret_cast<const T *>(a.msg_in.ptr[s]) + idx);
      }
      template <typename T, flamegpu::size_type N, unsigned int M>
      __device__ __forceinline__ T getVariable(const char (&name)[M], unsigned int index) const {
        const int s = detail::find_slot(a.msg_in, detail::name_hash(name));
        if (s < 0 || index >= N) return T{};
        return __ldg(reinterpret_cast<const T *>(a.msg_in.ptr[s]) + static_cast<size_t>(idx) * N + index);
      }
    };
    __device__ __forceinline__ explicit In(const detail::FunctionArgs &args, int = 0) : a(args) {}
    __device__ __forceinline__ unsigned int size() const { return a.d_msg_in_count ? __ldg(a.d_msg_in_count) : 0u; }
    __device__ __forceinline__ Message begin() const { return Message(a, 0); }
    __device__ __forceinline__ Message end() const { return Message(a, size()); }

   private:
    const detail::FunctionArgs &a;
  };

  // Writes go straight to the output list at the thread's own slot (+ append offset).  Whether the
  // thread wrote anything is kept in a register and stored ONCE by the kernel wrapper; the reference
  // stores scan_flag[index] = 1 on every setVariable call (MessageBruteForceDevice.cuh:270-275) and
  // zeroes the flag array with two memsets before each function (CUDAScanCompaction.cu:49-56).
  class Out {
   public:
    __device__ __forceinline__ Out(const detail::FunctionArgs &args, unsigned int index)
        : a(args), slot(index + (args.d_msg_out_offset ? __ldg(args.d_msg_out_offset) : 0u)), wrote(false) {}
    template <typename T, unsigned int N>
    __device__ __forceinline__ void setVariable(const char (&name)[N], T value) const {
      const int s = detail::find_slot(a.msg_out, detail::name_hash(name));
      if (s >= 0) reinterpret_cast<T *>(a.msg_out.ptr[s])[slot] = value;
      wrote = true;
    }
    template <typename T, flamegpu::size_type N, unsigned int M>
    __device__ __forceinline__ void setVariable(const char (&name)[M], unsigned int index, T value) const {
      const int s = detail::find_slot(a.msg_out, detail::name_hash(name));
      if (s >= 0 && index < N) reinterpret_cast<T *>(a.msg_out.ptr[s])[static_cast<size_t>(slot) * N + index] = value;
      wrote = true;
    }
    __device__ __forceinline__ bool written() const { return wrote; }

   protected:
    const detail::FunctionArgs &a;
    unsigned int slot;
    mutable bool wrote;
  };
#endif
};

}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGEBRUTEFORCE_CUH_
