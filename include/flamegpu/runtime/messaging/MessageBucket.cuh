// flamegpu/runtime/messaging/MessageBucket.cuh -- device side of bucket messaging (integer keys lower..upper, every
// key a bucket), API compatible with the reference's include/flamegpu/runtime/messaging/MessageBucket/MessageBucketDevice.cuh.
// The index is the same PBM pipeline as spatial messaging with hash = key - lower (reference MessageBucket.cu:49-64,
// 105-137); here it is fgb_build_index_keys.  The bounds arrive by value in kernel-parameter space.
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGEBUCKET_CUH_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGEBUCKET_CUH_

#include "flamegpu/runtime/detail/FunctionArgs.h"
#include "flamegpu/runtime/messaging/MessageBruteForce.cuh"

namespace flamegpu {

typedef int IntT;  // reference MessageBucket.h:14

class MessageBucket {
 public:
  class Description;  // host side, flamegpu/model/ModelDescription.h
  static constexpr int DIMS = 0;
  static constexpr bool SPATIAL = false;
  static constexpr bool HAS_OUTPUT = true;
  // reference MessageBucket.h:40-55: min inclusive, max exclusive (= upper bound + 1)
  struct MetaData {
    IntT min;
    IntT max;
    unsigned int *PBM;
  };

#if defined(__CUDACC__)
  class In {
   public:
    class Filter {
     public:
      class Message {
        const detail::FunctionArgs &a;
        unsigned int idx;

       public:
        __device__ __forceinline__ Message(const detail::FunctionArgs &args, unsigned int i) : a(args), idx(i) {}
        __device__ __forceinline__ bool operator!=(const Message &rhs) const { return idx != rhs.idx; }
        __device__ __forceinline__ bool operator==(const Message &rhs) const { return idx == rhs.idx; }
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
      // Same acceptance test as the reference (MessageBucketDevice.cuh:264-274): note `endKey < max` with the
      // exclusive max, so the single-key form operator()(key) = Filter(key, key + 1) is empty for key == upper bound.
      __device__ __forceinline__ Filter(const detail::FunctionArgs &args, IntT beginKey, IntT endKey)
          : a(args), bucket_begin(0), bucket_end(0) {
        const IntT mn = args.in_meta.grid_dim[0], mx = args.in_meta.grid_dim[1];  // bucket lists: {min, max exclusive}
        if (beginKey >= mn && endKey < mx && beginKey <= endKey) {
          bucket_begin = __ldg(args.in_meta.pbm + (beginKey - mn));
          bucket_end = __ldg(args.in_meta.pbm + (endKey - mn));
        }
      }
      __device__ __forceinline__ Message begin() const { return Message(a, bucket_begin); }
      __device__ __forceinline__ Message end() const { return Message(a, bucket_end); }
      __device__ __forceinline__ unsigned int size() const { return bucket_end - bucket_begin; }

     private:
      const detail::FunctionArgs &a;
      unsigned int bucket_begin, bucket_end;
    };

    __device__ __forceinline__ explicit In(const detail::FunctionArgs &args, int = 0) : a(args) {}
    // the reference validates the keys only with FLAMEGPU_SEATBELTS (:201-232); this is the seatbelts-off build
    __device__ __forceinline__ Filter operator()(const IntT &key) const { return Filter(a, key, key + 1); }
    __device__ __forceinline__ Filter operator()(const IntT &beginKey, const IntT &endKey) const { return Filter(a, beginKey, endKey); }

   private:
    const detail::FunctionArgs &a;
  };

  class Out : public MessageBruteForce::Out {
   public:
    __device__ __forceinline__ Out(const detail::FunctionArgs &args, unsigned int index) : MessageBruteForce::Out(args, index) {}
    // reference MessageBucketDevice.cuh:280-294: the key is the message variable "_key"
    __device__ __forceinline__ void setKey(const IntT &key) const { this->template setVariable<IntT>("_key", key); }
  };
#endif  // __CUDACC__
};

}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGEBUCKET_CUH_
