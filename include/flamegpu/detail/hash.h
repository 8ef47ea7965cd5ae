// flamegpu/detail/hash.h -- compile-time name hashing for device-side variable lookup.
//
// The reference resolves getVariable<T>("name") through cuRVE: a 512-slot open-addressed hash
// table that every block copies into shared memory (runtime/detail/curve/DeviceCurve.cuh:111-122)
// and probes on EVERY access (:342-352).  Here the name is hashed at compile time and matched
// against a small per-function table that lives in kernel parameter space (constant bank): the
// search is loop-invariant, so the compiler hoists it out of the message loop and each access
// costs one indexed load.
#ifndef FGB_INCLUDE_FLAMEGPU_DETAIL_HASH_H_
#define FGB_INCLUDE_FLAMEGPU_DETAIL_HASH_H_

#include <cstdint>
#include <string>

#if defined(__CUDACC__)
#define FGB_HD __host__ __device__ __forceinline__
#else
#define FGB_HD inline
#endif

namespace flamegpu {
namespace detail {

// FNV-1a, 32 bit, over the characters of a string literal (terminator excluded)
template <unsigned int N>
FGB_HD constexpr uint32_t name_hash(const char (&s)[N]) {
  uint32_t h = 2166136261u;
  for (unsigned int i = 0; i + 1 < N; ++i) h = (h ^ static_cast<uint32_t>(static_cast<unsigned char>(s[i]))) * 16777619u;
  return h ? h : 1u;  // 0 is reserved for "empty slot"
}
inline uint32_t name_hash_rt(const std::string &s) {
  uint32_t h = 2166136261u;
  for (char c : s) h = (h ^ static_cast<uint32_t>(static_cast<unsigned char>(c))) * 16777619u;
  return h ? h : 1u;
}

}  // namespace detail
}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_DETAIL_HASH_H_
