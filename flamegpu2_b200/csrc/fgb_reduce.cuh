// fgb_reduce.cuh -- device reductions behind HostAgentAPI::sum/min/max (reference
// include/flamegpu/runtime/agent/HostAgentAPI.cuh:540-700: cub::DeviceReduce + a D2H copy of the result).
// One kernel: grid-stride accumulation, block reduction, per-block partials, and the LAST block to finish folds
// the partials in block order -- so the result does not depend on scheduling (floating-point sums are
// accumulated in double and are reproducible run to run).  The result stays in a device word.
#pragma once
#include "fgb_common.cuh"

namespace fgb {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 1024;

enum { kOpSum = 0, kOpMin = 1, kOpMax = 2 };

#ifdef __CUDACC__

template <typename A, int OP>
__device__ __forceinline__ A red_combine(A a, A b) {
  if (OP == kOpSum) return a + b;
  if (OP == kOpMin) return b < a ? b : a;
  return b > a ? b : a;
}

// element transforms applied before a SUM: none, "== value" (thrust::count, HostAgentAPI.cuh:700-718) and
// "(x - mean)^2" (standard_deviation_subtract_mean, HostAgentAPI.cuh:598-600)
enum { kTrNone = 0, kTrEqual = 1, kTrSqDev = 2 };
template <typename T, typename A, int TR>
__device__ __forceinline__ A red_transform(T v, double param, T tvalue) {
  if (TR == kTrEqual) return static_cast<A>(v == tvalue ? 1 : 0);
  if (TR == kTrSqDev) {
    const double d = static_cast<double>(v) - param;
    return static_cast<A>(d * d);
  }
  return static_cast<A>(v);
}

// T: element type; A: accumulator (double for floating-point sums, else T widened to 64 bits for integer sums)
template <typename T, typename A, int OP, int TR = kTrNone>
__global__ void __launch_bounds__(kRedThreads)
k_reduce(const T *__restrict__ in, uint32_t n_max, const unsigned int *d_n, A identity, A *partial, uint32_t *done, A *out,
         double tparam = 0.0, T tvalue = T()) {
  __shared__ A s_warp[kRedThreads / 32];
  __shared__ uint32_t s_last;
  const uint32_t n = load_count(d_n, n_max);
  A acc = identity;
  for (uint32_t i = blockIdx.x * kRedThreads + threadIdx.x; i < n; i += gridDim.x * kRedThreads)
    acc = red_combine<A, OP>(acc, red_transform<T, A, TR>(in[i], tparam, tvalue));
  auto block_reduce = [&](A v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_combine<A, OP>(v, __shfl_down_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    A r = identity;
    if (threadIdx.x < kRedThreads / 32) r = s_warp[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r = red_combine<A, OP>(r, __shfl_down_sync(0xffffffffu, r, o));
    }
    __syncthreads();
    return r;  // valid in thread 0
  };
  const A b = block_reduce(acc);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = b;
    __threadfence();
    s_last = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // fixed order: thread t folds partials t, t+256, ... ; then the same block reduction
  A v = identity;
  for (uint32_t p = threadIdx.x; p < gridDim.x; p += kRedThreads) v = red_combine<A, OP>(v, *reinterpret_cast<volatile A *>(partial + p));
  const A r = block_reduce(v);
  if (threadIdx.x == 0) {
    *out = r;
    *done = 0u;  // self-cleaning, like the look-back words
  }
}

#endif  // __CUDACC__
}  // namespace fgb
