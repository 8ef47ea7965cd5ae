// fgb_scan.cuh -- single-pass exclusive scan (decoupled look-back) over u32 counters.
// Replaces cub::DeviceScan::ExclusiveSum at MessageSpatial3D.cu:134 / MessageSpatial2D.cu:134
// (histogram -> PBM) and the standalone scans of CUDAFatAgent.cu:118-132.
#pragma once
#include "fgb_common.cuh"

namespace fgb {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

inline unsigned int scan_num_tiles(unsigned int n) { return (n + kScanTile - 1) / kScanTile; }

#ifdef __CUDACC__
// in[0..n) -> out[shift + i] = sum_{j<i} in[j].  shift==0: additionally out[n] = total.
// shift==1: additionally out[0] = 0 (so out[1..n] holds the exclusive prefix and, after n
// atomicAdd(&out[k+1], c_k) increments, out[] is exactly the unshifted prefix array).
// zero_in: write zeros back over in[] (keeps the histogram clean for the next build).
// state[] must be all-zero on entry (one word per tile).
template <bool VEC>
__global__ void __launch_bounds__(kScanThreads) k_exclusive_scan(uint32_t *in, uint32_t *out, uint32_t n,
                                                                  unsigned long long *state, int shift, int zero_in) {
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t s_excl;
  const int tile = blockIdx.x;
  const uint32_t base = static_cast<uint32_t>(tile) * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  if (VEC && base + kScanItems <= n) {
    uint4 q = *reinterpret_cast<const uint4 *>(in + base);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    if (zero_in) *reinterpret_cast<uint4 *>(in + base) = make_uint4(0, 0, 0, 0);
  } else {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      v[j] = base + j < n ? in[base + j] : 0u;
      if (zero_in && base + j < n) in[base + j] = 0u;
    }
  }
  const uint32_t tsum = v[0] + v[1] + v[2] + v[3];
  uint32_t agg;
  uint32_t texcl = block_exclusive_scan(tsum, warp_sums, &agg);
  if (threadIdx.x == 0) {
    st_state(state + tile, (tile == 0 ? kStInclusive : kStAggregate) | agg);
    if (tile == 0) s_excl = 0;
  }
  if (tile > 0 && threadIdx.x < 32) {
    uint32_t e = lookback_exclusive(state, tile);
    if (threadIdx.x == 0) {
      st_state(state + tile, kStInclusive | static_cast<unsigned long long>(e + agg));
      s_excl = e;
    }
  }
  __syncthreads();
  const uint32_t excl = s_excl + texcl;
  uint32_t o[kScanItems];
  o[0] = excl;
  o[1] = o[0] + v[0];
  o[2] = o[1] + v[1];
  o[3] = o[2] + v[2];
  if (VEC && shift == 0 && base + kScanItems <= n) {
    *reinterpret_cast<uint4 *>(out + base) = make_uint4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
      if (base + j < n) out[base + j + shift] = o[j];
  }
  if (shift == 1 && tile == 0 && threadIdx.x == 0) out[0] = 0u;
  if (shift == 0 && tile == static_cast<int>(gridDim.x) - 1 && threadIdx.x == 0) out[n] = s_excl + agg;
}
#endif

}  // namespace fgb
