// fgb_radix.cuh -- stable LSD radix sort of (key, index) pairs for the automatic agent sort.
//
// Replaces cub::DeviceRadixSort::SortPairs as called by HostAgentAPI::sort_async
// (reference include/flamegpu/runtime/agent/HostAgentAPI.cuh:900-909).  The sort keys of the auto
// sort have few distinct values with hundreds of agents each (the reference's 3D key collapses to an
// (x,y) column index, see CUDASimulation.cu:487), which is the regime of a digit-partitioning sort:
//   one histogram kernel (all digit histograms in one read of the keys), then one single-pass
//   "onesweep" kernel per digit: per-tile stable ranking with warp match, decoupled look-back per
//   digit across tiles, direct scatter.  Digits are up to 9 bits wide (512 counters fit a block), so
//   a 12-bit key needs 2 passes and a 17-bit key 2 passes (CUB: 8-bit digits, 2 and 3 passes).
#pragma once
#include "fgb_common.cuh"
#include "fgb_compact.cuh"  // SortGeo, sort_key

namespace fgb {

constexpr int kRsThreads = 512;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsRounds = 8;                               // items per thread
constexpr int kRsWarpItems = 32 * kRsRounds;               // 256 consecutive items per warp
constexpr int kRsTile = kRsThreads * kRsRounds;            // 4096 items per tile
constexpr int kRsMaxBits = 9;
constexpr int kRsMaxDigits = 1 << kRsMaxBits;              // 512
constexpr int kRsMaxPasses = 4;
constexpr int kRsWindow = 4;                               // predecessors fetched together by the look-back

inline unsigned int radix_num_tiles(unsigned int n) { return (n + kRsTile - 1) / kRsTile; }

struct RadixPlan {
  int passes;
  int bits[kRsMaxPasses];
  int shift[kRsMaxPasses];
};
inline RadixPlan make_radix_plan(int max_bit) {
  RadixPlan p{};
  p.passes = (max_bit + kRsMaxBits - 1) / kRsMaxBits;
  if (p.passes < 1) p.passes = 1;
  int left = max_bit, sh = 0;
  for (int i = 0; i < p.passes; ++i) {
    const int b = (left + (p.passes - i) - 1) / (p.passes - i);  // spread the bits evenly
    p.bits[i] = b;
    p.shift[i] = sh;
    sh += b;
    left -= b;
  }
  return p;
}

#ifdef __CUDACC__

// ghist[pass][digit] += counts of every digit of every pass, one read of the keys.
__global__ void __launch_bounds__(kRsThreads) k_radix_hist(const uint32_t *__restrict__ keys, uint32_t n_max,
                                                           const unsigned int *d_n, RadixPlan plan, uint32_t *ghist) {
  __shared__ uint32_t sh[kRsMaxPasses][kRsMaxDigits];
  for (int i = threadIdx.x; i < kRsMaxPasses * kRsMaxDigits; i += kRsThreads) (&sh[0][0])[i] = 0u;
  __syncthreads();
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t stride = gridDim.x * kRsThreads;
  // warp-aggregated: lists arrive nearly sorted (agents were sorted a step ago), so the lanes of a warp mostly hold
  // the same digit; one shared-memory atomic per distinct digit of the warp instead of 32 on the same address
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t base = blockIdx.x * kRsThreads; base < n; base += stride) {  // block-uniform trip count
    const uint32_t i = base + threadIdx.x;
    const bool valid = i < n;
    const uint32_t k = valid ? ld_stream_u32(keys + i) : 0u;
    for (int p = 0; p < plan.passes; ++p) {
      const uint32_t d = (k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u);
      const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 0xFFFFFFFFu);
      if (valid && lane == static_cast<uint32_t>(__ffs(peers) - 1)) atomicAdd(&sh[p][d], static_cast<uint32_t>(__popc(peers)));
    }
  }
  __syncthreads();
  for (int p = 0; p < plan.passes; ++p)
    for (int d = threadIdx.x; d < (1 << plan.bits[p]); d += kRsThreads) {
      const uint32_t c = sh[p][d];
      if (c) atomicAdd(ghist + p * kRsMaxDigits + d, c);
    }
}

// calculateSpatialHash (CUDASimulation.cu:376-408) fused with the digit histograms: the key of every agent is
// computed, stored (it is the agent's _auto_sort_bin_index) and counted in the same pass.
template <int DIMS>
__global__ void __launch_bounds__(kRsThreads) k_sort_keys_hist(const float *__restrict__ x, const float *__restrict__ y,
                                                               const float *__restrict__ z, SortGeo g, uint32_t n_max,
                                                               const unsigned int *d_n, uint32_t *__restrict__ keys,
                                                               uint32_t key_mask, RadixPlan plan, uint32_t *ghist) {
  __shared__ uint32_t sh[kRsMaxPasses][kRsMaxDigits];
  for (int i = threadIdx.x; i < kRsMaxPasses * kRsMaxDigits; i += kRsThreads) (&sh[0][0])[i] = 0u;
  __syncthreads();
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t stride = gridDim.x * kRsThreads;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t base = blockIdx.x * kRsThreads; base < n; base += stride) {
    const uint32_t i = base + threadIdx.x;
    const bool valid = i < n;
    uint32_t k = 0u;
    if (valid) {
      k = sort_key<DIMS>(g, __ldg(x + i), __ldg(y + i), DIMS == 3 ? __ldg(z + i) : 0.f);
      keys[i] = k;
      k &= key_mask;
    }
    for (int p = 0; p < plan.passes; ++p) {
      const uint32_t d = (k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u);
      const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 0xFFFFFFFFu);
      if (valid && lane == static_cast<uint32_t>(__ffs(peers) - 1)) atomicAdd(&sh[p][d], static_cast<uint32_t>(__popc(peers)));
    }
  }
  __syncthreads();
  for (int p = 0; p < plan.passes; ++p)
    for (int d = threadIdx.x; d < (1 << plan.bits[p]); d += kRsThreads) {
      const uint32_t c = sh[p][d];
      if (c) atomicAdd(ghist + p * kRsMaxDigits + d, c);
    }
}

// One digit pass.  keys_in/idx_in -> keys_out/idx_out, stable.  idx_in == NULL: index = position.
// state: [tiles][digits] look-back words (2 flag bits | 30 value bits), all-zero on entry.
#ifndef FGB_ONESWEEP_MIN_BLOCKS
#define FGB_ONESWEEP_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(kRsThreads, FGB_ONESWEEP_MIN_BLOCKS) k_radix_onesweep(const uint32_t *__restrict__ keys_in,
                                                               const uint32_t *__restrict__ idx_in, uint32_t *keys_out,
                                                               uint32_t *idx_out, uint32_t n_max, const unsigned int *d_n,
                                                               int shift, int bits, uint32_t key_mask,
                                                               const uint32_t *__restrict__ ghist, uint32_t *state) {
  __shared__ uint32_t warp_cnt[kRsWarps][kRsMaxDigits];
  __shared__ uint32_t digit_base[kRsMaxDigits];
  __shared__ uint32_t scan_tmp[33];
  const int D = 1 << bits;
  const uint32_t dmask = static_cast<uint32_t>(D - 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n = load_count(d_n, n_max);
  const int tile = blockIdx.x;
  const uint32_t warp_base = static_cast<uint32_t>(tile) * kRsTile + warp * kRsWarpItems;

  for (int i = threadIdx.x; i < kRsWarps * D; i += kRsThreads) warp_cnt[i / D][i % D] = 0u;
  __syncthreads();

  // ---- stable rank of every item among the items of its digit inside the warp's 256-item slice
  uint32_t key[kRsRounds], idx[kRsRounds], rank[kRsRounds];
#pragma unroll
  for (int r = 0; r < kRsRounds; ++r) {
    const uint32_t i = warp_base + r * 32 + lane;
    const bool valid = i < n;
    key[r] = valid ? (ld_stream_u32(keys_in + i) & key_mask) : 0u;
    idx[r] = valid ? (idx_in ? ld_stream_u32(idx_in + i) : i) : 0u;
    const uint32_t d = (key[r] >> shift) & dmask;
    // invalid lanes form their own group (digit id D) and never touch the counters
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : static_cast<uint32_t>(D));
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) {
      base = warp_cnt[warp][d];
      warp_cnt[warp][d] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[r] = base + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();

  // ---- per digit: exclusive prefix over the warps of this tile, tile total, look-back across tiles
  uint32_t total = 0;
  if (threadIdx.x < D) {
    const int d = threadIdx.x;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      const uint32_t c = warp_cnt[w][d];
      warp_cnt[w][d] = total;
      total += c;
    }
    uint32_t *st = state + static_cast<size_t>(tile) * D + d;
    uint32_t excl = 0;
    if (tile == 0) {
      atomicExch(st, (2u << 30) | total);
    } else {
      atomicExch(st, (1u << 30) | total);
      // look-back in windows of kRsWindow predecessors: their words are fetched together (one L2 round trip per
      // window instead of one per predecessor: the serial chain of dependent loads was what bounded this kernel,
      // 30 % issue utilisation and 14 % DRAM at 16.8 M keys) and then consumed nearest first
      int p = tile - 1;
      bool done = false;
      while (!done) {
        uint32_t s[kRsWindow];
#pragma unroll
        for (int j = 0; j < kRsWindow; ++j)
          s[j] = p - j >= 0 ? *(reinterpret_cast<const volatile uint32_t *>(state) + static_cast<size_t>(p - j) * D + d) : (2u << 30);
#pragma unroll
        for (int j = 0; j < kRsWindow; ++j) {
          if (!done) {
            while ((s[j] >> 30) == 0u) s[j] = *(reinterpret_cast<const volatile uint32_t *>(state) + static_cast<size_t>(p - j) * D + d);
            excl += s[j] & 0x3FFFFFFFu;
            done = (s[j] >> 30) == 2u;
          }
        }
        p -= kRsWindow;
      }
      atomicExch(st, (2u << 30) | (excl + total));
    }
    digit_base[d] = excl;
  }
  // global start of every digit = exclusive scan of the digit totals (D <= 512 = blockDim)
  {
    const uint32_t g = threadIdx.x < D ? ghist[threadIdx.x] : 0u;
    uint32_t all;
    const uint32_t gexcl = block_exclusive_scan(g, scan_tmp, &all);
    if (threadIdx.x < D) digit_base[threadIdx.x] += gexcl;
  }
  __syncthreads();

  // ---- scatter
#pragma unroll
  for (int r = 0; r < kRsRounds; ++r) {
    const uint32_t i = warp_base + r * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[r] >> shift) & dmask;
      const uint32_t pos = digit_base[d] + warp_cnt[warp][d] + rank[r];
      if (keys_out) keys_out[pos] = key[r];
      idx_out[pos] = idx[r];
    }
  }
}

#endif  // __CUDACC__
}  // namespace fgb
