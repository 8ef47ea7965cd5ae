// fgb_compact.cuh -- single-pass stable flag compaction over all SoA variables, plus the plain
// data-movement kernels (append copy, default broadcast, sort-key computation).
//
// The reference compacts in three steps with two host round trips: cub ExclusiveSum over the
// scan flags (CUDAFatAgent.cu:118-132), scatter_generic (CUDAScatter.cu:67-88), then a D2H copy
// of position[n] and a stream sync (CUDAScatter.cu:175-178).  Here one kernel reads the flags,
// resolves the global rank of every kept item with a decoupled look-back over tile aggregates and
// moves every variable; the survivor count stays on the device.
#pragma once
#include "fgb_common.cuh"

namespace fgb {

constexpr int kCmpThreads = 256;
constexpr int kCmpItems = 8;
constexpr int kCmpTile = kCmpThreads * kCmpItems;

inline unsigned int compact_num_tiles(unsigned int n) { return (n + kCmpTile - 1) / kCmpTile; }

#ifdef __CUDACC__

// state[]: one look-back word per tile, all-zero on entry; done: zero on entry.  The last block
// to finish its look-back re-zeroes both, so consecutive launches (and CUDA-graph replays) need
// no memset in between.
//
// Layout: a tile is 8 warps x 256 items; in round r (0..7) lane l of warp w handles item
// tile0 + w*256 + r*32 + l.  Every load is a fully coalesced 128-byte warp access and, because the
// kept items of one round are ranked with a ballot, every store instruction writes one contiguous
// run of the output: no shared-memory staging and no alignment cases.  (The first version gave each
// thread 8 consecutive items: vector loads, but eight strided 4-byte stores per variable.)
// BULK variant (sm_90+ bulk async copies, "TMA" without a tensor map): the kept items of a tile form ONE contiguous
// range of the output, so instead of eight partially filled store instructions per thread and variable the tile is
// compacted in shared memory -- at the same offset modulo 16 bytes as its destination -- and leaves the SM as a
// single cp.async.bulk.global.shared::cta of up to 8 KB per variable (plus at most 3 + 3 scalar stores for the
// unaligned head and tail).  Two staging buffers alternate between consecutive variables; the issuing thread waits
// for the previous copy to have READ its buffer before the barrier that lets the block refill it.
__device__ __forceinline__ void bulk_store_s2g(void *gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef FGB_COMPACT_MIN_BLOCKS
#define FGB_COMPACT_MIN_BLOCKS 5  // 48 registers, no spills; 4 / 5 / 6 blocks per SM: 290 / 269 / 337 us at 16.8 M agents
#endif
template <bool BULK>
__global__ void __launch_bounds__(kCmpThreads, BULK ? 4 : FGB_COMPACT_MIN_BLOCKS)
k_compact(const uint32_t *__restrict__ flags, int invert, uint32_t n_max, const unsigned int *d_n, uint32_t keep_front,
          uint32_t out_offset, const unsigned int *d_out_offset, uint32_t out_limit, const __grid_constant__ VarTable vt,
          unsigned long long *state, uint32_t *done, uint32_t *d_out_count, uint32_t *d_out_total) {
  __shared__ uint32_t s_warp[kCmpThreads / 32];
  __shared__ uint32_t s_excl;
  __shared__ __align__(128) uint32_t s_stage[BULK ? 2 : 1][BULK ? kCmpTile + 8 : 1];
  // d_n and d_out_offset may alias d_out_total (callers pass a list's own count word for all three), which the last
  // tile of this very kernel writes: plain (volatile) loads, never the non-coherent path.  The write cannot overtake
  // a read: the last tile resolves its prefix only after every other tile has published, i.e. has read both words.
  const uint32_t n = load_count_coherent(d_n, n_max);
  const int tile = blockIdx.x;
  const bool last_tile = tile == static_cast<int>(gridDim.x) - 1;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t w0 = static_cast<uint32_t>(tile) * kCmpTile + warp * (32u * kCmpItems);
  if (d_out_offset) out_offset = *reinterpret_cast<const volatile unsigned int *>(d_out_offset);

  // ---- keep ballots of the warp's 8 rounds
  uint32_t bal[kCmpItems];
  uint32_t wcount = 0;
#pragma unroll
  for (int r = 0; r < kCmpItems; ++r) {
    const uint32_t i = w0 + r * 32u + lane;
    bool k = false;
    if (i < n) {
      k = i < keep_front;
      if (!k) k = (ld_stream_u32(flags + (i - keep_front)) == 1u) != (invert != 0);
    }
    bal[r] = __ballot_sync(0xFFFFFFFFu, k);
    wcount += __popc(bal[r]);
  }
  // the loads of the payload do not depend on the output rank: fetch the first variable now, so that the
  // tile-aggregate exchange and the look-back below overlap with memory latency instead of adding to it
  uint32_t keptbits = 0;  // bit r: this lane's item of round r is kept
#pragma unroll
  for (int r = 0; r < kCmpItems; ++r) keptbits |= ((bal[r] >> lane) & 1u) << r;
  const bool first4 = vt.n > 0 && vt.len[0] == 4;
  uint32_t val0[kCmpItems];
  if (first4) {
    const uint32_t *in0 = reinterpret_cast<const uint32_t *>(vt.in[0]);
#pragma unroll
    for (int r = 0; r < kCmpItems; ++r)
      if (keptbits & (1u << r)) val0[r] = ld_stream_u32(in0 + w0 + r * 32u + lane);
  }
  if (lane == 0) s_warp[warp] = wcount;
  __syncthreads();
  uint32_t wexcl = 0, agg = 0;
#pragma unroll
  for (uint32_t w = 0; w < kCmpThreads / 32; ++w) {
    const uint32_t c = s_warp[w];
    if (w < warp) wexcl += c;
    agg += c;
  }

  // ---- publish the tile aggregate, resolve the exclusive prefix
  if (threadIdx.x == 0) {
    st_state(state + tile, (tile == 0 ? kStInclusive : kStAggregate) | agg);
    if (tile == 0) s_excl = 0;
  }
  const bool need_prefix = agg != 0 || last_tile;  // empty tiles only publish
  if (threadIdx.x < 32 && tile > 0 && need_prefix) {
    const uint32_t e = lookback_exclusive(state, tile);
    if (threadIdx.x == 0) {
      st_state(state + tile, kStInclusive | static_cast<unsigned long long>(e + agg));
      s_excl = e;
    }
  }
  __syncthreads();
  const uint32_t tile_excl = s_excl;
  // ---- self-clean the look-back words.  This block will not read another block's word again, so thread 0 counts the
  // block as arrived now, with a fire-and-forget RED (no result, no register, nobody waits), and only looks at the
  // counter at the very end of the kernel: a block that then sees every block arrived knows that every look-back has
  // finished and re-zeroes the words with its first warp.  The block whose RED came last is guaranteed to see the full
  // count (unless an earlier finisher already cleaned up and reset it); several late finishers may all clean, which is
  // idempotent.  Release/acquire around the counter: this block's st_state() words must be visible before its arrival
  // is, and a cleaning block must observe every such word before it overwrites them (otherwise a zero could land first
  // and a stale {inclusive|value} word would survive into the next launch).
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(done, 1u);
  }
  if (last_tile && threadIdx.x == 0) {
    if (d_out_count) *d_out_count = tile_excl + agg;
    if (d_out_total) *d_out_total = out_offset + tile_excl + agg;
  }

  // ---- move the kept items of every variable
  // out_limit: only the first out_limit kept items are written (the counts still report all of them, so a
  // caller with a fixed-capacity destination can detect the overflow instead of corrupting memory)
  if (BULK ? agg != 0u : wcount != 0u) {  // BULK: block-uniform (the variable loop contains a barrier)
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t rank[kCmpItems];
    uint32_t run = tile_excl + wexcl;
    uint32_t mine = 0;  // bit r: this lane's item of round r is kept and fits under out_limit
#pragma unroll
    for (int r = 0; r < kCmpItems; ++r) {
      rank[r] = run + __popc(bal[r] & lt);
      if (((bal[r] >> lane) & 1u) && rank[r] < out_limit) mine |= 1u << r;
      run += __popc(bal[r]);
    }
    // software pipeline over the variables: the loads of variable v+1 are in flight while v is stored.  (Two variables
    // ahead needs 64 registers, i.e. 4 blocks per SM instead of 5: measured 281 us against 269 us at 16.8 M agents.)
    uint32_t cur[kCmpItems];
#pragma unroll
    for (int r = 0; r < kCmpItems; ++r) cur[r] = val0[r];
    bool cur4 = first4;
    // BULK: slot a + j of the staging buffer holds the tile's j-th kept item, a = destination index modulo 4, so the
    // 16-byte aligned middle [s_lo, s_hi) of the buffer and of the destination range coincide
    const uint32_t dst0 = out_offset + tile_excl;
    const uint32_t a = dst0 & 3u;
    const uint32_t nw = tile_excl >= out_limit ? 0u : (agg < out_limit - tile_excl ? agg : out_limit - tile_excl);
    const uint32_t s_lo = (a + 3u) & ~3u, s_hi = (a + nw) & ~3u;
    const bool bulk = BULK && s_hi > s_lo;
    for (uint32_t v = 0; v < vt.n; ++v) {
      const bool next4 = v + 1 < vt.n && vt.len[v + 1] == 4;
      uint32_t nxt[kCmpItems];
      if (next4) {
        const uint32_t *in = reinterpret_cast<const uint32_t *>(vt.in[v + 1]);
#pragma unroll
        for (int r = 0; r < kCmpItems; ++r)
          if (mine & (1u << r)) nxt[r] = ld_stream_u32(in + w0 + r * 32u + lane);
      }
      if (BULK && cur4) {
        uint32_t *stage = s_stage[v & 1u];
#pragma unroll
        for (int r = 0; r < kCmpItems; ++r)
          if (mine & (1u << r)) stage[a + (rank[r] - tile_excl)] = cur[r];
        fence_proxy_async_smem();
        if (threadIdx.x == 0) bulk_wait_read_all();  // the previous variable's copy has read the other buffer
        __syncthreads();
        uint32_t *o = reinterpret_cast<uint32_t *>(vt.out[v]);
        if (bulk) {
          if (threadIdx.x == 0) {
            bulk_store_s2g(o + (dst0 - a + s_lo), static_cast<uint32_t>(__cvta_generic_to_shared(stage + s_lo)), (s_hi - s_lo) * 4u);
            bulk_commit();
          } else if (threadIdx.x >= 32u && threadIdx.x < 40u) {
            // unaligned head [a, s_lo) and tail [s_hi, a + nw): at most three items each
            const uint32_t t = threadIdx.x - 32u;
            const uint32_t slot = t < 4u ? a + t : s_hi + (t - 4u);
            const bool ok = t < 4u ? slot < s_lo : slot < a + nw;
            if (ok) o[dst0 - a + slot] = stage[slot];
          }
        } else {
          for (uint32_t slot = a + threadIdx.x; slot < a + nw; slot += kCmpThreads) o[dst0 - a + slot] = stage[slot];
        }
      } else if (cur4) {
        uint32_t *o = reinterpret_cast<uint32_t *>(vt.out[v]) + out_offset;
#pragma unroll
        for (int r = 0; r < kCmpItems; ++r)
          if (mine & (1u << r)) o[rank[r]] = cur[r];
      } else {
#pragma unroll
        for (int r = 0; r < kCmpItems; ++r)
          if (mine & (1u << r)) copy_item(vt, v, w0 + r * 32u + lane, static_cast<size_t>(out_offset) + rank[r]);
      }
#pragma unroll
      for (int r = 0; r < kCmpItems; ++r) cur[r] = nxt[r];
      cur4 = next4;
    }
  }
  if (BULK && threadIdx.x == 0) bulk_wait_all();  // the staging buffers must outlive the copies
  if (threadIdx.x < 32) {
    uint32_t arrived = 0;
    if (threadIdx.x == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(arrived) : "l"(done) : "memory");
    arrived = __shfl_sync(0xFFFFFFFFu, arrived, 0);
    if (arrived == gridDim.x) {
      __threadfence();
      for (uint32_t t = threadIdx.x; t < gridDim.x; t += 32u) state[t] = 0ull;
      __syncwarp();
      if (threadIdx.x == 0) *done = 0u;
    }
  }
}

// scatter_all_generic (CUDAScatter.cu:105-117): out[off + i] = in[i]
template <bool VEC>
__global__ void __launch_bounds__(256) k_scatter_all(uint32_t n_max, const unsigned int *d_n, uint32_t out_offset,
                                                     const unsigned int *d_out_offset,
                                                     const __grid_constant__ VarTable vt) {
  const uint32_t n = load_count(d_n, n_max);
  if (d_out_offset) out_offset = __ldg(d_out_offset);
  const uint32_t i0 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (i0 >= n) return;
  const int cnt = (n - i0) < 4u ? static_cast<int>(n - i0) : 4;
  for (uint32_t v = 0; v < vt.n; ++v) {
    if (VEC && vt.len[v] == 4 && cnt == 4 && (out_offset & 3u) == 0) {
      const uint4 q = ld_stream_u4(vt.in[v] + static_cast<size_t>(i0) * 4);
      *reinterpret_cast<uint4 *>(vt.out[v] + (static_cast<size_t>(out_offset) + i0) * 4) = q;
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (t < cnt) copy_item(vt, v, i0 + t, static_cast<size_t>(out_offset) + i0 + t);
    }
  }
}

// broadcastInitKernel (CUDAScatter.cu:404-422): every item gets the default value at vt.in[v]
__global__ void __launch_bounds__(256) k_broadcast_init(uint32_t n, uint32_t out_offset,
                                                        const __grid_constant__ VarTable vt) {
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  for (uint32_t v = 0; v < vt.n; ++v) copy_item(vt, v, 0, static_cast<size_t>(out_offset) + i);
}

struct SortGeo {
  float min0, min1, min2;
  float w0, w1, w2;
  uint32_t g0, g1, g2;
};

// calculateSpatialHash (CUDASimulation.cu:376-408): floorf(((p-min)/width)*gridDim), no clamp.
template <int DIMS>
__device__ __forceinline__ uint32_t sort_key(const SortGeo &g, float x, float y, float z) {
  const int gx = static_cast<int>(floorf(__fmul_rn(__fdiv_rn(x - g.min0, g.w0), static_cast<float>(g.g0))));
  const int gy = static_cast<int>(floorf(__fmul_rn(__fdiv_rn(y - g.min1, g.w1), static_cast<float>(g.g1))));
  if (DIMS == 3) {
    const int gz = static_cast<int>(floorf(__fmul_rn(__fdiv_rn(z - g.min2, g.w2), static_cast<float>(g.g2))));
    return static_cast<uint32_t>(gz) * g.g0 * g.g1 + static_cast<uint32_t>(gy) * g.g0 + static_cast<uint32_t>(gx);
  }
  return static_cast<uint32_t>(gy) * g.g0 + static_cast<uint32_t>(gx);
}

template <int DIMS>
__global__ void __launch_bounds__(256) k_sort_keys(const float *__restrict__ x, const float *__restrict__ y,
                                                   const float *__restrict__ z, SortGeo g, uint32_t n_max,
                                                   const unsigned int *d_n, uint32_t *__restrict__ keys) {
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  keys[i] = sort_key<DIMS>(g, __ldg(x + i), __ldg(y + i), DIMS == 3 ? __ldg(z + i) : 0.f);
}

#endif  // __CUDACC__
}  // namespace fgb
