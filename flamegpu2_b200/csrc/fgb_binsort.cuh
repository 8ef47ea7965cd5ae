// fgb_binsort.cuh -- device kernels shared by the PBM build (fgb_build_index) and the automatic
// agent sort (fgb_sort_by_key): key histogram, cursor scatter, per-bin order fix-up, gather.
//
// Pipeline (every arrow is one kernel launch; nothing returns to the host):
//   keys -> k_bin_keys (bin key of every item stored once, RLE-aggregated RED atomics into a clean histogram; inside a
//           simulation step the list's writer publishes both and this launch does not exist)
//        -> k_scan_scatter: blocks [0,S) scan the histogram (shift=1: cursor[k+1] = start of bin k; histogram re-zeroed),
//           blocks [S,S+T) scatter one 2048-item tile each (dst = atomicAdd(&cursor[k+1], run); afterwards cursor[] IS
//           the prefix array, i.e. the PBM -- no separate cursor array, no second memset)
//   default order : the scatter moves the payload directly (arrival order inside a bin), any input order
//   stable order  : the scatter writes source indices (unordered tiles through k_bin_scatter_staged), k_fix_* sorts
//                   each bin's indices ascending (== source order), k_gather applies the permutation.
#pragma once
#include "fgb_common.cuh"

namespace fgb {

struct Geo {
  float min0, min1, min2, radius;
  int g0, g1, g2;     // GLOBAL grid dimensions (the clamp of the reference's getGridPosition3D)
  int z_off, z_cnt;   // slab window: planes [z_off, z_off + z_cnt) of the slowest axis are stored locally
};

#ifdef __CUDACC__

// getGridPosition3D + getHash3D (MessageSpatial3DDevice.cuh:646-672): IEEE divide, floorf, clamp.
template <int DIMS>
__device__ __forceinline__ uint32_t bin_key(const Geo &g, float x, float y, float z) {
  int cx = static_cast<int>(floorf(__fdiv_rn(x - g.min0, g.radius)));
  int cy = static_cast<int>(floorf(__fdiv_rn(y - g.min1, g.radius)));
  cx = cx < 0 ? 0 : (cx >= g.g0 ? g.g0 - 1 : cx);
  cy = cy < 0 ? 0 : (cy >= g.g1 ? g.g1 - 1 : cy);
  if (DIMS == 3) {
    int cz = static_cast<int>(floorf(__fdiv_rn(z - g.min2, g.radius)));
    cz = cz < 0 ? 0 : (cz >= g.g2 ? g.g2 - 1 : cz);
    cz -= g.z_off;  // bin arithmetic stays the global one; only the plane index is rebased
    cz = cz < 0 ? 0 : (cz >= g.z_cnt ? g.z_cnt - 1 : cz);
    return (static_cast<uint32_t>(cz) * g.g1 + cy) * g.g0 + cx;
  }
  cy -= g.z_off;
  cy = cy < 0 ? 0 : (cy >= g.z_cnt ? g.z_cnt - 1 : cy);
  return static_cast<uint32_t>(cy) * g.g0 + cx;
}

constexpr int kBinThreads = 256;
constexpr int kBinItems = 4;

// Key source: either positions (PBM build) or a ready key array masked to max_bit bits (agent sort).
template <int DIMS>  // DIMS 2/3: positions; DIMS 0: key array
struct KeySrc {
  const float *x, *y, *z;
  // DIMS == 0 (bucket lists, MessageBucket.cu:49-64): bin = keys[i] - key_min, clamped into [0, key_span) so that
  // an out-of-range key (a seatbelts-only error in the reference) cannot write outside the histogram
  const uint32_t *keys;
  uint32_t key_min, key_span;
  Geo g;
  __device__ __forceinline__ uint32_t int_key(uint32_t raw) const {
    const uint32_t k = raw - key_min;  // below the minimum wraps to a huge value and is clamped as well
    return k < key_span ? k : key_span - 1u;
  }
  template <bool VEC>
  __device__ __forceinline__ void load4(uint32_t i0, uint32_t n, uint32_t k[4]) const {
    if (VEC && i0 + 4 <= n) {
      if (DIMS == 0) {
        uint4 q = ld_stream_u4(keys + i0);
        k[0] = int_key(q.x); k[1] = int_key(q.y); k[2] = int_key(q.z); k[3] = int_key(q.w);
      } else {
        const float4 X = __ldg(reinterpret_cast<const float4 *>(x + i0));
        const float4 Y = __ldg(reinterpret_cast<const float4 *>(y + i0));
        float4 Z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (DIMS == 3) Z = __ldg(reinterpret_cast<const float4 *>(z + i0));
        k[0] = bin_key<DIMS == 0 ? 3 : DIMS>(g, X.x, Y.x, Z.x);
        k[1] = bin_key<DIMS == 0 ? 3 : DIMS>(g, X.y, Y.y, Z.y);
        k[2] = bin_key<DIMS == 0 ? 3 : DIMS>(g, X.z, Y.z, Z.z);
        k[3] = bin_key<DIMS == 0 ? 3 : DIMS>(g, X.w, Y.w, Z.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t i = i0 + j;
        if (i < n) {
          if (DIMS == 0) k[j] = int_key(__ldg(keys + i));
          else k[j] = bin_key<DIMS == 0 ? 3 : DIMS>(g, __ldg(x + i), __ldg(y + i), DIMS == 3 ? __ldg(z + i) : 0.f);
        } else {
          k[j] = 0xFFFFFFFFu;
        }
      }
    }
  }
};

// ---- block-local aggregation table ---------------------------------------------------------
// Every block first aggregates the keys of its tile in a shared-memory open-addressing table
// (shared-memory atomics), then touches global memory once per DISTINCT key of the tile.  Lists
// arrive grouped (agents are sorted every step), so a 2048-item tile holds a few hundred distinct
// bins: ~8x fewer L2 atomics than one per item, and none of them on the per-item critical path.
constexpr int kTileItems = 8;                               // items per thread
constexpr int kTile = kBinThreads * kTileItems;             // 2048 items per block
constexpr int kTabBits = 12;
constexpr int kTabSlots = 1 << kTabBits;                    // 4096 slots (load factor <= 0.5)
constexpr uint32_t kTabEmpty = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t tab_insert(uint32_t *s_key, uint32_t key) {
  // low key bits keep x-adjacent bins in adjacent slots (coalesced write-out of the staged tile); the high
  // bits are mixed in additively so that keys that differ by a multiple of the table size (the same column
  // in another z plane of a power-of-two grid) do not pile up on one probe chain
  uint32_t slot = (key + (key >> kTabBits) * 0x9E3779B1u) & (kTabSlots - 1);
  // Double hashing with a long odd stride.  Because adjacent keys sit in adjacent slots, the keys of a tile form
  // runs of hundreds of occupied slots; with linear probing two overlapping runs (2D lists: neighbouring rows
  // of the grid) cost ~250 probes per insert (measured: 20x slower builds).  A long stride leaves the run at once.
  const uint32_t step = ((key * 0x85EBCA6Bu) >> (32 - kTabBits)) | 0x401u;
  while (true) {
    const uint32_t prev = atomicCAS(s_key + slot, kTabEmpty, key);
    if (prev == kTabEmpty || prev == key) return slot;
    slot = (slot + step) & (kTabSlots - 1);
  }
}

template <int DIMS, bool VEC>
__device__ __forceinline__ void load_tile_keys(const KeySrc<DIMS> &src, uint32_t i0, uint32_t n, uint32_t k[kTileItems]) {
  // a thread's 8 consecutive items with ONE 256-bit load per axis: a warp request then covers 1 KB contiguously
  // (two 128-bit loads per thread touch every sector twice; ncu showed the L1 at 70-85 % in these kernels)
  if (VEC && DIMS != 0 && i0 + 8 <= n && ((reinterpret_cast<uintptr_t>(src.x + i0) | reinterpret_cast<uintptr_t>(src.y + i0) |
                                          (DIMS == 3 ? reinterpret_cast<uintptr_t>(src.z + i0) : 0)) & 31u) == 0) {
    float X[8], Y[8], Z[8];
    ld_nc_f8(src.x + i0, X);
    ld_nc_f8(src.y + i0, Y);
    if (DIMS == 3) ld_nc_f8(src.z + i0, Z);
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] = bin_key<DIMS == 0 ? 3 : DIMS>(src.g, X[j], Y[j], DIMS == 3 ? Z[j] : 0.f);
    return;
  }
  src.template load4<VEC>(i0, n, k);
  src.template load4<VEC>(i0 + 4, n, k + 4);
}

// Number of runs of equal keys among a thread's consecutive items, summed over the block.  A tile
// of a bin-sorted list has few runs (neighbouring items share a bin, and output positions of
// consecutive items are consecutive); a tile of a list in some other order has one run per item.
__device__ __forceinline__ uint32_t block_run_count(const uint32_t k[kTileItems], int cnt, uint32_t *s_counter) {
  uint32_t runs = cnt > 0 ? 1u : 0u;
#pragma unroll
  for (int t = 1; t < kTileItems; ++t)
    if (t < cnt && k[t] != k[t - 1]) ++runs;
  if (threadIdx.x == 0) *s_counter = 0u;
  __syncthreads();
  const uint32_t w = __reduce_add_sync(0xffffffffu, runs);
  if ((threadIdx.x & 31) == 0) atomicAdd(s_counter, w);
  __syncthreads();
  return *s_counter;
}

// ---- phase 1: keys + histogram --------------------------------------------------------------
// keys[i] = bin of item i (stored once; the scatter kernels never re-hash positions) and hist[key] += 1 for the items
// [first, n), first = *d_first (NULL: 0): the range form lets a list whose leading part was already counted by its
// writer (agent_function_wrapper publishes key + histogram for mandatory spatial output, AgentFunction.cuh) take
// further items (ghost messages appended by the slab exchange) without recounting.  One RED per run of equal keys of
// a thread's 8 consecutive items: lists arrive grouped (the writer runs in bin order), so ~1 RED per bin visit.
// Also zeroes the look-back words of the scan that follows and the big-bin counter.  hist[] must hold only the
// counts of items [0, first) on entry; the scan re-zeroes it.
template <int DIMS, bool VEC>
__global__ void __launch_bounds__(kBinThreads) k_bin_keys(KeySrc<DIMS> src, uint32_t n_max, const unsigned int *d_n,
                                                          const unsigned int *d_first, uint32_t *__restrict__ keys,
                                                          uint32_t *hist, unsigned long long *state, uint32_t n_state,
                                                          uint32_t *ctrl) {
  const uint32_t gtid = blockIdx.x * kBinThreads + threadIdx.x;
  const uint32_t total = gridDim.x * kBinThreads;
  for (uint32_t s = gtid; s < n_state; s += total) state[s] = 0ull;
  if (gtid == 0 && ctrl) ctrl[0] = 0u;
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t first = d_first ? __ldg(d_first) : 0u;
  // the first item of a thread stays a multiple of 8 (aligned vector accesses); items below `first` are skipped
  const uint32_t i0 = (first & ~7u) + gtid * kTileItems;
  if (i0 >= n) return;
  const int cnt = (n - i0) < static_cast<uint32_t>(kTileItems) ? static_cast<int>(n - i0) : kTileItems;
  uint32_t k[kTileItems];
  load_tile_keys<DIMS, VEC>(src, i0, n, k);
  const int lo = first > i0 ? static_cast<int>(first - i0) : 0;  // > 0 only in the thread that straddles `first`
  if (VEC && lo == 0 && cnt == kTileItems && (reinterpret_cast<uintptr_t>(keys + i0) & 31u) == 0) {
    st_u8(keys + i0, k);
  } else {
#pragma unroll
    for (int t = 0; t < kTileItems; ++t)
      if (t >= lo && t < cnt) keys[i0 + t] = k[t];
  }
  int j = lo;
#pragma unroll
  for (int r = 0; r < kTileItems; ++r) {
    if (r == j && j < cnt) {
      int e = j + 1;
#pragma unroll
      for (int t = 1; t < kTileItems; ++t)
        if (t < kTileItems - r && r + t < cnt && e == r + t && k[(r + t) & (kTileItems - 1)] == k[r]) e = r + t + 1;
      atomicAdd(hist + k[r], static_cast<uint32_t>(e - j));
      j = e;
    }
  }
}

// ---- phases 2 + 3: scan and scatter in ONE launch -------------------------------------------
// k_scan_scatter: blocks [0, scan_tiles) turn the histogram into the shifted prefix array (cursor[k+1] = start of bin
// k; single-pass decoupled look-back, the histogram is re-zeroed on the way), blocks [scan_tiles, grid) scatter one
// 2048-item tile each.  A scatter block loads and classifies its tile and has the first payload variable in flight
// BEFORE it needs the cursors; only then does it wait for the scan (one acquire-load spin by thread 0 on the count of
// finished scan tiles).  Blocks are dispatched in index order, so every scan tile is resident or done before the
// first scatter block exists: the wait cannot deadlock.  Against two launches this removes one kernel boundary and
// hides the scan behind the scatter's own loads (1 M messages: 8 + 14 us -> one kernel).
//
// Grouped tile (the steady state of a bin-sorted list): a warp owns 256 consecutive items and walks them in 8 rounds
// of 32; runs of equal keys inside a round are found with one shuffle and one ballot, the first lane of a run claims
// the run's output range with ONE global atomic (dst = atomicAdd(&cursor[k+1], run); afterwards cursor[] IS the PBM)
// and the lanes of the run store to consecutive addresses.  Ungrouped tile (any other order: one run per item): its
// index is queued for k_bin_scatter_staged, which runs next over that worklist.
// ctrl[]: [0] big-bin counter of the stable fix-up, [1] finished scan tiles, [2] finished scatter blocks,
//         [3] length of the staged worklist.  The last scatter block re-zeroes [1], [2] and the look-back words.
constexpr int kFsItems = 16;                          // counters per thread of a scan block
constexpr int kFsTile = kBinThreads * kFsItems;       // 4096 counters per scan tile (== kScanTile)

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void scan_tile_256x16(uint32_t *hist, uint32_t *out, uint32_t bins, unsigned long long *state,
                                                 int tile, uint32_t *warp_sums, uint32_t *s_excl) {
  const uint32_t base = static_cast<uint32_t>(tile) * kFsTile + threadIdx.x * kFsItems;
  uint32_t v[kFsItems];
  const bool full = base + kFsItems <= bins;
  if (full) {
#pragma unroll
    for (int q = 0; q < kFsItems / 4; ++q) {
      const uint4 w = *reinterpret_cast<const uint4 *>(hist + base + 4 * q);
      v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
      *reinterpret_cast<uint4 *>(hist + base + 4 * q) = make_uint4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int j = 0; j < kFsItems; ++j) {
      v[j] = base + j < bins ? hist[base + j] : 0u;
      if (base + j < bins) hist[base + j] = 0u;
    }
  }
  uint32_t tsum = 0;
#pragma unroll
  for (int j = 0; j < kFsItems; ++j) tsum += v[j];
  uint32_t agg;
  const uint32_t texcl = block_exclusive_scan(tsum, warp_sums, &agg);
  if (threadIdx.x == 0) {
    st_state(state + tile, (tile == 0 ? kStInclusive : kStAggregate) | agg);
    if (tile == 0) *s_excl = 0;
  }
  if (tile > 0 && threadIdx.x < 32) {
    const uint32_t e = lookback_exclusive(state, tile);
    if (threadIdx.x == 0) {
      st_state(state + tile, kStInclusive | static_cast<unsigned long long>(e + agg));
      *s_excl = e;
    }
  }
  __syncthreads();
  // out[base + j + 1] = exclusive prefix of item j (shifted by one): o[j] below is the value for out[base + 1 + j]
  uint32_t run = *s_excl + texcl;
  uint32_t o[kFsItems];
#pragma unroll
  for (int j = 0; j < kFsItems; ++j) {
    o[j] = run;
    run += v[j];
  }
  if (full) {
    // out[base+1 .. base+3], three aligned 128-bit stores for out[base+4 .. base+15], out[base+16]
    out[base + 1] = o[0];
    out[base + 2] = o[1];
    out[base + 3] = o[2];
#pragma unroll
    for (int q = 1; q < kFsItems / 4; ++q)
      *reinterpret_cast<uint4 *>(out + base + 4 * q) = make_uint4(o[4 * q - 1], o[4 * q], o[4 * q + 1], o[4 * q + 2]);
    out[base + kFsItems] = o[kFsItems - 1];
  } else {
#pragma unroll
    for (int j = 0; j < kFsItems; ++j)
      if (base + j < bins) out[base + j + 1] = o[j];
  }
  if (tile == 0 && threadIdx.x == 0) out[0] = 0u;
}

#ifndef FGB_SCATTER_MIN_BLOCKS
#define FGB_SCATTER_MIN_BLOCKS 5  // 48 registers instead of 64, no spills
#endif
template <bool IDX_ONLY>
__global__ void __launch_bounds__(kBinThreads, FGB_SCATTER_MIN_BLOCKS) k_scan_scatter(uint32_t *hist, uint32_t *cursor, uint32_t bins,
                                                              unsigned long long *state, uint32_t scan_tiles,
                                                              const uint32_t *__restrict__ keys, uint32_t n_max,
                                                              const unsigned int *d_n, const __grid_constant__ VarTable vt,
                                                              uint32_t *perm, uint32_t *worklist, uint32_t *ctrl, uint32_t inline_ungrouped) {
  __shared__ uint32_t s_runs;
  __shared__ uint32_t s_scan[33];
  __shared__ uint32_t s_excl;
  if (blockIdx.x < scan_tiles) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      ctrl[0] = 0u;  // big-bin counter of the stable fix-up that may follow
      ctrl[3] = 0u;  // staged worklist (scatter blocks append only after the scan has finished)
    }
    scan_tile_256x16(hist, cursor, bins, state, static_cast<int>(blockIdx.x), s_scan, &s_excl);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(ctrl + 1, 1u);
    return;
  }
  const uint32_t tile = blockIdx.x - scan_tiles, scatter_blocks = gridDim.x - scan_tiles;
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t tile0 = tile * kTile;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t w0 = tile0 + warp * (32u * kTileItems);
  bool grouped = false;
  uint32_t k[kTileItems], head[kTileItems];  // head: ballot of the lanes that start a run (bit 0 always set)
  uint32_t cur[kTileItems];
  bool cur4 = false;
  if (tile0 < n) {
    if (threadIdx.x == 0) s_runs = 0u;
    __syncthreads();
    const uint32_t tile_n = (n - tile0) < static_cast<uint32_t>(kTile) ? n - tile0 : kTile;
    uint32_t runs = 0;
#pragma unroll
    for (int r = 0; r < kTileItems; ++r) {
      const uint32_t i = w0 + r * 32u + lane;
      k[r] = i < n ? __ldg(keys + i) : 0xFFFFFFFFu;
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, k[r], 1);
      head[r] = __ballot_sync(0xFFFFFFFFu, lane == 0 || prev != k[r]);
      runs += __popc(head[r] & __ballot_sync(0xFFFFFFFFu, i < n));
    }
    if (lane == 0) atomicAdd(&s_runs, runs);
    __syncthreads();
    grouped = s_runs * 2u <= tile_n;
    // the first payload variable does not depend on the cursors: its loads overlap the wait for the scan
    if (!IDX_ONLY && grouped) {
      cur4 = vt.n > 0 && vt.len[0] == 4;
      if (cur4) {
        const uint32_t *in = reinterpret_cast<const uint32_t *>(vt.in[0]);
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) {
          const uint32_t i = w0 + r * 32u + lane;
          if (i < n) cur[r] = ld_stream_u32(in + i);
        }
      }
    }
    // wait for the scan
    if (threadIdx.x == 0)
      while (ld_acquire_u32(ctrl + 1) < scan_tiles) __nanosleep(40);
    __syncthreads();
    if (!grouped && !inline_ungrouped && threadIdx.x == 0) worklist[atomicAdd(ctrl + 3, 1u)] = tile;
  }
  if (tile0 < n && (grouped || inline_ungrouped)) {
    // claim: the first lane of every run adds the run's length to the bin's cursor.  All eight rounds' atomics are
    // issued before the first result is consumed (interleaved with the shuffles below, each round waited a full L2
    // round trip for its atomic before the next one was issued: 8 serial round trips per tile)
    // (An ungrouped tile handled inline -- FGB_BUILD_EXPECT_GROUPED -- takes the same code: with almost every lane a run
    // head it degenerates to one atomic and one scattered store per message, correct for any input.)
    uint32_t dst[kTileItems];
    if (!IDX_ONLY && !grouped) {  // nothing was prefetched for this tile
      cur4 = vt.n > 0 && vt.len[0] == 4;
      if (cur4) {
        const uint32_t *in = reinterpret_cast<const uint32_t *>(vt.in[0]);
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) {
          const uint32_t i = w0 + r * 32u + lane;
          if (i < n) cur[r] = ld_stream_u32(in + i);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kTileItems; ++r) {
      const uint32_t i = w0 + r * 32u + lane;
      const uint32_t below = head[r] & ((2u << lane) - 1u);            // run heads at or below this lane
      const int first = 31 - __clz(static_cast<int>(below));           // lane that starts this lane's run
      const uint32_t above = head[r] & ~((2u << lane) - 1u);           // run heads above this lane
      const int next = above ? __ffs(static_cast<int>(above)) - 1 : 32;  // first lane of the next run
      dst[r] = 0;
      if (static_cast<int>(lane) == first && i < n) {
        // the run ends at the next head or at the end of the list
        const uint32_t last_valid = (n - (w0 + r * 32u)) < 32u ? n - (w0 + r * 32u) : 32u;
        const uint32_t len = (static_cast<uint32_t>(next) < last_valid ? static_cast<uint32_t>(next) : last_valid) - lane;
        dst[r] = atomicAdd(cursor + k[r] + 1, len);
      }
    }
#pragma unroll
    for (int r = 0; r < kTileItems; ++r) {
      const uint32_t below = head[r] & ((2u << lane) - 1u);
      const int first = 31 - __clz(static_cast<int>(below));
      dst[r] = __shfl_sync(0xFFFFFFFFu, dst[r], first) + (lane - static_cast<uint32_t>(first));
    }
    if constexpr (IDX_ONLY) {
#pragma unroll
      for (int r = 0; r < kTileItems; ++r) {
        const uint32_t i = w0 + r * 32u + lane;
        if (i < n) perm[dst[r]] = i;
      }
    } else {
      // software pipeline over the variables: the loads of variable v+1 are in flight while v is stored
      for (uint32_t v = 0; v < vt.n; ++v) {
        const bool next4 = v + 1 < vt.n && vt.len[v + 1] == 4;
        uint32_t nxt[kTileItems];
        if (next4) {
          const uint32_t *in = reinterpret_cast<const uint32_t *>(vt.in[v + 1]);
#pragma unroll
          for (int r = 0; r < kTileItems; ++r) {
            const uint32_t i = w0 + r * 32u + lane;
            if (i < n) nxt[r] = ld_stream_u32(in + i);
          }
        }
        if (cur4) {
          uint32_t *o = reinterpret_cast<uint32_t *>(vt.out[v]);
#pragma unroll
          for (int r = 0; r < kTileItems; ++r) {
            const uint32_t i = w0 + r * 32u + lane;
            if (i < n) o[dst[r]] = cur[r];
          }
        } else {
#pragma unroll
          for (int r = 0; r < kTileItems; ++r) {
            const uint32_t i = w0 + r * 32u + lane;
            if (i < n) copy_item(vt, v, i, dst[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) cur[r] = nxt[r];
        cur4 = next4;
      }
      if (perm) {  // source slot of every sorted item
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) {
          const uint32_t i = w0 + r * 32u + lane;
          if (i < n) perm[dst[r]] = i;
        }
      }
    }
  }
  // the last scatter block to finish cleans up for the next build (it first makes sure the scan has finished: a
  // list whose device count is 0 lets every scatter block get here without having waited)
  __syncthreads();
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(ctrl + 2, 1u) == scatter_blocks - 1u ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0)
      while (ld_acquire_u32(ctrl + 1) < scan_tiles) __nanosleep(40);
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < scan_tiles; t += kBinThreads) state[t] = 0ull;
    if (threadIdx.x == 0) {
      ctrl[1] = 0u;
      ctrl[2] = 0u;
    }
  }
}

// Ungrouped tiles: shared-memory table + staged, coalesced write-out.  (1) rank the items per key in the
// shared-memory table, (2) claim one contiguous output range per distinct key with a single global atomic, (3) lay
// the tile out in TABLE order in shared memory (the table is indexed by the low key bits, so x-adjacent bins --
// adjacent in the output -- sit in adjacent slots), (4) write it out with consecutive lanes on consecutive staged
// items, so stores fill whole 32-byte sectors instead of one sector per 4 bytes.
// Persistent blocks walk the worklist that k_scan_scatter left (ctrl[3] entries): for a bin-ordered list the list is
// empty and the launch costs a few hundred threads reading one word.
template <bool VEC, bool IDX_ONLY>
__device__ __forceinline__ void staged_tile(uint32_t tile, const uint32_t *__restrict__ keys, uint32_t n, uint32_t *cursor,
                                            const VarTable &vt, uint32_t *perm, uint32_t *s_key, uint32_t *s_cnt, uint32_t *s_dst,
                                            uint16_t *s_src, uint32_t *s_scan);

template <bool VEC, bool IDX_ONLY>
__global__ void __launch_bounds__(kBinThreads) k_bin_scatter_staged(const uint32_t *__restrict__ keys, uint32_t n_max,
                                                                    const unsigned int *d_n, uint32_t *cursor,
                                                                    const __grid_constant__ VarTable vt, uint32_t *perm,
                                                                    const uint32_t *__restrict__ worklist, const uint32_t *ctrl) {
  __shared__ uint32_t s_key[kTabSlots];   // key of the slot, later the slot's offset in the staged tile
  __shared__ uint32_t s_cnt[kTabSlots];   // count of the key in this tile, later its global base
  __shared__ uint32_t s_dst[kTile];       // staged tile: destination index ...
  __shared__ uint16_t s_src[kTile];       // ... and source item (offset inside the tile)
  __shared__ uint32_t s_scan[33];
  const uint32_t count = ctrl[3];
  const uint32_t n = load_count(d_n, n_max);
  for (uint32_t w = blockIdx.x; w < count; w += gridDim.x) {
    staged_tile<VEC, IDX_ONLY>(worklist[w], keys, n, cursor, vt, perm, s_key, s_cnt, s_dst, s_src, s_scan);
    __syncthreads();
  }
}

template <bool VEC, bool IDX_ONLY>
__device__ __forceinline__ void staged_tile(uint32_t tile, const uint32_t *__restrict__ keys, uint32_t n, uint32_t *cursor,
                                            const VarTable &vt, uint32_t *perm, uint32_t *s_key, uint32_t *s_cnt, uint32_t *s_dst,
                                            uint16_t *s_src, uint32_t *s_scan) {
  const uint32_t tile0 = tile * kTile;
  const uint32_t i0 = tile0 + threadIdx.x * kTileItems;
  const int cnt = i0 < n ? ((n - i0) < static_cast<uint32_t>(kTileItems) ? static_cast<int>(n - i0) : kTileItems) : 0;
  const uint32_t tile_n = (n - tile0) < static_cast<uint32_t>(kTile) ? n - tile0 : kTile;
  uint32_t k[kTileItems];
  if (cnt == kTileItems && VEC) {
    ld_nc_u8(keys + i0, k);
  } else {
#pragma unroll
    for (int t = 0; t < kTileItems; ++t) k[t] = t < cnt ? __ldg(keys + i0 + t) : 0xFFFFFFFFu;
  }

  for (int s = threadIdx.x; s < kTabSlots; s += kBinThreads) {
    s_key[s] = kTabEmpty;
    s_cnt[s] = 0u;
  }
  __syncthreads();
  uint32_t slot[kTileItems], rank[kTileItems];
  {
    int j = 0;
#pragma unroll
    for (int r = 0; r < kTileItems; ++r) {
      if (r == j && j < cnt) {
        int e = j + 1;
#pragma unroll
        for (int t = 1; t < kTileItems; ++t)
          if (t < kTileItems - r && r + t < cnt && e == r + t && k[(r + t) & (kTileItems - 1)] == k[r]) e = r + t + 1;
        const uint32_t sl = tab_insert(s_key, k[r]);
        const uint32_t b = atomicAdd(s_cnt + sl, static_cast<uint32_t>(e - j));
#pragma unroll
        for (int t = 0; t < kTileItems; ++t)
          if (r + t < e && t < kTileItems - r) {
            slot[r + t] = sl;
            rank[r + t] = b + t;
          }
        j = e;
      }
    }
  }
  __syncthreads();
  // per slot: claim the global range (one atomic per distinct key) and compute the staging offset
  {
    constexpr int kPer = kTabSlots / kBinThreads;  // 16 consecutive slots per thread
    uint32_t c[kPer], sum = 0;
    const int s0 = threadIdx.x * kPer;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const uint32_t key = s_key[s0 + q];
      c[q] = key != kTabEmpty ? s_cnt[s0 + q] : 0u;
      if (c[q]) s_cnt[s0 + q] = atomicAdd(cursor + key + 1, c[q]);
      sum += c[q];
    }
    uint32_t tile_total;
    uint32_t off = block_exclusive_scan(sum, s_scan, &tile_total);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      s_key[s0 + q] = off;
      off += c[q];
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < kTileItems; ++t)
    if (t < cnt) {
      const uint32_t e = s_key[slot[t]] + rank[t];
      s_dst[e] = s_cnt[slot[t]] + rank[t];
      s_src[e] = static_cast<uint16_t>(threadIdx.x * kTileItems + t);
    }
  __syncthreads();
  if constexpr (IDX_ONLY) {
    for (uint32_t e = threadIdx.x; e < tile_n; e += kBinThreads) perm[s_dst[e]] = tile0 + s_src[e];
  } else {
    // Payload: read in source order (coalesced, 8 consecutive items per thread), permuted inside shared memory,
    // written in staged order.  (Gathering `in[s_src[e]]` straight from global memory costs one L1 sector per
    // 4-byte item: ncu showed 29 sectors per request in that version.)
    uint32_t pos[kTileItems];
#pragma unroll
    for (int t = 0; t < kTileItems; ++t) pos[t] = t < cnt ? s_key[slot[t]] + rank[t] : 0u;
    uint32_t *s_val = s_cnt;  // the per-slot global bases were folded into s_dst above: the array is free again
    for (uint32_t v = 0; v < vt.n; ++v) {
      if (vt.len[v] == 4) {
        const uint32_t *in = reinterpret_cast<const uint32_t *>(vt.in[v]);
        uint32_t *o = reinterpret_cast<uint32_t *>(vt.out[v]);
        uint32_t w[kTileItems];
        if (VEC && cnt == kTileItems && (reinterpret_cast<uintptr_t>(in + i0) & 31u) == 0) {
          ld_nc_u8(in + i0, w);
        } else {
#pragma unroll
          for (int t = 0; t < kTileItems; ++t)
            if (t < cnt) w[t] = __ldg(in + i0 + t);
        }
        __syncthreads();  // the previous variable's write-out has finished reading s_val
#pragma unroll
        for (int t = 0; t < kTileItems; ++t)
          if (t < cnt) s_val[pos[t]] = w[t];
        __syncthreads();
#pragma unroll 4
        for (uint32_t e = threadIdx.x; e < tile_n; e += kBinThreads) o[s_dst[e]] = s_val[e];
      } else {
        for (uint32_t e = threadIdx.x; e < tile_n; e += kBinThreads) copy_item(vt, v, tile0 + s_src[e], s_dst[e]);
      }
    }
    if (perm)
      for (uint32_t e = threadIdx.x; e < tile_n; e += kBinThreads) perm[s_dst[e]] = tile0 + s_src[e];
  }
}

// Tile-local grouping (FGB_BUILD_TILE_LOCAL): perm[tile0 + e] = index of the item at grouped position e of its
// 2048-item tile, equal bins adjacent (table order).  One pass, no global atomics, no histogram: enough to give
// the lanes of a warp common message strips when the list is already coarsely ordered (the reference's auto sort
// leaves agents in (x,y)-column order), at a quarter of the cost of the global permutation.
template <int DIMS, bool VEC>
__global__ void __launch_bounds__(kBinThreads) k_group_tile(KeySrc<DIMS> src, uint32_t n_max, const unsigned int *d_n,
                                                            uint32_t *__restrict__ perm) {
  __shared__ uint32_t s_key[kTabSlots];   // key of the slot, later the slot's offset in the tile
  __shared__ uint32_t s_cnt[kTabSlots];
  __shared__ uint16_t s_src[kTile];
  __shared__ uint32_t s_scan[33];
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t tile0 = blockIdx.x * kTile;
  if (tile0 >= n) return;
  const uint32_t i0 = tile0 + threadIdx.x * kTileItems;
  const int cnt = i0 < n ? ((n - i0) < static_cast<uint32_t>(kTileItems) ? static_cast<int>(n - i0) : kTileItems) : 0;
  const uint32_t tile_n = (n - tile0) < static_cast<uint32_t>(kTile) ? n - tile0 : kTile;
  uint32_t k[kTileItems];
  if (cnt) load_tile_keys<DIMS, VEC>(src, i0, n, k);
  for (int s = threadIdx.x; s < kTabSlots; s += kBinThreads) {
    s_key[s] = kTabEmpty;
    s_cnt[s] = 0u;
  }
  __syncthreads();
  uint32_t slot[kTileItems], rank[kTileItems];
  {
    int j = 0;
#pragma unroll
    for (int r = 0; r < kTileItems; ++r) {
      if (r == j && j < cnt) {
        int e = j + 1;
#pragma unroll
        for (int t = 1; t < kTileItems; ++t)
          if (t < kTileItems - r && r + t < cnt && e == r + t && k[(r + t) & (kTileItems - 1)] == k[r]) e = r + t + 1;
        const uint32_t sl = tab_insert(s_key, k[r]);
        const uint32_t b = atomicAdd(s_cnt + sl, static_cast<uint32_t>(e - j));
#pragma unroll
        for (int t = 0; t < kTileItems; ++t)
          if (r + t < e && t < kTileItems - r) {
            slot[r + t] = sl;
            rank[r + t] = b + t;
          }
        j = e;
      }
    }
  }
  __syncthreads();
  {
    constexpr int kPer = kTabSlots / kBinThreads;
    uint32_t c[kPer], sum = 0;
    const int s0 = threadIdx.x * kPer;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      c[q] = s_key[s0 + q] != kTabEmpty ? s_cnt[s0 + q] : 0u;
      sum += c[q];
    }
    uint32_t tile_total;
    uint32_t off = block_exclusive_scan(sum, s_scan, &tile_total);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      s_key[s0 + q] = off;
      off += c[q];
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < kTileItems; ++t)
    if (t < cnt) s_src[s_key[slot[t]] + rank[t]] = static_cast<uint16_t>(threadIdx.x * kTileItems + t);
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < tile_n; e += kBinThreads) perm[tile0 + e] = tile0 + s_src[e];
}

// ---- stable order: per-bin fix-up --------------------------------------------------------
constexpr int kFixSmall = 32;

// One thread per bin: bins of <= kFixSmall items are insertion-sorted in place, larger bins are
// queued for k_fix_big.  pbm[] is the final prefix array (bin b = [pbm[b], pbm[b+1])).
__global__ void __launch_bounds__(256) k_fix_small(const uint32_t *__restrict__ pbm, uint32_t bins, uint32_t *perm,
                                                   uint32_t *worklist, uint32_t *ctrl) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= bins) return;
  const uint32_t s = pbm[b], e = pbm[b + 1];
  const uint32_t n = e - s;
  if (n <= 1) return;
  if (n > kFixSmall) {
    worklist[atomicAdd(ctrl, 1u)] = b;
    return;
  }
  uint32_t a[kFixSmall];
  for (uint32_t i = 0; i < n; ++i) a[i] = perm[s + i];
  bool sorted = true;
  for (uint32_t i = 1; i < n; ++i) {
    const uint32_t key = a[i];
    uint32_t p = i;
    while (p > 0 && a[p - 1] > key) {
      a[p] = a[p - 1];
      --p;
      sorted = false;
    }
    a[p] = key;
  }
  if (!sorted)
    for (uint32_t i = 0; i < n; ++i) perm[s + i] = a[i];
}

// Arbitrary-n bitonic network in the all-ascending form (partner i^(k-1) for the first step of a
// merge, i^j afterwards); items at virtual indices >= n act as +inf and never move.
__device__ __forceinline__ void block_bitonic(uint32_t *a, uint32_t n) {
  for (uint32_t k = 2; (k >> 1) < n; k <<= 1) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t l = i ^ (k - 1);
      if (l > i && l < n) {
        const uint32_t ai = a[i], al = a[l];
        if (ai > al) { a[i] = al; a[l] = ai; }
      }
    }
    __syncthreads();
    for (uint32_t j = k >> 2; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t l = i ^ j;
        if (l > i && l < n) {
          const uint32_t ai = a[i], al = a[l];
          if (ai > al) { a[i] = al; a[l] = ai; }
        }
      }
      __syncthreads();
    }
  }
}

constexpr int kFixBigSmem = 8192;
// Persistent blocks walk the big-bin worklist; a bin that fits shared memory is sorted there,
// a larger one in place in global memory (correct for any size, slow only for degenerate inputs).
__global__ void __launch_bounds__(1024) k_fix_big(const uint32_t *__restrict__ pbm, uint32_t *perm,
                                                  const uint32_t *worklist, const uint32_t *ctrl) {
  __shared__ uint32_t buf[kFixBigSmem];
  const uint32_t count = *ctrl;
  for (uint32_t w = blockIdx.x; w < count; w += gridDim.x) {
    const uint32_t b = worklist[w];
    const uint32_t s = pbm[b], n = pbm[b + 1] - s;
    if (n <= kFixBigSmem) {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) buf[i] = perm[s + i];
      __syncthreads();
      block_bitonic(buf, n);
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) perm[s + i] = buf[i];
      __syncthreads();
    } else {
      block_bitonic(perm + s, n);
    }
  }
}

// ---- slab decomposition: classify items by the plane (slowest grid axis) of their position ----
// plane = clamp(floorf((p - min)/radius), 0, dim-1) exactly as the bin arithmetic.
// flag_lo[i] = plane < lo, flag_hi[i] = plane >= hi, flag_mid[i] = neither (any output may be NULL).
__global__ void __launch_bounds__(256) k_plane_flags(const float *__restrict__ p, uint32_t n_max, const unsigned int *d_n,
                                                     float mn, float radius, int dim, int lo, int hi, uint32_t *flag_lo,
                                                     uint32_t *flag_mid, uint32_t *flag_hi) {
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  int c = static_cast<int>(floorf(__fdiv_rn(__ldg(p + i) - mn, radius)));
  c = c < 0 ? 0 : (c >= dim ? dim - 1 : c);
  const bool l = c < lo, h = c >= hi;
  if (flag_lo) flag_lo[i] = l ? 1u : 0u;
  if (flag_hi) flag_hi[i] = h ? 1u : 0u;
  if (flag_mid) flag_mid[i] = (!l && !h) ? 1u : 0u;
}

// ---- gather: out[j] = in[perm[j]] (scatter_position_generic, CUDAScatter.cu:89-104) ------
template <bool VEC>
__global__ void __launch_bounds__(kBinThreads) k_gather(const uint32_t *__restrict__ perm, uint32_t n_max,
                                                        const unsigned int *d_n, const __grid_constant__ VarTable vt) {
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t j0 = (blockIdx.x * kBinThreads + threadIdx.x) * 4;
  if (j0 >= n) return;
  uint32_t src[4];
  const int cnt = (n - j0) < 4u ? static_cast<int>(n - j0) : 4;
  if (VEC && cnt == 4) {
    const uint4 q = ld_stream_u4(perm + j0);
    src[0] = q.x; src[1] = q.y; src[2] = q.z; src[3] = q.w;
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t) src[t] = t < cnt ? perm[j0 + t] : 0u;
  }
  for (uint32_t v = 0; v < vt.n; ++v) {
    const uint32_t len = vt.len[v];
    if (VEC && len == 4 && cnt == 4) {
      const uint32_t *in = reinterpret_cast<const uint32_t *>(vt.in[v]);
      uint4 q;
      q.x = __ldg(in + src[0]);
      q.y = __ldg(in + src[1]);
      q.z = __ldg(in + src[2]);
      q.w = __ldg(in + src[3]);
      *reinterpret_cast<uint4 *>(vt.out[v] + static_cast<size_t>(j0) * 4) = q;
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (t < cnt) copy_item(vt, v, src[t], j0 + t);
    }
  }
}

#endif  // __CUDACC__

// blocks for the 4-items-per-thread kernels (gather) and for the tile kernels (histogram, scatter)
inline unsigned int bin_grid(unsigned int n) {
  const unsigned int per_block = kBinThreads * kBinItems;
  return (n + per_block - 1) / per_block;
}
inline unsigned int tile_grid(unsigned int n) { return (n + kTile - 1) / kTile; }

}  // namespace fgb
