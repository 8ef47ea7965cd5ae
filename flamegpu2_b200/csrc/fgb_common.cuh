// fgb_common.cuh -- shared device/host helpers of the sm_100a hot-path library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "flamegpu2_b200.h"

#define FGB_CHECK(expr)                                  \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return static_cast<int>(_e);  \
  } while (0)

namespace fgb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ------------------------------------------------------------------------------------------
// SoA variable table passed BY VALUE in kernel parameter space (the reference uploads a
// ScatterData array with a cudaMemcpyAsync before every scatter launch, CUDAScatter.cu:160-161).
// ------------------------------------------------------------------------------------------
struct VarTable {
  uint32_t n;
  uint32_t len[FGB_MAX_VARS];
  const char *in[FGB_MAX_VARS];
  char *out[FGB_MAX_VARS];
};

inline int make_var_table(const fgb_var *vars, unsigned int nvars, VarTable *vt) {
  if (nvars > FGB_MAX_VARS) return FGB_ERR_TOO_MANY_VARS;
  if (nvars && !vars) return FGB_ERR_INVALID_ARG;
  vt->n = nvars;
  for (unsigned int v = 0; v < nvars; ++v) {
    if (vars[v].type_len == 0 || vars[v].type_len > 0xFFFFFFFFull) return FGB_ERR_INVALID_ARG;
    vt->len[v] = static_cast<uint32_t>(vars[v].type_len);
    vt->in[v] = static_cast<const char *>(vars[v].in);
    vt->out[v] = static_cast<char *>(vars[v].out);
  }
  return FGB_OK;
}

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t load_count(const unsigned int *d_n, uint32_t n_max) {
  if (d_n) {
    uint32_t v = __ldg(d_n);
    return v < n_max ? v : n_max;
  }
  return n_max;
}

// same, for a count word that the calling kernel may also WRITE through another argument (aliased device words)
__device__ __forceinline__ uint32_t load_count_coherent(const unsigned int *d_n, uint32_t n_max) {
  if (d_n) {
    const uint32_t v = *reinterpret_cast<const volatile unsigned int *>(d_n);
    return v < n_max ? v : n_max;
  }
  return n_max;
}

// streaming (read-once) loads / stores: keep them out of L1
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// 256-bit read-only load (sm_100: LDG.E.ENL2.256); p must be 32-byte aligned
__device__ __forceinline__ void ld_nc_f8(const float *p, float v[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void ld_nc_u8(const uint32_t *p, uint32_t v[8]) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
// 256-bit store; p must be 32-byte aligned
__device__ __forceinline__ void st_u8(uint32_t *p, const uint32_t v[8]) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
  uint4 r = ld_stream_u4(p);
  return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
}

// Copy one item of variable v from item index src to item index dst.
__device__ __forceinline__ void copy_item(const VarTable &vt, uint32_t v, size_t src, size_t dst) {
  const uint32_t len = vt.len[v];
  const char *in = vt.in[v];
  char *out = vt.out[v];
  if (len == 4) {
    reinterpret_cast<uint32_t *>(out)[dst] = __ldg(reinterpret_cast<const uint32_t *>(in) + src);
  } else if (len == 8) {
    reinterpret_cast<uint2 *>(out)[dst] = __ldg(reinterpret_cast<const uint2 *>(in) + src);
  } else if (len == 16) {
    reinterpret_cast<uint4 *>(out)[dst] = __ldg(reinterpret_cast<const uint4 *>(in) + src);
  } else if ((len & 3u) == 0) {
    const uint32_t *s = reinterpret_cast<const uint32_t *>(in + src * len);
    uint32_t *d = reinterpret_cast<uint32_t *>(out + dst * len);
    for (uint32_t w = 0; w < (len >> 2); ++w) d[w] = __ldg(s + w);
  } else {
    const char *s = in + src * len;
    char *d = out + dst * len;
    for (uint32_t b = 0; b < len; ++b) d[b] = s[b];
  }
}

// ------------------------------------------------------------------------------------------
// Decoupled look-back tile state: one 64-bit word per tile, {status:2 | value:32}, written and
// read with single 64-bit relaxed accesses, so value and status always travel together.
// ------------------------------------------------------------------------------------------
constexpr unsigned long long kStInvalid = 0ull;
constexpr unsigned long long kStAggregate = 1ull << 32;
constexpr unsigned long long kStInclusive = 2ull << 32;

__device__ __forceinline__ void st_state(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_state(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
  return __reduce_add_sync(0xffffffffu, v);
}

// Called by ONE full warp of tile `tile` (> 0) after the tile aggregate has been published.
// Returns the sum of the aggregates of tiles [0, tile).
__device__ __forceinline__ uint32_t lookback_exclusive(const unsigned long long *state, int tile) {
  const int lane = threadIdx.x & 31;
  uint32_t excl = 0;
  int pred = tile - 1 - lane;
  while (true) {
    unsigned long long s = pred >= 0 ? ld_state(state + pred) : kStInclusive;
    while (__any_sync(0xffffffffu, (s >> 32) == 0ull)) {
      if ((s >> 32) == 0ull) s = ld_state(state + pred);
    }
    const unsigned incl = __ballot_sync(0xffffffffu, (s >> 32) == 2ull);
    const uint32_t val = static_cast<uint32_t>(s);
    if (incl) {
      const int first = __ffs(incl) - 1;  // nearest predecessor holding an inclusive prefix
      excl += warp_sum(lane <= first ? val : 0u);
      break;
    }
    excl += warp_sum(val);
    pred -= 32;
  }
  return excl;
}

// Block-wide exclusive scan of one value per thread (blockDim.x multiple of 32, <= 1024).
// Returns the exclusive prefix of `v`; *block_total (all threads) gets the block sum.
// `warp_sums` is shared memory with >= 33 words.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t *block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nwarps ? warp_sums[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    warp_sums[lane] = wi - w;  // exclusive prefix of warp sums
    if (lane == 31) warp_sums[32] = wi;
  }
  __syncthreads();
  const uint32_t base = warp_sums[warp];
  *block_total = warp_sums[32];
  return base + incl - v;
}

#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------
// host side: grow-only device scratch
// ------------------------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  unsigned long long *gen = nullptr;  // owner's allocation generation: bumped whenever this buffer moves
  int reserve(size_t need) {
    if (need <= bytes) return 0;
    if (gen) ++*gen;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = need + need / 4;  // the reference grows lists by 1.25-1.5x as well
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, need);
      want = need;
    }
    if (e != cudaSuccess) return static_cast<int>(e);
    bytes = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace fgb

// Scratch is keyed by (ctx, stream_id) like CUDAScanCompaction's [type][stream] configs
// (include/flamegpu/simulation/detail/CUDAScanCompaction.h:77-89,137): functions of one layer run
// concurrently on different streams and must not share look-back words or histograms.
struct fgb_stream_scratch {
  fgb::DevBuf tile_state;   // look-back words + done counter (self-cleaning, zero between calls)
  fgb::DevBuf sort_hist;    // histogram of fgb_sort_by_key (all-zero between calls)
  fgb::DevBuf sort_cursor;  // scanned cursors
  fgb::DevBuf perm;         // permutation scratch
  fgb::DevBuf worklist;     // big-bin worklist
  fgb::DevBuf ctrl;         // small control words
  fgb::DevBuf rs_state;     // radix sort: digit histograms + per-pass look-back words
  fgb::DevBuf rs_keys[2];   // radix sort: key ping-pong
  fgb::DevBuf rs_idx[2];    // radix sort: index ping-pong
  fgb::DevBuf red;          // reductions: per-block partials (8 B each) + done counter
  fgb::DevBuf slab;         // slab migration: control block, leaver index lists, hole / tail pairing
};

#define FGB_MAX_STREAMS 128

struct fgb_ctx {
  int device = 0;
  unsigned long long launches = 0;
  // Counts every (re)allocation of scratch owned by this context or by its fgb_spatial handles.  A caller that
  // captured launches into a CUDA graph compares it before a replay: the graph holds the old pointers.
  unsigned long long generation = 0;
  fgb_stream_scratch slot[FGB_MAX_STREAMS];
  fgb_ctx() {
    for (auto &s : slot) {
      fgb::DevBuf *all[] = {&s.tile_state, &s.sort_hist, &s.sort_cursor, &s.perm, &s.worklist, &s.ctrl, &s.rs_state,
                            &s.rs_keys[0], &s.rs_keys[1], &s.rs_idx[0], &s.rs_idx[1], &s.red, &s.slab};
      for (fgb::DevBuf *b : all) b->gen = &generation;
    }
  }
};

struct fgb_spatial {
  fgb_ctx *ctx = nullptr;
  int dims = 3;
  fgb_spatial_metadata md{};
  unsigned int bin_count = 0;
  int win_begin = 0, win_count = 0;  // slab window on the slowest axis (planes stored locally)
  int key_min = 0;                   // bucket lists (dims == 0): bin = key - key_min
  unsigned int *d_hist = nullptr;       // bin_count + 1, all-zero between builds
  unsigned long long *d_state = nullptr;  // look-back words for the PBM scan
  unsigned int n_state = 0;
  fgb_spatial_metadata *d_md = nullptr;
  fgb::DevBuf perm;      // stable mode: permutation
  fgb::DevBuf worklist;  // stable mode: big bins
  unsigned int *d_ctrl = nullptr;  // [0] big-bin count, [1..3] k_scan_scatter (fgb_binsort.cuh)
  bool pbm_external = false;       // md.PBM belongs to the caller (fgb_spatial_use_pbm)
  fgb::DevBuf tile_mode;           // worklist of the ungrouped 2048-item tiles of the build in flight (staged scatter)
  fgb::DevBuf keys;                // bin of every item of the list being indexed (written once, by k_bin_keys or by the
                                   // list's writer: fgb_spatial_writer_args)
};
