// fgb_api.cu -- C ABI entry points (include/flamegpu2_b200.h) of the sm_100a hot-path library.
// Host launchers only: every entry point enqueues kernels on the caller's stream and returns.
// There is no CPU fallback; without a CUDA device every call fails with FGB_ERR_NO_DEVICE.
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <limits>
#include <new>

#include "fgb_binsort.cuh"
#include "fgb_common.cuh"
#include "fgb_compact.cuh"
#include "fgb_radix.cuh"
#include "fgb_reduce.cuh"
#include "fgb_scan.cuh"
#include "fgb_scatter_misc.cuh"
#include "fgb_slab.cuh"

using namespace fgb;

#ifndef FGB_BUILD_STAGED_DEFAULT
#define FGB_BUILD_STAGED_DEFAULT 0
#endif
#ifndef FGB_COMPACT_BULK_DEFAULT
#define FGB_COMPACT_BULK_DEFAULT 0
#endif

namespace {

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline bool vars_in_aligned(const fgb_var *vars, unsigned int nvars) {
  for (unsigned int v = 0; v < nvars; ++v)
    if (!aligned16(vars[v].in)) return false;
  return true;
}
inline bool vars_out_aligned(const fgb_var *vars, unsigned int nvars) {
  for (unsigned int v = 0; v < nvars; ++v)
    if (!aligned16(vars[v].out)) return false;
  return true;
}

int reserve_zeroed(DevBuf &b, size_t need) {
  if (need <= b.bytes) return 0;
  int r = b.reserve(need);
  if (r) return r;
  return static_cast<int>(cudaMemset(b.p, 0, b.bytes));
}

// approxExactlyDivisible<float> (include/flamegpu/detail/numeric.h:25-32)
bool approx_exactly_divisible(float x, float y) {
  const float scaled_eps = std::max(std::fabs(x), std::fabs(y)) * 1.1920928955078125e-07f;
  const float v = std::fmod(x, y);
  return v <= scaled_eps || v > y - scaled_eps;
}

inline int launch_ok() { return static_cast<int>(cudaPeekAtLastError()); }

inline Geo make_geo(const fgb_spatial *sp) {
  return Geo{sp->md.min[0], sp->md.min[1], sp->md.min[2], sp->md.radius, static_cast<int>(sp->md.grid_dim[0]),
             static_cast<int>(sp->md.grid_dim[1]), static_cast<int>(sp->md.grid_dim[2]), sp->win_begin, sp->win_count};
}

// worklist entries needed for n items: bins with more than kFixSmall items
inline size_t worklist_bytes(unsigned int n) { return (static_cast<size_t>(n) / (kFixSmall + 1) + 1) * sizeof(uint32_t); }

// The stable tail shared by fgb_build_index(STABLE) and fgb_sort_by_key:
// per-bin index fix-up, then the gather that applies the permutation.
int stable_tail(fgb_ctx *ctx, const uint32_t *pbm, unsigned int bins, uint32_t *perm, uint32_t *worklist,
                uint32_t *ctrl, unsigned int n, const unsigned int *d_n, const fgb_var *vars, unsigned int nvars,
                cudaStream_t st) {
  k_fix_small<<<(bins + 255) / 256, 256, 0, st>>>(pbm, bins, perm, worklist, ctrl);
  k_fix_big<<<kNumSMs, 1024, 0, st>>>(pbm, perm, worklist, ctrl);
  ctx->launches += 2;
  if (nvars) {
    VarTable vt;
    int r = make_var_table(vars, nvars, &vt);
    if (r) return r;
    const bool vec = aligned16(perm) && vars_out_aligned(vars, nvars);
    if (vec)
      k_gather<true><<<bin_grid(n), kBinThreads, 0, st>>>(perm, n, d_n, vt);
    else
      k_gather<false><<<bin_grid(n), kBinThreads, 0, st>>>(perm, n, d_n, vt);
    ctx->launches += 1;
  }
  return launch_ok();
}

// keys + histogram for the items [*d_keyed, n) of a list (d_keyed NULL: all of them)
template <int DIMS>
int launch_bin_keys(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const unsigned int *d_keyed, const float *x, const float *y,
                    const float *z, cudaStream_t st) {
  KeySrc<DIMS> src{};
  src.x = x;
  src.y = y;
  src.z = z;
  src.keys = reinterpret_cast<const uint32_t *>(x);  // DIMS == 0: `x` carries the integer key array
  src.key_min = static_cast<uint32_t>(sp->key_min);
  src.key_span = sp->bin_count;
  src.g = make_geo(sp);
  const bool vec = aligned16(x) && (DIMS == 0 || aligned16(y)) && (DIMS != 3 || aligned16(z));
  uint32_t *keys = static_cast<uint32_t *>(sp->keys.p);
  const unsigned int grid = tile_grid(n);
  if (vec)
    k_bin_keys<DIMS, true><<<grid, kBinThreads, 0, st>>>(src, n, d_n, d_keyed, keys, sp->d_hist, sp->d_state, sp->n_state, sp->d_ctrl);
  else
    k_bin_keys<DIMS, false><<<grid, kBinThreads, 0, st>>>(src, n, d_n, d_keyed, keys, sp->d_hist, sp->d_state, sp->n_state, sp->d_ctrl);
  sp->ctx->launches += 1;
  return launch_ok();
}

// scan + scatter of the stored keys in one launch, then the (normally empty) worklist of ungrouped tiles
static_assert(kFsTile == kScanTile, "k_scan_scatter and scan_num_tiles() must agree on the scan tile size");
template <bool IDX_ONLY>
void launch_scan_scatter(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const VarTable &vt, unsigned int *perm, bool vec,
                         cudaStream_t st, bool expect_grouped = false) {
  const unsigned int tiles = tile_grid(n), scan_tiles = scan_num_tiles(sp->bin_count);
  const uint32_t *keys = static_cast<const uint32_t *>(sp->keys.p);
  uint32_t *wl = static_cast<uint32_t *>(sp->tile_mode.p);
  k_scan_scatter<IDX_ONLY><<<scan_tiles + tiles, kBinThreads, 0, st>>>(sp->d_hist, sp->md.PBM, sp->bin_count, sp->d_state, scan_tiles, keys, n,
                                                                       d_n, vt, perm, wl, sp->d_ctrl, expect_grouped ? 1u : 0u);
  if (expect_grouped) {  // ungrouped tiles were scattered inside that launch: nothing is queued
    sp->ctx->launches += 1;
    return;
  }
  const unsigned int sgrid = std::min<unsigned int>(tiles, 4u * kNumSMs);  // 46 KB of shared memory per block: 4 per SM
  if (vec)
    k_bin_scatter_staged<true, IDX_ONLY><<<sgrid, kBinThreads, 0, st>>>(keys, n, d_n, sp->md.PBM, vt, perm, wl, sp->d_ctrl);
  else
    k_bin_scatter_staged<false, IDX_ONLY><<<sgrid, kBinThreads, 0, st>>>(keys, n, d_n, sp->md.PBM, vt, perm, wl, sp->d_ctrl);
  sp->ctx->launches += 2;
}

// scan + scatter from the stored keys (sp->keys, sp->d_hist complete for [0, n))
int scatter_from_keys(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const fgb_var *vars, unsigned int nvars, unsigned int flags,
                      unsigned int *src_slot_out, cudaStream_t st) {
  fgb_ctx *ctx = sp->ctx;
  VarTable vt;
  int r = make_var_table(vars, nvars, &vt);
  if (r) return r;
  const bool vec = vars_in_aligned(vars, nvars);
  const unsigned int B = sp->bin_count;
  const bool stable = (flags & FGB_BUILD_STABLE) != 0;
  uint32_t *perm = static_cast<uint32_t *>(sp->perm.p);
  if (!stable) {
    // Ungrouped tiles are scattered inside k_scan_scatter (one atomic and one scattered store per message) unless the
    // library is built with FGB_BUILD_STAGED_DEFAULT=1: measured on B200 the in-kernel path beats the shared-memory staged
    // kernel on every input order (16.8 M messages, uniformly random: 2.50 ms against 2.74 ms; 1 M: 60 against 91 us; nearly
    // ordered Boids-2D lists: 224 against 668 us), so FGB_BUILD_EXPECT_GROUPED only matters for that optional build
    launch_scan_scatter<false>(sp, n, d_n, vt, src_slot_out, vec, st, (flags & FGB_BUILD_EXPECT_GROUPED) != 0 || !FGB_BUILD_STAGED_DEFAULT);
    return launch_ok();
  }
  launch_scan_scatter<true>(sp, n, d_n, vt, perm, true, st);
  r = stable_tail(ctx, sp->md.PBM, B, perm, static_cast<uint32_t *>(sp->worklist.p), sp->d_ctrl, n, d_n, vars, nvars, st);
  if (r == 0 && src_slot_out) r = static_cast<int>(cudaMemcpyAsync(src_slot_out, perm, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, st));
  return r;
}

int reserve_index(fgb_spatial *sp, unsigned int n, bool stable) {
  int r = sp->keys.reserve((static_cast<size_t>(n) + 8) * 4);
  if (r) return r;
  r = sp->tile_mode.reserve((static_cast<size_t>(tile_grid(n)) + 1) * 4);
  if (r) return r;
  if (stable) {
    r = sp->perm.reserve(static_cast<size_t>(n) * 4);
    if (r) return r;
    r = sp->worklist.reserve(worklist_bytes(n));
  }
  return r;
}

template <int DIMS>
int build_index_impl(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const float *x, const float *y,
                     const float *z, const fgb_var *vars, unsigned int nvars, unsigned int flags, const unsigned int *d_keyed,
                     unsigned int *src_slot_out, cudaStream_t st) {
  int r = reserve_index(sp, n, (flags & FGB_BUILD_STABLE) != 0);
  if (r) return r;
  // keys + histogram: skipped when the list's writer published them for every item (FGB_BUILD_KEYS_READY without a
  // d_keyed word); with a d_keyed word only the items behind it are keyed here
  if (!(flags & FGB_BUILD_KEYS_READY) || d_keyed) {
    r = launch_bin_keys<DIMS>(sp, n, d_n, (flags & FGB_BUILD_KEYS_READY) ? d_keyed : nullptr, x, y, z, st);
    if (r) return r;
  }
  return scatter_from_keys(sp, n, d_n, vars, nvars, flags, src_slot_out, st);
}

}  // namespace

namespace {
template <int DIMS>
int bin_permutation_impl(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const float *x, const float *y,
                         const float *z, unsigned int *perm, unsigned int flags, cudaStream_t st) {
  fgb_ctx *ctx = sp->ctx;
  if (flags & FGB_BUILD_TILE_LOCAL) {
    KeySrc<DIMS> src{};
    src.x = x;
    src.y = y;
    src.z = z;
    src.g = make_geo(sp);
    const bool vec = aligned16(x) && aligned16(y) && (DIMS == 2 || aligned16(z));
    const unsigned int grid = tile_grid(n);
    if (vec)
      k_group_tile<DIMS, true><<<grid, kBinThreads, 0, st>>>(src, n, d_n, perm);
    else
      k_group_tile<DIMS, false><<<grid, kBinThreads, 0, st>>>(src, n, d_n, perm);
    ctx->launches += 1;
    return launch_ok();
  }
  const bool stable = (flags & FGB_BUILD_STABLE) != 0;
  int r = reserve_index(sp, n, stable);
  if (r) return r;
  r = launch_bin_keys<DIMS>(sp, n, d_n, nullptr, x, y, z, st);
  if (r) return r;
  const unsigned int B = sp->bin_count;
  VarTable none{};
  none.n = 0;
  launch_scan_scatter<true>(sp, n, d_n, none, perm, true, st);
  if (stable) return stable_tail(ctx, sp->md.PBM, B, perm, static_cast<uint32_t *>(sp->worklist.p), sp->d_ctrl, n, d_n, nullptr, 0, st);
  return launch_ok();
}
}  // namespace

namespace {
// stable LSD radix sort of (key, index) + gather of every variable.  geo != NULL: the keys are computed from the
// positions in the same pass that builds the digit histograms (and stored to `keys`); otherwise `keys` is read.
int sort_impl(fgb_ctx *ctx, unsigned int stream_id, unsigned int *keys, int max_bit, unsigned int n, const unsigned int *d_n,
              const fgb_var *vars, unsigned int nvars, unsigned int *position_out, cudaStream_t st, const float *x, const float *y,
              const float *z, const SortGeo *geo) {
  int r = fgb_ctx_reserve(ctx, stream_id, n, max_bit);
  if (r) return r;
  fgb_stream_scratch &s = ctx->slot[stream_id];
  const RadixPlan plan = make_radix_plan(max_bit);
  const unsigned int tiles = radix_num_tiles(n);
  uint32_t *ghist = static_cast<uint32_t *>(s.rs_state.p);
  uint32_t *perm = position_out ? position_out : static_cast<uint32_t *>(s.perm.p);
  size_t words = static_cast<size_t>(kRsMaxPasses) * kRsMaxDigits;
  size_t state_off[kRsMaxPasses];
  for (int p = 0; p < plan.passes; ++p) {
    state_off[p] = words;
    words += static_cast<size_t>(tiles) * (static_cast<size_t>(1) << plan.bits[p]);
  }
  FGB_CHECK(cudaMemsetAsync(ghist, 0, words * 4, st));
  const unsigned int hgrid = std::min<unsigned int>(tiles, 4u * kNumSMs);
  const uint32_t key_mask = (max_bit >= 32) ? 0xFFFFFFFFu : ((1u << max_bit) - 1u);
  if (geo) {
    if (z)
      k_sort_keys_hist<3><<<hgrid, kRsThreads, 0, st>>>(x, y, z, *geo, n, d_n, keys, key_mask, plan, ghist);
    else
      k_sort_keys_hist<2><<<hgrid, kRsThreads, 0, st>>>(x, y, nullptr, *geo, n, d_n, keys, key_mask, plan, ghist);
  } else {
    k_radix_hist<<<hgrid, kRsThreads, 0, st>>>(keys, n, d_n, plan, ghist);
  }
  for (int p = 0; p < plan.passes; ++p) {
    const bool last = p == plan.passes - 1;
    const uint32_t *kin = p == 0 ? keys : static_cast<const uint32_t *>(s.rs_keys[(p - 1) & 1].p);
    const uint32_t *iin = p == 0 ? nullptr : static_cast<const uint32_t *>(s.rs_idx[(p - 1) & 1].p);
    uint32_t *kout = last ? nullptr : static_cast<uint32_t *>(s.rs_keys[p & 1].p);
    uint32_t *iout = last ? perm : static_cast<uint32_t *>(s.rs_idx[p & 1].p);
    k_radix_onesweep<<<tiles, kRsThreads, 0, st>>>(kin, iin, kout, iout, n, d_n, plan.shift[p], plan.bits[p], key_mask,
                                                    ghist + p * kRsMaxDigits, ghist + state_off[p]);
  }
  ctx->launches += 1 + plan.passes;
  if (nvars) {
    VarTable vt;
    r = make_var_table(vars, nvars, &vt);
    if (r) return r;
    if (aligned16(perm) && vars_out_aligned(vars, nvars))
      k_gather<true><<<bin_grid(n), kBinThreads, 0, st>>>(perm, n, d_n, vt);
    else
      k_gather<false><<<bin_grid(n), kBinThreads, 0, st>>>(perm, n, d_n, vt);
    ctx->launches += 1;
  }
  return launch_ok();
}

// count of elements equal to `value` (as unsigned long long) or sum of squared deviations from `mean` (as double)
template <typename T>
int launch_transform_reduce(int transform, const void *in, unsigned int n, const unsigned int *d_n, double mean, T value, void *partial,
                            uint32_t *done, void *d_out, unsigned int blocks, cudaStream_t st) {
  const T *p = static_cast<const T *>(in);
  if (transform == FGB_TRANSFORM_COUNT_EQUAL)
    k_reduce<T, unsigned long long, kOpSum, kTrEqual><<<blocks, kRedThreads, 0, st>>>(
        p, n, d_n, 0ull, static_cast<unsigned long long *>(partial), done, static_cast<unsigned long long *>(d_out), 0.0, value);
  else
    k_reduce<T, double, kOpSum, kTrSqDev><<<blocks, kRedThreads, 0, st>>>(p, n, d_n, 0.0, static_cast<double *>(partial), done,
                                                                             static_cast<double *>(d_out), mean, T());
  return launch_ok();
}

template <typename T, typename A>
int launch_reduce(int op, const void *in, unsigned int n, const unsigned int *d_n, A id_min, A id_max, void *partial, uint32_t *done,
                  void *d_out, unsigned int blocks, cudaStream_t st) {
  const T *p = static_cast<const T *>(in);
  A *pa = static_cast<A *>(partial), *po = static_cast<A *>(d_out);
  if (op == FGB_REDUCE_SUM)
    k_reduce<T, A, kOpSum><<<blocks, kRedThreads, 0, st>>>(p, n, d_n, static_cast<A>(0), pa, done, po);
  else if (op == FGB_REDUCE_MIN)
    k_reduce<T, A, kOpMin><<<blocks, kRedThreads, 0, st>>>(p, n, d_n, id_min, pa, done, po);
  else
    k_reduce<T, A, kOpMax><<<blocks, kRedThreads, 0, st>>>(p, n, d_n, id_max, pa, done, po);
  return launch_ok();
}
}  // namespace

namespace {
// histogram, PBM, scan look-back words, device metadata copy; frees sp on failure
int alloc_index_buffers(fgb_spatial *sp) {
  fgb_spatial_metadata &md = sp->md;
  const size_t words = static_cast<size_t>(sp->bin_count) + 1;
  sp->n_state = scan_num_tiles(sp->bin_count);
  cudaError_t e = cudaMalloc(&sp->d_hist, words * 4);
  if (e == cudaSuccess) e = cudaMalloc(&md.PBM, words * 4);
  if (e == cudaSuccess) e = cudaMalloc(&sp->d_state, static_cast<size_t>(sp->n_state) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&sp->d_md, sizeof(fgb_spatial_metadata));
  if (e == cudaSuccess) e = cudaMalloc(&sp->d_ctrl, 64);
  if (e == cudaSuccess) e = cudaMemset(sp->d_hist, 0, words * 4);
  if (e == cudaSuccess) e = cudaMemset(md.PBM, 0, words * 4);  // MessageSpatial3D.cu:77, MessageBucket.cu:69
  if (e == cudaSuccess) e = cudaMemset(sp->d_state, 0, static_cast<size_t>(sp->n_state) * 8);
  if (e == cudaSuccess) e = cudaMemset(sp->d_ctrl, 0, 64);
  if (e == cudaSuccess) e = cudaMemcpy(sp->d_md, &md, sizeof(md), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    fgb_spatial_destroy(sp);
    return static_cast<int>(e);
  }
  return FGB_OK;
}
}  // namespace

extern "C" {

int fgb_version(void) { return FGB_VERSION; }

const char *fgb_error_string(fgb_status s) {
  if (s > 0) return cudaGetErrorString(static_cast<cudaError_t>(s));
  switch (s) {
    case FGB_OK: return "ok";
    case FGB_ERR_INVALID_ARG: return "invalid argument";
    case FGB_ERR_TOO_MANY_VARS: return "too many variables (FGB_MAX_VARS)";
    case FGB_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case FGB_ERR_ALLOC: return "allocation failed";
    case FGB_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown";
  }
}

fgb_status fgb_ctx_create(int device, fgb_ctx **out) {
  if (!out) return FGB_ERR_INVALID_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return FGB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) return FGB_ERR_INVALID_ARG;
  FGB_CHECK(cudaSetDevice(device));
  fgb_ctx *c = new (std::nothrow) fgb_ctx();
  if (!c) return FGB_ERR_ALLOC;
  c->device = device;
  *out = c;
  return FGB_OK;
}

fgb_status fgb_ctx_destroy(fgb_ctx *ctx) {
  if (!ctx) return FGB_OK;
  for (auto &s : ctx->slot) {
    s.tile_state.release();
    s.sort_hist.release();
    s.sort_cursor.release();
    s.perm.release();
    s.worklist.release();
    s.ctrl.release();
    s.rs_state.release();
    for (auto &b : s.rs_keys) b.release();
    for (auto &b : s.rs_idx) b.release();
    s.red.release();
    s.slab.release();
  }
  delete ctx;
  return FGB_OK;
}

unsigned long long fgb_launch_count(const fgb_ctx *ctx) { return ctx ? ctx->launches : 0ull; }

unsigned long long fgb_alloc_generation(const fgb_ctx *ctx) { return ctx ? ctx->generation : 0ull; }

fgb_status fgb_spatial_create(fgb_ctx *ctx, int dims, const float *env_min, const float *env_max, float radius,
                              fgb_spatial **out) {
  return fgb_spatial_create_window(ctx, dims, env_min, env_max, radius, 0, -1, out);
}

fgb_status fgb_spatial_create_window(fgb_ctx *ctx, int dims, const float *env_min, const float *env_max, float radius,
                                     int plane_begin, int plane_count, fgb_spatial **out) {
  if (!ctx || !out || !env_min || !env_max || (dims != 2 && dims != 3) || !(radius > 0.f)) return FGB_ERR_INVALID_ARG;
  fgb_spatial *sp = new (std::nothrow) fgb_spatial();
  if (!sp) return FGB_ERR_ALLOC;
  sp->ctx = ctx;
  sp->dims = dims;
  sp->perm.gen = sp->worklist.gen = sp->tile_mode.gen = sp->keys.gen = &ctx->generation;
  fgb_spatial_metadata &md = sp->md;
  std::memset(&md, 0, sizeof(md));
  md.radius = radius;
  md.wrap_compatible = true;
  unsigned long long bins = 1;
  for (int a = 0; a < 3; ++a) md.grid_dim[a] = 1;
  for (int a = 0; a < dims; ++a) {
    md.min[a] = env_min[a];
    md.max[a] = env_max[a];
    md.environment_width[a] = md.max[a] - md.min[a];
    // static_cast<unsigned int>(ceil(environmentWidth / radius)), MessageSpatial3D.cu:47
    md.grid_dim[a] = static_cast<unsigned int>(std::ceil(md.environment_width[a] / md.radius));
    bins *= md.grid_dim[a];
    md.wrap_compatible = md.wrap_compatible && approx_exactly_divisible(md.environment_width[a], md.radius);
  }
  {
    // slab window on the slowest axis: the PBM covers planes [plane_begin, plane_begin + plane_count) only
    const int slow = dims - 1;
    const int gslow = static_cast<int>(md.grid_dim[slow]);
    if (plane_count < 0) {
      plane_begin = 0;
      plane_count = gslow;
    }
    if (plane_begin < 0 || plane_count < 1 || plane_begin + plane_count > gslow) {
      delete sp;
      return FGB_ERR_INVALID_ARG;
    }
    sp->win_begin = plane_begin;
    sp->win_count = plane_count;
    bins = bins / static_cast<unsigned long long>(gslow) * static_cast<unsigned long long>(plane_count);
  }
  if (bins == 0 || bins >= 0x7FFFFFFFull) {
    delete sp;
    return FGB_ERR_INVALID_ARG;
  }
  sp->bin_count = static_cast<unsigned int>(bins);
  const int r = alloc_index_buffers(sp);
  if (r) return r;
  *out = sp;
  return FGB_OK;
}

/* MessageBucket::CUDAModelHandler (MessageBucket.cu:36-78): keys lower..upper inclusive, bucketCount = upper - lower + 1 */
fgb_status fgb_bucket_create(fgb_ctx *ctx, int lower_bound, int upper_bound, fgb_spatial **out) {
  if (!ctx || !out || upper_bound <= lower_bound) return FGB_ERR_INVALID_ARG;
  const long long bins = static_cast<long long>(upper_bound) - static_cast<long long>(lower_bound) + 1;
  if (bins >= 0x7FFFFFFFll) return FGB_ERR_INVALID_ARG;
  fgb_spatial *sp = new (std::nothrow) fgb_spatial();
  if (!sp) return FGB_ERR_ALLOC;
  sp->ctx = ctx;
  sp->dims = 0;
  sp->perm.gen = sp->worklist.gen = sp->tile_mode.gen = sp->keys.gen = &ctx->generation;
  std::memset(&sp->md, 0, sizeof(sp->md));
  sp->md.grid_dim[0] = static_cast<unsigned int>(bins);
  sp->md.grid_dim[1] = sp->md.grid_dim[2] = 1;
  sp->key_min = lower_bound;
  sp->win_begin = 0;
  sp->win_count = 1;
  sp->bin_count = static_cast<unsigned int>(bins);
  const int r = alloc_index_buffers(sp);
  if (r) return r;
  *out = sp;
  return FGB_OK;
}

fgb_status fgb_bucket_get_bounds(const fgb_spatial *sp, int *min_key, int *max_key_exclusive, const unsigned int **d_pbm) {
  if (!sp || sp->dims != 0) return FGB_ERR_INVALID_ARG;
  if (min_key) *min_key = sp->key_min;
  if (max_key_exclusive) *max_key_exclusive = sp->key_min + static_cast<int>(sp->bin_count);  // MetaData::max, MessageBucket.cu:44
  if (d_pbm) *d_pbm = sp->md.PBM;
  return FGB_OK;
}

fgb_status fgb_spatial_destroy(fgb_spatial *sp) {
  if (!sp) return FGB_OK;
  if (sp->d_hist) cudaFree(sp->d_hist);
  if (sp->md.PBM && !sp->pbm_external) cudaFree(sp->md.PBM);
  if (sp->d_state) cudaFree(sp->d_state);
  if (sp->d_md) cudaFree(sp->d_md);
  if (sp->d_ctrl) cudaFree(sp->d_ctrl);
  sp->perm.release();
  sp->worklist.release();
  sp->tile_mode.release();
  sp->keys.release();
  delete sp;
  return FGB_OK;
}

fgb_status fgb_spatial_get_metadata(const fgb_spatial *sp, fgb_spatial_metadata *host_out, unsigned int *bin_count) {
  if (!sp) return FGB_ERR_INVALID_ARG;
  if (host_out) *host_out = sp->md;
  if (bin_count) *bin_count = sp->bin_count;
  return FGB_OK;
}

unsigned int fgb_spatial_bin_count(const fgb_spatial *sp) { return sp ? sp->bin_count : 0u; }

fgb_status fgb_spatial_use_pbm(fgb_spatial *sp, unsigned int *pbm) {
  if (!sp || !pbm) return FGB_ERR_INVALID_ARG;
  FGB_CHECK(cudaDeviceSynchronize());
  if (sp->md.PBM && !sp->pbm_external) cudaFree(sp->md.PBM);
  sp->md.PBM = pbm;
  sp->pbm_external = true;
  FGB_CHECK(cudaMemcpy(sp->d_md, &sp->md, sizeof(sp->md), cudaMemcpyHostToDevice));
  return FGB_OK;
}

const void *fgb_spatial_metadata_device_ptr(const fgb_spatial *sp) { return sp ? sp->d_md : nullptr; }

fgb_status fgb_spatial_get_window(const fgb_spatial *sp, int *plane_begin, int *plane_count) {
  if (!sp) return FGB_ERR_INVALID_ARG;
  if (plane_begin) *plane_begin = sp->win_begin;
  if (plane_count) *plane_count = sp->win_count;
  return FGB_OK;
}

fgb_status fgb_plane_flags(fgb_ctx *ctx, const float *pos, unsigned int n, const unsigned int *d_n, float env_min,
                           float radius, int grid_dim, int lo, int hi, unsigned int *flag_lo, unsigned int *flag_mid,
                           unsigned int *flag_hi, void *stream) {
  if (!ctx || (n && !pos)) return FGB_ERR_INVALID_ARG;
  if (n == 0) return FGB_OK;
  k_plane_flags<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(pos, n, d_n, env_min, radius, grid_dim, lo, hi,
                                                                              flag_lo, flag_mid, flag_hi);
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_slab_signal(fgb_ctx *ctx, unsigned long long *peer_flag_lo, unsigned long long *peer_flag_hi, const unsigned int *d_epoch,
                           void *stream) {
  if (!ctx || !d_epoch) return FGB_ERR_INVALID_ARG;
  if (!peer_flag_lo && !peer_flag_hi) return FGB_OK;
  k_slab_signal<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(peer_flag_lo, peer_flag_hi, d_epoch);
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_slab_wait(fgb_ctx *ctx, const unsigned long long *flag_lo, const unsigned long long *flag_hi, const unsigned int *count_lo,
                         const unsigned int *count_hi, unsigned int capacity, const unsigned int *d_epoch, unsigned int *d_err,
                         unsigned int timeout_ms, void *stream) {
  if (!ctx || !d_epoch || !d_err || (flag_lo && !count_lo) || (flag_hi && !count_hi)) return FGB_ERR_INVALID_ARG;
  if (!flag_lo && !flag_hi) return FGB_OK;
  k_slab_wait<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flag_lo, flag_hi, count_lo, count_hi, capacity, d_epoch, d_err,
                                                              static_cast<unsigned long long>(timeout_ms) * 1000000ull);
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_slab_reserve(fgb_ctx *ctx, unsigned int stream_id, unsigned int capacity) {
  if (!ctx || stream_id >= FGB_MAX_STREAMS) return FGB_ERR_INVALID_ARG;
  return reserve_zeroed(ctx->slot[stream_id].slab, 256 + static_cast<size_t>(capacity) * 4 * 6);
}

fgb_status fgb_slab_migrate_out(fgb_ctx *ctx, unsigned int stream_id, const float *pos, unsigned int n, const unsigned int *d_n, float env_min,
                                float radius, int grid_dim, int lo_plane, int hi_plane, unsigned int capacity, const fgb_var *list_vars,
                                unsigned int nvars, void *const *peer_lo, void *const *peer_hi, unsigned int *peer_count_lo,
                                unsigned int *peer_count_hi, unsigned int *d_n_inout, unsigned int *d_err, void *stream) {
  if (!ctx || stream_id >= FGB_MAX_STREAMS || !d_n_inout || !d_err || (n && !pos) || capacity == 0 ||
      2ull * capacity > static_cast<unsigned long long>(kSlabHoleBits) || (peer_lo && !peer_count_lo) || (peer_hi && !peer_count_hi))
    return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  fgb_stream_scratch &s = ctx->slot[stream_id];
  const size_t head = 256;
  int r = reserve_zeroed(s.slab, head + static_cast<size_t>(capacity) * 4 * 6);
  if (r) return r;
  char *base = static_cast<char *>(s.slab.p);
  SlabSel *sel = reinterpret_cast<SlabSel *>(base);
  unsigned int *done = reinterpret_cast<unsigned int *>(base + 64);
  uint32_t *idx_lo = reinterpret_cast<uint32_t *>(base + head);
  uint32_t *idx_hi = idx_lo + capacity;
  uint32_t *to = idx_hi + capacity;
  uint32_t *from = to + 2 * static_cast<size_t>(capacity);
  VarTable vt_list, vt_lo{}, vt_hi{};
  r = make_var_table(list_vars, nvars, &vt_list);
  if (r) return r;
  vt_lo.n = vt_hi.n = 0;
  for (int side = 0; side < 2; ++side) {
    void *const *peer = side == 0 ? peer_lo : peer_hi;
    if (!peer) continue;
    VarTable &vt = side == 0 ? vt_lo : vt_hi;
    vt = vt_list;
    for (unsigned int v = 0; v < nvars; ++v) vt.out[v] = static_cast<char *>(peer[v]);
  }
  if (n) k_slab_select<<<(n + 255) / 256, 256, 0, st>>>(pos, n, d_n, env_min, radius, grid_dim, lo_plane, hi_plane, capacity, sel, idx_lo, idx_hi);
  const unsigned int pgrid = std::max(1u, std::min((capacity + 255u) / 256u, 4u * kNumSMs));
  k_slab_pack<<<dim3(pgrid, 2), 256, 0, st>>>(sel, idx_lo, idx_hi, capacity, vt_lo, vt_hi, peer_count_lo, peer_count_hi, nullptr);
  k_slab_holes<<<1, 1024, 0, st>>>(sel, idx_lo, idx_hi, capacity, n, d_n, to, from, d_err);
  // in == out == the list's own columns: tail agents move into the holes
  VarTable vt_move = vt_list;
  for (unsigned int v = 0; v < nvars; ++v) vt_move.out[v] = const_cast<char *>(vt_list.in[v]);
  const unsigned int fgrid = std::max(1u, std::min((2u * capacity + 255u) / 256u, 2u * kNumSMs));
  k_slab_fill<<<fgrid, 256, 0, st>>>(sel, to, from, vt_move, d_n_inout, done);
  ctx->launches += n ? 4 : 3;
  return launch_ok();
}

fgb_status fgb_slab_pack_planes(fgb_ctx *ctx, unsigned int stream_id, const float *pos, unsigned int n, const unsigned int *d_n, float env_min,
                                float radius, int grid_dim, int lo_plane, int hi_plane, unsigned int capacity, const fgb_var *list_vars,
                                unsigned int nvars, void *const *peer_lo, void *const *peer_hi, unsigned int *peer_count_lo,
                                unsigned int *peer_count_hi, void *stream) {
  if (!ctx || stream_id >= FGB_MAX_STREAMS || (n && !pos) || capacity == 0 || (peer_lo && !peer_count_lo) || (peer_hi && !peer_count_hi))
    return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  fgb_stream_scratch &s = ctx->slot[stream_id];
  const size_t head = 256;
  int r = reserve_zeroed(s.slab, head + static_cast<size_t>(capacity) * 4 * 6);
  if (r) return r;
  char *base = static_cast<char *>(s.slab.p);
  SlabSel *sel = reinterpret_cast<SlabSel *>(base);
  unsigned int *done = reinterpret_cast<unsigned int *>(base + 64);
  uint32_t *idx_lo = reinterpret_cast<uint32_t *>(base + head);
  uint32_t *idx_hi = idx_lo + capacity;
  VarTable vt_list, vt_lo{}, vt_hi{};
  r = make_var_table(list_vars, nvars, &vt_list);
  if (r) return r;
  vt_lo.n = vt_hi.n = 0;
  for (int side = 0; side < 2; ++side) {
    void *const *peer = side == 0 ? peer_lo : peer_hi;
    if (!peer) continue;
    VarTable &vt = side == 0 ? vt_lo : vt_hi;
    vt = vt_list;
    for (unsigned int v = 0; v < nvars; ++v) vt.out[v] = static_cast<char *>(peer[v]);
  }
  if (n) k_slab_select<<<(n + 255) / 256, 256, 0, st>>>(pos, n, d_n, env_min, radius, grid_dim, lo_plane, hi_plane, capacity, sel, idx_lo, idx_hi);
  const unsigned int pgrid = std::max(1u, std::min((capacity + 255u) / 256u, 4u * kNumSMs));
  k_slab_pack<<<dim3(pgrid, 2), 256, 0, st>>>(sel, idx_lo, idx_hi, capacity, vt_lo, vt_hi, peer_count_lo, peer_count_hi, done);
  ctx->launches += n ? 2 : 1;
  return launch_ok();
}

fgb_status fgb_slab_check_bound(fgb_ctx *ctx, const unsigned int *d_count, unsigned int bound, unsigned int *d_err, void *stream) {
  if (!ctx || !d_count || !d_err) return FGB_ERR_INVALID_ARG;
  k_slab_check_bound<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(d_count, bound, d_err);
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_slab_allreduce(fgb_ctx *ctx, int op, int dtype, void *d_value_inout, void *const *mailboxes, int rank, int world,
                              unsigned long long *d_epoch, unsigned int *d_err, unsigned int timeout_ms, void *stream) {
  if (!ctx || !d_value_inout || !mailboxes || !d_err || !d_epoch || world < 1 || world > kSlabMaxWorld || rank < 0 || rank >= world ||
      op < FGB_REDUCE_SUM || op > FGB_REDUCE_MAX || dtype < FGB_F32 || dtype > FGB_U64)
    return FGB_ERR_INVALID_ARG;
  SlabMailboxes mb{};
  for (int r = 0; r < world; ++r) {
    if (!mailboxes[r]) return FGB_ERR_INVALID_ARG;
    mb.box[r] = static_cast<SlabMailSlot *>(mailboxes[r]);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned long long to = static_cast<unsigned long long>(timeout_ms) * 1000000ull;
  switch (dtype) {
    case FGB_F32: k_slab_allreduce<float><<<1, 32, 0, st>>>(static_cast<float *>(d_value_inout), mb, rank, world, op, d_epoch, d_err, to); break;
    case FGB_F64: k_slab_allreduce<double><<<1, 32, 0, st>>>(static_cast<double *>(d_value_inout), mb, rank, world, op, d_epoch, d_err, to); break;
    case FGB_I32: k_slab_allreduce<int><<<1, 32, 0, st>>>(static_cast<int *>(d_value_inout), mb, rank, world, op, d_epoch, d_err, to); break;
    case FGB_U32: k_slab_allreduce<unsigned int><<<1, 32, 0, st>>>(static_cast<unsigned int *>(d_value_inout), mb, rank, world, op, d_epoch, d_err, to); break;
    case FGB_I64: k_slab_allreduce<long long><<<1, 32, 0, st>>>(static_cast<long long *>(d_value_inout), mb, rank, world, op, d_epoch, d_err, to); break;
    default: k_slab_allreduce<unsigned long long><<<1, 32, 0, st>>>(static_cast<unsigned long long *>(d_value_inout), mb, rank, world, op, d_epoch, d_err, to); break;
  }
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_spatial_read_pbm(const fgb_spatial *sp, unsigned int *host_out, void *stream) {
  if (!sp || !host_out) return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FGB_CHECK(cudaMemcpyAsync(host_out, sp->md.PBM, (static_cast<size_t>(sp->bin_count) + 1) * 4, cudaMemcpyDeviceToHost,
                            st));
  FGB_CHECK(cudaStreamSynchronize(st));
  return FGB_OK;
}

fgb_status fgb_spatial_reserve(fgb_spatial *sp, unsigned int n_max) { return sp ? reserve_index(sp, n_max, true) : FGB_ERR_INVALID_ARG; }

fgb_status fgb_spatial_clear_histogram(fgb_spatial *sp, void *stream) {
  if (!sp) return FGB_ERR_INVALID_ARG;
  FGB_CHECK(cudaMemsetAsync(sp->d_hist, 0, (static_cast<size_t>(sp->bin_count) + 1) * 4, static_cast<cudaStream_t>(stream)));
  return FGB_OK;
}

fgb_status fgb_spatial_writer_args(fgb_spatial *sp, unsigned int n_max, unsigned int **d_keys, unsigned int **d_hist) {
  if (!sp || !d_keys || !d_hist) return FGB_ERR_INVALID_ARG;
  const int r = reserve_index(sp, n_max, false);
  if (r) return r;
  *d_keys = static_cast<unsigned int *>(sp->keys.p);
  *d_hist = sp->d_hist;
  return FGB_OK;
}

fgb_status fgb_build_index(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const float *x, const float *y,
                           const float *z, const fgb_var *vars, unsigned int nvars, unsigned int flags, void *stream) {
  return fgb_build_index_ex(sp, n, d_n, x, y, z, vars, nvars, flags & ~static_cast<unsigned int>(FGB_BUILD_KEYS_READY), nullptr, nullptr, stream);
}

fgb_status fgb_build_index_ex(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const float *x, const float *y,
                              const float *z, const fgb_var *vars, unsigned int nvars, unsigned int flags,
                              const unsigned int *d_keyed, unsigned int *src_slot_out, void *stream) {
  if (!sp) return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {  // MessageSpatial3D.cu:116-120
    FGB_CHECK(cudaMemsetAsync(sp->md.PBM, 0, (static_cast<size_t>(sp->bin_count) + 1) * 4, st));
    return FGB_OK;
  }
  if (sp->dims == 0 || !x || !y || (sp->dims == 3 && !z)) return FGB_ERR_INVALID_ARG;  // bucket lists: fgb_build_index_keys
  if (sp->dims == 3) return build_index_impl<3>(sp, n, d_n, x, y, z, vars, nvars, flags, d_keyed, src_slot_out, st);
  return build_index_impl<2>(sp, n, d_n, x, y, nullptr, vars, nvars, flags, d_keyed, src_slot_out, st);
}

/* MessageBucket::CUDAModelHandler::buildIndex (MessageBucket.cu:105-137) */
fgb_status fgb_build_index_keys(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const int *keys, const fgb_var *vars,
                                unsigned int nvars, unsigned int flags, void *stream) {
  if (!sp || sp->dims != 0) return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    FGB_CHECK(cudaMemsetAsync(sp->md.PBM, 0, (static_cast<size_t>(sp->bin_count) + 1) * 4, st));
    return FGB_OK;
  }
  if (!keys) return FGB_ERR_INVALID_ARG;
  return build_index_impl<0>(sp, n, d_n, reinterpret_cast<const float *>(keys), nullptr, nullptr, vars, nvars,
                             flags & ~static_cast<unsigned int>(FGB_BUILD_KEYS_READY), nullptr, nullptr, st);
}

fgb_status fgb_bin_permutation(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const float *x, const float *y,
                               const float *z, unsigned int *perm_out, unsigned int flags, void *stream) {
  if (!sp || !perm_out || sp->dims == 0) return FGB_ERR_INVALID_ARG;
  if (n == 0) return FGB_OK;
  if (!x || !y || (sp->dims == 3 && !z)) return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (sp->dims == 3) return bin_permutation_impl<3>(sp, n, d_n, x, y, z, perm_out, flags, st);
  return bin_permutation_impl<2>(sp, n, d_n, x, y, nullptr, perm_out, flags, st);
}

fgb_status fgb_ctx_reserve(fgb_ctx *ctx, unsigned int stream_id, unsigned int n_max, int max_bit) {
  if (!ctx || stream_id >= FGB_MAX_STREAMS || max_bit < 0 || max_bit > 30) return FGB_ERR_INVALID_ARG;
  fgb_stream_scratch &s = ctx->slot[stream_id];
  int r = reserve_zeroed(s.ctrl, 64);
  if (r) return r;
  const size_t tiles = std::max<size_t>(compact_num_tiles(n_max), 1);
  size_t state_words = tiles;
  if (max_bit > 0) {
    const RadixPlan plan = make_radix_plan(max_bit);
    const size_t tiles = std::max<size_t>(radix_num_tiles(n_max), 1);
    size_t words = static_cast<size_t>(kRsMaxPasses) * kRsMaxDigits;
    for (int p = 0; p < plan.passes; ++p) words += tiles * (static_cast<size_t>(1) << plan.bits[p]);
    r = s.rs_state.reserve(words * 4);
    if (r) return r;
    const size_t nb = static_cast<size_t>(std::max(n_max, 1u)) * 4;
    for (int i = 0; i < 2; ++i) {
      r = s.rs_keys[i].reserve(nb);
      if (r) return r;
      r = s.rs_idx[i].reserve(nb);
      if (r) return r;
    }
    r = s.perm.reserve(nb);
    if (r) return r;
  }
  return reserve_zeroed(s.tile_state, state_words * 8);
}

fgb_status fgb_exclusive_scan_u32(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *in, unsigned int *out,
                                  unsigned int n, void *stream) {
  if (!ctx || !out || (n && !in) || stream_id >= FGB_MAX_STREAMS) return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    FGB_CHECK(cudaMemsetAsync(out, 0, 4, st));
    return FGB_OK;
  }
  fgb_stream_scratch &s = ctx->slot[stream_id];
  const unsigned int tiles = scan_num_tiles(n);
  int r = reserve_zeroed(s.tile_state, static_cast<size_t>(tiles) * 8);
  if (r) return r;
  unsigned long long *state = static_cast<unsigned long long *>(s.tile_state.p);
  FGB_CHECK(cudaMemsetAsync(state, 0, static_cast<size_t>(tiles) * 8, st));
  uint32_t *inp = const_cast<uint32_t *>(in);
  if (aligned16(in) && aligned16(out))
    k_exclusive_scan<true><<<tiles, kScanThreads, 0, st>>>(inp, out, n, state, 0, 0);
  else
    k_exclusive_scan<false><<<tiles, kScanThreads, 0, st>>>(inp, out, n, state, 0, 0);
  // the compaction kernels expect clean look-back words
  FGB_CHECK(cudaMemsetAsync(state, 0, static_cast<size_t>(tiles) * 8, st));
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_compact(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *flags, int invert, unsigned int n,
                       const unsigned int *d_n, unsigned int keep_front, unsigned int out_offset,
                       const unsigned int *d_out_offset, const fgb_var *vars, unsigned int nvars,
                       unsigned int *d_out_count, unsigned int *d_out_total, void *stream) {
  return fgb_compact_limited(ctx, stream_id, flags, invert, n, d_n, keep_front, out_offset, d_out_offset, 0xFFFFFFFFu, vars, nvars,
                             d_out_count, d_out_total, stream);
}

fgb_status fgb_compact_limited(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *flags, int invert, unsigned int n,
                               const unsigned int *d_n, unsigned int keep_front, unsigned int out_offset,
                               const unsigned int *d_out_offset, unsigned int out_limit, const fgb_var *vars, unsigned int nvars,
                               unsigned int *d_out_count, unsigned int *d_out_total, void *stream) {
  if (!ctx || stream_id >= FGB_MAX_STREAMS || (n > keep_front && !flags)) return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VarTable vt;
  int r = make_var_table(vars, nvars, &vt);
  if (r) return r;
  if (n == 0) {
    if (d_out_count) FGB_CHECK(cudaMemsetAsync(d_out_count, 0, 4, st));
    if (d_out_total) {
      if (d_out_offset) FGB_CHECK(cudaMemcpyAsync(d_out_total, d_out_offset, 4, cudaMemcpyDeviceToDevice, st));
      else FGB_CHECK(cudaMemcpyAsync(d_out_total, &out_offset, 4, cudaMemcpyHostToDevice, st));
    }
    return FGB_OK;
  }
  fgb_stream_scratch &s = ctx->slot[stream_id];
  const unsigned int tiles = compact_num_tiles(n);
  r = reserve_zeroed(s.tile_state, static_cast<size_t>(tiles) * 8);
  if (r) return r;
  r = reserve_zeroed(s.ctrl, 64);
  if (r) return r;
  unsigned long long *state = static_cast<unsigned long long *>(s.tile_state.p);
  uint32_t *done = static_cast<uint32_t *>(s.ctrl.p) + 1;
  // bulk-copy write-out (fgb_compact.cuh) needs 16-byte aligned output arrays; FGB_COMPACT_BULK=0/1 overrides for A/B runs
  static const int bulk_env = [] {
    const char *e = std::getenv("FGB_COMPACT_BULK");
    return e ? std::atoi(e) : FGB_COMPACT_BULK_DEFAULT;
  }();
  if (bulk_env && vars_out_aligned(vars, nvars))
    k_compact<true><<<tiles, kCmpThreads, 0, st>>>(flags, invert, n, d_n, keep_front, out_offset, d_out_offset, out_limit, vt, state, done,
                                                   d_out_count, d_out_total);
  else
    k_compact<false><<<tiles, kCmpThreads, 0, st>>>(flags, invert, n, d_n, keep_front, out_offset, d_out_offset, out_limit, vt, state, done,
                                                    d_out_count, d_out_total);
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_reduce(fgb_ctx *ctx, unsigned int stream_id, int op, int dtype, const void *in, unsigned int n,
                      const unsigned int *d_n, void *d_out, void *stream) {
  if (!ctx || !d_out || (n && !in) || stream_id >= FGB_MAX_STREAMS || op < FGB_REDUCE_SUM || op > FGB_REDUCE_MAX || dtype < FGB_F32 ||
      dtype > FGB_U64)
    return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  fgb_stream_scratch &s = ctx->slot[stream_id];
  int r = reserve_zeroed(s.red, static_cast<size_t>(kRedMaxBlocks) * 8 + 8);
  if (r) return r;
  void *partial = s.red.p;
  uint32_t *done = reinterpret_cast<uint32_t *>(static_cast<char *>(s.red.p) + static_cast<size_t>(kRedMaxBlocks) * 8);
  unsigned int blocks = (n + kRedThreads * 8 - 1) / (kRedThreads * 8);
  blocks = blocks < 1u ? 1u : (blocks > static_cast<unsigned int>(kRedMaxBlocks) ? static_cast<unsigned int>(kRedMaxBlocks) : blocks);
  ctx->launches += 1;
  const bool sum = op == FGB_REDUCE_SUM;
  switch (dtype) {
    case FGB_F32:
      if (sum) return launch_reduce<float, double>(op, in, n, d_n, 0.0, 0.0, partial, done, d_out, blocks, st);
      return launch_reduce<float, float>(op, in, n, d_n, std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity(),
                                         partial, done, d_out, blocks, st);
    case FGB_F64:
      return launch_reduce<double, double>(op, in, n, d_n, std::numeric_limits<double>::infinity(),
                                           -std::numeric_limits<double>::infinity(), partial, done, d_out, blocks, st);
    case FGB_I32:
      if (sum) return launch_reduce<int, long long>(op, in, n, d_n, 0ll, 0ll, partial, done, d_out, blocks, st);
      return launch_reduce<int, int>(op, in, n, d_n, std::numeric_limits<int>::max(), std::numeric_limits<int>::lowest(), partial, done,
                                     d_out, blocks, st);
    case FGB_U32:
      if (sum) return launch_reduce<unsigned int, unsigned long long>(op, in, n, d_n, 0ull, 0ull, partial, done, d_out, blocks, st);
      return launch_reduce<unsigned int, unsigned int>(op, in, n, d_n, std::numeric_limits<unsigned int>::max(), 0u, partial, done, d_out,
                                                       blocks, st);
    case FGB_I64:
      return launch_reduce<long long, long long>(op, in, n, d_n, std::numeric_limits<long long>::max(),
                                                 std::numeric_limits<long long>::lowest(), partial, done, d_out, blocks, st);
    default:
      return launch_reduce<unsigned long long, unsigned long long>(op, in, n, d_n, std::numeric_limits<unsigned long long>::max(), 0ull,
                                                                   partial, done, d_out, blocks, st);
  }
}

fgb_status fgb_transform_reduce(fgb_ctx *ctx, unsigned int stream_id, int transform, int dtype, const void *in, unsigned int n,
                                const unsigned int *d_n, const void *param, void *d_out, void *stream) {
  if (!ctx || !d_out || !param || (n && !in) || stream_id >= FGB_MAX_STREAMS || dtype < FGB_F32 || dtype > FGB_U64 ||
      (transform != FGB_TRANSFORM_COUNT_EQUAL && transform != FGB_TRANSFORM_SUM_SQ_DEV))
    return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  fgb_stream_scratch &s = ctx->slot[stream_id];
  int r = reserve_zeroed(s.red, static_cast<size_t>(kRedMaxBlocks) * 8 + 8);
  if (r) return r;
  void *partial = s.red.p;
  uint32_t *done = reinterpret_cast<uint32_t *>(static_cast<char *>(s.red.p) + static_cast<size_t>(kRedMaxBlocks) * 8);
  unsigned int blocks = (n + kRedThreads * 8 - 1) / (kRedThreads * 8);
  blocks = blocks < 1u ? 1u : (blocks > static_cast<unsigned int>(kRedMaxBlocks) ? static_cast<unsigned int>(kRedMaxBlocks) : blocks);
  ctx->launches += 1;
  const double mean = transform == FGB_TRANSFORM_SUM_SQ_DEV ? *static_cast<const double *>(param) : 0.0;
  const bool cnt = transform == FGB_TRANSFORM_COUNT_EQUAL;
  switch (dtype) {
    case FGB_F32: return launch_transform_reduce<float>(transform, in, n, d_n, mean, cnt ? *static_cast<const float *>(param) : 0.f, partial, done, d_out, blocks, st);
    case FGB_F64: return launch_transform_reduce<double>(transform, in, n, d_n, mean, cnt ? *static_cast<const double *>(param) : 0.0, partial, done, d_out, blocks, st);
    case FGB_I32: return launch_transform_reduce<int>(transform, in, n, d_n, mean, cnt ? *static_cast<const int *>(param) : 0, partial, done, d_out, blocks, st);
    case FGB_U32: return launch_transform_reduce<unsigned int>(transform, in, n, d_n, mean, cnt ? *static_cast<const unsigned int *>(param) : 0u, partial, done, d_out, blocks, st);
    case FGB_I64: return launch_transform_reduce<long long>(transform, in, n, d_n, mean, cnt ? *static_cast<const long long *>(param) : 0ll, partial, done, d_out, blocks, st);
    default: return launch_transform_reduce<unsigned long long>(transform, in, n, d_n, mean, cnt ? *static_cast<const unsigned long long *>(param) : 0ull, partial, done, d_out, blocks, st);
  }
}

fgb_status fgb_scatter_all(fgb_ctx *ctx, const fgb_var *vars, unsigned int nvars, unsigned int n,
                           const unsigned int *d_n, unsigned int out_offset, const unsigned int *d_out_offset,
                           void *stream) {
  if (!ctx) return FGB_ERR_INVALID_ARG;
  if (n == 0 || nvars == 0) return FGB_OK;  // CUDAScatter.cu:225-226
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VarTable vt;
  int r = make_var_table(vars, nvars, &vt);
  if (r) return r;
  const unsigned int grid = (n + 1023) / 1024;
  if (vars_in_aligned(vars, nvars) && vars_out_aligned(vars, nvars))
    k_scatter_all<true><<<grid, 256, 0, st>>>(n, d_n, out_offset, d_out_offset, vt);
  else
    k_scatter_all<false><<<grid, 256, 0, st>>>(n, d_n, out_offset, d_out_offset, vt);
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_gather(fgb_ctx *ctx, const unsigned int *position, const fgb_var *vars, unsigned int nvars,
                      unsigned int n, const unsigned int *d_n, void *stream) {
  if (!ctx || (n && !position)) return FGB_ERR_INVALID_ARG;
  if (n == 0 || nvars == 0) return FGB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VarTable vt;
  int r = make_var_table(vars, nvars, &vt);
  if (r) return r;
  if (aligned16(position) && vars_out_aligned(vars, nvars))
    k_gather<true><<<bin_grid(n), kBinThreads, 0, st>>>(position, n, d_n, vt);
  else
    k_gather<false><<<bin_grid(n), kBinThreads, 0, st>>>(position, n, d_n, vt);
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_broadcast_init(fgb_ctx *ctx, const fgb_var *vars, unsigned int nvars, unsigned int n,
                              unsigned int out_offset, void *stream) {
  if (!ctx) return FGB_ERR_INVALID_ARG;
  if (n == 0 || nvars == 0) return FGB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VarTable vt;
  int r = make_var_table(vars, nvars, &vt);
  if (r) return r;
  k_broadcast_init<<<(n + 255) / 256, 256, 0, st>>>(n, out_offset, vt);
  ctx->launches += 1;
  return launch_ok();
}

/* CUDAScatter::arrayMessageReorder (CUDAScatter.cu:567-655) */
fgb_status fgb_array_reorder(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *index, unsigned int array_length, const fgb_var *vars,
                             unsigned int nvars, unsigned int n, const unsigned int *d_n, unsigned int *d_write_count,
                             unsigned int *d_max_writes, void *stream) {
  if (!ctx || stream_id >= FGB_MAX_STREAMS || (n && !index) || (d_max_writes && !d_write_count)) return FGB_ERR_INVALID_ARG;
  if (n > array_length) return FGB_ERR_INVALID_ARG;  // "Too many messages output for array message structure" (:579-581)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VarTable vt;
  int r = make_var_table(vars, nvars, &vt);
  if (r) return r;
  if (n) {
    k_array_reorder<<<(n + 255) / 256, 256, 0, st>>>(index, array_length, n, d_n, vt, d_write_count);
    ctx->launches += 1;
  }
  if (d_max_writes) {
    fgb_stream_scratch &s = ctx->slot[stream_id];
    r = reserve_zeroed(s.red, static_cast<size_t>(kRedMaxBlocks) * 8 + 8);
    if (r) return r;
    uint32_t *partial = static_cast<uint32_t *>(s.red.p);
    uint32_t *done = reinterpret_cast<uint32_t *>(static_cast<char *>(s.red.p) + static_cast<size_t>(kRedMaxBlocks) * 8);
    unsigned int blocks = (array_length + 2047) / 2048;
    blocks = blocks < 1u ? 1u : (blocks > static_cast<unsigned int>(kRedMaxBlocks) ? static_cast<unsigned int>(kRedMaxBlocks) : blocks);
    k_array_conflicts<<<blocks, 256, 0, st>>>(d_write_count, array_length, partial, done, d_max_writes);
    ctx->launches += 1;
  }
  return launch_ok();
}

/* CUDAScatter::scatterNewAgents (CUDAScatter.cu:367-395) */
fgb_status fgb_scatter_new_agents(fgb_ctx *ctx, const void *d_aos, unsigned int agent_size, const fgb_var *vars, unsigned int nvars,
                                  unsigned int n, unsigned int out_offset, const unsigned int *d_out_offset, void *stream) {
  if (!ctx || (n && !d_aos) || agent_size == 0) return FGB_ERR_INVALID_ARG;
  if (n == 0 || nvars == 0) return FGB_OK;
  VarTable vt;
  int r = make_var_table(vars, nvars, &vt);
  if (r) return r;
  const char *base = static_cast<const char *>(d_aos);
  for (unsigned int v = 0; v < nvars; ++v)
    if (vt.in[v] < base || vt.in[v] + vt.len[v] > base + agent_size) return FGB_ERR_INVALID_ARG;  // every variable lies inside the struct
  const size_t tile_bytes = static_cast<size_t>(kNaTile) * agent_size;
  const int staged = tile_bytes <= 48 * 1024 ? 1 : 0;
  k_new_agents<<<(n + kNaTile - 1) / kNaTile, 256, staged ? tile_bytes : 0, static_cast<cudaStream_t>(stream)>>>(n, agent_size, out_offset, d_out_offset,
                                                                                                                  base, vt, staged);
  ctx->launches += 1;
  return launch_ok();
}

/* cub::DeviceHistogram::HistogramEven as called by HostAgentAPI::histogramEven (HostAgentAPI.cuh:720-745) */
fgb_status fgb_histogram_even(fgb_ctx *ctx, int dtype, const void *in, unsigned int n, const unsigned int *d_n, unsigned int bins,
                              double lower, double upper, unsigned int *d_counts, void *stream) {
  if (!ctx || !d_counts || (n && !in) || bins == 0 || !(upper > lower) || dtype < FGB_F32 || dtype > FGB_U64) return FGB_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FGB_CHECK(cudaMemsetAsync(d_counts, 0, static_cast<size_t>(bins) * 4, st));
  if (n == 0) return FGB_OK;
  unsigned int blocks = (n + 2047) / 2048;
  blocks = blocks > 4u * kNumSMs ? 4u * kNumSMs : blocks;
  ctx->launches += 1;
  switch (dtype) {
    case FGB_F32: k_histogram_even<float><<<blocks, 256, 0, st>>>(static_cast<const float *>(in), n, d_n, bins, static_cast<float>(lower), static_cast<float>(upper), d_counts); break;
    case FGB_F64: k_histogram_even<double><<<blocks, 256, 0, st>>>(static_cast<const double *>(in), n, d_n, bins, lower, upper, d_counts); break;
    case FGB_I32: k_histogram_even<int><<<blocks, 256, 0, st>>>(static_cast<const int *>(in), n, d_n, bins, static_cast<int>(lower), static_cast<int>(upper), d_counts); break;
    case FGB_U32: k_histogram_even<unsigned int><<<blocks, 256, 0, st>>>(static_cast<const unsigned int *>(in), n, d_n, bins, static_cast<unsigned int>(lower), static_cast<unsigned int>(upper), d_counts); break;
    case FGB_I64: k_histogram_even<long long><<<blocks, 256, 0, st>>>(static_cast<const long long *>(in), n, d_n, bins, static_cast<long long>(lower), static_cast<long long>(upper), d_counts); break;
    default: k_histogram_even<unsigned long long><<<blocks, 256, 0, st>>>(static_cast<const unsigned long long *>(in), n, d_n, bins, static_cast<unsigned long long>(lower), static_cast<unsigned long long>(upper), d_counts); break;
  }
  return launch_ok();
}

fgb_status fgb_sort_keys(fgb_ctx *ctx, const float *x, const float *y, const float *z, const float *env_min,
                         const float *env_width, const unsigned int *grid_dim, unsigned int n,
                         const unsigned int *d_n, unsigned int *keys_out, void *stream) {
  if (!ctx || !x || !y || !env_min || !env_width || !grid_dim || !keys_out) return FGB_ERR_INVALID_ARG;
  if (n == 0) return FGB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SortGeo g{};
  g.min0 = env_min[0]; g.min1 = env_min[1]; g.min2 = z ? env_min[2] : 0.f;
  g.w0 = env_width[0]; g.w1 = env_width[1]; g.w2 = z ? env_width[2] : 1.f;
  g.g0 = grid_dim[0]; g.g1 = grid_dim[1]; g.g2 = z ? grid_dim[2] : 1u;
  if (z)
    k_sort_keys<3><<<(n + 255) / 256, 256, 0, st>>>(x, y, z, g, n, d_n, keys_out);
  else
    k_sort_keys<2><<<(n + 255) / 256, 256, 0, st>>>(x, y, nullptr, g, n, d_n, keys_out);
  ctx->launches += 1;
  return launch_ok();
}

fgb_status fgb_sort_by_key(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *keys, int max_bit,
                           unsigned int n, const unsigned int *d_n, const fgb_var *vars, unsigned int nvars,
                           unsigned int *position_out, void *stream) {
  if (!ctx || stream_id >= FGB_MAX_STREAMS || max_bit < 1 || max_bit > 30 || (n && !keys)) return FGB_ERR_INVALID_ARG;
  if (n == 0) return FGB_OK;
  return sort_impl(ctx, stream_id, const_cast<unsigned int *>(keys), max_bit, n, d_n, vars, nvars, position_out,
                   static_cast<cudaStream_t>(stream), nullptr, nullptr, nullptr, nullptr);
}

fgb_status fgb_sort_spatial(fgb_ctx *ctx, unsigned int stream_id, const float *x, const float *y, const float *z,
                            const float *env_min, const float *env_width, const unsigned int *grid_dim, int max_bit,
                            unsigned int n, const unsigned int *d_n, unsigned int *keys_out, const fgb_var *vars,
                            unsigned int nvars, unsigned int *position_out, void *stream) {
  if (!ctx || stream_id >= FGB_MAX_STREAMS || max_bit < 1 || max_bit > 30 || !x || !y || !env_min || !env_width || !grid_dim ||
      !keys_out)
    return FGB_ERR_INVALID_ARG;
  if (n == 0) return FGB_OK;
  SortGeo g{};
  g.min0 = env_min[0]; g.min1 = env_min[1]; g.min2 = z ? env_min[2] : 0.f;
  g.w0 = env_width[0]; g.w1 = env_width[1]; g.w2 = z ? env_width[2] : 1.f;
  g.g0 = grid_dim[0]; g.g1 = grid_dim[1]; g.g2 = z ? grid_dim[2] : 1u;
  return sort_impl(ctx, stream_id, keys_out, max_bit, n, d_n, vars, nvars, position_out, static_cast<cudaStream_t>(stream), x, y, z, &g);
}

}  // extern "C"
