/*
 * flamegpu2_b200.h -- C ABI of the B200-native (sm_100a) implementation of FLAME GPU 2's
 * per-step spatial hot path: spatial message binning (PBM build), stable compaction for agent
 * death / birth / optional messages / function conditions, the automatic spatial agent sort,
 * and the generic SoA data movement those need.
 *
 * The reference (FLAME GPU 2, v2.0.0-rc.5) has no C ABI: its seams are C++ virtual dispatch
 * (MessageSpecialisationHandler) and the CUDAScatter class.  Each entry point below names the
 * reference interface (file:line, relative to the reference root) it replaces; the C++ shim
 * classes that a maintainer drops into the reference are shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer marked "device" is a CUDA device pointer;
 *  - `stream` is a cudaStream_t passed as void*;
 *  - all calls are asynchronous on `stream` unless stated otherwise and never synchronise the
 *    device (the reference synchronises after most of these, e.g. CUDAScatter.cu:177,336);
 *  - item counts: `n` is the host-known upper bound used to size the launch; when `d_n` is
 *    non-NULL the kernels read the actual count from that device word (<= n), so a captured
 *    CUDA graph stays valid while the population changes;
 *  - `stream_id` (< 128) selects the per-stream scratch slot, exactly like the reference's
 *    streamResourceId argument; calls that may overlap in time must use different slots;
 *  - no process-global state: all scratch belongs to an fgb_ctx (one per simulation, any number
 *    per process, as CUDAEnsemble runs many simulations concurrently, CUDAEnsemble.cu:214-260);
 *  - return value: 0 on success, a positive cudaError_t, or a negative fgb_error.
 *  - there is NO CPU fallback: every entry point fails with FGB_ERR_NO_DEVICE without a GPU.
 */
#ifndef FLAMEGPU2_B200_H_
#define FLAMEGPU2_B200_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGB_VERSION 100

typedef int fgb_status;
enum fgb_error {
  FGB_OK = 0,
  FGB_ERR_INVALID_ARG = -1,
  FGB_ERR_TOO_MANY_VARS = -2,
  FGB_ERR_NO_DEVICE = -3,
  FGB_ERR_ALLOC = -4,
  FGB_ERR_UNSUPPORTED = -5
};

#define FGB_MAX_VARS 32

/* One SoA variable of a list being moved: same layout and meaning as
 * CUDAScatter::ScatterData {size_t typeLen; char *in; char *out;}
 * (include/flamegpu/simulation/detail/CUDAScatter.cuh:58-62).  type_len = type_size * elements. */
typedef struct fgb_var {
  size_t type_len;
  const void *in; /* device */
  void *out;      /* device */
} fgb_var;

/* Device-visible message-list metadata.  Field-for-field the layout of
 * MessageSpatial3D::MetaData (include/flamegpu/runtime/messaging/MessageSpatial3D.h:38-68); a
 * 2D list uses the first two entries of each array, grid_dim[2] == 1. */
typedef struct fgb_spatial_metadata {
  float min[3];
  float max[3];
  float radius;
  unsigned int *PBM; /* device, bin_count + 1 entries */
  unsigned int grid_dim[3];
  float environment_width[3];
  bool wrap_compatible;
} fgb_spatial_metadata;

typedef struct fgb_ctx fgb_ctx;         /* per-simulation scratch owner */
typedef struct fgb_spatial fgb_spatial; /* per message list: histogram, PBM, metadata */

/* ---- context ------------------------------------------------------------------------------ */
/* Replaces the Singletons{scatter, ...} bundle a CUDASimulation owns (CUDASimulation.h:567-591).
 * Synchronous. */
fgb_status fgb_ctx_create(int device, fgb_ctx **out);
/* Frees all scratch.  Like CUDAScatter::purge (CUDAScatter.cu:46-54) it must be called before
 * device reset, not from a static destructor.  Synchronous. */
fgb_status fgb_ctx_destroy(fgb_ctx *ctx);
const char *fgb_error_string(fgb_status s);
int fgb_version(void);

/* ---- spatial message list handler ----------------------------------------------------------- */
/* MessageSpatial3D::CUDAModelHandler ctor + init + allocateMetaDataDevicePtr
 * (src/flamegpu/runtime/messaging/MessageSpatial3D.cu:31-52,74-90); dims==2 gives the
 * MessageSpatial2D twin (MessageSpatial2D.cu:35-52,74-90).  Computes grid_dim/bin_count/
 * wrap_compatible exactly as the reference, allocates PBM (zeroed), histogram and the device
 * MetaData.  Synchronous. */
fgb_status fgb_spatial_create(fgb_ctx *ctx, int dims, const float *env_min, const float *env_max, float radius,
                              fgb_spatial **out);
/* Multi-GPU slab decomposition (no reference counterpart: a reference simulation never spans GPUs,
 * SURVEY.md section 8e): same as fgb_spatial_create, but the handler stores only planes
 * [plane_begin, plane_begin + plane_count) of the slowest grid axis (z in 3D, y in 2D).  Bin arithmetic
 * is the GLOBAL one (identical to the single-GPU build); only the plane index is rebased, so the slab
 * PBMs concatenate to the single-GPU PBM.  plane_count < 0 means the whole grid.  The metadata keeps
 * the GLOBAL grid_dim; the window is queried with fgb_spatial_get_window. */
fgb_status fgb_spatial_create_window(fgb_ctx *ctx, int dims, const float *env_min, const float *env_max, float radius,
                                     int plane_begin, int plane_count, fgb_spatial **out);
fgb_status fgb_spatial_get_window(const fgb_spatial *sp, int *plane_begin, int *plane_count);
/* Classifies n positions along one axis by grid plane, with the bin arithmetic of the PBM
 * (plane = clamp(floorf((p - env_min) / radius), 0, grid_dim - 1)):
 * flag_lo[i] = plane < lo, flag_hi[i] = plane >= hi, flag_mid[i] = neither.  Outputs may be NULL.
 * Halo selection and agent migration of the slab decomposition are fgb_compact over these flags. */
fgb_status fgb_plane_flags(fgb_ctx *ctx, const float *pos, unsigned int n, const unsigned int *d_n, float env_min,
                           float radius, int grid_dim, int lo, int hi, unsigned int *flag_lo, unsigned int *flag_mid,
                           unsigned int *flag_hi, void *stream);
/* CUDAModelHandler::freeMetaDataDevicePtr (MessageSpatial3D.cu:92-111).  Synchronous. */
fgb_status fgb_spatial_destroy(fgb_spatial *sp);
/* Host copy of the metadata (PBM is the device pointer) and the bin count. */
fgb_status fgb_spatial_get_metadata(const fgb_spatial *sp, fgb_spatial_metadata *host_out, unsigned int *bin_count);
/* MessageSpecialisationHandler::getMetaDataDevicePtr
 * (include/flamegpu/runtime/messaging/MessageSpecialisationHandler.h:44) */
const void *fgb_spatial_metadata_device_ptr(const fgb_spatial *sp);
unsigned int fgb_spatial_bin_count(const fgb_spatial *sp);  /* bins of the (windowed) PBM */
/* Drop-in use inside the reference: build into a PBM the CALLER owns -- the hd_data.PBM that
 * MessageSpatial3D::CUDAModelHandler::allocateMetaDataDevicePtr allocated (MessageSpatial3D.cu:81-90, binCount + 1
 * words) and that the reference's device MetaData already points at -- instead of the handler's own array, which
 * is freed.  fgb_spatial_destroy then leaves `pbm` alone.  Synchronous. */
fgb_status fgb_spatial_use_pbm(fgb_spatial *sp, unsigned int *pbm);

/* Inspection helper (no reference counterpart: "The PBM is never stored on the host",
 * MessageSpatial3D.h:55-58): copies the bin_count+1 PBM entries to host memory after
 * synchronising `stream`.  Used by tests and by the parity harness. */
fgb_status fgb_spatial_read_pbm(const fgb_spatial *sp, unsigned int *host_out, void *stream);

enum fgb_build_flags {
  FGB_BUILD_DEFAULT = 0,
  /* within-bin order = source order (deterministic).  Default order is arrival order of the
   * scatter, like the reference's atomicInc sub-index (MessageSpatial3D.cu:70). */
  FGB_BUILD_STABLE = 1,
  /* fgb_bin_permutation only: group equal bins inside tiles of 2048 consecutive points instead of globally
   * (one pass, the PBM is not written).  For lists that are already coarsely ordered. */
  FGB_BUILD_TILE_LOCAL = 2,
  /* fgb_build_index_ex only: the bin key of the leading items and their histogram contribution were already written
   * by the list's writer (fgb_spatial_writer_args); see fgb_build_index_ex */
  FGB_BUILD_KEYS_READY = 4,
  /* The caller expects the list to arrive (nearly) bin-grouped -- its writer ran in bin order: tiles that are not grouped
   * are scattered inside the scan + scatter launch itself (one atomic and one scattered store per message, correct for
   * any input) and the launch that walks the worklist of unordered tiles is not issued.  This is the library's default
   * behaviour (measured faster on every input order); the flag only matters for a -DFGB_BUILD_STAGED_DEFAULT=1 build.
   * Ignored with FGB_BUILD_STABLE. */
  FGB_BUILD_EXPECT_GROUPED = 8
};

/* MessageSpatial3D::CUDAModelHandler::buildIndex (MessageSpatial3D.cu:113-146) and
 * MessageSpatial2D::CUDAModelHandler::buildIndex (MessageSpatial2D.cu:113-146), including the
 * CUDAScatter::pbm_reorder they call (CUDAScatter.cu:276-337):
 *   bins every message by its location (x,y[,z] device float arrays, z ignored for 2D),
 *   writes the PBM (exclusive scan of the bin counts, bin_count+1 entries, PBM[bin_count] = n),
 *   and copies every variable vars[v].in -> vars[v].out grouped by bin.
 * x/y/z are the `in` arrays of the location variables and must also be listed in vars.
 * The caller swaps its read/write lists afterwards (CUDAMessage::swap, MessageSpatial3D.cu:139).
 * n == 0 (and d_n == NULL) zeroes the PBM like MessageSpatial3D.cu:116-120. */
fgb_status fgb_build_index(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const float *x, const float *y,
                           const float *z, const fgb_var *vars, unsigned int nvars, unsigned int flags, void *stream);

/* ---- bucket lists (MessageBucket): integer keys instead of positions -------------------------------------
 * MessageBucket::CUDAModelHandler (src/flamegpu/runtime/messaging/MessageBucket.cu:36-78): keys
 * lower_bound..upper_bound (inclusive), bucketCount = upper_bound - lower_bound + 1, PBM of bucketCount + 1 words,
 * zeroed (:69).  The handle is an fgb_spatial (destroy / read_pbm / reserve apply); positional entry points reject it. */
/* b200 extension: the fused form of buildIndex.  The reference's atomicHistogram3D (MessageSpatial3D.cu:54-72) re-reads
 * the location of every message right after the output function wrote it; here the kernel that WRITES a list can
 * publish, per message slot i, keys[i] = bin of the message (getGridPosition3D + getHash3D arithmetic, window-rebased)
 * and add 1 to hist[keys[i]] (fgb_spatial_writer_args returns both device arrays, sized for n_max items; hist must
 * only be touched between two builds).  fgb_build_index_ex(FGB_BUILD_KEYS_READY) then starts at the scan:
 *   d_keyed == NULL : every item of the list was keyed by its writer;
 *   d_keyed != NULL : only the first *d_keyed items were (device word); the items behind them -- e.g. ghost messages
 *                     appended by the slab exchange -- are keyed and counted here, the others are not read again.
 * src_slot_out (may be NULL): src_slot_out[j] = slot the message at sorted position j came from. */
fgb_status fgb_spatial_writer_args(fgb_spatial *sp, unsigned int n_max, unsigned int **d_keys, unsigned int **d_hist);
/* zeroes the histogram: only needed when a writer published counts that no fgb_build_index_ex consumed */
fgb_status fgb_spatial_clear_histogram(fgb_spatial *sp, void *stream);
fgb_status fgb_build_index_ex(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const float *x, const float *y,
                              const float *z, const fgb_var *vars, unsigned int nvars, unsigned int flags,
                              const unsigned int *d_keyed, unsigned int *src_slot_out, void *stream);

fgb_status fgb_bucket_create(fgb_ctx *ctx, int lower_bound, int upper_bound, fgb_spatial **out);
/* MessageBucket::MetaData {min, max (exclusive), PBM} (include/flamegpu/runtime/messaging/MessageBucket.h:40-55) */
fgb_status fgb_bucket_get_bounds(const fgb_spatial *sp, int *min_key, int *max_key_exclusive, const unsigned int **d_pbm);
/* MessageBucket::CUDAModelHandler::buildIndex (MessageBucket.cu:105-137): atomicHistogram1D over
 * keys[i] - min (:49-64), exclusive scan, pbm_reorder of every variable.  keys is the `in` array of the "_key"
 * variable (also listed in vars).  A key outside the bounds (a seatbelts-only error in the reference, :283-288 of
 * MessageBucketDevice.cuh) is counted in the nearest valid bucket instead of writing out of bounds. */
fgb_status fgb_build_index_keys(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const int *keys, const fgb_var *vars,
                                unsigned int nvars, unsigned int flags, void *stream);

/* B200 extension (no reference counterpart): the permutation that WOULD group `n` points by bin,
 * without moving any payload: perm_out[j] = index of the point at grouped position j, and the
 * handler's PBM receives the bin offsets of the points.  Used by the step scheduler to run an agent
 * function in bin order (warp lanes share message strips) while the agent list itself stays in the
 * reference's order.  Same kernels as fgb_build_index (histogram, scan, cursor scatter). */
fgb_status fgb_bin_permutation(fgb_spatial *sp, unsigned int n, const unsigned int *d_n, const float *x, const float *y,
                               const float *z, unsigned int *perm_out, unsigned int flags, void *stream);

/* ---- scan / compaction / data movement ------------------------------------------------------ */
/* cub::DeviceScan::ExclusiveSum as called at CUDAFatAgent.cu:118-132, CUDAAgentStateList.cu:195-210,
 * CUDAMessage.cu:180-194: out[i] = sum_{j<i} in[j] for i in [0,n]  (n+1 outputs, out[n] = total),
 * in[] has n valid entries.  Single pass, decoupled look-back. */
fgb_status fgb_exclusive_scan_u32(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *in, unsigned int *out,
                                  unsigned int n, void *stream);

/* Scan-flag compaction in ONE pass (flags -> positions -> move), replacing the pair
 * cub::DeviceScan::ExclusiveSum + CUDAScatter::scatter / scatter_generic
 * (CUDAFatAgent.cu:105-137 death, CUDAFatAgent.cu:186-236 function condition,
 *  CUDAMessage.cu:171-208 optional messages, CUDAAgentStateList.cu:189-250 birth append;
 *  kernel CUDAScatter.cu:67-88, wrapper CUDAScatter.cu:138-179).
 * Item i in [0,n) is kept if i < keep_front (scatter_all_count) or
 * (flags[i - keep_front] == 1) != invert (InversionIterator, CUDAScatter.cuh:32-49);
 * kept items keep their relative order and go to vars[v].out[out_offset + rank].
 * The kept count (including keep_front) is written to *d_out_count (device, may be NULL);
 * when d_out_total is non-NULL it receives out_offset + kept (the new list size after an append).
 * out_offset may also be read from the device word d_out_offset when that is non-NULL. */
fgb_status fgb_compact(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *flags, int invert, unsigned int n,
                       const unsigned int *d_n, unsigned int keep_front, unsigned int out_offset, const unsigned int *d_out_offset,
                       const fgb_var *vars, unsigned int nvars, unsigned int *d_out_count, unsigned int *d_out_total,
                       void *stream);

/* fgb_compact with a bounded destination: only the first out_limit kept items are written, the counts
 * still report every kept item (overflow is detectable, memory is never overrun).  Used for the
 * fixed-capacity halo / migration staging buffers of the multi-GPU slab decomposition. */
fgb_status fgb_compact_limited(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *flags, int invert, unsigned int n,
                               const unsigned int *d_n, unsigned int keep_front, unsigned int out_offset,
                               const unsigned int *d_out_offset, unsigned int out_limit, const fgb_var *vars, unsigned int nvars,
                               unsigned int *d_out_count, unsigned int *d_out_total, void *stream);

/* CUDAScatter::scatterAll / scatter_all_generic (CUDAScatter.cu:105-117,219-262):
 * out[out_offset + i] = in[i] for every variable (state transition append, message append). */
fgb_status fgb_scatter_all(fgb_ctx *ctx, const fgb_var *vars, unsigned int nvars, unsigned int n,
                           const unsigned int *d_n, unsigned int out_offset, const unsigned int *d_out_offset,
                           void *stream);

/* CUDAScatter::scatterPosition / scatter_position_generic (CUDAScatter.cu:89-104,180-218):
 * out[i] = in[position[i]] for every variable (the gather that applies a sort). */
fgb_status fgb_gather(fgb_ctx *ctx, const unsigned int *position, const fgb_var *vars, unsigned int nvars,
                      unsigned int n, const unsigned int *d_n, void *stream);

/* CUDAScatter::broadcastInit / broadcastInitKernel (CUDAScatter.cu:404-422,490-539):
 * fills out[out_offset + i] of every variable with the type_len bytes at vars[v].in (a DEVICE
 * pointer to one default value). */
fgb_status fgb_broadcast_init(fgb_ctx *ctx, const fgb_var *vars, unsigned int nvars, unsigned int n,
                              unsigned int out_offset, void *stream);

/* ---- reductions behind HostAgentAPI::sum / min / max -------------------------------------------------------
 * include/flamegpu/runtime/agent/HostAgentAPI.cuh:540-700 (cub::DeviceReduce::{Sum,Min,Max} + copy of the result).
 * d_out receives ONE 8-byte value: sums of FGB_F32/FGB_F64 as double, sums of the integer types as 64-bit integers
 * of the same signedness; min / max in the element type (low bytes of the word).  An empty input gives the
 * operation's identity (0 for sums).  Nothing returns to the host; the caller copies the word when it needs it. */
enum fgb_reduce_op { FGB_REDUCE_SUM = 0, FGB_REDUCE_MIN = 1, FGB_REDUCE_MAX = 2 };
enum fgb_dtype { FGB_F32 = 0, FGB_F64 = 1, FGB_I32 = 2, FGB_U32 = 3, FGB_I64 = 4, FGB_U64 = 5 };
fgb_status fgb_reduce(fgb_ctx *ctx, unsigned int stream_id, int op, int dtype, const void *in, unsigned int n,
                      const unsigned int *d_n, void *d_out, void *stream);

/* HostAgentAPI::count(variable, value) (thrust::count, HostAgentAPI.cuh:700-718) and the second pass of
 * HostAgentAPI::meanStandardDeviation (sum of (x - mean)^2, :598-600).  param points to HOST memory: the value to
 * count in the element type, or the mean as a double.  d_out receives 8 bytes: an unsigned 64-bit count, or a double. */
enum fgb_transform { FGB_TRANSFORM_COUNT_EQUAL = 1, FGB_TRANSFORM_SUM_SQ_DEV = 2 };
fgb_status fgb_transform_reduce(fgb_ctx *ctx, unsigned int stream_id, int transform, int dtype, const void *in, unsigned int n,
                                const unsigned int *d_n, const void *param, void *d_out, void *stream);

/* ---- automatic spatial agent sort ----------------------------------------------------------- */
/* calculateSpatialHash kernels (src/flamegpu/simulation/CUDASimulation.cu:335-408):
 * key = floorf(((p-min)/width)*grid_dim) linearised, NOT clamped.  z == NULL for 2D.
 * grid_dim as CUDASimulation.cu:498-505 computes it (ceilf(width/radius), 1 when width == 0). */
fgb_status fgb_sort_keys(fgb_ctx *ctx, const float *x, const float *y, const float *z, const float *env_min,
                         const float *env_width, const unsigned int *grid_dim, unsigned int n,
                         const unsigned int *d_n, unsigned int *keys_out, void *stream);

/* HostAgentAPI::sort_async<unsigned int> (include/flamegpu/runtime/agent/HostAgentAPI.cuh:858-914:
 * fillTIDArray + cub::DeviceRadixSort::SortPairs on bits [0,max_bit) + scatterSort_async) fused:
 * stable sort of the n items by (keys[i] & ((1<<max_bit)-1)), applied to every variable
 * (vars[v].in -> vars[v].out).  If position_out is non-NULL it receives the permutation
 * (position_out[j] = source index of sorted item j). */
fgb_status fgb_sort_by_key(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *keys, int max_bit,
                           unsigned int n, const unsigned int *d_n, const fgb_var *vars, unsigned int nvars,
                           unsigned int *position_out, void *stream);

/* fgb_sort_keys + fgb_sort_by_key in one call (the whole of CUDASimulation::spatialSortAgent_async,
 * CUDASimulation.cu:463-573): the key of every agent is computed, stored to keys_out (the agent's
 * _auto_sort_bin_index, which is also listed in vars and therefore travels with the sort) and counted into the digit
 * histograms in ONE pass; results are identical to the two separate calls. */
fgb_status fgb_sort_spatial(fgb_ctx *ctx, unsigned int stream_id, const float *x, const float *y, const float *z,
                            const float *env_min, const float *env_width, const unsigned int *grid_dim, int max_bit,
                            unsigned int n, const unsigned int *d_n, unsigned int *keys_out, const fgb_var *vars,
                            unsigned int nvars, unsigned int *position_out, void *stream);

/* Scratch (look-back words, histograms, permutations) grows on demand with cudaMalloc, which is
 * not allowed during CUDA stream capture: reserve for the largest n / max_bit up front (or run
 * one warm-up call outside the capture).  stream_id < 128 selects the scratch slot, like the
 * reference's streamResourceId (CUDAScatter.cuh:100-317). */
fgb_status fgb_ctx_reserve(fgb_ctx *ctx, unsigned int stream_id, unsigned int n_max, int max_bit);
fgb_status fgb_spatial_reserve(fgb_spatial *sp, unsigned int n_max);

/* ---- the remaining CUDAScatter kernels and the histogram behind HostAgentAPI::histogramEven ------------- */
/* CUDAScatter::arrayMessageReorder + reorder_array_messages (CUDAScatter.cu:540-655): array messages (MessageArray /
 * MessageArray2D / MessageArray3D) are written at the sender's thread index together with their target element
 * (`index`, the "___INDEX" variable); this moves every variable to out[index[i]].  An index >= array_length is dropped
 * (:553-554); n > array_length is an error (:579-581, FGB_ERR_INVALID_ARG).  d_write_count (array_length words, all-zero
 * on entry, re-zeroed by the call; may be NULL) counts the writes per element; d_max_writes (may be NULL) receives their
 * maximum -- > 1 is the reference's ArrayMessageWriteConflict (:629-651), left to the caller to raise, so nothing here
 * synchronises. */
fgb_status fgb_array_reorder(fgb_ctx *ctx, unsigned int stream_id, const unsigned int *index, unsigned int array_length, const fgb_var *vars,
                             unsigned int nvars, unsigned int n, const unsigned int *d_n, unsigned int *d_write_count,
                             unsigned int *d_max_writes, void *stream);
/* CUDAScatter::scatterNewAgents + scatter_new_agents (CUDAScatter.cu:348-395): host-created agents arrive as `n` structs
 * of `agent_size` bytes at d_aos (device memory); vars[v].in points at variable v inside the FIRST struct, vars[v].out is
 * the state list's SoA column; agent i lands at out[out_offset + i] (offset optionally from a device word). */
fgb_status fgb_scatter_new_agents(fgb_ctx *ctx, const void *d_aos, unsigned int agent_size, const fgb_var *vars, unsigned int nvars,
                                  unsigned int n, unsigned int out_offset, const unsigned int *d_out_offset, void *stream);
/* cub::DeviceHistogram::HistogramEven as HostAgentAPI::histogramEven calls it (include/flamegpu/runtime/agent/HostAgentAPI.cuh:
 * 720-745): d_counts[b] (bins words, overwritten) = number of items with lower <= v < upper in even bin b. */
fgb_status fgb_histogram_even(fgb_ctx *ctx, int dtype, const void *in, unsigned int n, const unsigned int *d_n, unsigned int bins,
                              double lower, double upper, unsigned int *d_counts, void *stream);

/* ---- multi-GPU z-slab exchange (b200 extension; the reference has no multi-GPU simulation) -----------
 * One process per GPU.  Halo messages / migrating agents are packed by fgb_compact_limited straight into the
 * NEIGHBOUR's staging buffer (peer memory mapped with cudaIpcOpenMemHandle: vars[v].out and d_out_count are peer
 * pointers), so the pack is the transfer.  These entry points are the synchronisation around it; epochs advance by one
 * per simulation step and are read from a device word, so every launch is CUDA-graph replayable.
 *   fgb_slab_signal : peer_flag_{lo,hi} (peer memory, NULL = no neighbour) <- *d_epoch + 1, ordered after all earlier
 *                     writes of the stream (system-scope release)
 *   fgb_slab_wait   : waits until the LOCAL flag words reach *d_epoch + 1; then *count_{lo,hi} > capacity raises
 *                     FGB_SLAB_ERR_OVERFLOW in *d_err; gives up after timeout_ms with FGB_SLAB_ERR_TIMEOUT (a dead
 *                     peer must not wedge the GPU)
 *   fgb_slab_check_bound : *d_count > bound raises FGB_SLAB_ERR_BOUND (agents beyond a launch bound would not execute)
 *   fgb_slab_allreduce   : all-reduce (op: fgb_reduce_op, dtype: fgb_dtype of the 4/8-byte value) over per-rank
 *                     mailboxes (mailboxes[r] = rank r's array of 2 * world 16-byte slots, peer memory for r != rank);
 *                     folded in rank order, so every rank obtains the identical value; *d_epoch is a device word (zero
 *                     before the first call) that every call advances by one, identically on every rank, so the launch is
 *                     CUDA-graph replayable
 *   fgb_slab_migrate_out : removes the items of a list whose position `pos` lies in planes < lo_plane / >= hi_plane
 *                     (plane arithmetic of fgb_plane_flags) and writes them -- and their counts -- into the neighbours'
 *                     staging columns peer_lo[v] / peer_hi[v] (peer memory; NULL = no neighbour on that side).  Only the
 *                     position column is read in full: the few leavers are gathered by index, the holes they leave are
 *                     filled with agents from the tail of the list (list_vars[v].in is the column, .out is ignored), and
 *                     *d_n_inout becomes the new count.  More than `capacity` leavers on a side (2 * capacity must not
 *                     exceed 262144) raise FGB_SLAB_ERR_OVERFLOW here and on the receiver. */
enum { FGB_SLAB_ERR_TIMEOUT = 1, FGB_SLAB_ERR_OVERFLOW = 2, FGB_SLAB_ERR_BOUND = 4 };
/* scratch of fgb_slab_migrate_out for staging buffers of `capacity` items (call outside stream capture) */
fgb_status fgb_slab_reserve(fgb_ctx *ctx, unsigned int stream_id, unsigned int capacity);
fgb_status fgb_slab_migrate_out(fgb_ctx *ctx, unsigned int stream_id, const float *pos, unsigned int n, const unsigned int *d_n, float env_min,
                                float radius, int grid_dim, int lo_plane, int hi_plane, unsigned int capacity, const fgb_var *list_vars,
                                unsigned int nvars, void *const *peer_lo, void *const *peer_hi, unsigned int *peer_count_lo,
                                unsigned int *peer_count_hi, unsigned int *d_n_inout, unsigned int *d_err, void *stream);
/* The halo form of fgb_slab_migrate_out: the items in planes < lo_plane / >= hi_plane are COPIED to the neighbours' staging
 * columns (and their counts to *peer_count_*); the list itself is left alone.  One read of the position column + a gather
 * of the few selected items, instead of scan flags and two compactions over the whole list.  The receiver validates the
 * count against its capacity (fgb_slab_wait).  An item can be sent to one side only: the slab must own >= 2 planes. */
fgb_status fgb_slab_pack_planes(fgb_ctx *ctx, unsigned int stream_id, const float *pos, unsigned int n, const unsigned int *d_n, float env_min,
                                float radius, int grid_dim, int lo_plane, int hi_plane, unsigned int capacity, const fgb_var *list_vars,
                                unsigned int nvars, void *const *peer_lo, void *const *peer_hi, unsigned int *peer_count_lo,
                                unsigned int *peer_count_hi, void *stream);
fgb_status fgb_slab_signal(fgb_ctx *ctx, unsigned long long *peer_flag_lo, unsigned long long *peer_flag_hi, const unsigned int *d_epoch,
                           void *stream);
fgb_status fgb_slab_wait(fgb_ctx *ctx, const unsigned long long *flag_lo, const unsigned long long *flag_hi, const unsigned int *count_lo,
                         const unsigned int *count_hi, unsigned int capacity, const unsigned int *d_epoch, unsigned int *d_err,
                         unsigned int timeout_ms, void *stream);
fgb_status fgb_slab_check_bound(fgb_ctx *ctx, const unsigned int *d_count, unsigned int bound, unsigned int *d_err, void *stream);
fgb_status fgb_slab_allreduce(fgb_ctx *ctx, int op, int dtype, void *d_value_inout, void *const *mailboxes, int rank, int world,
                              unsigned long long *d_epoch, unsigned int *d_err, unsigned int timeout_ms, void *stream);

/* Number of kernels this library has launched through `ctx` (bench.py's gpu_launches). */
unsigned long long fgb_launch_count(const fgb_ctx *ctx);

/* Allocation generation of `ctx`: incremented whenever scratch owned by the context or by one of its
 * fgb_spatial handles is (re)allocated.  Kernels captured into a CUDA graph hold the old pointers, so a
 * caller that replays graphs compares this value before each replay and re-captures when it moved (the
 * reference keeps its scratch in CubTemporaryMemory / CUDAScanCompaction, which it resizes between
 * launches, never under a graph: src/flamegpu/simulation/detail/CubTemporaryMemory.cu:23-32). */
unsigned long long fgb_alloc_generation(const fgb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FLAMEGPU2_B200_H_ */
