/*
 * fgb_oracle.h -- CPU restatement of FLAME GPU 2's per-step spatial hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may include, link or call this
 * file: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / reported CPU baseline.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_golden.py) against
 * the known-answer properties of the reference's own tests (test_spatial_3d.cu Mandatory /
 * Wrapped / bounds_not_factor_radius, test_spatial_2d.cu twins, test_cuda_simulation.cu
 * AgentDeath, test_spatial_agent_sort.cu, test_device_agent_creation.cu) and, on the GPU box
 * (tests/test_ref_parity_gpu.py), against the reference's own CUDA code built unmodified from
 * /root/reference into oracle/_ref/ref_sim (PBM bit-exact, bins multiset-equal, agent order
 * bit-exact, floats within tolerance).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef FGB_ORACLE_H_
#define FGB_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/flamegpu/runtime/messaging/MessageSpatial3D.h:38-68 (MetaData), 2D twin MessageSpatial2D.h */
typedef struct orc_grid {
  int dims; /* 2 or 3 */
  float min[3];
  float max[3];
  float radius;
  uint32_t grid_dim[3];
  float env_width[3];
  int wrap_compatible;
  uint32_t bin_count;
} orc_grid;

/* src/flamegpu/runtime/messaging/MessageSpatial3D.cu:31-52, MessageSpatial2D.cu:35-52 */
void orc_grid_init(orc_grid *g, int dims, const float *mn, const float *mx, float radius);

/* MessageSpatial3DDevice.cuh:646-659 / MessageSpatial2DDevice.cuh:605-616 : clamped cell of a point */
void orc_grid_pos(const orc_grid *g, float x, float y, float z, int cell[3]);
/* MessageSpatial3DDevice.cuh:660-672 / 2D :617-628 : linear bin index (x re-clamped) */
uint32_t orc_hash(const orc_grid *g, int cx, int cy, int cz);
/* bin index of every message (what atomicHistogram3D writes to bin_index, MessageSpatial3D.cu:54-72) */
void orc_bin_keys(const orc_grid *g, uint32_t n, const float *x, const float *y, const float *z, uint32_t *keys);

/* MessageSpatial3D.cu:113-146 + CUDAScatter.cu:276-294.  pbm has bin_count+1 entries.
 * perm[j] = source index of the message that lands at sorted position j; the reference's
 * within-bin order is atomicInc arrival order, the oracle uses source order (stable), which is
 * one legal instance. */
void orc_build_index(const orc_grid *g, uint32_t n, const float *x, const float *y, const float *z,
                     uint32_t *pbm, uint32_t *perm);

/* MessageBucket (src/flamegpu/runtime/messaging/MessageBucket.cu:36-64,105-137): keys lower..upper inclusive,
 * bucketCount = upper - lower + 1, hash = key - lower (atomicHistogram1D :49-64), exclusive scan, reorder.
 * pbm has bucketCount + 1 entries; perm as in orc_build_index (stable).  Keys are expected inside the bounds. */
void orc_bucket_build(int32_t lower, int32_t upper, uint32_t n, const int32_t *keys, uint32_t *pbm, uint32_t *perm);
/* MessageBucket::In::Filter (MessageBucketDevice.cuh:264-274): messages [PBM[begin-min], PBM[end-min]) when
 * begin >= min && end < max && begin <= end, with max = upper + 1 -- so operator()(key) = Filter(key, key + 1)
 * yields nothing for key == upper (reference behaviour, kept).  Returns the count, *first = first index. */
uint32_t orc_bucket_range(int32_t lower, int32_t upper, const uint32_t *pbm, int32_t begin_key, int32_t end_key,
                          uint32_t *first);

/* out[j] = in[perm[j]] for one variable of type_len bytes per item (CUDAScatter.cu:89-104 order) */
void orc_gather(const uint32_t *perm, uint32_t n, uint32_t type_len, const void *in, void *out);

/* Indices visited by In::Filter (MessageSpatial3DDevice.cuh:693-719, 2D :644-672) in visit order.
 * Returns the count; writes at most cap indices. */
uint32_t orc_filter(const orc_grid *g, const uint32_t *pbm, float x, float y, float z, uint32_t *out_idx, uint32_t cap);
/* In::WrapFilter (MessageSpatial3DDevice.cuh:727-749, 2D :679-698) */
uint32_t orc_wrap_filter(const orc_grid *g, const uint32_t *pbm, float x, float y, float z, uint32_t *out_idx,
                         uint32_t cap);
/* WrapFilter::Message::getVirtualX/Y/Z (MessageSpatial3DDevice.cuh:373-397) */
float orc_virtual(float x2, float x1, float env_width);

/* Flag compaction, CUDAScatter.cu:67-88 + CUDAFatAgentStateList.cu:155-178: the first keep_front
 * items are copied unconditionally, the rest if (flag==1) != invert.  perm[j] = source index of
 * output j (stable).  Returns the output count. */
uint32_t orc_compact(const uint32_t *flags, int invert, uint32_t n, uint32_t keep_front, uint32_t *perm);

/* Geometry the reference hands to its sort-key kernel (CUDASimulation.cu:480-506), including its
 * quirk for 3D lists (z extent collapses to 0, see fgb_oracle.c); true3d != 0 gives the intended 3D key. */
void orc_sort_geometry(const orc_grid *g, int true3d, float mn[3], float width[3], uint32_t gd[3]);
/* Auto agent sort key, calculateSpatialHash CUDASimulation.cu:376-408 (no clamp); z may be NULL for 2D. */
void orc_sort_keys(const orc_grid *g, int true3d, uint32_t n, const float *x, const float *y, const float *z,
                   uint32_t *keys);
/* max_bit = floor(log2(bins))+1, CUDASimulation.cu:571 */
int orc_sort_max_bit(const orc_grid *g, int true3d);
/* Stable sort on the low max_bit bits (cub::DeviceRadixSort::SortPairs, HostAgentAPI.cuh:900-909):
 * perm[j] = source index of the agent at sorted position j. */
void orc_sort_perm(const uint32_t *keys, uint32_t n, int max_bit, uint32_t *perm);

/* ---- model steps (examples restated; OpenMP over agents when built with -fopenmp) ---- */

/* examples/cpp/circles_spatial3D/src/main.cu:13-54 `move` over a built index.
 * Agent i reads messages (mid,mx,my,mz sorted by bin, pbm) and writes x/y/z/drift in place.  */
void orc_circles_move(const orc_grid *g, const uint32_t *pbm, uint32_t n_msg, const uint32_t *mid, const float *mx,
                      const float *my, const float *mz, uint32_t n_agent, const uint32_t *aid, float *ax, float *ay,
                      float *az, float *adrift, float repulse);

/* One whole Circles step as CUDASimulation::step() runs it:
 * output_message -> auto sort of agents -> buildIndex -> move.  Arrays are permuted in place
 * into the post-sort order.  scratch-free convenience used by the CPU baseline. */
void orc_circles_step(const orc_grid *g, uint32_t n, uint32_t *id, float *x, float *y, float *z, float *drift,
                      float repulse, int do_sort, uint32_t *pbm_out /* may be NULL */);

/* Number of messages with id != own id and strictly inside the radius, per agent, visiting the
 * reference Filter's bins (examples/stress_model.cuh stress_update).  Each product and sum is
 * rounded separately (no FMA contraction), matching the __fmul_rn/__fadd_rn device code. */
void orc_neighbour_count(const orc_grid *g, const uint32_t *pbm, const uint32_t *mid, const float *mx, const float *my,
                         const float *mz, uint32_t n_agent, const uint32_t *aid, const float *ax, const float *ay,
                         const float *az, uint32_t *out);

/* integer hash used by the birth/death stress model (ours) */
uint32_t orc_hash32(uint32_t a, uint32_t b);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* FGB_ORACLE_H_ */
