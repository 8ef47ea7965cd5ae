/*
 * fgb_oracle.c -- CPU restatement of FLAME GPU 2's spatial hot path (see fgb_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into, or called from, the product path.
 * Parity status: PINNED (see header).  Paths cited are relative to /root/reference.
 */
#include "fgb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* include/flamegpu/detail/numeric.h:25-32 approxExactlyDivisible<float> */
static int approx_exactly_divisible(float x, float y) {
  float ax = fabsf(x), ay = fabsf(y);
  float scaled_eps = (ax > ay ? ax : ay) * 1.1920928955078125e-07f; /* FLT_EPSILON */
  float v = fmodf(x, y);
  return v <= scaled_eps || v > y - scaled_eps;
}

/* src/flamegpu/runtime/messaging/MessageSpatial3D.cu:31-52 (2D: MessageSpatial2D.cu:35-52) */
void orc_grid_init(orc_grid *g, int dims, const float *mn, const float *mx, float radius) {
  memset(g, 0, sizeof(*g));
  g->dims = dims;
  g->radius = radius;
  g->wrap_compatible = 1;
  g->bin_count = 1;
  for (int a = 0; a < 3; ++a) g->grid_dim[a] = 1;
  for (int a = 0; a < dims; ++a) {
    g->min[a] = mn[a];
    g->max[a] = mx[a];
    g->env_width[a] = g->max[a] - g->min[a];
    /* static_cast<unsigned int>(ceil(environmentWidth / radius)): float divide, then ceil */
    g->grid_dim[a] = (uint32_t)ceil((double)(g->env_width[a] / g->radius));
    g->bin_count *= g->grid_dim[a];
    g->wrap_compatible = g->wrap_compatible && approx_exactly_divisible(g->env_width[a], g->radius);
  }
}

/* MessageSpatial3DDevice.cuh:646-659: floorf((p-min)/radius) then clamp to [0, dim-1] */
void orc_grid_pos(const orc_grid *g, float x, float y, float z, int cell[3]) {
  const float p[3] = {x, y, z};
  cell[2] = 0;
  for (int a = 0; a < g->dims; ++a) {
    int c = (int)floorf((p[a] - g->min[a]) / g->radius);
    int d = (int)g->grid_dim[a];
    cell[a] = c < 0 ? 0 : (c >= d ? d - 1 : c);
  }
}

/* MessageSpatial3DDevice.cuh:660-672: only x is re-clamped (the iterator passes cell.x +- 1) */
uint32_t orc_hash(const orc_grid *g, int cx, int cy, int cz) {
  int dx = (int)g->grid_dim[0];
  uint32_t hx = (uint32_t)(cx < 0 ? 0 : (cx >= dx - 1 ? dx - 1 : cx));
  if (g->dims == 2) return (uint32_t)cy * g->grid_dim[0] + hx;
  return (uint32_t)cz * g->grid_dim[0] * g->grid_dim[1] + (uint32_t)cy * g->grid_dim[0] + hx;
}

void orc_bin_keys(const orc_grid *g, uint32_t n, const float *x, const float *y, const float *z, uint32_t *keys) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    int c[3];
    orc_grid_pos(g, x[i], y[i], g->dims == 3 ? z[i] : 0.0f, c);
    keys[i] = orc_hash(g, c[0], c[1], c[2]);
  }
}

/* MessageSpatial3D.cu:113-146: histogram -> exclusive scan (PBM, binCount+1 entries) -> reorder
 * (CUDAScatter.cu:276-294: dst = PBM[bin] + sub_index).  Source order inside a bin. */
void orc_build_index(const orc_grid *g, uint32_t n, const float *x, const float *y, const float *z, uint32_t *pbm,
                     uint32_t *perm) {
  const uint32_t B = g->bin_count;
  uint32_t *keys = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
  uint32_t *cursor = (uint32_t *)calloc((size_t)B + 1, sizeof(uint32_t));
  orc_bin_keys(g, n, x, y, z, keys);
  for (uint32_t i = 0; i < n; ++i) cursor[keys[i]]++;
  uint32_t run = 0;
  for (uint32_t b = 0; b < B; ++b) {
    pbm[b] = run;
    run += cursor[b];
    cursor[b] = pbm[b];
  }
  pbm[B] = run; /* == n; an empty list gives all zeros (MessageSpatial3D.cu:116-120) */
  if (perm)
    for (uint32_t i = 0; i < n; ++i) perm[cursor[keys[i]]++] = i;
  free(keys);
  free(cursor);
}

void orc_bucket_build(int32_t lower, int32_t upper, uint32_t n, const int32_t *keys, uint32_t *pbm, uint32_t *perm) {
  const uint32_t B = (uint32_t)(upper - lower + 1); /* MessageBucket.cu:45 */
  uint32_t *cursor = (uint32_t *)calloc((size_t)B + 1, sizeof(uint32_t));
  for (uint32_t i = 0; i < n; ++i) cursor[(uint32_t)(keys[i] - lower)]++; /* :60-62 */
  uint32_t run = 0;
  for (uint32_t b = 0; b < B; ++b) {
    pbm[b] = run;
    run += cursor[b];
    cursor[b] = pbm[b];
  }
  pbm[B] = run;
  if (perm)
    for (uint32_t i = 0; i < n; ++i) perm[cursor[(uint32_t)(keys[i] - lower)]++] = i;
  free(cursor);
}

uint32_t orc_bucket_range(int32_t lower, int32_t upper, const uint32_t *pbm, int32_t begin_key, int32_t end_key,
                          uint32_t *first) {
  const int32_t mn = lower, mx = upper + 1; /* MetaData::max is exclusive, MessageBucket.cu:42-44 */
  uint32_t b = 0, e = 0;
  if (begin_key >= mn && end_key < mx && begin_key <= end_key) { /* MessageBucketDevice.cuh:269 */
    b = pbm[begin_key - mn];
    e = pbm[end_key - mn];
  }
  if (first) *first = b;
  return e - b;
}

void orc_gather(const uint32_t *perm, uint32_t n, uint32_t type_len, const void *in, void *out) {
  const char *src = (const char *)in;
  char *dst = (char *)out;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < (int64_t)n; ++j) memcpy(dst + (size_t)j * type_len, src + (size_t)perm[j] * type_len, type_len);
}

/* In::Filter::Message::operator++ (MessageSpatial3DDevice.cuh:693-719): strips (dy,dz) in the order
 * (-1,-1),(-1,0),(-1,1),(0,-1),...,(1,1); out-of-grid strips skipped; a strip is the contiguous
 * range [PBM[hash(cx-1)], PBM[hash(cx+1)+1]).  2D (MessageSpatial2DDevice.cuh:644-672): dy=-1,0,1. */
uint32_t orc_filter(const orc_grid *g, const uint32_t *pbm, float x, float y, float z, uint32_t *out_idx, uint32_t cap) {
  int c[3];
  uint32_t cnt = 0;
  orc_grid_pos(g, x, y, z, c);
  const int zlo = g->dims == 3 ? -1 : 0, zhi = g->dims == 3 ? 1 : 0;
  for (int dy = -1; dy <= 1; ++dy) {
    for (int dz = zlo; dz <= zhi; ++dz) {
      int ay = c[1] + dy, az = c[2] + dz;
      if (ay < 0 || ay >= (int)g->grid_dim[1]) continue;
      if (g->dims == 3 && (az < 0 || az >= (int)g->grid_dim[2])) continue;
      uint32_t s = pbm[orc_hash(g, c[0] - 1, ay, az)];
      uint32_t e = pbm[orc_hash(g, c[0] + 1, ay, az) + 1];
      for (uint32_t i = s; i < e; ++i) {
        if (cnt < cap) out_idx[cnt] = i;
        ++cnt;
      }
    }
  }
  return cnt;
}

/* In::WrapFilter::Message::operator++ (MessageSpatial3DDevice.cuh:727-749; nextCell :264-276):
 * single bins, x slowest, then y, z fastest, each wrapped (c+rel+dim)%dim.  2D: x slowest, y fastest. */
uint32_t orc_wrap_filter(const orc_grid *g, const uint32_t *pbm, float x, float y, float z, uint32_t *out_idx,
                         uint32_t cap) {
  int c[3];
  uint32_t cnt = 0;
  orc_grid_pos(g, x, y, z, c);
  const int zlo = g->dims == 3 ? -1 : 0, zhi = g->dims == 3 ? 1 : 0;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dz = zlo; dz <= zhi; ++dz) {
        int ax = (c[0] + dx + (int)g->grid_dim[0]) % (int)g->grid_dim[0];
        int ay = (c[1] + dy + (int)g->grid_dim[1]) % (int)g->grid_dim[1];
        int az = g->dims == 3 ? (c[2] + dz + (int)g->grid_dim[2]) % (int)g->grid_dim[2] : 0;
        uint32_t h = orc_hash(g, ax, ay, az);
        for (uint32_t i = pbm[h]; i < pbm[h + 1]; ++i) {
          if (cnt < cap) out_idx[cnt] = i;
          ++cnt;
        }
      }
  return cnt;
}

/* MessageSpatial3DDevice.cuh:373-379 */
float orc_virtual(float x2, float x1, float env_width) {
  const float x21 = x2 - x1;
  return fabsf(x21) > env_width / 2.0f ? x2 - (x21 / fabsf(x21) * env_width) : x2;
}

/* CUDAScatter.cu:67-88 driven by CUDAFatAgent.cu:105-137 (exclusive scan of the flags) */
uint32_t orc_compact(const uint32_t *flags, int invert, uint32_t n, uint32_t keep_front, uint32_t *perm) {
  uint32_t out = 0;
  for (uint32_t i = 0; i < n; ++i) {
    int keep = i < keep_front;
    if (!keep) {
      int f = flags[i - keep_front] == 1;
      keep = invert ? !f : f;
    }
    if (keep) perm[out++] = i;
  }
  return out;
}

/* CUDA's float -> int conversion (cvt.rzi.s32.f32) saturates and maps NaN to 0; C leaves it undefined */
static int cuda_f2i(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.0f) return 2147483647;
  if (v <= -2147483648.0f) return (-2147483647 - 1);
  return (int)v;
}

/* Geometry spatialSortAgent_async hands to the key kernel (CUDASimulation.cu:480-506).  REFERENCE QUIRK,
 * reproduced because agent order is a parity gate: a MessageSpatial3D::Data IS-A MessageSpatial2D::Data
 * (MessageSpatial3DHost.h:120), so the dynamic_cast at CUDASimulation.cu:487 always takes the 2D branch and
 * envMin.z = envMax.z = 0 => envWidth.z = 0, gridDim.z = 1, while the kernel still reads z (mode Agent3D).
 * The z term becomes floorf(((z-0)/0)*1) = +-inf/NaN -> saturated int, and only x,y order the agents. */
void orc_sort_geometry(const orc_grid *g, int true3d, float mn[3], float width[3], uint32_t gd[3]) {
  for (int a = 0; a < 3; ++a) {
    mn[a] = 0.0f;
    width[a] = 0.0f;
    gd[a] = 1;
  }
  const int dims_used = (g->dims == 3 && !true3d) ? 2 : g->dims;
  for (int a = 0; a < dims_used; ++a) {
    mn[a] = g->min[a];
    width[a] = g->env_width[a];
    gd[a] = width[a] ? (uint32_t)ceilf(width[a] / g->radius) : 1; /* CUDASimulation.cu:498-505 */
  }
}

/* calculateSpatialHash, CUDASimulation.cu:376-408: floorf(((p-min)/width)*gridDim), no clamp */
void orc_sort_keys(const orc_grid *g, int true3d, uint32_t n, const float *x, const float *y, const float *z,
                   uint32_t *keys) {
  float mn[3], w[3];
  uint32_t gd[3];
  orc_sort_geometry(g, true3d, mn, w, gd);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    int gx = cuda_f2i(floorf(((x[i] - mn[0]) / w[0]) * gd[0]));
    int gy = cuda_f2i(floorf(((y[i] - mn[1]) / w[1]) * gd[1]));
    if (g->dims == 3 && z) {
      int gz = cuda_f2i(floorf(((z[i] - mn[2]) / w[2]) * gd[2]));
      /* (gridPos[2] * gridDim.x * gridDim.y + gridPos[1] * gridDim.x + gridPos[0]) in unsigned arithmetic */
      keys[i] = (uint32_t)gz * gd[0] * gd[1] + (uint32_t)gy * gd[0] + (uint32_t)gx;
    } else {
      keys[i] = (uint32_t)gy * gd[0] + (uint32_t)gx;
    }
  }
}

/* max_bit = floor(log2(gridDim.x*gridDim.y*gridDim.z)) + 1, CUDASimulation.cu:571 */
int orc_sort_max_bit(const orc_grid *g, int true3d) {
  float mn[3], w[3];
  uint32_t gd[3];
  orc_sort_geometry(g, true3d, mn, w, gd);
  return (int)floor(log2((double)(gd[0] * gd[1] * gd[2]))) + 1;
}

/* LSD radix on bits [0,max_bit) == any stable sort on the masked key */
void orc_sort_perm(const uint32_t *keys, uint32_t n, int max_bit, uint32_t *perm) {
  uint32_t *a = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
  uint32_t *b = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
  for (uint32_t i = 0; i < n; ++i) a[i] = i;
  for (int shift = 0; shift < max_bit; shift += 8) {
    int bits = max_bit - shift < 8 ? max_bit - shift : 8;
    uint32_t mask = (1u << bits) - 1u;
    uint32_t cnt[257] = {0};
    for (uint32_t i = 0; i < n; ++i) cnt[((keys[a[i]] >> shift) & mask) + 1]++;
    for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
    for (uint32_t i = 0; i < n; ++i) b[cnt[(keys[a[i]] >> shift) & mask]++] = a[i];
    uint32_t *t = a;
    a = b;
    b = t;
  }
  memcpy(perm, a, sizeof(uint32_t) * n);
  free(a);
  free(b);
}

/* examples/cpp/circles_spatial3D/src/main.cu:13-54 */
void orc_circles_move(const orc_grid *g, const uint32_t *pbm, uint32_t n_msg, const uint32_t *mid, const float *mx,
                      const float *my, const float *mz, uint32_t n_agent, const uint32_t *aid, float *ax, float *ay,
                      float *az, float *adrift, float repulse) {
  (void)n_msg;
  const float RADIUS = g->radius;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t)n_agent; ++i) {
    const uint32_t ID = aid[i];
    const float x1 = ax[i], y1 = ay[i], z1 = az[i];
    float fx = 0.0f, fy = 0.0f, fz = 0.0f;
    int count = 0;
    int c[3];
    orc_grid_pos(g, x1, y1, z1, c);
    for (int dy = -1; dy <= 1; ++dy)
      for (int dz = -1; dz <= 1; ++dz) {
        int cy = c[1] + dy, cz = c[2] + dz;
        if (cy < 0 || cz < 0 || cy >= (int)g->grid_dim[1] || cz >= (int)g->grid_dim[2]) continue;
        uint32_t s = pbm[orc_hash(g, c[0] - 1, cy, cz)];
        uint32_t e = pbm[orc_hash(g, c[0] + 1, cy, cz) + 1];
        for (uint32_t m = s; m < e; ++m) {
          if (mid[m] != ID) {
            float x21 = mx[m] - x1, y21 = my[m] - y1, z21 = mz[m] - z1;
            const float separation = sqrtf(x21 * x21 + y21 * y21 + z21 * z21);
            if (separation < RADIUS && separation > 0.0f) {
              float k = sinf((separation / RADIUS) * 3.141f * -2) * repulse;
              x21 /= separation;
              y21 /= separation;
              z21 /= separation;
              fx += k * x21;
              fy += k * y21;
              fz += k * z21;
              count++;
            }
          }
        }
      }
    fx /= count > 0 ? count : 1;
    fy /= count > 0 ? count : 1;
    fz /= count > 0 ? count : 1;
    /* agents only read messages, never other agents, so in-place update is safe */
    ax[i] = x1 + fx;
    ay[i] = y1 + fy;
    az[i] = z1 + fz;
    adrift[i] = sqrtf(fx * fx + fy * fy + fz * fz);
  }
}

/* CUDASimulation::step() for the Circles model: layer 1 output_message (messages
 * in current agent order), layer 2: auto-sort agents (CUDASimulation.cu:671-681), buildIndex, move */
void orc_circles_step(const orc_grid *g, uint32_t n, uint32_t *id, float *x, float *y, float *z, float *drift,
                      float repulse, int do_sort, uint32_t *pbm_out) {
  const size_t nn = n ? n : 1;
  uint32_t *mid = (uint32_t *)malloc(4 * nn), *perm = (uint32_t *)malloc(4 * nn), *keys = (uint32_t *)malloc(4 * nn);
  float *mx = (float *)malloc(4 * nn), *my = (float *)malloc(4 * nn), *mz = (float *)malloc(4 * nn);
  uint32_t *tmp = (uint32_t *)malloc(4 * nn);
  uint32_t *pbm = (uint32_t *)malloc(4 * ((size_t)g->bin_count + 1));
  /* layer 1: output_message, unsorted message list = agent order */
  orc_build_index(g, n, x, y, z, pbm, perm);
  orc_gather(perm, n, 4, id, mid);
  orc_gather(perm, n, 4, x, mx);
  orc_gather(perm, n, 4, y, my);
  orc_gather(perm, n, 4, z, mz);
  /* layer 2: auto sort of the agents (positions unchanged by output_message) */
  if (do_sort) {
    orc_sort_keys(g, 0, n, x, y, z, keys);
    orc_sort_perm(keys, n, orc_sort_max_bit(g, 0), perm);
    void *vars[5] = {id, x, y, z, drift};
    for (int v = 0; v < 5; ++v) {
      orc_gather(perm, n, 4, vars[v], tmp);
      memcpy(vars[v], tmp, 4 * (size_t)n);
    }
  }
  orc_circles_move(g, pbm, n, mid, mx, my, mz, n, id, x, y, z, drift, repulse);
  if (pbm_out) memcpy(pbm_out, pbm, 4 * ((size_t)g->bin_count + 1));
  free(mid); free(perm); free(keys); free(mx); free(my); free(mz); free(tmp); free(pbm);
}

void orc_neighbour_count(const orc_grid *g, const uint32_t *pbm, const uint32_t *mid, const float *mx, const float *my,
                         const float *mz, uint32_t n_agent, const uint32_t *aid, const float *ax, const float *ay,
                         const float *az, uint32_t *out) {
  const float r2 = g->radius * g->radius;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t)n_agent; ++i) {
    int c[3];
    uint32_t cnt = 0;
    orc_grid_pos(g, ax[i], ay[i], az[i], c);
    for (int dy = -1; dy <= 1; ++dy)
      for (int dz = -1; dz <= 1; ++dz) {
        int cy = c[1] + dy, cz = c[2] + dz;
        if (cy < 0 || cz < 0 || cy >= (int)g->grid_dim[1] || cz >= (int)g->grid_dim[2]) continue;
        uint32_t s = pbm[orc_hash(g, c[0] - 1, cy, cz)];
        uint32_t e = pbm[orc_hash(g, c[0] + 1, cy, cz) + 1];
        for (uint32_t m = s; m < e; ++m) {
          if (mid[m] == aid[i]) continue;
          const float dx = mx[m] - ax[i], ddy = my[m] - ay[i], ddz = mz[m] - az[i];
          const float px = dx * dx, py = ddy * ddy, pz = ddz * ddz;
          const float s1 = px + py;
          const float s2 = s1 + pz;
          if (s2 < r2) ++cnt;
        }
      }
    out[i] = cnt;
  }
}

/* Stress model decisions (ours): a 32-bit mix of (id, step); independent of thread order */
uint32_t orc_hash32(uint32_t a, uint32_t b) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u + (a << 6) + (a >> 2));
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
