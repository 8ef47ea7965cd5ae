// flamegpu/runtime/detail/StripWalk.cuh -- the message iterator behind MessageSpatial2D/3D::In::Filter.
//
// One class for both dimensionalities: a Moore neighbourhood is 3 (2D) or 9 (3D) x-strips of the PBM, each strip
// one contiguous range [PBM[hash(cx-1,..)], PBM[hash(cx+1,..)+1]) of the bin-sorted message list (reference
// MessageSpatial3DDevice.cuh:693-719, MessageSpatial2DDevice.cuh:579-600).  Two compile-time variants
// (agent_function_wrapper<..., ITER_MODE>):
//
//  mode 0  reference visit order: every message of every strip, strips in the reference's (dy,dz) order.
//
//  mode 1  radius-filtered (b200 extension): the lanes of a warp alternate between
//    walk : every lane tests chunks of <= 32 consecutive messages of its own strips against the radius with packed
//           fp32x2 arithmetic (two messages per instruction) and queues {first index, accepted mask} per non-empty
//           chunk in shared memory; the round ends when some lane's queue is full or every lane has walked all strips;
//    drain: the lanes return to the agent function once per queued message, `left` = max over the lanes of the number
//           of messages queued in the round (one REDUX per round, so the hot path is a decrement and a compare).
//           A lane whose queue ran dry is shown the list's PADDING message: the slot one past the capacity of
//           every spatial message list holds a location far outside any environment (DevList::pad_slot), so
//           padding needs no select on the loads and no branch in getVariable.
//    Contract: every message within the radius is presented exactly once, in the reference's relative order;
//    messages beyond the radius may be skipped and padding messages (beyond any radius) may appear.
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_DETAIL_STRIPWALK_CUH_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_DETAIL_STRIPWALK_CUH_

#include "flamegpu/runtime/detail/FunctionArgs.h"

namespace flamegpu {
namespace detail {

#if defined(__CUDACC__)
// Loads of one location array at a BYTE offset kept in an unsigned 32-bit register: the array bases are kernel
// parameters (uniform registers), so the load takes the form [R.U32 + UR] with no per-load address arithmetic.
// (The scheduler selects the radius-filtered mode only for lists below 2^30 items.)
__device__ __forceinline__ unsigned long long ldg_pair(const char *base, uint32_t byte_off) {
  return __ldg(reinterpret_cast<const unsigned long long *>(base + byte_off));
}
__device__ __forceinline__ float ldg_one(const char *base, uint32_t byte_off) {
  return __ldg(reinterpret_cast<const float *>(base + byte_off));
}

// bit k of the result: message i0+k (k < n <= 32, n > 0) lies within sqrt(r2) of the origin.  The location arrays
// are 8-byte aligned at even indices (one cudaMalloc per variable), so aligned pairs are fetched with one 64-bit load.
template <int DIMS>
__device__ __forceinline__ uint32_t radius_mask(const char *px, const char *py, const char *pz, uint32_t i0, uint32_t n,
                                                float ox, float oy, float oz, float r2) {
  auto one = [&](uint32_t o) {
    const float dx = ldg_one(px, o) - ox, dy = ldg_one(py, o) - oy;
    float d2 = dx * dx + dy * dy;
    if (DIMS == 3) {
      const float dz = ldg_one(pz, o) - oz;
      d2 += dz * dz;
    }
    return d2 <= r2;
  };
  uint32_t m = 0;  // filled from the top: after c messages they occupy bits [32-c, 32)
  uint32_t o = i0 * 4u;
  const uint32_t oe = (i0 + n) * 4u;
  if (o & 4u) {
    m = one(o) ? 0x80000000u : 0u;
    o += 4u;
  }
  const unsigned long long OX = f32x2(ox, ox), OY = f32x2(oy, oy), OZ = f32x2(oz, oz);
  // squared distances of the message pair at byte offset `off` of the row pointers
  auto pair = [&](const char *bx, const char *by, const char *bz, uint32_t off) {
    const unsigned long long X = __ldg(reinterpret_cast<const unsigned long long *>(bx + off));
    const unsigned long long Y = __ldg(reinterpret_cast<const unsigned long long *>(by + off));
    const unsigned long long dx = f32x2_sub(X, OX), dy = f32x2_sub(Y, OY);
    unsigned long long d = f32x2_fma(dy, dy, f32x2_mul(dx, dx));
    if (DIMS == 3) {
      const unsigned long long Z = __ldg(reinterpret_cast<const unsigned long long *>(bz + off));
      const unsigned long long dz = f32x2_sub(Z, OZ);
      d = f32x2_fma(dz, dz, d);
    }
    return d;
  };
  auto lo = [](unsigned long long d) { return __uint_as_float(static_cast<uint32_t>(d)); };
  auto hi = [](unsigned long long d) { return __uint_as_float(static_cast<uint32_t>(d >> 32)); };
  // four messages per iteration: one address per array, the second pair at an immediate offset
  for (; o + 16u <= oe; o += 16u) {
    const char *bx = px + o, *by = py + o, *bz = pz + o;
    const unsigned long long d0 = pair(bx, by, bz, 0u), d1 = pair(bx, by, bz, 8u);
    m = (m >> 4) | (lo(d0) <= r2 ? 0x10000000u : 0u) | (hi(d0) <= r2 ? 0x20000000u : 0u) | (lo(d1) <= r2 ? 0x40000000u : 0u) |
        (hi(d1) <= r2 ? 0x80000000u : 0u);
  }
  if (o + 8u <= oe) {
    const unsigned long long d0 = pair(px + o, py + o, pz + o, 0u);
    m = (m >> 2) | (lo(d0) <= r2 ? 0x40000000u : 0u) | (hi(d0) <= r2 ? 0x80000000u : 0u);
    o += 8u;
  }
  if (o < oe) m = (m >> 1) | (one(o) ? 0x80000000u : 0u);
  return m >> (32u - n);
}

template <int DIMS>
class SpatialFilterMessage {
  static constexpr int kStrips = DIMS == 3 ? 9 : 3;
  const FunctionArgs &a;
  const char *px, *py, *pz;  // location arrays of the input list, pinned in registers (make_loc)
  float ox, oy, oz;          // search origin (mode 1)
  int cx, cy, cz;
  int strip;                 // current strip; kStrips == all strips walked
  uint32_t idx, idx_end;     // message presented to the agent function; mode 0: one past the last message of the strip
  uint32_t nxt, nxt_end;     // prefetched bounds of strip + 1
  // mode 1
  uint32_t sidx, send;       // scan cursor, end of the strip being scanned
  uint32_t cbase, cmask;     // chunk being handed out: first message index - 1, accepted bits not yet presented
  uint32_t qbase;            // shared-memory address of this lane's queue column (entry k at qbase + k * 8 * blockDim.x)
  uint32_t qpos, qcount;     // shared-memory addresses of the next entry to pop / to push
  int left;                  // drain iterations left in this round (warp-uniform); < 0: end of the iteration
  uint32_t lanes;            // the lanes that walk together
  const int mode;            // compile-time constant after inlining

  // [PBM[hash(cx-1,y,z)], PBM[hash(cx+1,y,z)+1]) of strip s; empty if outside the grid (reference 3D :704, 2D :590)
  __device__ __forceinline__ void fetch(int s, uint32_t &b, uint32_t &e) const {
    b = 0;
    e = 0;
    if (s < kStrips) {
      const int gx = a.in_meta.grid_dim[0];
      int row;
      bool inside;
      if (DIMS == 3) {
        const int y = cy + (s / 3) - 1, z = cz + (s % 3) - 1;
        const int gy = a.in_meta.grid_dim[1], gz = a.in_meta.win_count;
        inside = y >= 0 && z >= 0 && y < gy && z < gz;
        row = (z * gy + y) * gx;
      } else {
        const int y = cy + s - 1;
        inside = y >= 0 && y < a.in_meta.win_count;
        row = y * gx;
      }
      if (inside) {
        const int x0 = cx > 0 ? cx - 1 : 0;  // getHash clamps x (reference 3D :660-672)
        const int x1 = cx + 1 < gx ? cx + 1 : gx - 1;
        b = __ldg(a.in_meta.pbm + row + x0);
        e = __ldg(a.in_meta.pbm + row + x1 + 1);
      }
    }
  }
  // mode 0: after the last strip this leaves an empty range, so `idx < idx_end` serves as the end test as well
  __device__ __forceinline__ void next_strip(uint32_t &cur, uint32_t &cur_end) {
    do {
      ++strip;
      cur = nxt;
      cur_end = nxt_end;
      fetch(strip + 1, nxt, nxt_end);
    } while (cur >= cur_end && strip < kStrips);
  }

  // the queue is addressed with 32-bit shared-memory addresses: one register, no generic-pointer arithmetic
  __device__ __forceinline__ static void lds_v2(uint32_t addr, uint32_t &x, uint32_t &y) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(addr));
  }
  __device__ __forceinline__ static void sts_v2(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y));
  }
  // one message off this lane's queue (padding if it ran dry); the caller has checked left > 0
  __device__ __forceinline__ void pop() {
    --left;
    const bool dry = cmask == 0u, more = qpos < qcount;
    if (dry && more) {
      lds_v2(qpos, cbase, cmask);
      qpos += blockDim.x * 8u;
    }
    if (dry && !more) cbase = a.in_meta.pad_index;  // __ffs(0) == 0: the lane is shown the padding message
    idx = cbase + static_cast<uint32_t>(__ffs(static_cast<int>(cmask)));
    cmask &= cmask - 1u;
  }
  // walk rounds until some lane has accepted messages (then present the first) or every lane has walked all strips
  __device__ __forceinline__ void refill() {
    const uint32_t qstride = blockDim.x * 8u, qend = qbase + kFilterChunks * qstride;
    for (;;) {
      if (__all_sync(lanes, strip >= kStrips)) {
        left = -1;
        return;
      }
      qpos = qbase;
      qcount = qbase;
      cmask = 0;
      uint32_t mine = 0;
      for (;;) {
        const bool walked = strip >= kStrips;
        if (__any_sync(lanes, !walked && qcount >= qend)) break;
        if (__all_sync(lanes, walked)) break;
        if (!walked) {
          const uint32_t n = send - sidx < 32u ? send - sidx : 32u;
          const uint32_t m = radius_mask<DIMS>(px, py, pz, sidx, n, ox, oy, oz, a.in_meta.radius2_eps);
          if (m) {
            sts_v2(qcount, sidx - 1u, m);
            qcount += qstride;
            mine += static_cast<uint32_t>(__popc(m));
          }
          sidx += n;
          if (sidx >= send) next_strip(sidx, send);
        }
      }
      left = static_cast<int>(__reduce_max_sync(lanes, mine));
      if (left > 0) {
        pop();
        return;
      }
    }
  }

 public:
  __device__ __forceinline__ SpatialFilterMessage(const FunctionArgs &args, float x, float y, float z, int _cx, int _cy, int _cz,
                                                  bool begin, int _mode)
      : a(args), ox(x), oy(y), oz(z), cx(_cx), cy(_cy), cz(_cz), strip(kStrips), idx(0), idx_end(0), nxt(0), nxt_end(0),
        sidx(0), send(0), cbase(0), cmask(0), qbase(0), qpos(0), qcount(0), left(-1), lanes(0), mode(_mode) {
    const LocPtrs loc = make_loc(args);
    px = loc.x;
    py = loc.y;
    pz = loc.z;
    if (begin) {
      strip = -1;
      fetch(0, nxt, nxt_end);
      if (mode != 0) {
        lanes = __activemask();
        qbase = static_cast<uint32_t>(__cvta_generic_to_shared(filter_queue())) + threadIdx.x * 8u;
        next_strip(sidx, send);
        left = 0;
        refill();
      } else {
        next_strip(idx, idx_end);
      }
    }
  }
  __device__ __forceinline__ bool operator!=(const SpatialFilterMessage &) const { return mode != 0 ? left >= 0 : idx < idx_end; }
  __device__ __forceinline__ bool operator==(const SpatialFilterMessage &rhs) const { return strip == rhs.strip && idx == rhs.idx; }
  __device__ __forceinline__ SpatialFilterMessage &operator++() {
    if (mode != 0) {
      if (left > 0) pop();
      else refill();
    } else if (++idx >= idx_end) {
      next_strip(idx, idx_end);
    }
    return *this;
  }
  template <typename T, unsigned int N>
  __device__ __forceinline__ T getVariable(const char (&name)[N]) const {
    const uint32_t h = name_hash(name);  // folds to a constant after inlining
    if (h == kHashX) return __ldg(reinterpret_cast<const T *>(px + static_cast<unsigned long long>(idx) * sizeof(T)));
    if (h == kHashY) return __ldg(reinterpret_cast<const T *>(py + static_cast<unsigned long long>(idx) * sizeof(T)));
    if (DIMS == 3 && h == kHashZ) return __ldg(reinterpret_cast<const T *>(pz + static_cast<unsigned long long>(idx) * sizeof(T)));
    const int s = find_slot(a.msg_in, h);
    if (s < 0) return T{};
    return __ldg(reinterpret_cast<const T *>(a.msg_in.ptr[s] + static_cast<unsigned long long>(idx) * sizeof(T)));
  }
  template <typename T, flamegpu::size_type N, unsigned int M>
  __device__ __forceinline__ T getVariable(const char (&name)[M], unsigned int index) const {
    const int s = find_slot(a.msg_in, name_hash(name));
    if (s < 0 || index >= N) return T{};
    return __ldg(reinterpret_cast<const T *>(a.msg_in.ptr[s]) + static_cast<size_t>(idx) * N + index);
  }
  __device__ __forceinline__ unsigned int getIndex() const { return idx; }
};
#endif  // __CUDACC__

}  // namespace detail
}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_RUNTIME_DETAIL_STRIPWALK_CUH_
