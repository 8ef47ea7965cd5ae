// flamegpu/runtime/messaging/MessageSpatial3D.cuh -- device side of 3D spatially partitioned
// messaging: In (Moore-neighbourhood iterators) and Out (setLocation), API compatible with the
// reference's include/flamegpu/runtime/messaging/MessageSpatial3D/MessageSpatial3DDevice.cuh.
//
// What is different underneath (B200-first, see DESIGN.md "neighbour iterator"):
//  * the grid metadata arrives by value in kernel-parameter space (no global re-reads of MetaData
//    per step, reference :704-709 reads gridDim/PBM through a pointer for every strip);
//  * the PBM bounds of the NEXT strip are fetched while the current strip is being walked, so the
//    dependent PBM -> message load chain of the reference (:708-709) is off the critical path;
//  * variable names resolve through a compile-time hash against a constant-bank table (hoisted
//    out of the loop) instead of a shared-memory hash probe per access (DeviceCurve.cuh:342-352);
//  * message loads use the read-only path (ld.global.nc); lanes of a warp that sit in the same
//    bin (agents are bin-sorted) read the same address and are served by one broadcast.
// Visit order is exactly the reference's: 9 x-strips in (dy,dz) order (-1,-1),(-1,0),...,(1,1)
// (nextStrip, reference :52-59), strips outside the grid skipped (:704), no radius filtering.
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGESPATIAL3D_CUH_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGESPATIAL3D_CUH_

#include "flamegpu/runtime/detail/FunctionArgs.h"
#include "flamegpu/runtime/messaging/MessageSpatial2D.cuh"

namespace flamegpu {

class MessageSpatial3D {
 public:
  class Description;  // host side, flamegpu/model/MessageDescriptions.h
  static constexpr int DIMS = 3;
  static constexpr bool SPATIAL = true;
  static constexpr bool HAS_OUTPUT = true;
  // MetaData keeps the reference's layout (MessageSpatial3D.h:38-68); it is what
  // fgb_spatial_metadata_device_ptr() points at.
  struct MetaData {
    float min[3];
    float max[3];
    float radius;
    unsigned int *PBM;
    unsigned int gridDim[3];
    float environmentWidth[3];
    bool wrapCompatible;
  };
  struct GridPos3D {
    int x, y, z;
  };

#if defined(__CUDACC__)
  class In {
   public:
    class Filter {
     public:
      class Message {
        const detail::FunctionArgs &a;
        const detail::LocPtrs loc;
        float ox, oy, oz;   // search origin (used by the radius modes only)
        int cx, cy, cz;
        int strip;          // 0..8 current strip, 9 == end
        int idx, idx_end;   // current message, one past the last message of the strip
        int nxt, nxt_end;   // prefetched bounds of strip+1
        int phase;          // radius-first mode: 0 = in-radius pass, 1 = the rest

        // [PBM[hash(cx-1,y,z)], PBM[hash(cx+1,y,z)+1]) of strip s; empty if outside the grid
        __device__ __forceinline__ void fetch(int s, int &b, int &e) const {
          b = 0;
          e = 0;
          if (s < 9) {
            const int y = cy + (s / 3) - 1, z = cz + (s % 3) - 1;
            const int gx = a.in_meta.grid_dim[0], gy = a.in_meta.grid_dim[1], gz = a.in_meta.win_count;
            if (y >= 0 && z >= 0 && y < gy && z < gz) {
              const int row = (z * gy + y) * gx;
              const int x0 = cx > 0 ? cx - 1 : 0;                 // getHash3D clamps x (reference :660-672)
              const int x1 = cx + 1 < gx ? cx + 1 : gx - 1;
              b = static_cast<int>(__ldg(a.in_meta.pbm + row + x0));
              e = static_cast<int>(__ldg(a.in_meta.pbm + row + x1 + 1));
            }
          }
        }
        __device__ __forceinline__ void next_strip() {
          do {
            ++strip;
            idx = nxt;
            idx_end = nxt_end;
            fetch(strip + 1, nxt, nxt_end);
          } while (idx >= idx_end && strip < 9);
        }
        __device__ __forceinline__ void restart() {
          strip = -1;
          fetch(0, nxt, nxt_end);
          next_strip();
        }
        // conservative superset of the user's usual `sqrtf(dx*dx+dy*dy+dz*dz) < radius`
        __device__ __forceinline__ bool in_radius() const {
          const float dx = __ldg(reinterpret_cast<const float *>(loc.x) + idx) - ox;
          const float dy = __ldg(reinterpret_cast<const float *>(loc.y) + idx) - oy;
          const float dz = __ldg(reinterpret_cast<const float *>(loc.z) + idx) - oz;
          return dx * dx + dy * dy + dz * dz <= a.in_meta.radius2_eps;
        }
        // radius modes: move on until the current message belongs to the current pass
        __device__ __forceinline__ void settle() {
          for (;;) {
            if (strip >= 9) {
              if (a.in_meta.iter_mode == 1 && phase == 0) {
                phase = 1;
                restart();
                continue;
              }
              return;
            }
            if (in_radius() == (phase == 0)) return;
            if (++idx >= idx_end) next_strip();
          }
        }

       public:
        __device__ __forceinline__ Message(const detail::FunctionArgs &args, float x, float y, float z, int _cx, int _cy, int _cz,
                                           bool begin)
            : a(args), loc(detail::make_loc(args)), ox(x), oy(y), oz(z), cx(_cx), cy(_cy), cz(_cz), strip(9), idx(0), idx_end(0),
              nxt(0), nxt_end(0), phase(0) {
          if (begin) {
            restart();
            if (a.in_meta.iter_mode != 0) settle();
          }
        }
        __device__ __forceinline__ bool operator!=(const Message &) const { return strip < 9; }
        __device__ __forceinline__ bool operator==(const Message &rhs) const {
          return strip == rhs.strip && idx == rhs.idx;
        }
        __device__ __forceinline__ Message &operator++() {
          if (++idx >= idx_end) next_strip();
          if (a.in_meta.iter_mode != 0) settle();
          return *this;
        }
        template <typename T, unsigned int N>
        __device__ __forceinline__ T getVariable(const char (&name)[N]) const {
          const uint32_t h = detail::name_hash(name);  // folds to a constant after inlining
          if (h == detail::kHashX) return __ldg(reinterpret_cast<const T *>(loc.x) + idx);
          if (h == detail::kHashY) return __ldg(reinterpret_cast<const T *>(loc.y) + idx);
          if (h == detail::kHashZ) return __ldg(reinterpret_cast<const T *>(loc.z) + idx);
          const int s = detail::find_slot(a.msg_in, h);
          if (s < 0) return T{};
          return __ldg(reinterpret_cast<const T *>(a.msg_in.ptr[s]) + idx);
        }
        template <typename T, flamegpu::size_type N, unsigned int M>
        __device__ __forceinline__ T getVariable(const char (&name)[M], unsigned int index) const {
          const int s = detail::find_slot(a.msg_in, detail::name_hash(name));
          if (s < 0 || index >= N) return T{};
          return __ldg(reinterpret_cast<const T *>(a.msg_in.ptr[s]) + static_cast<size_t>(idx) * N + index);
        }
        __device__ __forceinline__ unsigned int getIndex() const { return static_cast<unsigned int>(idx); }
      };
      class iterator {
        Message m;

       public:
        __device__ __forceinline__ iterator(const detail::FunctionArgs &args, float x, float y, float z, int cx, int cy, int cz,
                                            bool begin)
            : m(args, x, y, z, cx, cy, cz, begin) {}
        __device__ __forceinline__ iterator &operator++() {
          ++m;
          return *this;
        }
        __device__ __forceinline__ bool operator!=(const iterator &rhs) const { return m != rhs.m; }
        __device__ __forceinline__ bool operator==(const iterator &rhs) const { return m == rhs.m; }
        __device__ __forceinline__ Message &operator*() { return m; }
        __device__ __forceinline__ Message *operator->() { return &m; }
      };
      __device__ __forceinline__ Filter(const detail::FunctionArgs &args, float x, float y, float z)
          : a(args), lx(x), ly(y), lz(z) {
        cx = detail::grid_cell(args.in_meta, 0, x);
        cy = detail::grid_cell(args.in_meta, 1, y);
        cz = detail::grid_cell(args.in_meta, 2, z) - args.in_meta.win_begin;  // plane index inside the slab window
      }
      __device__ __forceinline__ iterator begin() const { return iterator(a, lx, ly, lz, cx, cy, cz, true); }
      __device__ __forceinline__ iterator end() const { return iterator(a, lx, ly, lz, cx, cy, cz, false); }

     private:
      const detail::FunctionArgs &a;
      float lx, ly, lz;
      int cx, cy, cz;
    };

    // 27 single bins, x slowest, z fastest, toroidal wrap (reference nextCell :264-276, :727-749)
    class WrapFilter {
     public:
      class Message {
        const detail::FunctionArgs &a;
        const detail::LocPtrs loc;
        float lx, ly, lz;
        int cx, cy, cz;
        int cell;  // 0..26, 27 == end
        int idx, idx_end, nxt, nxt_end;
        __device__ __forceinline__ void fetch(int c, int &b, int &e) const {
          b = 0;
          e = 0;
          if (c < 27) {
            const int gx = a.in_meta.grid_dim[0], gy = a.in_meta.grid_dim[1], gz = a.in_meta.grid_dim[2];
            const int x = (cx + (c / 9) - 1 + gx) % gx;
            const int y = (cy + ((c / 3) % 3) - 1 + gy) % gy;
            const int z = (cz + (c % 3) - 1 + gz) % gz;
            const int h = (z * gy + y) * gx + x;
            b = static_cast<int>(__ldg(a.in_meta.pbm + h));
            e = static_cast<int>(__ldg(a.in_meta.pbm + h + 1));
          }
        }
        __device__ __forceinline__ void next_cell() {
          do {
            ++cell;
            idx = nxt;
            idx_end = nxt_end;
            fetch(cell + 1, nxt, nxt_end);
          } while (idx >= idx_end && cell < 27);
        }
        __device__ __forceinline__ float virt(float p2, float p1, int axis) const {
          // reference getVirtualX :373-379
          const float d = p2 - p1;
          const float w = a.in_meta.env_width[axis];
          return fabsf(d) > w / 2.0f ? p2 - (d / fabsf(d) * w) : p2;
        }

       public:
        __device__ __forceinline__ Message(const detail::FunctionArgs &args, float x, float y, float z, int _cx, int _cy,
                                           int _cz, bool begin)
            : a(args), loc(detail::make_loc(args)), lx(x), ly(y), lz(z), cx(_cx), cy(_cy), cz(_cz), cell(27), idx(0), idx_end(0), nxt(0), nxt_end(0) {
          if (begin) {
            cell = -1;
            fetch(0, nxt, nxt_end);
            next_cell();
          }
        }
        __device__ __forceinline__ bool operator!=(const Message &) const { return cell < 27; }
        __device__ __forceinline__ bool operator==(const Message &rhs) const { return cell == rhs.cell && idx == rhs.idx; }
        __device__ __forceinline__ Message &operator++() {
          if (++idx >= idx_end) next_cell();
          return *this;
        }
        template <typename T, unsigned int N>
        __device__ __forceinline__ T getVariable(const char (&name)[N]) const {
          const uint32_t h = detail::name_hash(name);  // folds to a constant after inlining
          if (h == detail::kHashX) return __ldg(reinterpret_cast<const T *>(loc.x) + idx);
          if (h == detail::kHashY) return __ldg(reinterpret_cast<const T *>(loc.y) + idx);
          if (h == detail::kHashZ) return __ldg(reinterpret_cast<const T *>(loc.z) + idx);
          const int s = detail::find_slot(a.msg_in, h);
          if (s < 0) return T{};
          return __ldg(reinterpret_cast<const T *>(a.msg_in.ptr[s]) + idx);
        }
        template <typename T, flamegpu::size_type N, unsigned int M>
        __device__ __forceinline__ T getVariable(const char (&name)[M], unsigned int index) const {
          const int s = detail::find_slot(a.msg_in, detail::name_hash(name));
          if (s < 0 || index >= N) return T{};
          return __ldg(reinterpret_cast<const T *>(a.msg_in.ptr[s]) + static_cast<size_t>(idx) * N + index);
        }
        __device__ __forceinline__ float getVirtualX(float x1) const { return virt(getVariable<float>("x"), x1, 0); }
        __device__ __forceinline__ float getVirtualY(float y1) const { return virt(getVariable<float>("y"), y1, 1); }
        __device__ __forceinline__ float getVirtualZ(float z1) const { return virt(getVariable<float>("z"), z1, 2); }
        __device__ __forceinline__ float getVirtualX() const { return getVirtualX(lx); }
        __device__ __forceinline__ float getVirtualY() const { return getVirtualY(ly); }
        __device__ __forceinline__ float getVirtualZ() const { return getVirtualZ(lz); }
      };
      class iterator {
        Message m;

       public:
        __device__ __forceinline__ iterator(const detail::FunctionArgs &args, float x, float y, float z, int cx, int cy,
                                            int cz, bool begin)
            : m(args, x, y, z, cx, cy, cz, begin) {}
        __device__ __forceinline__ iterator &operator++() {
          ++m;
          return *this;
        }
        __device__ __forceinline__ bool operator!=(const iterator &rhs) const { return m != rhs.m; }
        __device__ __forceinline__ bool operator==(const iterator &rhs) const { return m == rhs.m; }
        __device__ __forceinline__ Message &operator*() { return m; }
        __device__ __forceinline__ Message *operator->() { return &m; }
      };
      __device__ __forceinline__ WrapFilter(const detail::FunctionArgs &args, float x, float y, float z)
          : a(args), lx(x), ly(y), lz(z) {
        cx = detail::grid_cell(args.in_meta, 0, x);
        cy = detail::grid_cell(args.in_meta, 1, y);
        cz = detail::grid_cell(args.in_meta, 2, z);
      }
      __device__ __forceinline__ iterator begin() const { return iterator(a, lx, ly, lz, cx, cy, cz, true); }
      __device__ __forceinline__ iterator end() const { return iterator(a, lx, ly, lz, cx, cy, cz, false); }

     private:
      const detail::FunctionArgs &a;
      float lx, ly, lz;
      int cx, cy, cz;
    };

    __device__ __forceinline__ explicit In(const detail::FunctionArgs &args) : a(args) {}
    __device__ __forceinline__ Filter operator()(float x, float y, float z) const { return Filter(a, x, y, z); }
    // The reference checks bounds / wrapCompatible only with FLAMEGPU_SEATBELTS (:527-551); this build
    // is the seatbelts-off configuration.
    __device__ __forceinline__ WrapFilter wrap(float x, float y, float z) const { return WrapFilter(a, x, y, z); }
    __device__ __forceinline__ float radius() const { return a.in_meta.radius; }

   private:
    const detail::FunctionArgs &a;
  };

  class Out : public MessageSpatial2D::Out {
   public:
    __device__ __forceinline__ Out(const detail::FunctionArgs &args, unsigned int index)
        : MessageSpatial2D::Out(args, index) {}
    __device__ __forceinline__ void setLocation(float x, float y, float z) const {
      this->template setVariable<float>("x", x);
      this->template setVariable<float>("y", y);
      this->template setVariable<float>("z", z);
    }
  };
#endif  // __CUDACC__
};

}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGESPATIAL3D_CUH_
