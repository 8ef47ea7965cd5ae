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
#include "flamegpu/runtime/detail/StripWalk.cuh"
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
      // strip walk in the reference's order or radius-filtered: flamegpu/runtime/detail/StripWalk.cuh
      typedef detail::SpatialFilterMessage<3> Message;
      class iterator {
        Message m;

       public:
        __device__ __forceinline__ iterator(const detail::FunctionArgs &args, float x, float y, float z, int cx, int cy, int cz,
                                            bool begin, int mode)
            : m(args, x, y, z, cx, cy, cz, begin, mode) {}
        __device__ __forceinline__ iterator &operator++() {
          ++m;
          return *this;
        }
        __device__ __forceinline__ bool operator!=(const iterator &rhs) const { return m != rhs.m; }
        __device__ __forceinline__ bool operator==(const iterator &rhs) const { return m == rhs.m; }
        __device__ __forceinline__ Message &operator*() { return m; }
        __device__ __forceinline__ Message *operator->() { return &m; }
      };
      __device__ __forceinline__ Filter(const detail::FunctionArgs &args, float x, float y, float z, int _mode)
          : a(args), lx(x), ly(y), lz(z), mode(_mode) {
        cx = detail::grid_cell(args.in_meta, 0, x);
        cy = detail::grid_cell(args.in_meta, 1, y);
        cz = detail::grid_cell(args.in_meta, 2, z) - args.in_meta.win_begin;  // plane index inside the slab window
      }
      __device__ __forceinline__ iterator begin() const { return iterator(a, lx, ly, lz, cx, cy, cz, true, mode); }
      __device__ __forceinline__ iterator end() const { return iterator(a, lx, ly, lz, cx, cy, cz, false, mode); }

     private:
      const detail::FunctionArgs &a;
      float lx, ly, lz;
      int cx, cy, cz;
      int mode;
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
        __device__ __forceinline__ bool operator!=(const Message &) const { return idx < idx_end; }  // empty range after the last cell
        __device__ __forceinline__ bool operator==(const Message &rhs) const { return cell == rhs.cell && idx == rhs.idx; }
        __device__ __forceinline__ Message &operator++() {
          if (++idx >= idx_end) next_cell();
          return *this;
        }
        template <typename T, unsigned int N>
        __device__ __forceinline__ T getVariable(const char (&name)[N]) const {
          const uint32_t h = detail::name_hash(name);  // folds to a constant after inlining
          if (h == detail::kHashX) return __ldg(reinterpret_cast<const T *>(loc.x) + static_cast<unsigned int>(idx));
          if (h == detail::kHashY) return __ldg(reinterpret_cast<const T *>(loc.y) + static_cast<unsigned int>(idx));
          if (h == detail::kHashZ) return __ldg(reinterpret_cast<const T *>(loc.z) + static_cast<unsigned int>(idx));
          const int s = detail::find_slot(a.msg_in, h);
          if (s < 0) return T{};
          return __ldg(reinterpret_cast<const T *>(a.msg_in.ptr[s]) + static_cast<unsigned int>(idx));
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

    // mode: 0 reference visit order, 1 radius-filtered; a compile-time constant of the kernel instance
    __device__ __forceinline__ explicit In(const detail::FunctionArgs &args, int _mode = 0) : a(args), mode(_mode) {}
    __device__ __forceinline__ Filter operator()(float x, float y, float z) const { return Filter(a, x, y, z, mode); }
    // The reference checks bounds / wrapCompatible only with FLAMEGPU_SEATBELTS (:527-551); this build
    // is the seatbelts-off configuration.
    __device__ __forceinline__ WrapFilter wrap(float x, float y, float z) const { return WrapFilter(a, x, y, z); }
    __device__ __forceinline__ float radius() const { return a.in_meta.radius; }

   private:
    const detail::FunctionArgs &a;
    const int mode;
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
