// flamegpu/runtime/messaging/MessageSpatial2D.cuh -- device side of 2D spatially partitioned
// messaging, API compatible with the reference's MessageSpatial2D/MessageSpatial2DDevice.cuh.
// Same design as the 3D twin (MessageSpatial3D.cuh): 3 x-strips in dy order -1,0,1 (reference
// :644-672), 9-bin toroidal WrapFilter with x slowest / y fastest (:253-260, :679-698).
#ifndef FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGESPATIAL2D_CUH_
#define FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGESPATIAL2D_CUH_

#include <type_traits>

#include "flamegpu/runtime/detail/FunctionArgs.h"
#include "flamegpu/runtime/detail/StripWalk.cuh"
#include "flamegpu/runtime/messaging/MessageBruteForce.cuh"

namespace flamegpu {
namespace detail {
#if defined(__CUDACC__)
// getGridPosition2D/3D (reference MessageSpatial3DDevice.cuh:646-659): IEEE divide, floorf, clamp
__device__ __forceinline__ int grid_cell(const SpatialMeta &m, int axis, float p) {
  const int c = static_cast<int>(floorf(__fdiv_rn(p - m.min[axis], m.radius)));
  const int d = m.grid_dim[axis];
  return c < 0 ? 0 : (c >= d ? d - 1 : c);
}
#endif
}  // namespace detail

class MessageSpatial2D {
 public:
  class Description;  // host side
  static constexpr int DIMS = 2;
  static constexpr bool SPATIAL = true;
  static constexpr bool HAS_OUTPUT = true;
  struct MetaData {  // reference MessageSpatial2D.h:36-66
    float min[2];
    float max[2];
    float radius;
    unsigned int *PBM;
    unsigned int gridDim[2];
    float environmentWidth[2];
    bool wrapCompatible;
  };
  struct GridPos2D {
    int x, y;
  };

#if defined(__CUDACC__)
  class In {
   public:
    class Filter {
     public:
      // strip walk in the reference's order or radius-filtered: flamegpu/runtime/detail/StripWalk.cuh
      typedef detail::SpatialFilterMessage<2> Message;
      class iterator {
        Message m;

       public:
        __device__ __forceinline__ iterator(const detail::FunctionArgs &args, float x, float y, int cx, int cy, bool begin, int mode)
            : m(args, x, y, 0.f, cx, cy, 0, begin, mode) {}
        __device__ __forceinline__ iterator &operator++() {
          ++m;
          return *this;
        }
        __device__ __forceinline__ bool operator!=(const iterator &rhs) const { return m != rhs.m; }
        __device__ __forceinline__ bool operator==(const iterator &rhs) const { return m == rhs.m; }
        __device__ __forceinline__ Message &operator*() { return m; }
        __device__ __forceinline__ Message *operator->() { return &m; }
      };
      __device__ __forceinline__ Filter(const detail::FunctionArgs &args, float x, float y, int _mode) : a(args), lx(x), ly(y), mode(_mode) {
        cx = detail::grid_cell(args.in_meta, 0, x);
        cy = detail::grid_cell(args.in_meta, 1, y) - args.in_meta.win_begin;  // row index inside the slab window
      }
      __device__ __forceinline__ iterator begin() const { return iterator(a, lx, ly, cx, cy, true, mode); }
      __device__ __forceinline__ iterator end() const { return iterator(a, lx, ly, cx, cy, false, mode); }

     private:
      const detail::FunctionArgs &a;
      float lx, ly;
      int cx, cy;
      int mode;
    };

    class WrapFilter {
     public:
      class Message {
        const detail::FunctionArgs &a;
        const detail::LocPtrs loc;
        float lx, ly;
        int cx, cy;
        int cell;  // 0..8, 9 == end
        int idx, idx_end, nxt, nxt_end;
        __device__ __forceinline__ void fetch(int c, int &b, int &e) const {
          b = 0;
          e = 0;
          if (c < 9) {
            const int gx = a.in_meta.grid_dim[0], gy = a.in_meta.grid_dim[1];
            const int x = (cx + (c / 3) - 1 + gx) % gx;
            const int y = (cy + (c % 3) - 1 + gy) % gy;
            const int h = y * gx + x;
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
          } while (idx >= idx_end && cell < 9);
        }
        __device__ __forceinline__ float virt(float p2, float p1, int axis) const {
          const float d = p2 - p1;
          const float w = a.in_meta.env_width[axis];
          return fabsf(d) > w / 2.0f ? p2 - (d / fabsf(d) * w) : p2;
        }

       public:
        __device__ __forceinline__ Message(const detail::FunctionArgs &args, float x, float y, int _cx, int _cy, bool begin)
            : a(args), loc(detail::make_loc(args)), lx(x), ly(y), cx(_cx), cy(_cy), cell(9), idx(0), idx_end(0), nxt(0), nxt_end(0) {
          if (begin) {
            cell = -1;
            fetch(0, nxt, nxt_end);
            next_cell();
          }
        }
        __device__ __forceinline__ bool operator!=(const Message &) const { return idx < idx_end; }
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
        __device__ __forceinline__ float getVirtualX() const { return getVirtualX(lx); }
        __device__ __forceinline__ float getVirtualY() const { return getVirtualY(ly); }
      };
      class iterator {
        Message m;

       public:
        __device__ __forceinline__ iterator(const detail::FunctionArgs &args, float x, float y, int cx, int cy, bool begin)
            : m(args, x, y, cx, cy, begin) {}
        __device__ __forceinline__ iterator &operator++() {
          ++m;
          return *this;
        }
        __device__ __forceinline__ bool operator!=(const iterator &rhs) const { return m != rhs.m; }
        __device__ __forceinline__ bool operator==(const iterator &rhs) const { return m == rhs.m; }
        __device__ __forceinline__ Message &operator*() { return m; }
        __device__ __forceinline__ Message *operator->() { return &m; }
      };
      __device__ __forceinline__ WrapFilter(const detail::FunctionArgs &args, float x, float y) : a(args), lx(x), ly(y) {
        cx = detail::grid_cell(args.in_meta, 0, x);
        cy = detail::grid_cell(args.in_meta, 1, y);
      }
      __device__ __forceinline__ iterator begin() const { return iterator(a, lx, ly, cx, cy, true); }
      __device__ __forceinline__ iterator end() const { return iterator(a, lx, ly, cx, cy, false); }

     private:
      const detail::FunctionArgs &a;
      float lx, ly;
      int cx, cy;
    };

    __device__ __forceinline__ explicit In(const detail::FunctionArgs &args, int _mode = 0) : a(args), mode(_mode) {}
    __device__ __forceinline__ Filter operator()(float x, float y) const { return Filter(a, x, y, mode); }
    __device__ __forceinline__ WrapFilter wrap(float x, float y) const { return WrapFilter(a, x, y); }
    __device__ __forceinline__ float radius() const { return a.in_meta.radius; }

   private:
    const detail::FunctionArgs &a;
    const int mode;
  };

  class Out : public MessageBruteForce::Out {
   public:
    __device__ __forceinline__ Out(const detail::FunctionArgs &args, unsigned int index)
        : MessageBruteForce::Out(args, index), lx(0.f), ly(0.f), lz(0.f) {}
    // the location is remembered in registers as it is written (through setLocation or setVariable("x"|"y"|"z")),
    // so that the kernel wrapper can publish the message's bin without reading it back (publish_index)
    template <typename T, unsigned int N>
    __device__ __forceinline__ void setVariable(const char (&name)[N], T value) const {
      MessageBruteForce::Out::setVariable<T>(name, value);
      if constexpr (std::is_same<T, float>::value) {
        const uint32_t h = detail::name_hash(name);  // folds to a constant
        if (h == detail::kHashX) lx = value;
        if (h == detail::kHashY) ly = value;
        if (h == detail::kHashZ) lz = value;
      }
    }
    template <typename T, flamegpu::size_type N, unsigned int M>
    __device__ __forceinline__ void setVariable(const char (&name)[M], unsigned int index, T value) const {
      MessageBruteForce::Out::setVariable<T, N>(name, index, value);
    }
    __device__ __forceinline__ void setLocation(float x, float y) const {
      this->template setVariable<float>("x", x);
      this->template setVariable<float>("y", y);
    }
    // bin of the written message: getGridPosition + getHash (reference MessageSpatial3DDevice.cuh:646-672), plane index
    // rebased to the slab window exactly as fgb::bin_key (csrc/fgb_binsort.cuh); one warp-aggregated RED per distinct
    // bin of the warp (the writer runs in bin order, so neighbouring lanes share bins)
    template <int DIMS>
    __device__ __forceinline__ void publish_index() const {
      if (!a.out_keys) return;
      auto cell = [&](int axis, float p) {
        const int c = static_cast<int>(floorf(__fdiv_rn(p - a.out_min[axis], a.out_radius)));
        const int d = a.out_grid_dim[axis];
        return c < 0 ? 0 : (c >= d ? d - 1 : c);
      };
      const int cx = cell(0, lx);
      int cy = cell(1, ly);
      unsigned int key;
      if (DIMS == 3) {
        int cz = cell(2, lz) - a.out_win_begin;
        cz = cz < 0 ? 0 : (cz >= a.out_win_count ? a.out_win_count - 1 : cz);
        key = (static_cast<unsigned int>(cz) * a.out_grid_dim[1] + cy) * a.out_grid_dim[0] + cx;
      } else {
        cy -= a.out_win_begin;
        cy = cy < 0 ? 0 : (cy >= a.out_win_count ? a.out_win_count - 1 : cy);
        key = static_cast<unsigned int>(cy) * a.out_grid_dim[0] + cx;
      }
      a.out_keys[slot] = key;
      const unsigned int active = __activemask();
      const unsigned int peers = __match_any_sync(active, key);
      if ((threadIdx.x & 31u) == static_cast<unsigned int>(__ffs(peers) - 1)) atomicAdd(a.out_hist + key, static_cast<unsigned int>(__popc(peers)));
    }

   protected:
    mutable float lx, ly, lz;
  };
#endif  // __CUDACC__
};

}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_RUNTIME_MESSAGING_MESSAGESPATIAL2D_CUH_
