// fgb_scatter_misc.cuh -- the remaining CUDAScatter kernels and the even-bin histogram behind
// HostAgentAPI::histogramEven (8f.3):
//   k_array_reorder      reorder_array_messages (CUDAScatter.cu:540-566): out[index[i]] = in[i] for every variable, one
//                        write counter per array element (the reference's "max bin size > 1" conflict check, :609-655)
//   k_array_conflicts    max over the write counters, folded by the last block -> one device word (no host round trip)
//   k_new_agents         scatter_new_agents (CUDAScatter.cu:348-366): host-created agents arrive as an array of structs
//                        and are transposed into the SoA state list; a tile of structs is staged in shared memory with
//                        coalesced 16-byte loads, then every variable is written with consecutive lanes on consecutive items
//   k_histogram_even     cub::DeviceHistogram::HistogramEven as called by HostAgentAPI.cuh:720-745
#pragma once
#include "fgb_common.cuh"

namespace fgb {

#ifdef __CUDACC__

// one thread per message; the payload moves with copy_item (4/8/16-byte fast paths)
__global__ void __launch_bounds__(256) k_array_reorder(const uint32_t *__restrict__ index, uint32_t array_length, uint32_t n_max,
                                                       const unsigned int *d_n, const __grid_constant__ VarTable vt,
                                                       uint32_t *write_count) {
  const uint32_t n = load_count(d_n, n_max);
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint32_t o = __ldg(index + i);
  if (o >= array_length) return;  // out of bounds: dropped (CUDAScatter.cu:553-554)
  for (uint32_t v = 0; v < vt.n; ++v) copy_item(vt, v, i, o);
  if (write_count) atomicAdd(write_count + o, 1u);
}

// *d_max = max(write_count[0..len)); write_count is re-zeroed for the next reorder; partial/done: block scratch
__global__ void __launch_bounds__(256) k_array_conflicts(uint32_t *write_count, uint32_t len, uint32_t *partial, uint32_t *done,
                                                         uint32_t *d_max) {
  __shared__ uint32_t s_max[8];
  __shared__ uint32_t s_last;
  uint32_t m = 0;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < len; i += gridDim.x * 256) {
    const uint32_t c = write_count[i];
    m = c > m ? c : m;
    if (c) write_count[i] = 0u;
  }
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = s_max[w] > m ? s_max[w] : m;
    partial[blockIdx.x] = m;
    __threadfence();
    s_last = atomicAdd(done, 1u) == gridDim.x - 1 ? 1u : 0u;
    __threadfence();
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    uint32_t all = 0;
    for (uint32_t b = 0; b < gridDim.x; ++b) {
      const uint32_t p = *reinterpret_cast<volatile uint32_t *>(partial + b);
      all = p > all ? p : all;
    }
    *d_max = all;
    *done = 0u;
  }
}

// AoS -> SoA.  vt.in[v] = address of variable v inside the FIRST struct, vt.out[v] = the SoA column, agent_size = bytes
// per struct.  Tile of kNaTile structs per block, staged in dynamic shared memory when it fits.
constexpr int kNaTile = 256;
__global__ void __launch_bounds__(256) k_new_agents(uint32_t n, uint32_t agent_size, uint32_t out_offset, const unsigned int *d_out_offset,
                                                    const char *__restrict__ aos_base, const __grid_constant__ VarTable vt, int staged) {
  extern __shared__ __align__(16) unsigned char s_tile[];
  if (d_out_offset) out_offset = __ldg(d_out_offset);
  const uint32_t a0 = blockIdx.x * kNaTile;
  if (a0 >= n) return;
  const uint32_t cnt = (n - a0) < static_cast<uint32_t>(kNaTile) ? n - a0 : kNaTile;
  const char *src = aos_base + static_cast<size_t>(a0) * agent_size;
  if (staged) {
    const size_t bytes = static_cast<size_t>(cnt) * agent_size;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      for (size_t b = threadIdx.x * 16; b + 16 <= bytes; b += 256 * 16) *reinterpret_cast<uint4 *>(s_tile + b) = ld_stream_u4(src + b);
      for (size_t b = (bytes & ~static_cast<size_t>(15)) + threadIdx.x; b < bytes; b += 256) s_tile[b] = src[b];
    } else {
      for (size_t b = threadIdx.x; b < bytes; b += 256) s_tile[b] = src[b];
    }
    __syncthreads();
  }
  for (uint32_t v = 0; v < vt.n; ++v) {
    const uint32_t len = vt.len[v];
    const size_t voff = static_cast<size_t>(vt.in[v] - aos_base);
    char *out = vt.out[v] + (static_cast<size_t>(out_offset) + a0) * len;
    if ((len & 3u) == 0 && (voff & 3u) == 0 && (agent_size & 3u) == 0) {
      const uint32_t words = len >> 2;
      for (uint32_t w = threadIdx.x; w < cnt * words; w += 256) {  // consecutive lanes write consecutive words of the column
        const uint32_t a = w / words, k = w - a * words;
        const size_t from = static_cast<size_t>(a) * agent_size + voff + 4u * k;
        reinterpret_cast<uint32_t *>(out)[w] = staged ? *reinterpret_cast<const uint32_t *>(s_tile + from) : *reinterpret_cast<const uint32_t *>(src + from);
      }
    } else {
      for (uint32_t b = threadIdx.x; b < cnt * len; b += 256) {
        const uint32_t a = b / len, k = b - a * len;
        const size_t from = static_cast<size_t>(a) * agent_size + voff + k;
        out[b] = staged ? static_cast<char>(s_tile[from]) : src[from];
      }
    }
  }
}

// counts[bin] += 1 for lower <= v < upper, bin as cub::DeviceHistogram::HistogramEven computes it: floating point
// (int)((v - lower) * (bins / (upper - lower))) in the sample's own type, integers ((v - lower) * bins) / (upper - lower).
// Block-private shared-memory histogram (bins <= kHistSmemBins), flushed with one RED per non-empty bin.
constexpr int kHistSmemBins = 4096;
template <typename T>
__global__ void __launch_bounds__(256) k_histogram_even(const T *__restrict__ in, uint32_t n_max, const unsigned int *d_n, uint32_t bins,
                                                        T lower, T upper, uint32_t *counts) {
  __shared__ uint32_t s_hist[kHistSmemBins];
  const bool smem = bins <= static_cast<uint32_t>(kHistSmemBins);
  if (smem) {
    for (uint32_t b = threadIdx.x; b < bins; b += 256) s_hist[b] = 0u;
    __syncthreads();
  }
  const uint32_t n = load_count(d_n, n_max);
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const T v = in[i];
    if (!(v >= lower) || !(v < upper)) continue;
    uint32_t bin;
    if constexpr (sizeof(T) == 4 && !(static_cast<T>(0.5) == static_cast<T>(0))) {  // float
      const float scale = static_cast<float>(bins) / (static_cast<float>(upper) - static_cast<float>(lower));
      bin = static_cast<uint32_t>(static_cast<int>((static_cast<float>(v) - static_cast<float>(lower)) * scale));
    } else if constexpr (!(static_cast<T>(0.5) == static_cast<T>(0))) {  // double
      const double scale = static_cast<double>(bins) / (static_cast<double>(upper) - static_cast<double>(lower));
      bin = static_cast<uint32_t>(static_cast<int>((static_cast<double>(v) - static_cast<double>(lower)) * scale));
    } else {
      const unsigned long long d = static_cast<unsigned long long>(static_cast<long long>(v) - static_cast<long long>(lower));
      const unsigned long long w = static_cast<unsigned long long>(static_cast<long long>(upper) - static_cast<long long>(lower));
      bin = static_cast<uint32_t>((d * bins) / w);
    }
    if (bin >= bins) bin = bins - 1;  // rounding at the upper edge
    if (smem) atomicAdd(s_hist + bin, 1u);
    else atomicAdd(counts + bin, 1u);
  }
  if (smem) {
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += 256) {
      const uint32_t c = s_hist[b];
      if (c) atomicAdd(counts + b, c);
    }
  }
}

#endif  // __CUDACC__
}  // namespace fgb
