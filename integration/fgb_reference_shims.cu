// integration/fgb_reference_shims.cu -- the drop-in of INTEGRATION.md section A as a compiled translation unit.
//
// Compiled AGAINST THE REFERENCE'S HEADERS and linked into the reference library in place of four of its method
// bodies (oracle/ref_build/build_dropin.sh weakens the original symbols in the reference's objects, so these
// definitions win at link time and the vtables / call sites of every other reference TU bind to them):
//
//   MessageSpatial3D::CUDAModelHandler::buildIndex   src/flamegpu/runtime/messaging/MessageSpatial3D.cu:113-146
//   MessageSpatial2D::CUDAModelHandler::buildIndex   src/flamegpu/runtime/messaging/MessageSpatial2D.cu:113-146
//   MessageBucket::CUDAModelHandler::buildIndex      src/flamegpu/runtime/messaging/MessageBucket.cu:105-137
//   CUDAScatter::scatter (both overloads)            src/flamegpu/simulation/detail/CUDAScatter.cu:118-179
//
// Everything else of FLAME GPU 2 (model description, CUDASimulation::step, cuRVE, the device iterators, agent
// functions compiled against the reference) is untouched: this is the reference running with the sm_100a kernels
// of libflamegpu2_b200.so behind its own interfaces.  The class declarations are the reference's, so the handles a
// maintainer would add as members (fgb_ctx* in CUDAScatter, fgb_spatial* in the handlers) live in side tables here.
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "flamegpu/runtime/messaging/MessageBucket/MessageBucketHost.h"
#include "flamegpu/runtime/messaging/MessageSpatial2D/MessageSpatial2DHost.h"
#include "flamegpu/runtime/messaging/MessageSpatial3D/MessageSpatial3DHost.h"
#include "flamegpu/simulation/detail/CUDAErrorChecking.cuh"
#include "flamegpu/simulation/detail/CUDAMessage.h"
#include "flamegpu/simulation/detail/CUDAScatter.cuh"
#include "flamegpu2_b200.h"

namespace {

#define FGB_SHIM_CHECK(call)                                                                      \
  do {                                                                                            \
    const fgb_status fgb_shim_s = (call);                                                         \
    if (fgb_shim_s) { /* the reference's THROW is two statements */                               \
      THROW flamegpu::exception::CUDAError("%s: %s", #call, fgb_error_string(fgb_shim_s));         \
    }                                                                                             \
  } while (0)

std::mutex g_lock;
std::map<int, fgb_ctx *> g_ctx;                  // one context per device (CUDAScatter is one per simulation instance)
std::map<const void *, fgb_spatial *> g_spatial;  // handler -> fgb_spatial

fgb_ctx *ctx_of_current_device() {
  int dev = 0;
  gpuErrchk(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(g_lock);
  auto it = g_ctx.find(dev);
  if (it != g_ctx.end()) return it->second;
  fgb_ctx *c = nullptr;
  FGB_SHIM_CHECK(fgb_ctx_create(dev, &c));
  g_ctx[dev] = c;
  return c;
}

// {typeLen, in, out} of every message variable, in the reference's (alphabetical) variable order:
// what CUDAScatter::pbm_reorder builds at CUDAScatter.cu:316-323
std::vector<fgb_var> message_vars(flamegpu::detail::CUDAMessage &m) {
  std::vector<fgb_var> vars;
  for (const auto &v : m.getMessageData().variables) {
    fgb_var f;
    f.type_len = v.second.type_size * v.second.elements;
    f.in = static_cast<const char *>(m.getReadList().at(v.first));
    f.out = static_cast<char *>(m.getWriteList().at(v.first));
    vars.push_back(f);
  }
  return vars;
}

template <typename Make>
fgb_spatial *spatial_of(const void *handler, unsigned int *pbm, Make make) {
  {
    std::lock_guard<std::mutex> g(g_lock);
    auto it = g_spatial.find(handler);
    if (it != g_spatial.end()) return it->second;
  }
  fgb_spatial *sp = make();
  // build straight into the PBM the reference allocated: its device MetaData (read by the reference's own
  // MessageSpatial3D::In iterator) already points there
  FGB_SHIM_CHECK(fgb_spatial_use_pbm(sp, pbm));
  std::lock_guard<std::mutex> g(g_lock);
  g_spatial[handler] = sp;
  return sp;
}

}  // namespace

namespace flamegpu {

void MessageSpatial3D::CUDAModelHandler::buildIndex(detail::CUDAScatter &, unsigned int, cudaStream_t stream) {
  fgb_spatial *sp = spatial_of(this, hd_data.PBM, [&]() {
    fgb_spatial *s = nullptr;
    FGB_SHIM_CHECK(fgb_spatial_create(ctx_of_current_device(), 3, hd_data.min, hd_data.max, hd_data.radius, &s));
    return s;
  });
  const unsigned int n = sim_message.getMessageCount();
  std::vector<fgb_var> vars = message_vars(sim_message);
  FGB_SHIM_CHECK(fgb_build_index(sp, n, nullptr, static_cast<const float *>(sim_message.getReadPtr("x")),
                                 static_cast<const float *>(sim_message.getReadPtr("y")),
                                 static_cast<const float *>(sim_message.getReadPtr("z")), vars.data(),
                                 static_cast<unsigned int>(vars.size()), FGB_BUILD_DEFAULT, stream));
  if (n) sim_message.swap();  // reference :139; no stream synchronisation: the reader runs on the same stream
}

void MessageSpatial2D::CUDAModelHandler::buildIndex(detail::CUDAScatter &, unsigned int, cudaStream_t stream) {
  fgb_spatial *sp = spatial_of(this, hd_data.PBM, [&]() {
    fgb_spatial *s = nullptr;
    FGB_SHIM_CHECK(fgb_spatial_create(ctx_of_current_device(), 2, hd_data.min, hd_data.max, hd_data.radius, &s));
    return s;
  });
  const unsigned int n = sim_message.getMessageCount();
  std::vector<fgb_var> vars = message_vars(sim_message);
  FGB_SHIM_CHECK(fgb_build_index(sp, n, nullptr, static_cast<const float *>(sim_message.getReadPtr("x")),
                                 static_cast<const float *>(sim_message.getReadPtr("y")), nullptr, vars.data(),
                                 static_cast<unsigned int>(vars.size()), FGB_BUILD_DEFAULT, stream));
  if (n) sim_message.swap();
}

void MessageBucket::CUDAModelHandler::buildIndex(detail::CUDAScatter &, unsigned int, cudaStream_t stream) {
  fgb_spatial *sp = spatial_of(this, hd_data.PBM, [&]() {
    fgb_spatial *s = nullptr;
    FGB_SHIM_CHECK(fgb_bucket_create(ctx_of_current_device(), hd_data.min, hd_data.max - 1, &s));  // MetaData::max is exclusive
    return s;
  });
  const unsigned int n = sim_message.getMessageCount();
  std::vector<fgb_var> vars = message_vars(sim_message);
  FGB_SHIM_CHECK(fgb_build_index_keys(sp, n, nullptr, static_cast<const int *>(sim_message.getReadPtr("_key")), vars.data(),
                                      static_cast<unsigned int>(vars.size()), FGB_BUILD_DEFAULT, stream));
  if (n) sim_message.swap();
}

namespace detail {

// One pass over the scan flags (fgb_compact) instead of the scatter_generic kernel over a position array that the
// callers' cub::DeviceScan::ExclusiveSum produced (those calls stay where they are and simply go unused).
unsigned int CUDAScatter::scatter(const unsigned int streamResourceId, const cudaStream_t stream, const Type &messageOrAgent,
                                  const std::vector<ScatterData> &sd, const unsigned int itemCount,
                                  const unsigned int out_index_offset, const bool invert_scan_flag,
                                  const unsigned int scatter_all_count) {
  static_assert(sizeof(ScatterData) == sizeof(fgb_var), "ScatterData and fgb_var share the layout {size_t, char*, char*}");
  CUDAScanCompactionConfig &cfg = scan.Config(messageOrAgent, streamResourceId);
  // the kept count lands in the word the reference reads it from: position[itemCount - scatter_all_count]
  unsigned int *d_count = cfg.d_ptrs.position + itemCount - scatter_all_count;
  FGB_SHIM_CHECK(fgb_compact(ctx_of_current_device(), streamResourceId, cfg.d_ptrs.scan_flag, invert_scan_flag ? 1 : 0, itemCount,
                             nullptr, scatter_all_count, out_index_offset, nullptr, reinterpret_cast<const fgb_var *>(sd.data()),
                             static_cast<unsigned int>(sd.size()), d_count, nullptr, stream));
  unsigned int rtn = 0;  // the reference's callers expect the count on the host (CUDAScatter.cu:175-178)
  gpuErrchk(cudaMemcpyAsync(&rtn, d_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
  gpuErrchk(cudaStreamSynchronize(stream));
  return rtn;  // fgb_compact counts the scatter_all_count leading items as kept
}

unsigned int CUDAScatter::scatter(const unsigned int streamResourceId, const cudaStream_t stream, const Type &messageOrAgent,
                                  const VariableMap &vars, const std::map<std::string, void *> &in,
                                  const std::map<std::string, void *> &out, const unsigned int itemCount,
                                  const unsigned int out_index_offset, const bool invert_scan_flag,
                                  const unsigned int scatter_all_count) {
  std::vector<ScatterData> sd;
  for (const auto &v : vars)
    sd.push_back({v.second.type_size * v.second.elements, reinterpret_cast<char *>(in.at(v.first)), reinterpret_cast<char *>(out.at(v.first))});
  return scatter(streamResourceId, stream, messageOrAgent, sd, itemCount, out_index_offset, invert_scan_flag, scatter_all_count);
}

}  // namespace detail
}  // namespace flamegpu
