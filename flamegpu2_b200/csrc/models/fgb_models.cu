// fgb_models.cu -- the example / benchmark models (examples/*.cuh), compiled against this repo's
// C++ API layer (include/flamegpu) and exposed through a small C ABI so that the Python harness
// (tests, bench.py, multi-GPU driver) can drive whole CUDASimulation::step() runs.
// The same model headers are compiled against the REFERENCE library by oracle/ref_build/ref_sim.cu.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <string>

#include "flamegpu/flamegpu.h"

#include "../../../examples/boids_model.cuh"
#include "../../../examples/circles_model.cuh"
#include "../../../examples/stress_model.cuh"
#include "../../../examples/test_models.cuh"

// user functors of the reference's own host-reduction tests (tests/test_cases/runtime/agent/host_reduction/
// test_transform_reduce.cu:8-13, test_reduce.cu) for fgbm_agent_custom_reduce
FLAMEGPU_CUSTOM_REDUCTION(fgbm_custom_sum, a, b) { return a + b; }
FLAMEGPU_CUSTOM_REDUCTION(fgbm_custom_max, a, b) { return a > b ? a : b; }
FLAMEGPU_CUSTOM_TRANSFORM(fgbm_custom_nonpositive, a) { return a <= 0 ? 1 : 0; }

namespace {

thread_local std::string g_last_error;

struct Sim {
  std::unique_ptr<flamegpu::ModelDescription> model;
  std::unique_ptr<flamegpu::CUDASimulation> sim;
  std::map<std::string, std::unique_ptr<flamegpu::AgentVector>> staging;  // per agent, reused by the host-buffer path
};

std::map<std::string, std::string> parse_kv(const char *s) {
  std::map<std::string, std::string> kv;
  if (!s) return kv;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, ',')) {
    const size_t eq = item.find('=');
    if (eq == std::string::npos) continue;
    kv[item.substr(0, eq)] = item.substr(eq + 1);
  }
  return kv;
}
float getf(const std::map<std::string, std::string> &kv, const char *k, float d) {
  auto it = kv.find(k);
  return it == kv.end() ? d : std::strtof(it->second.c_str(), nullptr);
}
unsigned int getu(const std::map<std::string, std::string> &kv, const char *k, unsigned int d) {
  auto it = kv.find(k);
  return it == kv.end() ? d : static_cast<unsigned int>(std::strtoul(it->second.c_str(), nullptr, 10));
}

template <typename F>
int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    return -1;
  }
}

}  // namespace

extern "C" {

const char *fgbm_last_error(void) { return g_last_error.c_str(); }

// model: "circles" | "boids3d" | "boids2d" | "stress" | "test"; params: "key=value,key=value"
int fgbm_create(const char *model_name, const char *params, int device, void **out) {
  return guarded([&] {
    const std::string name = model_name ? model_name : "";
    const auto kv = parse_kv(params);
    auto s = std::make_unique<Sim>();
    s->model = std::make_unique<flamegpu::ModelDescription>(name);
    if (name == "circles") {
      fgb_examples::CirclesParams p;
      p.env_max = getf(kv, "env_max", p.env_max);
      p.radius = getf(kv, "radius", p.radius);
      p.repulse = getf(kv, "repulse", p.repulse);
      p.sort_period = getu(kv, "sort_period", p.sort_period);
      p.env_max_z = getf(kv, "env_max_z", p.env_max_z);
      p.validation = getu(kv, "validation", p.validation);
      p.radius_filtered = getu(kv, "radius_filtered", p.radius_filtered);
      fgb_examples::define_circles(*s->model, p);
    } else if (name == "boids3d" || name == "boids2d") {
      fgb_examples::BoidsParams p;
      p.dims = name == "boids3d" ? 3 : 2;
      p.min_position = getf(kv, "min_position", p.min_position);
      p.max_position = getf(kv, "max_position", p.max_position);
      p.interaction_radius = getf(kv, "interaction_radius", p.interaction_radius);
      p.separation_radius = getf(kv, "separation_radius", p.separation_radius);
      fgb_examples::define_boids(*s->model, p);
    } else if (name == "stress") {
      fgb_examples::StressParams p;
      p.env_max = getf(kv, "env_max", p.env_max);
      p.radius = getf(kv, "radius", p.radius);
      p.death_mod = getu(kv, "death_mod", p.death_mod);
      p.birth_mod = getu(kv, "birth_mod", p.birth_mod);
      fgb_examples::define_stress(*s->model, p);
    } else if (name == "test") {
      fgb_examples::TestParams p;
      p.which = static_cast<int>(getu(kv, "which", 0));
      p.mn[0] = getf(kv, "min_x", 0.f); p.mn[1] = getf(kv, "min_y", 0.f); p.mn[2] = getf(kv, "min_z", 0.f);
      p.mx[0] = getf(kv, "max_x", 5.f); p.mx[1] = getf(kv, "max_y", 5.f); p.mx[2] = getf(kv, "max_z", 5.f);
      p.radius = getf(kv, "radius", 1.f);
      p.sort_period = getu(kv, "sort_period", 1);
      p.bucket_upper = static_cast<int>(getu(kv, "bucket_upper", 12 + 512));
      p.birth_optional = static_cast<int>(getu(kv, "birth_optional", 0));
      p.birth_death = static_cast<int>(getu(kv, "birth_death", 0));
      p.birth_condition = static_cast<int>(getu(kv, "birth_condition", 0));
      p.birth_target = static_cast<int>(getu(kv, "birth_target", 0));
      p.append_optional = static_cast<int>(getu(kv, "append_optional", 0));
      p.host_init = static_cast<int>(getu(kv, "host_init", 0));
      fgb_examples::define_test_model(*s->model, p);
    } else {
      throw std::runtime_error("unknown model '" + name + "'");
    }
    s->sim = std::make_unique<flamegpu::CUDASimulation>(*s->model);
    s->sim->CUDAConfig().device_id = device;
    s->sim->CUDAConfig().inLayerConcurrency = getu(kv, "concurrency", 1) != 0;
    if (getu(kv, "slab_world", 1) > 1)
      s->sim->configureSlabs(static_cast<int>(getu(kv, "slab_rank", 0)), static_cast<int>(getu(kv, "slab_world", 1)),
                             kv.count("slab_message") ? kv.at("slab_message") : std::string("location"), getu(kv, "halo_cap", 65536),
                             getu(kv, "mig_cap", 16384));
    s->sim->CUDAConfig().slabRefreshPeriod = getu(kv, "slab_refresh", 16);
    s->sim->CUDAConfig().fusedIndexBuild = getu(kv, "fused_index", 1) != 0;
    s->sim->CUDAConfig().binOrderedOutput = getu(kv, "ordered_output", 1) != 0;
    if (kv.count("win_count")) s->sim->setMessageWindow("location", static_cast<int>(getu(kv, "win_begin", 0)), static_cast<int>(getu(kv, "win_count", 0)));
    s->sim->CUDAConfig().useCUDAGraphs = getu(kv, "graphs", 1) != 0;
    s->sim->CUDAConfig().stableMessageOrder = getu(kv, "stable", 0) != 0;
    s->sim->CUDAConfig().trueSpatialSortKey = getu(kv, "true3d_sort", 0) != 0;
    s->sim->SimulationConfig().timing = getu(kv, "timing", 0) != 0;
    s->sim->CUDAConfig().profile = getu(kv, "profile", 0) != 0;
    s->sim->CUDAConfig().binOrderExecution = getu(kv, "bin_order", 1) != 0;
    // -1 (default): per function, as declared by the model; 0: reference visit order everywhere; 1: radius-filtered everywhere
    s->sim->CUDAConfig().spatialIterationMode = kv.count("iter_mode") ? std::atoi(kv.at("iter_mode").c_str()) : -1;
    s->sim->CUDAConfig().overlapIndexBuild = getu(kv, "overlap", 1) != 0;
    s->sim->CUDAConfig().agentFunctionBlockSize = static_cast<int>(getu(kv, "block", 128));
    s->sim->CUDAConfig().tileLocalExecOrder = getu(kv, "tile_order", 1) != 0;
    *out = s.release();
  });
}

int fgbm_destroy(void *h) {
  return guarded([&] { delete static_cast<Sim *>(h); });
}

// Upload a population: n agents, nvars SoA host arrays named names[v] (any variable not listed keeps
// its default).  Replaces the state list.
int fgbm_set_population(void *h, const char *agent, const char *state, unsigned int n, unsigned int nvars,
                        const char **names, const void **host_ptrs) {
  return guarded([&] {
    Sim *s = static_cast<Sim *>(h);
    auto &stg = s->staging[agent];
    if (!stg) stg = std::make_unique<flamegpu::AgentVector>(s->model->Agent(agent), 0);
    stg->resize(0);
    stg->resize(n);
    for (unsigned int v = 0; v < nvars; ++v) {
      if (std::strcmp(names[v], "_n") == 0) continue;  // carries only the population size
      std::vector<char> &col = stg->raw(names[v]);
      std::memcpy(col.data(), host_ptrs[v], col.size());
    }
    s->sim->setPopulationData(*stg, state ? state : flamegpu::DEFAULT_STATE);
  });
}

int fgbm_get_count(void *h, const char *agent, const char *state, unsigned int *n) {
  return guarded([&] { *n = static_cast<Sim *>(h)->sim->getAgentCount(agent, state ? state : flamegpu::DEFAULT_STATE); });
}

// Host-side description validation, restating the reference's DescriptionValidation / DataValidation / reserved_name
// tests (test_bucket.cu:23-52, test_spatial_3d.cu DescriptionValidation).  Pure host code: returns 0 when every
// expectation holds, otherwise the 1-based index of the first failed check.  No GPU needed.
int fgbm_selftest_descriptions(void) {
  using namespace flamegpu;
  int check = 0;
#define EXPECT_THROWS(expr, exc)   \
  ++check;                         \
  try {                            \
    expr;                          \
    return check;                  \
  } catch (const exc &) {          \
  } catch (...) {                  \
    return check;                  \
  }
#define EXPECT_OK(expr) \
  ++check;              \
  try {                 \
    expr;               \
  } catch (...) {       \
    return check;       \
  }
  {
    ModelDescription model("BucketMessageTest");
    MessageBucket::Description message = model.newMessage<MessageBucket>("buckets");
    EXPECT_THROWS(message.setUpperBound(0), exception::InvalidArgument);  // min defaults to 0: no buckets
    EXPECT_OK(message.setLowerBound(10));
    EXPECT_OK(message.setUpperBound(11));
    EXPECT_THROWS(message.setUpperBound(0), exception::InvalidArgument);  // max < min
    EXPECT_OK(message.setUpperBound(12));
    EXPECT_THROWS(message.setLowerBound(13), exception::InvalidArgument);  // min > max
    EXPECT_THROWS(message.setBounds(12, 12), exception::InvalidArgument);
    EXPECT_THROWS(message.setBounds(13, 12), exception::InvalidArgument);
    EXPECT_OK(message.setBounds(12, 13));
    EXPECT_OK(message.newVariable<int>("somevar"));
    EXPECT_THROWS(message.newVariable<int>("_"), exception::ReservedName);
    EXPECT_THROWS(message.newVariable<int>("somevar"), exception::InvalidMessageVar);
  }
  {
    ModelDescription model("Spatial3DMessageTest");
    MessageSpatial3D::Description message = model.newMessage<MessageSpatial3D>("location");
    EXPECT_THROWS(message.setRadius(0), exception::InvalidArgument);
    EXPECT_THROWS(message.setRadius(-10), exception::InvalidArgument);
    EXPECT_OK(message.setRadius(1));
    EXPECT_OK(message.setMinX(0));
    EXPECT_THROWS(message.setMaxX(-1), exception::InvalidArgument);  // max <= min
    EXPECT_OK(message.setMax(5, 5, 5));
    EXPECT_THROWS(message.setMinZ(5), exception::InvalidArgument);   // min >= max
    EXPECT_THROWS(model.newMessage<MessageSpatial3D>("location"), exception::InvalidMessageName);
    AgentDescription agent = model.newAgent("agent");
    EXPECT_THROWS(model.newAgent("agent"), exception::InvalidAgentName);
    EXPECT_OK(agent.newVariable<float>("x"));
  }
#undef EXPECT_THROWS
#undef EXPECT_OK
  return 0;
}

// HostAgentAPI reductions (what a step function calls): op 0 sum, 1 min, 2 max, 3 count(value), 4 mean, 5 std;
// kind 'f' float, 'i' int, 'u' unsigned
int fgbm_agent_reduce(void *h, const char *agent, const char *var, int op, char kind, double *out) {
  return guarded([&] {
    flamegpu::HostAgentAPI api = static_cast<Sim *>(h)->sim->hostAPI().agent(agent);
    // op 3: count(variable, value = *out on entry); 4: mean; 5: population standard deviation
    const double arg = *out;
    auto run = [&](auto tag) {
      using T = decltype(tag);
      if (op == 3) return static_cast<double>(api.count<T>(var, static_cast<T>(arg)));
      if (op == 4) return api.meanStandardDeviation<T>(var).first;
      if (op == 5) return api.meanStandardDeviation<T>(var).second;
      return static_cast<double>(op == 0 ? api.sum<T>(var) : (op == 1 ? api.min<T>(var) : api.max<T>(var)));
    };
    *out = kind == 'f' ? run(float{}) : (kind == 'i' ? run(int{}) : run(static_cast<unsigned int>(0)));
  });
}

// what the Circles example's Validation step function saw last (a sum<float>("drift") per step, examples/circles_model.cuh)
int fgbm_circles_validation(double *last_total_drift, unsigned int *dropped, unsigned int *increased) {
  const fgb_examples::CirclesValidationState &v = fgb_examples::circles_validation_state();
  if (last_total_drift) *last_total_drift = static_cast<double>(v.prev_total_drift);
  if (dropped) *dropped = v.dropped;
  if (increased) *increased = v.increased;
  return 0;
}

// HostAgentAPI::histogramEven (kind 'f' float, 'i' int, 'u' unsigned) into out[bins]
int fgbm_agent_histogram(void *h, const char *agent, const char *var, char kind, unsigned int bins, double lower, double upper, unsigned int *out) {
  return guarded([&] {
    flamegpu::HostAgentAPI api = static_cast<Sim *>(h)->sim->hostAPI().agent(agent);
    std::vector<unsigned int> r = kind == 'f' ? api.histogramEven<float>(var, bins, static_cast<float>(lower), static_cast<float>(upper))
                                  : (kind == 'i' ? api.histogramEven<int>(var, bins, static_cast<int>(lower), static_cast<int>(upper))
                                                 : api.histogramEven<unsigned int>(var, bins, static_cast<unsigned int>(lower), static_cast<unsigned int>(upper)));
    for (unsigned int b = 0; b < bins; ++b) out[b] = r[b];
  });
}
// HostAgentAPI::reduce / transformReduce with user functors: which 0 = reduce(customSum, 0), 1 = reduce(customMax, lowest),
// 2 = transformReduce(customTransform "a <= 0 ? 1 : 0", customSum, 0) as in the reference's CustomTransformReduce tests
int fgbm_agent_custom_reduce(void *h, const char *agent, const char *var, int which, char kind, double *out) {
  return guarded([&] {
    flamegpu::HostAgentAPI api = static_cast<Sim *>(h)->sim->hostAPI().agent(agent);
    auto run = [&](auto tag) -> double {
      using T = decltype(tag);
      if (which == 0) return static_cast<double>(api.reduce<T>(var, fgbm_custom_sum, static_cast<T>(0)));
      if (which == 1) return static_cast<double>(api.reduce<T>(var, fgbm_custom_max, std::numeric_limits<T>::lowest()));
      return static_cast<double>(api.transformReduce<T, int>(var, fgbm_custom_nonpositive, fgbm_custom_sum, 0));
    };
    *out = kind == 'f' ? run(float{}) : (kind == 'i' ? run(int{}) : run(static_cast<unsigned int>(0)));
  });
}

// Copy one variable of a state list to host memory (bytes = count * type_len, checked).
int fgbm_get_variable(void *h, const char *agent, const char *state, const char *var, void *host_out, size_t bytes) {
  return guarded([&] {
    Sim *s = static_cast<Sim *>(h);
    const char *st = state ? state : flamegpu::DEFAULT_STATE;
    const unsigned int n = s->sim->getAgentCount(agent, st);
    const auto &vars = s->model->Agent(agent).agent->variables;
    auto it = vars.find(var);
    if (it == vars.end()) throw std::runtime_error(std::string("no variable ") + var);
    if (bytes != static_cast<size_t>(n) * it->second.bytes()) throw std::runtime_error("size mismatch in fgbm_get_variable");
    void *d = s->sim->getAgentVariableDevicePtr(agent, st, var);
    if (n && cudaMemcpy(host_out, d, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) throw std::runtime_error("cudaMemcpy failed");
  });
}

int fgbm_step(void *h, unsigned int steps) {
  return guarded([&] {
    Sim *s = static_cast<Sim *>(h);
    for (unsigned int i = 0; i < steps; ++i) s->sim->step();
  });
}

// CUDASimulation::simulate(): init functions, `steps` steps, exit functions
int fgbm_simulate(void *h, unsigned int steps) {
  return guarded([&] {
    Sim *s = static_cast<Sim *>(h);
    s->sim->SimulationConfig().steps = steps;
    s->sim->simulate();
  });
}

int fgbm_sync(void *h) {
  return guarded([&] { static_cast<Sim *>(h)->sim->synchronize(); });
}

void *fgbm_stream(void *h) { return static_cast<Sim *>(h)->sim->getStream(); }
unsigned long long fgbm_launch_count(void *h) { return static_cast<Sim *>(h)->sim->getLaunchCount(); }
unsigned int fgbm_graph_count(void *h) { return static_cast<Sim *>(h)->sim->getGraphCount(); }
// widest level (kernel nodes at equal depth) of the most recently captured step graph: > 1 means parallel branches
unsigned int fgbm_graph_width(void *h) { return static_cast<Sim *>(h)->sim->getGraphWidth(); }
// host-side launch bound and capacity of a state list (tests: both must stay bounded under conditional transitions)
int fgbm_list_bound(void *h, const char *agent, const char *state, unsigned int *bound, unsigned int *capacity) {
  return guarded([&] { static_cast<Sim *>(h)->sim->getListBound(agent, state ? state : flamegpu::DEFAULT_STATE, bound, capacity); });
}
unsigned int fgbm_step_counter(void *h) { return static_cast<Sim *>(h)->sim->getStepCounter(); }

// Per-step device times (seconds) recorded when the model was created with timing=1.
int fgbm_step_times(void *h, double *out, unsigned int cap, unsigned int *n) {
  return guarded([&] {
    std::vector<double> t = static_cast<Sim *>(h)->sim->getElapsedTimeSteps();
    *n = static_cast<unsigned int>(t.size());
    for (unsigned int i = 0; i < cap && i < t.size(); ++i) out[i] = t[i];
  });
}

// ---- multi-GPU slab decomposition behind step() (CUDASimulation::configureSlabs) -----------------------
// The harness (flamegpu2_b200/slab.py) only moves the 64-byte staging handles between the ranks.
unsigned int fgbm_slab_handle_bytes(void *h) { return static_cast<unsigned int>(static_cast<Sim *>(h)->sim->slabHandleBytes()); }
int fgbm_slab_export(void *h, void *out) {
  return guarded([&] { static_cast<Sim *>(h)->sim->slabExportHandle(out); });
}
int fgbm_slab_connect(void *h, const void *all_handles) {
  return guarded([&] { static_cast<Sim *>(h)->sim->slabConnect(all_handles); });
}
int fgbm_slab_planes(void *h, int *z0, int *z1) {
  return guarded([&] { static_cast<Sim *>(h)->sim->slabPlanes(z0, z1); });
}
int fgbm_slab_error(void *h, unsigned int *bits) {
  return guarded([&] { *bits = static_cast<Sim *>(h)->sim->slabError(); });
}

// ---- phase-wise multi-GPU driver hooks (round-1 path, kept for A/B) ------------------------------------
int fgbm_run_layers(void *h, unsigned int first, unsigned int last) {
  return guarded([&] { static_cast<Sim *>(h)->sim->runLayers(first, last); });
}
int fgbm_end_step(void *h) {
  return guarded([&] { static_cast<Sim *>(h)->sim->endStep(); });
}
int fgbm_end_step_pipelined(void *h) {
  return guarded([&] { static_cast<Sim *>(h)->sim->endStepPipelined(); });
}
int fgbm_begin_exchange(void *h) {
  return guarded([&] { static_cast<Sim *>(h)->sim->beginMessageExchange(); });
}
int fgbm_end_exchange(void *h, const char *message) {
  return guarded([&] { static_cast<Sim *>(h)->sim->endMessageExchange(message); });
}
void *fgbm_exchange_stream(void *h) { return static_cast<Sim *>(h)->sim->exchangeStream(); }
int fgbm_refresh_bounds(void *h) {
  return guarded([&] { static_cast<Sim *>(h)->sim->refreshBounds(); });
}
// JSON [["name", bytes_per_item], ...] of a message list (is_message) or of an agent's state lists
int fgbm_list_layout(void *h, int is_message, const char *name, char *buf, size_t cap) {
  return guarded([&] {
    auto lay = static_cast<Sim *>(h)->sim->listLayout(is_message != 0, name);
    std::string js = "[";
    for (size_t i = 0; i < lay.size(); ++i) js += (i ? ", [\"" : "[\"") + lay[i].first + "\", " + std::to_string(lay[i].second) + "]";
    js += "]";
    if (js.size() + 1 > cap) throw std::runtime_error("layout buffer too small");
    std::memcpy(buf, js.c_str(), js.size() + 1);
  });
}
int fgbm_slab_pack(void *h, int is_message, const char *name, const char *geometry_message, int lo, int hi, void *const *dst_lo,
                   void *const *dst_hi, unsigned int capacity, int remove, unsigned int *d_count_lo, unsigned int *d_count_hi) {
  return guarded([&] {
    static_cast<Sim *>(h)->sim->slabPack(is_message != 0, name, flamegpu::DEFAULT_STATE, geometry_message, lo, hi, dst_lo, dst_hi,
                                         capacity, remove != 0, d_count_lo, d_count_hi);
  });
}
int fgbm_list_append(void *h, int is_message, const char *name, unsigned int n_max, const unsigned int *d_n, const void *const *src) {
  return guarded([&] { static_cast<Sim *>(h)->sim->listAppend(is_message != 0, name, flamegpu::DEFAULT_STATE, n_max, d_n, src); });
}

// Phase profile (model created with profile=1): JSON {"phase": [total_ms, calls], ...} into buf.
int fgbm_profile(void *h, char *buf, size_t cap) {
  return guarded([&] {
    auto p = static_cast<Sim *>(h)->sim->getProfile();
    std::string js = "{";
    bool first = true;
    for (const auto &kv : p) {
      char tmp[256];
      std::snprintf(tmp, sizeof(tmp), "%s\"%s\": [%.6f, %u]", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
      js += tmp;
      first = false;
    }
    js += "}";
    if (js.size() + 1 > cap) throw std::runtime_error("profile buffer too small");
    std::memcpy(buf, js.c_str(), js.size() + 1);
  });
}

// Message list inspection for the parity harness: count, PBM (bin_count+1 words) and one variable.
int fgbm_message_count(void *h, const char *message, unsigned int *n) {
  return guarded([&] { *n = static_cast<Sim *>(h)->sim->getMessageCount(message); });
}
int fgbm_message_pbm(void *h, const char *message, unsigned int *host_out, unsigned int *bin_count) {
  return guarded([&] {
    Sim *s = static_cast<Sim *>(h);
    fgb_spatial *sp = s->sim->getSpatialHandler(message);
    if (!sp) throw std::runtime_error("not a spatial message list");
    fgb_spatial_metadata md;
    unsigned int bins = 0;
    fgb_spatial_get_metadata(sp, &md, &bins);
    if (bin_count) *bin_count = bins;
    if (host_out) {
      s->sim->synchronize();
      if (fgb_spatial_read_pbm(sp, host_out, s->sim->getStream()) != 0) throw std::runtime_error("fgb_spatial_read_pbm failed");
    }
  });
}
int fgbm_message_variable(void *h, const char *message, const char *var, void *host_out, size_t bytes) {
  return guarded([&] {
    Sim *s = static_cast<Sim *>(h);
    s->sim->synchronize();
    void *d = s->sim->getMessageVariableDevicePtr(message, var);
    if (bytes && cudaMemcpy(host_out, d, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) throw std::runtime_error("cudaMemcpy failed");
  });
}

// End-to-end Circles step with HOST buffers (bench.py "e2e"): upload x,y,z,drift of n agents from
// (pinned) host memory, run `steps` steps, download x,y,z,drift and the ids in device order.
int fgbm_circles_step_host(void *h, unsigned int n, const float *x, const float *y, const float *z, const float *drift,
                           unsigned int steps, float *x_out, float *y_out, float *z_out, float *drift_out,
                           unsigned int *id_out) {
  return guarded([&] {
    Sim *s = static_cast<Sim *>(h);
    const char *names[4] = {"x", "y", "z", "drift"};
    const void *ptrs[4] = {x, y, z, drift};
    s->sim->setPopulationDataSoA("Circle", flamegpu::DEFAULT_STATE, n, 4, names, ptrs);
    const char *onames[5] = {"x", "y", "z", "drift", "_id"};
    void *optrs[5] = {x_out, y_out, z_out, drift_out, id_out};
    for (unsigned int i = 0; i + 1 < steps; ++i) s->sim->step();
    // the download rides on the last step: finished chunks of `move` are copied while the rest still computes
    if (steps > 0) s->sim->streamPopulationDataSoA("Circle", flamegpu::DEFAULT_STATE, id_out ? 5 : 4, onames, optrs, 8);
    if (steps > 0) {
      s->sim->step();
      if (s->sim->finishStreamedPopulation() > n) throw std::runtime_error("population grew beyond the caller's buffers");
    } else {
      s->sim->getPopulationDataSoA("Circle", flamegpu::DEFAULT_STATE, id_out ? 5 : 4, onames, optrs, n);
    }
  });
}

}  // extern "C"
