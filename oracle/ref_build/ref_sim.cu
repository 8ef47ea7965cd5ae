// ref_sim.cu -- ORACLE SCAFFOLDING, not product code.
//
// Drives the UNMODIFIED reference library (FLAME GPU 2 built from /root/reference by build_ref.sh)
// on the example models of examples/*.cuh -- the very same model sources this repo's own API layer
// compiles -- so that the reference's real CUDASimulation::step() can be run on the GPU box beside
// ours: same seeded input state in, agent state + PBM + sorted message list + per-step times out.
//
//   ref_sim --model circles --params env_max=100,radius=2 --in state.bin --out prefix --steps 1
//           [--warmup W] [--dump-messages location] [--quiet]
//
// State container ("FGBS"): u32 magic, u32 nvars, u32 n, then per variable
//   char name[32]; u32 elem_size; u32 elements; raw SoA data (n * elem_size * elements bytes).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

// Private-member access for inspection only (PBM and sorted message list live behind
// CUDASimulation::getCUDAMessage, CUDASimulation.h:413).  Access specifiers do not change layout or
// mangling, so the unmodified library objects link unchanged.
#define private public
#define protected public
#include "flamegpu/flamegpu.h"
#include "flamegpu/simulation/detail/CUDAMessage.h"
#include "flamegpu/runtime/messaging/MessageSpatial2D/MessageSpatial2DHost.h"
#include "flamegpu/runtime/messaging/MessageSpatial3D/MessageSpatial3DHost.h"
#undef private
#undef protected

#include "../../examples/boids_model.cuh"
#include "../../examples/circles_model.cuh"
#include "../../examples/stress_model.cuh"
#include "../../examples/test_models.cuh"

namespace {

constexpr uint32_t kMagic = 0x53424746u;  // "FGBS"

struct Column {
  std::string name;
  uint32_t elem_size = 0, elements = 1;
  std::vector<char> data;
};

std::vector<Column> read_state(const std::string &path, uint32_t *n_out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  uint32_t magic = 0, nvars = 0, n = 0;
  f.read(reinterpret_cast<char *>(&magic), 4);
  f.read(reinterpret_cast<char *>(&nvars), 4);
  f.read(reinterpret_cast<char *>(&n), 4);
  if (magic != kMagic) throw std::runtime_error("bad magic in " + path);
  std::vector<Column> cols(nvars);
  for (auto &c : cols) {
    char name[32];
    f.read(name, 32);
    name[31] = 0;
    c.name = name;
    f.read(reinterpret_cast<char *>(&c.elem_size), 4);
    f.read(reinterpret_cast<char *>(&c.elements), 4);
    c.data.resize(static_cast<size_t>(n) * c.elem_size * c.elements);
    f.read(c.data.data(), c.data.size());
  }
  *n_out = n;
  return cols;
}

void write_state(const std::string &path, uint32_t n, const std::vector<Column> &cols) {
  std::ofstream f(path, std::ios::binary);
  const uint32_t nvars = static_cast<uint32_t>(cols.size());
  f.write(reinterpret_cast<const char *>(&kMagic), 4);
  f.write(reinterpret_cast<const char *>(&nvars), 4);
  f.write(reinterpret_cast<const char *>(&n), 4);
  for (const auto &c : cols) {
    char name[32] = {0};
    std::strncpy(name, c.name.c_str(), 31);
    f.write(name, 32);
    f.write(reinterpret_cast<const char *>(&c.elem_size), 4);
    f.write(reinterpret_cast<const char *>(&c.elements), 4);
    f.write(c.data.data(), c.data.size());
  }
}

std::map<std::string, std::string> parse_kv(const std::string &s) {
  std::map<std::string, std::string> kv;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, ',')) {
    const size_t eq = item.find('=');
    if (eq != std::string::npos) kv[item.substr(0, eq)] = item.substr(eq + 1);
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

void dump_agents(flamegpu::CUDASimulation &sim, flamegpu::ModelDescription &model, const std::string &agent, const std::string &prefix,
                 const std::string &state = "") {
  flamegpu::AgentVector pop(model.Agent(agent));
  if (state.empty()) sim.getPopulationData(pop);
  else sim.getPopulationData(pop, state);
  const flamegpu::AgentVector &cpop = pop;
  std::vector<Column> cols;
  for (const auto &v : cpop.getVariableMetaData()) {
    Column c;
    c.name = v.first;
    c.elem_size = static_cast<uint32_t>(v.second.type_size);
    c.elements = v.second.elements;
    c.data.resize(static_cast<size_t>(cpop.size()) * c.elem_size * c.elements);
    const void *src = cpop.data(v.first);
    if (src && !c.data.empty()) std::memcpy(c.data.data(), src, c.data.size());
    cols.push_back(std::move(c));
  }
  write_state(prefix + "." + agent + (state.empty() ? "" : "." + state) + ".bin", cpop.size(), cols);
}

// PBM + the bin-sorted message list of a spatial message, straight from the reference's device buffers
void dump_messages(flamegpu::CUDASimulation &sim, const std::string &message, int dims, const std::string &prefix) {
  flamegpu::detail::CUDAMessage &m = sim.getCUDAMessage(message);
  unsigned int bins = 1;
  unsigned int *d_pbm = nullptr;
  if (dims == 0) {  // bucket list: MetaData {min, max exclusive, PBM}
    flamegpu::MessageBucket::MetaData md;
    cudaMemcpy(&md, m.getMetaDataDevicePtr(), sizeof(md), cudaMemcpyDeviceToHost);
    bins = static_cast<unsigned int>(md.max - md.min);
    d_pbm = md.PBM;
  } else if (dims == 3) {
    flamegpu::MessageSpatial3D::MetaData md;
    cudaMemcpy(&md, m.getMetaDataDevicePtr(), sizeof(md), cudaMemcpyDeviceToHost);
    bins = md.gridDim[0] * md.gridDim[1] * md.gridDim[2];
    d_pbm = md.PBM;
  } else {
    flamegpu::MessageSpatial2D::MetaData md;
    cudaMemcpy(&md, m.getMetaDataDevicePtr(), sizeof(md), cudaMemcpyDeviceToHost);
    bins = md.gridDim[0] * md.gridDim[1];
    d_pbm = md.PBM;
  }
  std::vector<Column> cols;
  Column pbm;
  pbm.name = "_pbm";
  pbm.elem_size = 4;
  pbm.elements = 1;
  pbm.data.resize((static_cast<size_t>(bins) + 1) * 4);
  cudaMemcpy(pbm.data.data(), d_pbm, pbm.data.size(), cudaMemcpyDeviceToHost);
  const uint32_t count = reinterpret_cast<const uint32_t *>(pbm.data.data())[bins];
  // the container stores one n for all columns: write the PBM as its own file
  write_state(prefix + ".pbm." + message + ".bin", bins + 1, {pbm});
  for (const auto &v : m.getMessageData().variables) {
    Column c;
    c.name = v.first;
    c.elem_size = static_cast<uint32_t>(v.second.type_size);
    c.elements = v.second.elements;
    c.data.resize(static_cast<size_t>(count) * c.elem_size * c.elements);
    if (count) cudaMemcpy(c.data.data(), m.getReadPtr(v.first), c.data.size(), cudaMemcpyDeviceToHost);
    cols.push_back(std::move(c));
  }
  write_state(prefix + ".msg." + message + ".bin", count, cols);
}

}  // namespace

int main(int argc, const char **argv) {
  std::string model_name = "circles", params, in_path, out_prefix = "ref_out", dump_msg;
  std::vector<std::string> pops;   // --pop agent:state:path (state may be empty = the agent's initial state), repeatable
  std::vector<std::string> dumps;  // --dump agent:state, repeatable: <out>.<agent>.<state>.bin
  unsigned int steps = 1, warmup = 0;
  bool quiet = false, dump_steps = false;  // --dump-steps: <out>.s<k>.<agent>.bin after every step (teacher-forced parity)
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--model") model_name = next();
    else if (a == "--params") params = next();
    else if (a == "--in") in_path = next();
    else if (a == "--out") out_prefix = next();
    else if (a == "--steps") steps = static_cast<unsigned int>(std::strtoul(next().c_str(), nullptr, 10));
    else if (a == "--warmup") warmup = static_cast<unsigned int>(std::strtoul(next().c_str(), nullptr, 10));
    else if (a == "--dump-messages") dump_msg = next();
    else if (a == "--quiet") quiet = true;
    else if (a == "--dump-steps") dump_steps = true;
    else if (a == "--pop") pops.push_back(next());
    else if (a == "--dump") dumps.push_back(next());
  }
  try {
    flamegpu::io::Telemetry::disable();
    const auto kv = parse_kv(params);
    flamegpu::ModelDescription model(model_name);
    std::string agent_name = "Circle";
    int msg_dims = 3;
    if (model_name == "circles") {
      fgb_examples::CirclesParams p;
      p.env_max = getf(kv, "env_max", p.env_max);
      p.radius = getf(kv, "radius", p.radius);
      p.repulse = getf(kv, "repulse", p.repulse);
      p.sort_period = getu(kv, "sort_period", p.sort_period);
      p.validation = getu(kv, "validation", p.validation);
      p.env_max_z = getf(kv, "env_max_z", p.env_max_z);
      fgb_examples::define_circles(model, p);
    } else if (model_name == "boids3d" || model_name == "boids2d") {
      fgb_examples::BoidsParams p;
      p.dims = model_name == "boids3d" ? 3 : 2;
      msg_dims = p.dims;
      p.min_position = getf(kv, "min_position", p.min_position);
      p.max_position = getf(kv, "max_position", p.max_position);
      p.interaction_radius = getf(kv, "interaction_radius", p.interaction_radius);
      p.separation_radius = getf(kv, "separation_radius", p.separation_radius);
      fgb_examples::define_boids(model, p);
      agent_name = "Boid";
    } else if (model_name == "stress") {
      fgb_examples::StressParams p;
      p.env_max = getf(kv, "env_max", p.env_max);
      p.radius = getf(kv, "radius", p.radius);
      p.death_mod = getu(kv, "death_mod", p.death_mod);
      p.birth_mod = getu(kv, "birth_mod", p.birth_mod);
      fgb_examples::define_stress(model, p);
    } else if (model_name == "test") {
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
      msg_dims = (p.which == fgb_examples::TM_COUNT2D || p.which == fgb_examples::TM_WRAP2D) ? 2 : 3;
      if (p.which >= fgb_examples::TM_BUCKET && p.which <= fgb_examples::TM_BUCKET_RANGE) msg_dims = 0;
      fgb_examples::define_test_model(model, p);
      agent_name = "agent";
    } else {
      throw std::runtime_error("unknown model " + model_name);
    }
    flamegpu::CUDASimulation sim(model);
    sim.SimulationConfig().steps = 1;
    sim.SimulationConfig().verbosity = flamegpu::Verbosity::Quiet;
    sim.SimulationConfig().telemetry = false;
    sim.applyConfig();
    uint32_t n = 0;
    for (const std::string &spec : pops) {
      const size_t c1 = spec.find(':'), c2 = spec.find(':', c1 + 1);
      if (c1 == std::string::npos || c2 == std::string::npos) throw std::runtime_error("--pop expects agent:state:path");
      const std::string an = spec.substr(0, c1), st = spec.substr(c1 + 1, c2 - c1 - 1), path = spec.substr(c2 + 1);
      uint32_t m = 0;
      std::vector<Column> cols = read_state(path, &m);
      flamegpu::AgentVector pop(model.Agent(an), m);
      for (const auto &c : cols) {
        if (!c.name.empty() && c.name[0] == '_') continue;  // "_n": carries only the population size
        void *dst = pop.data(c.name);
        if (dst && !c.data.empty()) std::memcpy(dst, c.data.data(), c.data.size());
      }
      if (st.empty()) sim.setPopulationData(pop);
      else sim.setPopulationData(pop, st);
      n += m;
    }
    if (!in_path.empty()) {
      std::vector<Column> cols = read_state(in_path, &n);
      flamegpu::AgentVector pop(model.Agent(agent_name), n);
      for (const auto &c : cols) {
        void *dst = pop.data(c.name);
        if (dst && !c.data.empty()) std::memcpy(dst, c.data.data(), c.data.size());
      }
      sim.setPopulationData(pop);
    }
    for (unsigned int i = 0; i < warmup + steps; ++i) {
      sim.step();
      if (dump_steps) dump_agents(sim, model, agent_name, out_prefix + ".s" + std::to_string(i + 1));
    }
    cudaDeviceSynchronize();
    const std::vector<double> t = sim.getElapsedTimeSteps();
    if (dumps.empty()) {
      dump_agents(sim, model, agent_name, out_prefix);
      if (model_name == "test" && getu(kv, "which", 0) == fgb_examples::TM_BIRTH_OTHER_AGENT) dump_agents(sim, model, "agent2", out_prefix);
    }
    for (const std::string &spec : dumps) {
      const size_t c1 = spec.find(':');
      dump_agents(sim, model, spec.substr(0, c1), out_prefix, c1 == std::string::npos ? "" : spec.substr(c1 + 1));
    }
    if (!dump_msg.empty()) dump_messages(sim, dump_msg, msg_dims, out_prefix);
    // per-step seconds of the timed steps (after warm-up) as one JSON line
    std::printf("{\"impl\": \"flamegpu2-reference-cuda\", \"model\": \"%s\", \"n\": %u, \"steps\": %u, \"warmup\": %u, \"step_seconds\": [",
                model_name.c_str(), n, steps, warmup);
    for (size_t i = warmup; i < t.size(); ++i) std::printf("%s%.9f", i > warmup ? ", " : "", t[i]);
    std::printf("]}\n");
    if (!quiet) std::fflush(stdout);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "ref_sim: %s\n", e.what());
    return 1;
  }
  return 0;
}
