// examples/stress_model.cuh -- synthetic birth/death stress (BASELINE.json config 3): Circles layout,
// spatial3D messages output and read every step; each step every agent with
// hash32(_id, step) % death_mod == 0 dies and every agent with hash32(_id ^ 0x9e3779b9, step) %
// birth_mod == 0 emits one child at its own position through agent_out.  Decisions are integer
// hashes of (id, step), not FLAMEGPU->random, so they do not depend on thread order.
// FLAME GPU 2 API only: compiles against the reference headers and against include/flamegpu.
#pragma once
#include "flamegpu/flamegpu.h"

namespace fgb_examples {

FLAMEGPU_HOST_DEVICE_FUNCTION unsigned int stress_hash32(unsigned int a, unsigned int b) {
  unsigned int h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u + (a << 6) + (a >> 2));
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

FLAMEGPU_AGENT_FUNCTION(stress_output, flamegpu::MessageNone, flamegpu::MessageSpatial3D) {
  FLAMEGPU->message_out.setVariable<flamegpu::id_t>("id", FLAMEGPU->getID());
  FLAMEGPU->message_out.setLocation(FLAMEGPU->getVariable<float>("x"), FLAMEGPU->getVariable<float>("y"),
                                    FLAMEGPU->getVariable<float>("z"));
  return flamegpu::ALIVE;
}

// neighbour count within the radius (integer result: bit-exact against the oracle), then the
// birth / death decisions
FLAMEGPU_AGENT_FUNCTION(stress_update, flamegpu::MessageSpatial3D, flamegpu::MessageNone) {
  const flamegpu::id_t ID = FLAMEGPU->getID();
  const float RADIUS = FLAMEGPU->message_in.radius();
  const float x1 = FLAMEGPU->getVariable<float>("x");
  const float y1 = FLAMEGPU->getVariable<float>("y");
  const float z1 = FLAMEGPU->getVariable<float>("z");
  unsigned int neighbours = 0;
  for (const auto &message : FLAMEGPU->message_in(x1, y1, z1)) {
    if (message.getVariable<flamegpu::id_t>("id") != ID) {
      const float dx = message.getVariable<float>("x") - x1;
      const float dy = message.getVariable<float>("y") - y1;
      const float dz = message.getVariable<float>("z") - z1;
      // every product / sum rounded separately so the integer result is bit-exact against a CPU restatement
      if (__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)) < __fmul_rn(RADIUS, RADIUS)) ++neighbours;
    }
  }
  FLAMEGPU->setVariable<unsigned int>("neighbours", neighbours);
  const unsigned int step = FLAMEGPU->getStepCounter();
  const unsigned int death_mod = FLAMEGPU->environment.getProperty<unsigned int>("death_mod");
  const unsigned int birth_mod = FLAMEGPU->environment.getProperty<unsigned int>("birth_mod");
  if (birth_mod && stress_hash32(ID ^ 0x9e3779b9u, step) % birth_mod == 0) {
    FLAMEGPU->agent_out.setVariable<float>("x", x1);
    FLAMEGPU->agent_out.setVariable<float>("y", y1);
    FLAMEGPU->agent_out.setVariable<float>("z", z1);
    FLAMEGPU->agent_out.setVariable<unsigned int>("parent", ID);
  }
  if (death_mod && stress_hash32(ID, step) % death_mod == 0) return flamegpu::DEAD;
  return flamegpu::ALIVE;
}

struct StressParams {
  float env_max = 25.0f;
  float radius = 2.0f;
  unsigned int death_mod = 10;
  unsigned int birth_mod = 20;
};

inline void define_stress(flamegpu::ModelDescription &model, const StressParams &p) {
  {
    flamegpu::MessageSpatial3D::Description message = model.newMessage<flamegpu::MessageSpatial3D>("location");
    message.newVariable<flamegpu::id_t>("id");
    message.setRadius(p.radius);
    message.setMin(0, 0, 0);
    message.setMax(p.env_max, p.env_max, p.env_max);
  }
  flamegpu::AgentDescription agent = model.newAgent("Circle");
  agent.newVariable<float>("x");
  agent.newVariable<float>("y");
  agent.newVariable<float>("z");
  agent.newVariable<unsigned int>("neighbours", 0u);
  agent.newVariable<unsigned int>("parent", 0u);
  agent.newFunction("output_message", stress_output).setMessageOutput("location");
  flamegpu::AgentFunctionDescription upd = agent.newFunction("update", stress_update);
  upd.setMessageInput("location");
  upd.setAllowAgentDeath(true);
  upd.setAgentOutput(agent);
#ifdef FLAMEGPU2_B200
  upd.setMessageInputRadiusFiltered(true);  // b200 extension: `update` counts only messages within the radius
#endif
  flamegpu::EnvironmentDescription env = model.Environment();
  env.newProperty<unsigned int>("death_mod", p.death_mod);
  env.newProperty<unsigned int>("birth_mod", p.birth_mod);
  model.newLayer().addAgentFunction(stress_output);
  model.newLayer().addAgentFunction(stress_update);
}

}  // namespace fgb_examples
