// examples/circles_model.cuh -- the Circles-3D benchmark model (BASELINE.json configs 1, 2, 5),
// written against the FLAME GPU 2 API only, so the same file compiles against the reference's
// headers (oracle/ref_build/ref_sim.cu) and against this repo's include/flamegpu (fgb_models.cu).
// Model definition follows the reference example examples/cpp/circles_spatial3D/src/main.cu:5-54,
// 77-113 (same variables, same arithmetic); population size, extent and radius are parameters.
#pragma once
#include <cstdio>

#include "flamegpu/flamegpu.h"

namespace fgb_examples {

FLAMEGPU_AGENT_FUNCTION(circles_output, flamegpu::MessageNone, flamegpu::MessageSpatial3D) {
  FLAMEGPU->message_out.setVariable<flamegpu::id_t>("id", FLAMEGPU->getID());
  FLAMEGPU->message_out.setLocation(FLAMEGPU->getVariable<float>("x"), FLAMEGPU->getVariable<float>("y"),
                                    FLAMEGPU->getVariable<float>("z"));
  return flamegpu::ALIVE;
}

FLAMEGPU_AGENT_FUNCTION(circles_move, flamegpu::MessageSpatial3D, flamegpu::MessageNone) {
  const flamegpu::id_t ID = FLAMEGPU->getID();
  const float REPULSE_FACTOR = FLAMEGPU->environment.getProperty<float>("repulse");
  const float RADIUS = FLAMEGPU->message_in.radius();
  float fx = 0.0;
  float fy = 0.0;
  float fz = 0.0;
  const float x1 = FLAMEGPU->getVariable<float>("x");
  const float y1 = FLAMEGPU->getVariable<float>("y");
  const float z1 = FLAMEGPU->getVariable<float>("z");
  int count = 0;
  for (const auto &message : FLAMEGPU->message_in(x1, y1, z1)) {
    if (message.getVariable<flamegpu::id_t>("id") != ID) {
      const float x2 = message.getVariable<float>("x");
      const float y2 = message.getVariable<float>("y");
      const float z2 = message.getVariable<float>("z");
      float x21 = x2 - x1;
      float y21 = y2 - y1;
      float z21 = z2 - z1;
      const float separation = sqrtf(x21 * x21 + y21 * y21 + z21 * z21);
      if (separation < RADIUS && separation > 0.0f) {
        float k = sinf((separation / RADIUS) * 3.141f * -2) * REPULSE_FACTOR;
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
  fx /= count > 0 ? count : 1;
  fy /= count > 0 ? count : 1;
  fz /= count > 0 ? count : 1;
  FLAMEGPU->setVariable<float>("x", x1 + fx);
  FLAMEGPU->setVariable<float>("y", y1 + fy);
  FLAMEGPU->setVariable<float>("z", z1 + fz);
  FLAMEGPU->setVariable<float>("drift", sqrtf(fx * fx + fy * fy + fz * fz));
  return flamegpu::ALIVE;
}

// The example's own step function (reference examples/cpp/circles_spatial3D/src/main.cu:55-70): one device reduction
// per step whose result the host inspects.  The per-step printf of the example is kept behind `verbose`.
struct CirclesValidationState {
  float prev_total_drift = 3.402823466e+38f;
  unsigned int dropped = 0, increased = 0;
  bool verbose = false;
};
inline CirclesValidationState &circles_validation_state() {
  static CirclesValidationState s;
  return s;
}
FLAMEGPU_STEP_FUNCTION(circles_validation) {
  CirclesValidationState &v = circles_validation_state();
  const float totalDrift = FLAMEGPU->agent("Circle").sum<float>("drift");
  if (totalDrift <= v.prev_total_drift) v.dropped++;
  else v.increased++;
  v.prev_total_drift = totalDrift;
  if (v.verbose) printf("%.2f%% Drift correct\n", 100 * v.dropped / static_cast<float>(v.dropped + v.increased));
}

struct CirclesParams {
  float env_max = 25.0f;   // reference example: floor(cbrt(16384)) = 25
  float env_max_z = 0.0f;  // > 0: non-cubic box [0,env_max)^2 x [0,env_max_z) (multi-GPU weak scaling stacks slabs in z)
  float radius = 2.0f;
  float repulse = 0.05f;
  unsigned int sort_period = 1;
  unsigned int validation = 1;      // attach the example's Validation step function (sum of "drift" every step)
  unsigned int radius_filtered = 1; // this repo only: declare that `move` ignores messages beyond the radius (see below)
};

inline void define_circles(flamegpu::ModelDescription &model, const CirclesParams &p) {
  {
    flamegpu::MessageSpatial3D::Description message = model.newMessage<flamegpu::MessageSpatial3D>("location");
    message.newVariable<flamegpu::id_t>("id");
    message.setRadius(p.radius);
    message.setMin(0, 0, 0);
    message.setMax(p.env_max, p.env_max, p.env_max_z > 0.0f ? p.env_max_z : p.env_max);
  }
  {
    flamegpu::AgentDescription agent = model.newAgent("Circle");
    agent.newVariable<float>("x");
    agent.newVariable<float>("y");
    agent.newVariable<float>("z");
    agent.newVariable<float>("drift");
    agent.setSortPeriod(p.sort_period);
    agent.newFunction("output_message", circles_output).setMessageOutput("location");
    flamegpu::AgentFunctionDescription mv = agent.newFunction("move", circles_move);
    mv.setMessageInput("location");
#ifdef FLAMEGPU2_B200
    // b200 extension (no reference counterpart): `move` tests `separation < RADIUS` itself, so it may be shown only the
    // messages within the radius of its search origin (AgentFunctionDescription::setMessageInputRadiusFiltered).
    mv.setMessageInputRadiusFiltered(p.radius_filtered != 0);
#endif
  }
  model.Environment().newProperty("repulse", p.repulse);
  if (p.validation) model.addStepFunction(circles_validation);
  model.newLayer().addAgentFunction(circles_output);
  model.newLayer().addAgentFunction(circles_move);
}

}  // namespace fgb_examples
