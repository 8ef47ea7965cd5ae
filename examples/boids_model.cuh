// examples/boids_model.cuh -- Boids with spatial messaging, 3D (BASELINE.json config 0: the
// reference's examples/cpp/boids_spatial3D, same variables / arithmetic / constants, main.cu:121-342)
// and the 2D variant of BASELINE.json config 2 (no reference example exists: the 3D model with z / fz
// dropped).  Written against the FLAME GPU 2 API only; compiles against either header tree.
#pragma once
#include "flamegpu/flamegpu.h"

namespace fgb_examples {

FLAMEGPU_HOST_DEVICE_FUNCTION float boids_len3(const float x, const float y, const float z) { return sqrtf(x * x + y * y + z * z); }
FLAMEGPU_HOST_DEVICE_FUNCTION float boids_len2(const float x, const float y) { return sqrtf(x * x + y * y); }

FLAMEGPU_AGENT_FUNCTION(boids3d_output, flamegpu::MessageNone, flamegpu::MessageSpatial3D) {
  FLAMEGPU->message_out.setVariable<flamegpu::id_t>("id", FLAMEGPU->getID());
  FLAMEGPU->message_out.setVariable<float>("x", FLAMEGPU->getVariable<float>("x"));
  FLAMEGPU->message_out.setVariable<float>("y", FLAMEGPU->getVariable<float>("y"));
  FLAMEGPU->message_out.setVariable<float>("z", FLAMEGPU->getVariable<float>("z"));
  FLAMEGPU->message_out.setVariable<float>("fx", FLAMEGPU->getVariable<float>("fx"));
  FLAMEGPU->message_out.setVariable<float>("fy", FLAMEGPU->getVariable<float>("fy"));
  FLAMEGPU->message_out.setVariable<float>("fz", FLAMEGPU->getVariable<float>("fz"));
  return flamegpu::ALIVE;
}

FLAMEGPU_AGENT_FUNCTION(boids3d_input, flamegpu::MessageSpatial3D, flamegpu::MessageNone) {
  const flamegpu::id_t id = FLAMEGPU->getID();
  float agent_x = FLAMEGPU->getVariable<float>("x");
  float agent_y = FLAMEGPU->getVariable<float>("y");
  float agent_z = FLAMEGPU->getVariable<float>("z");
  float agent_fx = FLAMEGPU->getVariable<float>("fx");
  float agent_fy = FLAMEGPU->getVariable<float>("fy");
  float agent_fz = FLAMEGPU->getVariable<float>("fz");
  float perceived_centre_x = 0.0f, perceived_centre_y = 0.0f, perceived_centre_z = 0.0f;
  int perceived_count = 0;
  float global_velocity_x = 0.0f, global_velocity_y = 0.0f, global_velocity_z = 0.0f;
  float velocity_change_x = 0.f, velocity_change_y = 0.f, velocity_change_z = 0.f;
  const float INTERACTION_RADIUS = FLAMEGPU->environment.getProperty<float>("INTERACTION_RADIUS");
  const float SEPARATION_RADIUS = FLAMEGPU->environment.getProperty<float>("SEPARATION_RADIUS");
  for (const auto &message : FLAMEGPU->message_in(agent_x, agent_y, agent_z)) {
    if (message.getVariable<flamegpu::id_t>("id") != id) {
      const float message_x = message.getVariable<float>("x");
      const float message_y = message.getVariable<float>("y");
      const float message_z = message.getVariable<float>("z");
      float separation = boids_len3(agent_x - message_x, agent_y - message_y, agent_z - message_z);
      if (separation < INTERACTION_RADIUS) {
        perceived_centre_x += message_x;
        perceived_centre_y += message_y;
        perceived_centre_z += message_z;
        perceived_count++;
        const float message_fx = message.getVariable<float>("fx");
        const float message_fy = message.getVariable<float>("fy");
        const float message_fz = message.getVariable<float>("fz");
        global_velocity_x += message_fx;
        global_velocity_y += message_fy;
        global_velocity_z += message_fz;
        if (separation < (SEPARATION_RADIUS)) {
          float normalizedSeparation = (separation / SEPARATION_RADIUS);
          float invNormSep = (1.0f - normalizedSeparation);
          float invSqSep = invNormSep * invNormSep;
          const float collisionScale = FLAMEGPU->environment.getProperty<float>("COLLISION_SCALE");
          velocity_change_x += collisionScale * (agent_x - message_x) * invSqSep;
          velocity_change_y += collisionScale * (agent_y - message_y) * invSqSep;
          velocity_change_z += collisionScale * (agent_z - message_z) * invSqSep;
        }
      }
    }
  }
  if (perceived_count) {
    perceived_centre_x /= perceived_count;
    perceived_centre_y /= perceived_count;
    perceived_centre_z /= perceived_count;
    global_velocity_x /= perceived_count;
    global_velocity_y /= perceived_count;
    global_velocity_z /= perceived_count;
    const float STEER_SCALE = FLAMEGPU->environment.getProperty<float>("STEER_SCALE");
    velocity_change_x += (perceived_centre_x - agent_x) * STEER_SCALE;
    velocity_change_y += (perceived_centre_y - agent_y) * STEER_SCALE;
    velocity_change_z += (perceived_centre_z - agent_z) * STEER_SCALE;
    const float MATCH_SCALE = FLAMEGPU->environment.getProperty<float>("MATCH_SCALE");
    velocity_change_x += global_velocity_x * MATCH_SCALE - agent_fx;
    velocity_change_y += global_velocity_y * MATCH_SCALE - agent_fy;
    velocity_change_z += global_velocity_z * MATCH_SCALE - agent_fz;
  }
  const float GLOBAL_SCALE = FLAMEGPU->environment.getProperty<float>("GLOBAL_SCALE");
  velocity_change_x *= GLOBAL_SCALE;
  velocity_change_y *= GLOBAL_SCALE;
  velocity_change_z *= GLOBAL_SCALE;
  agent_fx += velocity_change_x;
  agent_fy += velocity_change_y;
  agent_fz += velocity_change_z;
  float agent_fscale = boids_len3(agent_fx, agent_fy, agent_fz);
  if (agent_fscale > 1) {
    agent_fx /= agent_fscale;
    agent_fy /= agent_fscale;
    agent_fz /= agent_fscale;
  }
  float minSpeed = 0.5f;
  if (agent_fscale < minSpeed) {
    agent_fx /= agent_fscale;
    agent_fy /= agent_fscale;
    agent_fz /= agent_fscale;
    agent_fx *= minSpeed;
    agent_fy *= minSpeed;
    agent_fz *= minSpeed;
  }
  const float wallInteractionDistance = 0.10f;
  const float wallSteerStrength = 0.05f;
  const float minPosition = FLAMEGPU->environment.getProperty<float>("MIN_POSITION");
  const float maxPosition = FLAMEGPU->environment.getProperty<float>("MAX_POSITION");
  if (agent_x - minPosition < wallInteractionDistance) agent_fx += wallSteerStrength;
  if (agent_y - minPosition < wallInteractionDistance) agent_fy += wallSteerStrength;
  if (agent_z - minPosition < wallInteractionDistance) agent_fz += wallSteerStrength;
  if (maxPosition - agent_x < wallInteractionDistance) agent_fx -= wallSteerStrength;
  if (maxPosition - agent_y < wallInteractionDistance) agent_fy -= wallSteerStrength;
  if (maxPosition - agent_z < wallInteractionDistance) agent_fz -= wallSteerStrength;
  const float TIME_SCALE = FLAMEGPU->environment.getProperty<float>("TIME_SCALE");
  agent_x += agent_fx * TIME_SCALE;
  agent_y += agent_fy * TIME_SCALE;
  agent_z += agent_fz * TIME_SCALE;
  agent_x = (agent_x < minPosition) ? minPosition : agent_x;
  agent_x = (agent_x > maxPosition) ? maxPosition : agent_x;
  agent_y = (agent_y < minPosition) ? minPosition : agent_y;
  agent_y = (agent_y > maxPosition) ? maxPosition : agent_y;
  agent_z = (agent_z < minPosition) ? minPosition : agent_z;
  agent_z = (agent_z > maxPosition) ? maxPosition : agent_z;
  FLAMEGPU->setVariable<float>("x", agent_x);
  FLAMEGPU->setVariable<float>("y", agent_y);
  FLAMEGPU->setVariable<float>("z", agent_z);
  FLAMEGPU->setVariable<float>("fx", agent_fx);
  FLAMEGPU->setVariable<float>("fy", agent_fy);
  FLAMEGPU->setVariable<float>("fz", agent_fz);
  return flamegpu::ALIVE;
}

FLAMEGPU_AGENT_FUNCTION(boids2d_output, flamegpu::MessageNone, flamegpu::MessageSpatial2D) {
  FLAMEGPU->message_out.setVariable<flamegpu::id_t>("id", FLAMEGPU->getID());
  FLAMEGPU->message_out.setLocation(FLAMEGPU->getVariable<float>("x"), FLAMEGPU->getVariable<float>("y"));
  FLAMEGPU->message_out.setVariable<float>("fx", FLAMEGPU->getVariable<float>("fx"));
  FLAMEGPU->message_out.setVariable<float>("fy", FLAMEGPU->getVariable<float>("fy"));
  return flamegpu::ALIVE;
}

FLAMEGPU_AGENT_FUNCTION(boids2d_input, flamegpu::MessageSpatial2D, flamegpu::MessageNone) {
  const flamegpu::id_t id = FLAMEGPU->getID();
  float agent_x = FLAMEGPU->getVariable<float>("x");
  float agent_y = FLAMEGPU->getVariable<float>("y");
  float agent_fx = FLAMEGPU->getVariable<float>("fx");
  float agent_fy = FLAMEGPU->getVariable<float>("fy");
  float perceived_centre_x = 0.0f, perceived_centre_y = 0.0f;
  int perceived_count = 0;
  float global_velocity_x = 0.0f, global_velocity_y = 0.0f;
  float velocity_change_x = 0.f, velocity_change_y = 0.f;
  const float INTERACTION_RADIUS = FLAMEGPU->environment.getProperty<float>("INTERACTION_RADIUS");
  const float SEPARATION_RADIUS = FLAMEGPU->environment.getProperty<float>("SEPARATION_RADIUS");
  for (const auto &message : FLAMEGPU->message_in(agent_x, agent_y)) {
    if (message.getVariable<flamegpu::id_t>("id") != id) {
      const float message_x = message.getVariable<float>("x");
      const float message_y = message.getVariable<float>("y");
      float separation = boids_len2(agent_x - message_x, agent_y - message_y);
      if (separation < INTERACTION_RADIUS) {
        perceived_centre_x += message_x;
        perceived_centre_y += message_y;
        perceived_count++;
        global_velocity_x += message.getVariable<float>("fx");
        global_velocity_y += message.getVariable<float>("fy");
        if (separation < (SEPARATION_RADIUS)) {
          float normalizedSeparation = (separation / SEPARATION_RADIUS);
          float invNormSep = (1.0f - normalizedSeparation);
          float invSqSep = invNormSep * invNormSep;
          const float collisionScale = FLAMEGPU->environment.getProperty<float>("COLLISION_SCALE");
          velocity_change_x += collisionScale * (agent_x - message_x) * invSqSep;
          velocity_change_y += collisionScale * (agent_y - message_y) * invSqSep;
        }
      }
    }
  }
  if (perceived_count) {
    perceived_centre_x /= perceived_count;
    perceived_centre_y /= perceived_count;
    global_velocity_x /= perceived_count;
    global_velocity_y /= perceived_count;
    const float STEER_SCALE = FLAMEGPU->environment.getProperty<float>("STEER_SCALE");
    velocity_change_x += (perceived_centre_x - agent_x) * STEER_SCALE;
    velocity_change_y += (perceived_centre_y - agent_y) * STEER_SCALE;
    const float MATCH_SCALE = FLAMEGPU->environment.getProperty<float>("MATCH_SCALE");
    velocity_change_x += global_velocity_x * MATCH_SCALE - agent_fx;
    velocity_change_y += global_velocity_y * MATCH_SCALE - agent_fy;
  }
  const float GLOBAL_SCALE = FLAMEGPU->environment.getProperty<float>("GLOBAL_SCALE");
  agent_fx += velocity_change_x * GLOBAL_SCALE;
  agent_fy += velocity_change_y * GLOBAL_SCALE;
  float agent_fscale = boids_len2(agent_fx, agent_fy);
  if (agent_fscale > 1) {
    agent_fx /= agent_fscale;
    agent_fy /= agent_fscale;
  }
  float minSpeed = 0.5f;
  if (agent_fscale < minSpeed) {
    agent_fx = agent_fx / agent_fscale * minSpeed;
    agent_fy = agent_fy / agent_fscale * minSpeed;
  }
  const float wallInteractionDistance = 0.10f;
  const float wallSteerStrength = 0.05f;
  const float minPosition = FLAMEGPU->environment.getProperty<float>("MIN_POSITION");
  const float maxPosition = FLAMEGPU->environment.getProperty<float>("MAX_POSITION");
  if (agent_x - minPosition < wallInteractionDistance) agent_fx += wallSteerStrength;
  if (agent_y - minPosition < wallInteractionDistance) agent_fy += wallSteerStrength;
  if (maxPosition - agent_x < wallInteractionDistance) agent_fx -= wallSteerStrength;
  if (maxPosition - agent_y < wallInteractionDistance) agent_fy -= wallSteerStrength;
  const float TIME_SCALE = FLAMEGPU->environment.getProperty<float>("TIME_SCALE");
  agent_x += agent_fx * TIME_SCALE;
  agent_y += agent_fy * TIME_SCALE;
  agent_x = (agent_x < minPosition) ? minPosition : agent_x;
  agent_x = (agent_x > maxPosition) ? maxPosition : agent_x;
  agent_y = (agent_y < minPosition) ? minPosition : agent_y;
  agent_y = (agent_y > maxPosition) ? maxPosition : agent_y;
  FLAMEGPU->setVariable<float>("x", agent_x);
  FLAMEGPU->setVariable<float>("y", agent_y);
  FLAMEGPU->setVariable<float>("fx", agent_fx);
  FLAMEGPU->setVariable<float>("fy", agent_fy);
  return flamegpu::ALIVE;
}

struct BoidsParams {
  int dims = 3;
  float min_position = -0.5f, max_position = 0.5f;
  float interaction_radius = 0.05f, separation_radius = 0.01f;
  float time_scale = 0.0005f, global_scale = 0.15f, steer_scale = 0.055f, collision_scale = 10.0f, match_scale = 0.015f;
};

inline void define_boids(flamegpu::ModelDescription &model, const BoidsParams &p) {
  flamegpu::EnvironmentDescription env = model.Environment();
  env.newProperty<float>("MIN_POSITION", p.min_position);
  env.newProperty<float>("MAX_POSITION", p.max_position);
  env.newProperty<float>("INTERACTION_RADIUS", p.interaction_radius);
  env.newProperty<float>("SEPARATION_RADIUS", p.separation_radius);
  env.newProperty<float>("TIME_SCALE", p.time_scale);
  env.newProperty<float>("GLOBAL_SCALE", p.global_scale);
  env.newProperty<float>("STEER_SCALE", p.steer_scale);
  env.newProperty<float>("COLLISION_SCALE", p.collision_scale);
  env.newProperty<float>("MATCH_SCALE", p.match_scale);
  flamegpu::AgentDescription agent = model.newAgent("Boid");
  agent.newVariable<float>("x");
  agent.newVariable<float>("y");
  agent.newVariable<float>("fx");
  agent.newVariable<float>("fy");
  if (p.dims == 3) {
    flamegpu::MessageSpatial3D::Description message = model.newMessage<flamegpu::MessageSpatial3D>("location");
    message.setRadius(p.interaction_radius);
    message.setMin(p.min_position, p.min_position, p.min_position);
    message.setMax(p.max_position, p.max_position, p.max_position);
    message.newVariable<flamegpu::id_t>("id");
    message.newVariable<float>("fx");
    message.newVariable<float>("fy");
    message.newVariable<float>("fz");
    agent.newVariable<float>("z");
    agent.newVariable<float>("fz");
    agent.newFunction("outputdata", boids3d_output).setMessageOutput("location");
    flamegpu::AgentFunctionDescription in3 = agent.newFunction("inputdata", boids3d_input);
    in3.setMessageInput("location");
#ifdef FLAMEGPU2_B200
    // b200 extension (no reference counterpart): `inputdata` tests `separation < INTERACTION_RADIUS` (== the message radius)
    // itself, so it may be shown only the messages within the radius of its search origin
    in3.setMessageInputRadiusFiltered(true);
#endif
    model.newLayer().addAgentFunction(boids3d_output);
    model.newLayer().addAgentFunction(boids3d_input);
  } else {
    flamegpu::MessageSpatial2D::Description message = model.newMessage<flamegpu::MessageSpatial2D>("location");
    message.setRadius(p.interaction_radius);
    message.setMin(p.min_position, p.min_position);
    message.setMax(p.max_position, p.max_position);
    message.newVariable<flamegpu::id_t>("id");
    message.newVariable<float>("fx");
    message.newVariable<float>("fy");
    agent.newFunction("outputdata", boids2d_output).setMessageOutput("location");
    flamegpu::AgentFunctionDescription in2 = agent.newFunction("inputdata", boids2d_input);
    in2.setMessageInput("location");
#ifdef FLAMEGPU2_B200
    in2.setMessageInputRadiusFiltered(true);
#endif
    model.newLayer().addAgentFunction(boids2d_output);
    model.newLayer().addAgentFunction(boids2d_input);
  }
}

}  // namespace fgb_examples
