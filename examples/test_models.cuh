// examples/test_models.cuh -- small models that restate the reference's own black-box tests of the
// hot path (tests/test_cases/runtime/messaging/test_spatial_{2,3}d.cu, simulation/test_cuda_simulation.cu,
// runtime/agent/test_device_agent_creation.cu, runtime/agent/detail/test_spatial_agent_sort.cu,
// runtime/messaging/test_bucket.cu) so the
// same scenario can run on the reference build and on this repo and be compared.
// FLAME GPU 2 API only.
#pragma once
#include "flamegpu/flamegpu.h"

namespace fgb_examples {

// ---- test_spatial_3d.cu:15-71 ---------------------------------------------------------------
FLAMEGPU_AGENT_FUNCTION(t_out3d, flamegpu::MessageNone, flamegpu::MessageSpatial3D) {
  FLAMEGPU->message_out.setVariable<flamegpu::id_t>("id", FLAMEGPU->getID());
  FLAMEGPU->message_out.setLocation(FLAMEGPU->getVariable<float>("x"), FLAMEGPU->getVariable<float>("y"),
                                    FLAMEGPU->getVariable<float>("z"));
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_out3d_optional, flamegpu::MessageNone, flamegpu::MessageSpatial3D) {
  if (FLAMEGPU->getVariable<int>("do_output")) {
    FLAMEGPU->message_out.setVariable<flamegpu::id_t>("id", FLAMEGPU->getID());
    FLAMEGPU->message_out.setLocation(FLAMEGPU->getVariable<float>("x"), FLAMEGPU->getVariable<float>("y"),
                                      FLAMEGPU->getVariable<float>("z"));
  }
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_in3d, flamegpu::MessageSpatial3D, flamegpu::MessageNone) {
  const float x1 = FLAMEGPU->getVariable<float>("x");
  const float y1 = FLAMEGPU->getVariable<float>("y");
  const float z1 = FLAMEGPU->getVariable<float>("z");
  unsigned int count = 0, badCount = 0, idsum = 0;
  const int myBin[3] = {static_cast<int>(x1), static_cast<int>(y1), static_cast<int>(z1)};
  for (const auto &message : FLAMEGPU->message_in(x1, y1, z1)) {
    const int mb[3] = {static_cast<int>(message.getVariable<float>("x")), static_cast<int>(message.getVariable<float>("y")),
                       static_cast<int>(message.getVariable<float>("z"))};
    bool bad = false;
    for (unsigned int i = 0; i < 3; ++i) {
      const int d = myBin[i] - mb[i];
      if (d > 1 || d < -1) bad = true;
    }
    ++count;
    badCount += bad ? 1u : 0u;
    idsum += message.getVariable<flamegpu::id_t>("id");
  }
  FLAMEGPU->setVariable<unsigned int>("count", count);
  FLAMEGPU->setVariable<unsigned int>("badCount", badCount);
  FLAMEGPU->setVariable<unsigned int>("idsum", idsum);  // order-independent checksum of what was read
  return flamegpu::ALIVE;
}
// test_spatial_3d.cu:855-902 inWrapped3D
FLAMEGPU_AGENT_FUNCTION(t_in3d_wrap, flamegpu::MessageSpatial3D, flamegpu::MessageNone) {
  const float x1 = FLAMEGPU->getVariable<float>("x");
  const float y1 = FLAMEGPU->getVariable<float>("y");
  const float z1 = FLAMEGPU->getVariable<float>("z");
  const flamegpu::id_t ID = FLAMEGPU->getID();
  unsigned int count = 0, badCount = 0;
  float xSum = 0, ySum = 0, zSum = 0;
  for (const auto &message : FLAMEGPU->message_in.wrap(x1, y1, z1)) {
    const float x2 = message.getVirtualX(x1);
    const float y2 = message.getVirtualY(y1);
    const float z2 = message.getVirtualZ(z1);
    float x21 = x2 - x1;
    float y21 = y2 - y1;
    float z21 = z2 - z1;
    const float distance = sqrtf(x21 * x21 + y21 * y21 + z21 * z21);
    if (distance > FLAMEGPU->message_in.radius() || (fabsf(x21) != 2.0f && x2 != x1) || (fabsf(y21) != 2.0f && y2 != y1) ||
        (fabsf(z21) != 2.0f && z2 != z1)) {
      badCount++;
    } else {
      count++;
      if (message.getVariable<flamegpu::id_t>("id") != ID) {
        xSum += x21;
        ySum += y21;
        zSum += z21;
      }
    }
  }
  FLAMEGPU->setVariable<unsigned int>("count", count);
  FLAMEGPU->setVariable<unsigned int>("badCount", badCount);
  FLAMEGPU->setVariable<float>("result_x", xSum);
  FLAMEGPU->setVariable<float>("result_y", ySum);
  FLAMEGPU->setVariable<float>("result_z", zSum);
  return flamegpu::ALIVE;
}

// ---- 2D twins (test_spatial_2d.cu:15-64, wrapped :700-760) --------------------------------------
FLAMEGPU_AGENT_FUNCTION(t_out2d, flamegpu::MessageNone, flamegpu::MessageSpatial2D) {
  FLAMEGPU->message_out.setVariable<flamegpu::id_t>("id", FLAMEGPU->getID());
  FLAMEGPU->message_out.setLocation(FLAMEGPU->getVariable<float>("x"), FLAMEGPU->getVariable<float>("y"));
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_in2d, flamegpu::MessageSpatial2D, flamegpu::MessageNone) {
  const float x1 = FLAMEGPU->getVariable<float>("x");
  const float y1 = FLAMEGPU->getVariable<float>("y");
  unsigned int count = 0, badCount = 0, idsum = 0;
  const int myBin[2] = {static_cast<int>(x1), static_cast<int>(y1)};
  for (const auto &message : FLAMEGPU->message_in(x1, y1)) {
    const int mb[2] = {static_cast<int>(message.getVariable<float>("x")), static_cast<int>(message.getVariable<float>("y"))};
    bool bad = false;
    for (unsigned int i = 0; i < 2; ++i) {
      const int d = myBin[i] - mb[i];
      if (d > 1 || d < -1) bad = true;
    }
    ++count;
    badCount += bad ? 1u : 0u;
    idsum += message.getVariable<flamegpu::id_t>("id");
  }
  FLAMEGPU->setVariable<unsigned int>("count", count);
  FLAMEGPU->setVariable<unsigned int>("badCount", badCount);
  FLAMEGPU->setVariable<unsigned int>("idsum", idsum);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_in2d_wrap, flamegpu::MessageSpatial2D, flamegpu::MessageNone) {
  const float x1 = FLAMEGPU->getVariable<float>("x");
  const float y1 = FLAMEGPU->getVariable<float>("y");
  const flamegpu::id_t ID = FLAMEGPU->getID();
  unsigned int count = 0, badCount = 0;
  float xSum = 0, ySum = 0;
  for (const auto &message : FLAMEGPU->message_in.wrap(x1, y1)) {
    const float x2 = message.getVirtualX(x1);
    const float y2 = message.getVirtualY(y1);
    float x21 = x2 - x1;
    float y21 = y2 - y1;
    const float distance = sqrtf(x21 * x21 + y21 * y21);
    if (distance > FLAMEGPU->message_in.radius() || (fabsf(x21) != 2.0f && x2 != x1) || (fabsf(y21) != 2.0f && y2 != y1)) {
      badCount++;
    } else {
      count++;
      if (message.getVariable<flamegpu::id_t>("id") != ID) {
        xSum += x21;
        ySum += y21;
      }
    }
  }
  FLAMEGPU->setVariable<unsigned int>("count", count);
  FLAMEGPU->setVariable<unsigned int>("badCount", badCount);
  FLAMEGPU->setVariable<float>("result_x", xSum);
  FLAMEGPU->setVariable<float>("result_y", ySum);
  return flamegpu::ALIVE;
}

// ---- test_cuda_simulation.cu:406-430 AgentDeath ---------------------------------------------
FLAMEGPU_AGENT_FUNCTION(t_death, flamegpu::MessageNone, flamegpu::MessageNone) {
  const unsigned int x = FLAMEGPU->getVariable<unsigned int>("x");
  FLAMEGPU->setVariable<unsigned int>("x", x + 12);  // reference DeathTestFunc also mutates before dying/surviving
  return (x % 2 == 0) ? flamegpu::DEAD : flamegpu::ALIVE;
}

// ---- test_device_agent_creation.cu:18-53 ------------------------------------------------------
FLAMEGPU_AGENT_FUNCTION(t_birth_mandatory, flamegpu::MessageNone, flamegpu::MessageNone) {
  const unsigned int id = FLAMEGPU->getVariable<unsigned int>("id") + 1;
  FLAMEGPU->agent_out.setVariable<float>("x", id + 12.0f);
  FLAMEGPU->agent_out.setVariable<unsigned int>("id", id);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_birth_optional, flamegpu::MessageNone, flamegpu::MessageNone) {
  const unsigned int id = FLAMEGPU->getVariable<unsigned int>("id") + 1;
  if (id % 2 == 1) {  // the reference keys this on threadIdx.x % 2; the id is thread-order independent
    FLAMEGPU->agent_out.setVariable<float>("x", id + 12.0f);
    FLAMEGPU->agent_out.setVariable<unsigned int>("id", id);
  }
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_birth_optional_death, flamegpu::MessageNone, flamegpu::MessageNone) {
  const unsigned int id = FLAMEGPU->getVariable<unsigned int>("id") + 1;
  if (id % 2 == 1) {
    FLAMEGPU->agent_out.setVariable<float>("x", id + 12.0f);
    FLAMEGPU->agent_out.setVariable<unsigned int>("id", id);
  } else {
    return flamegpu::DEAD;
  }
  return flamegpu::ALIVE;
}

// ---- test_agent_function_conditions.cu:24-47 ---------------------------------------------------
FLAMEGPU_AGENT_FUNCTION(t_cond_fn1, flamegpu::MessageNone, flamegpu::MessageNone) {
  FLAMEGPU->setVariable<int>("x", FLAMEGPU->getVariable<int>("x") + 1);
  FLAMEGPU->setVariable<int, 4>("y", 0, 3);
  FLAMEGPU->setVariable<int, 4>("y", 1, 4);
  FLAMEGPU->setVariable<int, 4>("y", 2, 5);
  FLAMEGPU->setVariable<int, 4>("y", 3, 6);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_cond_fn2, flamegpu::MessageNone, flamegpu::MessageNone) {
  FLAMEGPU->setVariable<int>("x", FLAMEGPU->getVariable<int>("x") - 1);
  FLAMEGPU->setVariable<int, 4>("y", 0, 23);
  FLAMEGPU->setVariable<int, 4>("y", 1, 24);
  FLAMEGPU->setVariable<int, 4>("y", 2, 25);
  FLAMEGPU->setVariable<int, 4>("y", 3, 26);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION_CONDITION(t_cond_is1) { return FLAMEGPU->getVariable<int>("x") == 1; }
FLAMEGPU_AGENT_FUNCTION_CONDITION(t_cond_not1) { return FLAMEGPU->getVariable<int>("x") != 1; }
// same-state condition + death: only agents with x % 3 == 0 run; of those, even x die
FLAMEGPU_AGENT_FUNCTION(t_cond_death_fn, flamegpu::MessageNone, flamegpu::MessageNone) {
  const int x = FLAMEGPU->getVariable<int>("x");
  FLAMEGPU->setVariable<int>("x", x + 1000);
  return (x % 2 == 0) ? flamegpu::DEAD : flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION_CONDITION(t_cond_mod3) { return FLAMEGPU->getVariable<int>("x") % 3 == 0; }

// ---- test_bucket.cu:54-98 (out_mandatory, out_optional, in, in_range), key base 12 as in the reference test ----
FLAMEGPU_AGENT_FUNCTION(t_bucket_out, flamegpu::MessageNone, flamegpu::MessageBucket) {
  const int id = FLAMEGPU->getVariable<int>("id");
  FLAMEGPU->message_out.setVariable<int>("id", id);
  FLAMEGPU->message_out.setKey(12 + (id / 2));
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_bucket_out_optional, flamegpu::MessageNone, flamegpu::MessageBucket) {
  if (FLAMEGPU->getVariable<int>("do_output")) {
    const int id = FLAMEGPU->getVariable<int>("id");
    FLAMEGPU->message_out.setVariable<int>("id", id);
    FLAMEGPU->message_out.setKey(12 + (id / 2));
  }
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_bucket_in, flamegpu::MessageBucket, flamegpu::MessageNone) {
  const int id = FLAMEGPU->getVariable<int>("id");
  const int id_m1 = id == 0 ? 0 : id - 1;
  unsigned int count = 0, sum = 0;
  for (auto &m : FLAMEGPU->message_in(12 + (id_m1 / 2))) {
    count++;
    sum += m.getVariable<int>("id");
  }
  FLAMEGPU->setVariable<unsigned int>("count1", count);
  FLAMEGPU->setVariable<unsigned int>("count2", FLAMEGPU->message_in(12 + (id_m1 / 2)).size());
  FLAMEGPU->setVariable<unsigned int>("sum", sum);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_bucket_in_range, flamegpu::MessageBucket, flamegpu::MessageNone) {
  const int id = FLAMEGPU->getVariable<int>("id");
  const int id_m4 = 12 + ((id / 8) * 4);
  unsigned int count = 0, sum = 0;
  for (auto &m : FLAMEGPU->message_in(id_m4, id_m4 + 4)) {
    count++;
    sum += m.getVariable<int>("id");
  }
  FLAMEGPU->setVariable<unsigned int>("count1", count);
  FLAMEGPU->setVariable<unsigned int>("count2", FLAMEGPU->message_in(12 + id / 2).size());
  FLAMEGPU->setVariable<unsigned int>("sum", sum);
  return flamegpu::ALIVE;
}

enum TestModel {
  TM_COUNT3D = 0,        // Spatial3DMessageTest.Mandatory
  TM_OPTIONAL3D = 1,     // Spatial3DMessageTest.Optional
  TM_WRAP3D = 2,         // Spatial3DMessageTest.Wrapped*
  TM_COUNT2D = 3,        // Spatial2DMessageTest.Mandatory
  TM_WRAP2D = 4,         // Spatial2DMessageTest.Wrapped
  TM_DEATH = 5,          // TestCUDASimulation.AgentDeath
  TM_BIRTH_MANDATORY = 6,       // DeviceAgentCreationTest.Mandatory_Output_SameState
  TM_BIRTH_OPTIONAL = 7,        // DeviceAgentCreationTest.Optional_Output_SameState
  TM_BIRTH_OPTIONAL_DEATH = 8,  // DeviceAgentCreationTest.Optional_Output_SameState_WithDeath
  TM_BIRTH_OTHER_AGENT = 9,     // DeviceAgentCreationTest.Mandatory_Output_DifferentAgent
  TM_CONDITION_SPLIT = 10,      // TestAgentFunctionConditions.SplitAgents
  TM_CONDITION_DEATH = 11,      // condition + death in the same state (order: disabled front, then survivors)
  TM_BUCKET = 12,               // BucketMessageTest.Mandatory
  TM_BUCKET_OPTIONAL = 13,      // BucketMessageTest.Optional / OptionalNone (do_output all zero)
  TM_BUCKET_RANGE = 14          // BucketMessageTest.Mandatory_Range
};

struct TestParams {
  int which = TM_COUNT3D;
  float mn[3] = {0, 0, 0};
  float mx[3] = {5, 5, 5};
  float radius = 1.0f;
  unsigned int sort_period = 1;
  int bucket_upper = 12 + 512;  // bucket models: bounds (12, bucket_upper), reference test: 12 + AGENT_COUNT / 2
};

inline void define_test_model(flamegpu::ModelDescription &model, const TestParams &p) {
  const bool is3d = p.which == TM_COUNT3D || p.which == TM_OPTIONAL3D || p.which == TM_WRAP3D;
  const bool is2d = p.which == TM_COUNT2D || p.which == TM_WRAP2D;
  if (is3d) {
    flamegpu::MessageSpatial3D::Description message = model.newMessage<flamegpu::MessageSpatial3D>("location");
    message.setMin(p.mn[0], p.mn[1], p.mn[2]);
    message.setMax(p.mx[0], p.mx[1], p.mx[2]);
    message.setRadius(p.radius);
    message.newVariable<flamegpu::id_t>("id");
  } else if (is2d) {
    flamegpu::MessageSpatial2D::Description message = model.newMessage<flamegpu::MessageSpatial2D>("location");
    message.setMin(p.mn[0], p.mn[1]);
    message.setMax(p.mx[0], p.mx[1]);
    message.setRadius(p.radius);
    message.newVariable<flamegpu::id_t>("id");
  }
  const bool isbucket = p.which == TM_BUCKET || p.which == TM_BUCKET_OPTIONAL || p.which == TM_BUCKET_RANGE;
  if (isbucket) {
    flamegpu::MessageBucket::Description message = model.newMessage<flamegpu::MessageBucket>("bucket");
    message.setBounds(12, p.bucket_upper);  // non-zero lower bound, as the reference test
    message.newVariable<int>("id");
  }
  flamegpu::AgentDescription agent = model.newAgent("agent");
  if (isbucket) {
    agent.newVariable<int>("id");
    agent.newVariable<int>("do_output", 1);
    agent.newVariable<unsigned int>("count1", 0);
    agent.newVariable<unsigned int>("count2", 0);
    agent.newVariable<unsigned int>("sum", 0);
  }
  if (is3d || is2d) {
    agent.newVariable<float>("x");
    agent.newVariable<float>("y");
    if (is3d) agent.newVariable<float>("z");
    agent.newVariable<unsigned int>("count");
    agent.newVariable<unsigned int>("badCount");
    agent.newVariable<unsigned int>("idsum");
    agent.newVariable<int>("do_output", 1);
    agent.newVariable<float>("result_x");
    agent.newVariable<float>("result_y");
    agent.newVariable<float>("result_z");
    agent.setSortPeriod(p.sort_period);
  }
  switch (p.which) {
    case TM_COUNT3D:
      agent.newFunction("out", t_out3d).setMessageOutput("location");
      agent.newFunction("in", t_in3d).setMessageInput("location");
      model.newLayer().addAgentFunction(t_out3d);
      model.newLayer().addAgentFunction(t_in3d);
      break;
    case TM_OPTIONAL3D: {
      flamegpu::AgentFunctionDescription af = agent.newFunction("out", t_out3d_optional);
      af.setMessageOutput("location");
      af.setMessageOutputOptional(true);
      agent.newFunction("in", t_in3d).setMessageInput("location");
      model.newLayer().addAgentFunction(t_out3d_optional);
      model.newLayer().addAgentFunction(t_in3d);
      break;
    }
    case TM_WRAP3D:
      agent.newFunction("out", t_out3d).setMessageOutput("location");
      agent.newFunction("in", t_in3d_wrap).setMessageInput("location");
      model.newLayer().addAgentFunction(t_out3d);
      model.newLayer().addAgentFunction(t_in3d_wrap);
      break;
    case TM_COUNT2D:
      agent.newFunction("out", t_out2d).setMessageOutput("location");
      agent.newFunction("in", t_in2d).setMessageInput("location");
      model.newLayer().addAgentFunction(t_out2d);
      model.newLayer().addAgentFunction(t_in2d);
      break;
    case TM_WRAP2D:
      agent.newFunction("out", t_out2d).setMessageOutput("location");
      agent.newFunction("in", t_in2d_wrap).setMessageInput("location");
      model.newLayer().addAgentFunction(t_out2d);
      model.newLayer().addAgentFunction(t_in2d_wrap);
      break;
    case TM_DEATH:
      agent.newVariable<unsigned int>("x");
      agent.newVariable<unsigned int, 3>("arr");  // AgentDeath_array (test_device_api.cu:15-58)
      agent.newFunction("DeathFunc", t_death).setAllowAgentDeath(true);
      model.newLayer().addAgentFunction(t_death);
      break;
    case TM_BIRTH_MANDATORY:
    case TM_BIRTH_OPTIONAL:
    case TM_BIRTH_OPTIONAL_DEATH: {
      agent.newVariable<float>("x");
      agent.newVariable<unsigned int>("id");
      agent.newVariable<float>("untouched", 15.0f);  // default-value check (test_device_agent_creation.cu:604)
      if (p.which == TM_BIRTH_MANDATORY) {
        flamegpu::AgentFunctionDescription f = agent.newFunction("output", t_birth_mandatory);
        f.setAgentOutput(agent);
        model.newLayer().addAgentFunction(f);
      } else if (p.which == TM_BIRTH_OPTIONAL) {
        flamegpu::AgentFunctionDescription f = agent.newFunction("output", t_birth_optional);
        f.setAgentOutput(agent);
        model.newLayer().addAgentFunction(f);
      } else {
        flamegpu::AgentFunctionDescription f = agent.newFunction("output", t_birth_optional_death);
        f.setAgentOutput(agent);
        f.setAllowAgentDeath(true);
        model.newLayer().addAgentFunction(f);
      }
      break;
    }
    case TM_BIRTH_OTHER_AGENT: {
      agent.newVariable<float>("x");
      agent.newVariable<unsigned int>("id");
      flamegpu::AgentDescription agent2 = model.newAgent("agent2");
      agent2.newVariable<float>("x");
      agent2.newVariable<unsigned int>("id");
      agent2.newVariable<float>("untouched", 15.0f);
      flamegpu::AgentFunctionDescription f = agent.newFunction("output", t_birth_mandatory);
      f.setAgentOutput(agent2);
      model.newLayer().addAgentFunction(f);
      break;
    }
    case TM_CONDITION_SPLIT: {
      agent.newVariable<int>("x");
      agent.newVariable<int, 4>("y");
      agent.newState("Start");
      agent.newState("End");
      agent.newState("End2");
      flamegpu::AgentFunctionDescription af1 = agent.newFunction("Function1", t_cond_fn1);
      af1.setInitialState("Start");
      af1.setEndState("End");
      af1.setFunctionCondition(t_cond_is1);
      flamegpu::AgentFunctionDescription af2 = agent.newFunction("Function2", t_cond_fn2);
      af2.setInitialState("Start");
      af2.setEndState("End2");
      af2.setFunctionCondition(t_cond_not1);
      model.newLayer().addAgentFunction(af1);
      model.newLayer().addAgentFunction(af2);
      break;
    }
    case TM_CONDITION_DEATH: {
      agent.newVariable<int>("x");
      flamegpu::AgentFunctionDescription af = agent.newFunction("Function1", t_cond_death_fn);
      af.setFunctionCondition(t_cond_mod3);
      af.setAllowAgentDeath(true);
      model.newLayer().addAgentFunction(af);
      break;
    }
    case TM_BUCKET:
    case TM_BUCKET_RANGE:
      agent.newFunction("out", t_bucket_out).setMessageOutput("bucket");
      if (p.which == TM_BUCKET) {
        agent.newFunction("in", t_bucket_in).setMessageInput("bucket");
        model.newLayer().addAgentFunction(t_bucket_out);
        model.newLayer().addAgentFunction(t_bucket_in);
      } else {
        agent.newFunction("in", t_bucket_in_range).setMessageInput("bucket");
        model.newLayer().addAgentFunction(t_bucket_out);
        model.newLayer().addAgentFunction(t_bucket_in_range);
      }
      break;
    case TM_BUCKET_OPTIONAL: {
      flamegpu::AgentFunctionDescription af = agent.newFunction("out", t_bucket_out_optional);
      af.setMessageOutput("bucket");
      af.setMessageOutputOptional(true);
      agent.newFunction("in", t_bucket_in).setMessageInput("bucket");
      model.newLayer().addAgentFunction(t_bucket_out_optional);
      model.newLayer().addAgentFunction(t_bucket_in);
      break;
    }
    default:
      break;
  }
}

}  // namespace fgb_examples
