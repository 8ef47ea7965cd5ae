// examples/test_models.cuh -- small models that restate the reference's own black-box tests of the
// hot path (tests/test_cases/runtime/messaging/test_spatial_{2,3}d.cu, simulation/test_cuda_simulation.cu,
// runtime/agent/test_device_agent_creation.cu, runtime/agent/detail/test_spatial_agent_sort.cu,
// runtime/messaging/test_bucket.cu) so the
// same scenario can run on the reference build and on this repo and be compared.
// FLAME GPU 2 API only.
#pragma once
#include <string>

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


// ---- test_device_agent_creation.cu:37-52 (WithDeath variants; id parity stands in for threadIdx.x parity) -------
FLAMEGPU_AGENT_FUNCTION(t_birth_mandatory_death, flamegpu::MessageNone, flamegpu::MessageNone) {
  const unsigned int id = FLAMEGPU->getVariable<unsigned int>("id") + 1;
  FLAMEGPU->agent_out.setVariable<float>("x", id + 12.0f);
  FLAMEGPU->agent_out.setVariable<unsigned int>("id", id);
  return flamegpu::DEAD;
}
// test_device_agent_creation.cu:654-656 EvenThreadsOnlyCdn selects half of the threads; here half of the agents by
// id pairs, so that the optional functions (which key on id parity) still birth from half of the EXECUTING agents
FLAMEGPU_AGENT_FUNCTION_CONDITION(t_birth_even_cdn) { return (FLAMEGPU->getVariable<unsigned int>("id") / 2) % 2 == 0; }

// ---- test_device_agent_creation.cu:1196-1204 (AgentID_MultipleStatesUniqueIDs) ---------------------------------
FLAMEGPU_AGENT_FUNCTION(t_copy_id, flamegpu::MessageNone, flamegpu::MessageNone) {
  FLAMEGPU->setVariable<flamegpu::id_t>("id_copy", FLAMEGPU->getID());
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_device_birth_ids, flamegpu::MessageNone, flamegpu::MessageNone) {
  FLAMEGPU->setVariable<flamegpu::id_t>("id_other", FLAMEGPU->agent_out.getID());
  FLAMEGPU->agent_out.setVariable<flamegpu::id_t>("id_other", FLAMEGPU->getID());
  return flamegpu::ALIVE;
}

// ---- test_append_truncate.cu:22-71,272-296 (brute-force list written by two functions of one step) -------------
FLAMEGPU_AGENT_FUNCTION(t_app_out0, flamegpu::MessageNone, flamegpu::MessageBruteForce) {
  FLAMEGPU->message_out.setVariable<int>("x", 0);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_app_out1, flamegpu::MessageNone, flamegpu::MessageBruteForce) {
  FLAMEGPU->message_out.setVariable<int>("x", 1);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_app_out2, flamegpu::MessageNone, flamegpu::MessageBruteForce) {
  FLAMEGPU->message_out.setVariable<int>("x", 2);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_app_optout0, flamegpu::MessageNone, flamegpu::MessageBruteForce) {
  if (FLAMEGPU->getVariable<unsigned int>("do_out") > 0) {
    FLAMEGPU->message_out.setVariable<int>("x", 0);
    FLAMEGPU->setVariable<unsigned int>("do_out", 0);
  } else {
    FLAMEGPU->setVariable<unsigned int>("do_out", 1);
  }
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_app_optout1, flamegpu::MessageNone, flamegpu::MessageBruteForce) {
  if (FLAMEGPU->getVariable<unsigned int>("do_out") > 0) FLAMEGPU->message_out.setVariable<int>("x", 1);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_app_in, flamegpu::MessageBruteForce, flamegpu::MessageNone) {
  unsigned int c0 = 0, c1 = 0, c2 = 0;
  for (const auto &message : FLAMEGPU->message_in) {
    const int x = message.getVariable<int>("x");
    if (x == 0) ++c0;
    else if (x == 1) ++c1;
    else ++c2;
  }
  FLAMEGPU->setVariable<unsigned int>("count0", c0);
  FLAMEGPU->setVariable<unsigned int>("count1", c1);
  FLAMEGPU->setVariable<unsigned int>("count2", c2);
  return flamegpu::ALIVE;
}

// ---- runtime/agent/detail/test_agent_state_transition.cu:30-66 ---------------------------------------------------
FLAMEGPU_AGENT_FUNCTION(t_tr_good, flamegpu::MessageNone, flamegpu::MessageNone) {
  FLAMEGPU->setVariable<int>("x", 11);
  FLAMEGPU->setVariable<int, 4>("y", 0, 23);
  FLAMEGPU->setVariable<int, 4>("y", 1, 24);
  FLAMEGPU->setVariable<int, 4>("y", 2, 25);
  FLAMEGPU->setVariable<int, 4>("y", 3, 26);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_tr_bad, flamegpu::MessageNone, flamegpu::MessageNone) {
  FLAMEGPU->setVariable<int>("x", 13);
  FLAMEGPU->setVariable<int, 4>("y", 0, 3);
  FLAMEGPU->setVariable<int, 4>("y", 1, 4);
  FLAMEGPU->setVariable<int, 4>("y", 2, 5);
  FLAMEGPU->setVariable<int, 4>("y", 3, 6);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_tr_decrement, flamegpu::MessageNone, flamegpu::MessageNone) {
  const unsigned int x = FLAMEGPU->getVariable<unsigned int>("x");
  FLAMEGPU->setVariable<unsigned int>("x", x == 0 ? 0 : x - 1);
  FLAMEGPU->setVariable<int, 4>("z", 0, 23);
  FLAMEGPU->setVariable<int, 4>("z", 1, 24);
  FLAMEGPU->setVariable<int, 4>("z", 2, 25);
  FLAMEGPU->setVariable<int, 4>("z", 3, 26);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION(t_tr_null, flamegpu::MessageNone, flamegpu::MessageNone) {
  FLAMEGPU->setVariable<unsigned int>("x", 0xFFFFFFFFu);
  FLAMEGPU->setVariable<int, 4>("z", 0, 3);
  FLAMEGPU->setVariable<int, 4>("z", 1, 4);
  FLAMEGPU->setVariable<int, 4>("z", 2, 5);
  FLAMEGPU->setVariable<int, 4>("z", 3, 6);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION_CONDITION(t_tr_zero_x) { return FLAMEGPU->getVariable<unsigned int>("x") == 0; }
// two-way conditional transitions (ping-pong between two states): a multi-step model whose list bounds must stay bounded
FLAMEGPU_AGENT_FUNCTION(t_tr_touch, flamegpu::MessageNone, flamegpu::MessageNone) {
  FLAMEGPU->setVariable<unsigned int>("x", FLAMEGPU->getVariable<unsigned int>("x") + 1);
  return flamegpu::ALIVE;
}
FLAMEGPU_AGENT_FUNCTION_CONDITION(t_tr_step_parity) {
  return (FLAMEGPU->getVariable<unsigned int>("y") + FLAMEGPU->getStepCounter()) % 3 == 0;
}

// ---- test_host_agent_creation.cu:19-43 (BasicOutput from init and step functions, OutputMultiAgent) ---------------
FLAMEGPU_STEP_FUNCTION(t_host_basic_output) {
  auto t = FLAMEGPU->agent("agent");
  for (unsigned int i = 0; i < 512; ++i) t.newAgent().setVariable<float>("x", 1.0f);
}
FLAMEGPU_STEP_FUNCTION(t_host_multi_output) {
  auto t = FLAMEGPU->agent("agent", "b");
  auto t2 = FLAMEGPU->agent("agent2");
  for (unsigned int i = 0; i < 512; ++i) {
    t.newAgent().setVariable<float>("x", 1.0f);
    t2.newAgent().setVariable<float>("y", 2.0f);
  }
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
  TM_BUCKET_RANGE = 14,         // BucketMessageTest.Mandatory_Range
  TM_BIRTH_COMBO = 15,          // DeviceAgentCreationTest.{Mandatory,Optional}_Output_{SameState,DifferentState,DifferentAgent}
                                //   [_WithDeath][_WithAgentFunctionCondition]: TestParams::birth_*
  TM_APPEND = 16,               // TestMessage_AppendTruncate.Append_KeepData / OptionalAppend_KeepData (append_optional)
  TM_APPEND_RESIZE = 17,        // TestMessage_AppendTruncate.Append_KeepData_Resize (agents a, b, c)
  TM_TRANSITION_CHAIN = 18,     // TestAgentStateTransitions.Src_0_Dest_10 / Src_10_Dest_0
  TM_TRANSITION_COND = 19,      // TestAgentStateTransitions.Src_10_Dest_10 (conditional transition)
  TM_UNIQUE_IDS = 20,           // DeviceAgentCreationTest.AgentID_MultipleStatesUniqueIDs (two functions in one layer)
  TM_CONCURRENT_SPATIAL = 21,   // TestCUDASimulationConcurrency.ConcurrentMessageOutputInputSpatial3D (4 agent types)
  TM_TRANSITION_PINGPONG = 22,  // conditional transitions both ways, many steps (list bounds must stay bounded)
  TM_HOST_CREATION = 23         // HostAgentCreationTest.FromInit / FromStep / FromStepMultiAgent (host_init: 1 init, 0 step, 2 multi)
};

struct TestParams {
  int which = TM_COUNT3D;
  float mn[3] = {0, 0, 0};
  float mx[3] = {5, 5, 5};
  float radius = 1.0f;
  unsigned int sort_period = 1;
  int bucket_upper = 12 + 512;  // bucket models: bounds (12, bucket_upper), reference test: 12 + AGENT_COUNT / 2
  // TM_BIRTH_COMBO
  int birth_optional = 0, birth_death = 0, birth_condition = 0;
  int birth_target = 0;  // 0 same state, 1 different state, 2 different agent
  int append_optional = 0;  // TM_APPEND
  int host_init = 0;        // TM_HOST_CREATION
};

constexpr int kConcurrentAgents = 4;  // TM_CONCURRENT_SPATIAL: agent_0..3, location_0..3

inline void define_concurrent_spatial(flamegpu::ModelDescription &model, const TestParams &p);

inline void define_test_model(flamegpu::ModelDescription &model, const TestParams &p) {
  if (p.which == TM_CONCURRENT_SPATIAL) {
    define_concurrent_spatial(model, p);
    return;
  }
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
    case TM_BIRTH_COMBO: {
      agent.newVariable<float>("x");
      agent.newVariable<unsigned int>("id");
      agent.newVariable<float>("untouched", 15.0f);
      flamegpu::AgentFunctionDescription f =
          p.birth_optional ? (p.birth_death ? agent.newFunction("output", t_birth_optional_death) : agent.newFunction("output", t_birth_optional))
                           : (p.birth_death ? agent.newFunction("output", t_birth_mandatory_death) : agent.newFunction("output", t_birth_mandatory));
      if (p.birth_death) f.setAllowAgentDeath(true);
      if (p.birth_target == 2) {
        flamegpu::AgentDescription agent2 = model.newAgent("agent2");
        agent2.newVariable<float>("x");
        agent2.newVariable<unsigned int>("id");
        agent2.newVariable<float>("untouched", 15.0f);
        f.setAgentOutput(agent2);
      } else if (p.birth_condition && !p.birth_death) {
        // test_device_agent_creation.cu:657-668 (same state: a -> b, births into b), :779-793 (different: a -> c, births into b)
        agent.newState("a");
        agent.newState("b");
        if (p.birth_target == 1) agent.newState("c");
        f.setInitialState("a");
        f.setEndState(p.birth_target == 1 ? "c" : "b");
        f.setAgentOutput(agent, "b");
      } else if (p.birth_target == 1) {
        // :133-146, :299-312, :1003-1016: a -> a, births into b
        agent.newState("a");
        agent.newState("b");
        f.setInitialState("a");
        f.setEndState("a");
        f.setAgentOutput(agent, "b");
      } else {
        f.setAgentOutput(agent);
      }
      if (p.birth_condition) f.setFunctionCondition(t_birth_even_cdn);
      model.newLayer().addAgentFunction(f);
      break;
    }
    case TM_APPEND: {
      flamegpu::MessageBruteForce::Description message = model.newMessage("msg");
      message.newVariable<int>("x");
      agent.newVariable<unsigned int>("count0", 0);
      agent.newVariable<unsigned int>("count1", 0);
      agent.newVariable<unsigned int>("count2", 0);
      agent.newVariable<unsigned int>("do_out", 0);
      flamegpu::AgentFunctionDescription fo = p.append_optional ? agent.newFunction("out", t_app_optout0) : agent.newFunction("out", t_app_out0);
      flamegpu::AgentFunctionDescription fo2 = p.append_optional ? agent.newFunction("out2", t_app_optout1) : agent.newFunction("out2", t_app_out1);
      fo.setMessageOutput("msg");
      fo2.setMessageOutput("msg");
      if (p.append_optional) {
        fo.setMessageOutputOptional(true);
        fo2.setMessageOutputOptional(true);
      }
      flamegpu::AgentFunctionDescription fi = agent.newFunction("in", t_app_in);
      fi.setMessageInput("msg");
      model.newLayer().addAgentFunction(fo);
      model.newLayer().addAgentFunction(fo2);
      model.newLayer().addAgentFunction(fi);
      break;
    }
    case TM_APPEND_RESIZE: {
      flamegpu::MessageBruteForce::Description message = model.newMessage("msg");
      message.newVariable<int>("x");
      flamegpu::AgentFunctionDescription f1 = agent.newFunction("Out_1", t_app_out1);
      f1.setMessageOutput("msg");
      flamegpu::AgentDescription b = model.newAgent("b");
      flamegpu::AgentFunctionDescription f2 = b.newFunction("Out_2", t_app_out2);
      f2.setMessageOutput("msg");
      flamegpu::AgentDescription c = model.newAgent("c");
      c.newVariable<unsigned int>("count0", 0);
      c.newVariable<unsigned int>("count1", 0);
      c.newVariable<unsigned int>("count2", 0);
      flamegpu::AgentFunctionDescription fi = c.newFunction("In", t_app_in);
      fi.setMessageInput("msg");
      model.newLayer().addAgentFunction(f1);
      model.newLayer().addAgentFunction(f2);
      model.newLayer().addAgentFunction(fi);
      break;
    }
    case TM_TRANSITION_CHAIN: {
      agent.newState("Start");
      agent.newState("End");
      agent.newState("End2");
      agent.setInitialState("Start");
      agent.newVariable<int>("x");
      agent.newVariable<int, 4>("y");
      flamegpu::AgentFunctionDescription af1 = agent.newFunction("Function1", t_tr_good);
      af1.setInitialState("Start");
      af1.setEndState("End");
      flamegpu::AgentFunctionDescription af2 = agent.newFunction("Function2", t_tr_bad);
      af2.setInitialState("End");
      af2.setEndState("End2");
      model.newLayer().addAgentFunction(af2);  // reference order: End -> End2 first, Start -> End second
      model.newLayer().addAgentFunction(af1);
      break;
    }
    case TM_TRANSITION_COND: {
      agent.newState("Start");
      agent.newState("End");
      agent.setInitialState("Start");
      agent.newVariable<unsigned int>("x");
      agent.newVariable<unsigned int>("y");
      agent.newVariable<int, 4>("z");
      flamegpu::AgentFunctionDescription af1 = agent.newFunction("Function1", t_tr_decrement);
      af1.setInitialState("Start");
      af1.setEndState("Start");
      flamegpu::AgentFunctionDescription af2 = agent.newFunction("Function2", t_tr_null);
      af2.setInitialState("Start");
      af2.setEndState("End");
      af2.setFunctionCondition(t_tr_zero_x);
      model.newLayer().addAgentFunction(af1);
      model.newLayer().addAgentFunction(af2);
      break;
    }
    case TM_TRANSITION_PINGPONG: {
      agent.newState("A");
      agent.newState("B");
      agent.setInitialState("A");
      agent.newVariable<unsigned int>("x");
      agent.newVariable<unsigned int>("y");
      flamegpu::AgentFunctionDescription ab = agent.newFunction("AtoB", t_tr_touch);
      ab.setInitialState("A");
      ab.setEndState("B");
      ab.setFunctionCondition(t_tr_step_parity);
      flamegpu::AgentFunctionDescription ba = agent.newFunction("BtoA", t_tr_touch);
      ba.setInitialState("B");
      ba.setEndState("A");
      ba.setFunctionCondition(t_tr_step_parity);
      model.newLayer().addAgentFunction(ab);
      model.newLayer().addAgentFunction(ba);
      break;
    }
    case TM_HOST_CREATION: {
      agent.newVariable<float>("x");
      agent.newVariable<float>("default", 15.0f);  // test_host_agent_creation.cu DefaultVariableValue
      if (p.host_init == 2) {
        agent.newState("a");
        agent.newState("b");
        flamegpu::AgentDescription agent2 = model.newAgent("agent2");
        agent2.newVariable<float>("y");
        model.addStepFunction(t_host_multi_output);
      } else if (p.host_init == 1) {
        model.addInitFunction(t_host_basic_output);
      } else {
        model.addStepFunction(t_host_basic_output);
      }
      break;
    }
    case TM_UNIQUE_IDS: {
      agent.newVariable<flamegpu::id_t>("id_copy", flamegpu::ID_NOT_SET);
      agent.newVariable<flamegpu::id_t>("id_other", flamegpu::ID_NOT_SET);
      agent.newState("a");
      agent.newState("b");
      flamegpu::AgentFunctionDescription af1_a = agent.newFunction("birth", t_device_birth_ids);
      af1_a.setAgentOutput(agent, "a");
      af1_a.setInitialState("a");
      af1_a.setEndState("a");
      flamegpu::AgentFunctionDescription af1_b = agent.newFunction("birth2", t_device_birth_ids);
      af1_b.setAgentOutput(agent, "b");
      af1_b.setInitialState("b");
      af1_b.setEndState("b");
      flamegpu::AgentFunctionDescription af_a = agent.newFunction("copy_id", t_copy_id);
      af_a.setInitialState("a");
      af_a.setEndState("a");
      flamegpu::AgentFunctionDescription af_b = agent.newFunction("copy_id2", t_copy_id);
      af_b.setInitialState("b");
      af_b.setEndState("b");
      model.newLayer().addAgentFunction(af1_a);
      model.newLayer().addAgentFunction(af1_b);
      flamegpu::LayerDescription layer2 = model.newLayer();  // two functions of one layer: one stream each
      layer2.addAgentFunction(af_a);
      layer2.addAgentFunction(af_b);
      break;
    }
    default:
      break;
  }
}

// TestCUDASimulationConcurrency.ConcurrentMessageOutputInputSpatial3D (test_cuda_simulation_concurrency.cu:779-853):
// kConcurrentAgents agent types, each with its own Spatial3D list; all outputs share layer 0, all inputs layer 1.
inline void define_concurrent_spatial(flamegpu::ModelDescription &model, const TestParams &p) {
  flamegpu::LayerDescription layer0 = model.newLayer();
  flamegpu::LayerDescription layer1 = model.newLayer();
  for (int i = 0; i < kConcurrentAgents; ++i) {
    const std::string an = "agent_" + std::to_string(i), mn = "location_" + std::to_string(i);
    flamegpu::MessageSpatial3D::Description message = model.newMessage<flamegpu::MessageSpatial3D>(mn);
    message.setMin(p.mn[0], p.mn[1], p.mn[2]);
    message.setMax(p.mx[0], p.mx[1], p.mx[2]);
    message.setRadius(p.radius);
    message.newVariable<flamegpu::id_t>("id");
    flamegpu::AgentDescription a = model.newAgent(an);
    a.newVariable<float>("x");
    a.newVariable<float>("y");
    a.newVariable<float>("z");
    a.newVariable<unsigned int>("count");
    a.newVariable<unsigned int>("badCount");
    a.newVariable<unsigned int>("idsum");
    a.setSortPeriod(p.sort_period);
    flamegpu::AgentFunctionDescription fo = a.newFunction("out", t_out3d);
    fo.setMessageOutput(mn);
    flamegpu::AgentFunctionDescription fi = a.newFunction("in", t_in3d);
    fi.setMessageInput(mn);
    layer0.addAgentFunction(fo);
    layer1.addAgentFunction(fi);
  }
}

}  // namespace fgb_examples
