// flamegpu/defines.h -- basic types of the API layer (same names/values as the reference's
// include/flamegpu/defines.h:8-36 so that user code is source compatible).
#ifndef FGB_INCLUDE_FLAMEGPU_DEFINES_H_
#define FGB_INCLUDE_FLAMEGPU_DEFINES_H_

#include <cstdint>

// Defined by THIS implementation of the FLAME GPU 2 API: model sources shared with the reference build use it to
// guard the few b200 extension calls (e.g. AgentFunctionDescription::setMessageInputRadiusFiltered).
#define FLAMEGPU2_B200 1

namespace flamegpu {

typedef unsigned int id_t;         // reference defines.h:8
typedef unsigned int size_type;    // reference defines.h:21
constexpr id_t ID_NOT_SET = 0;     // reference defines.h:12
constexpr const char *ID_VARIABLE_NAME = "_id";
constexpr const char *DEFAULT_STATE = "default";

// reference include/flamegpu/runtime/AgentFunction.cuh:19
enum AGENT_STATUS { ALIVE = 1, DEAD = 0 };

// B200-native extension knobs (no reference counterpart); see DESIGN.md
namespace b200 {
constexpr int kMaxVars = 24;       // variables per device lookup table
constexpr int kMaxEnvProps = 32;   // environment properties visible on the device
}  // namespace b200

}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_DEFINES_H_
