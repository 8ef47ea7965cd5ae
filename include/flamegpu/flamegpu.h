// flamegpu/flamegpu.h -- umbrella header of the B200-native hot-path API layer.
// User code written against the reference's "flamegpu/flamegpu.h" (models built from agents,
// MessageSpatial2D/3D / MessageBruteForce lists, agent functions with births / deaths / optional
// message output, layers, environment properties) compiles against this header unchanged.
#ifndef FGB_INCLUDE_FLAMEGPU_FLAMEGPU_H_
#define FGB_INCLUDE_FLAMEGPU_FLAMEGPU_H_

#include <random>

#include "flamegpu/defines.h"
#include "flamegpu/runtime/AgentFunction.cuh"
#include "flamegpu/model/ModelDescription.h"
#include "flamegpu/simulation/AgentVector.h"
#include "flamegpu/simulation/CUDASimulation.h"

#define FLAMEGPU_B200 1

namespace flamegpu {
namespace util {
inline void cleanup() { cudaDeviceSynchronize(); }
}  // namespace util
}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_FLAMEGPU_H_
