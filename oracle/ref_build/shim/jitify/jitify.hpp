// Oracle scaffolding (NOT product code): the smallest jitify2 surface that lets the
// reference's non-RTC translation units compile offline.  The real Jitify2 is an
// un-vendored FetchContent dependency (cmake/dependencies/Jitify.cmake:12-13) and
// there is no network here.  Every entry point reports "rtc disabled".
#pragma once
#include <cuda.h>
#include <memory>
#include <string>
#include <vector>
namespace jitify2 {
using ErrorMsg = std::string;
struct LoadedProgramDataStub {
  ErrorMsg get_global_ptr(const char*, CUdeviceptr*) const { return "rtc disabled"; }
};
struct ConfiguredKernelStub {
  ErrorMsg launch_raw(const std::vector<void*>&) const { return "rtc disabled"; }
  ConfiguredKernelStub* operator->() { return this; }
  const ConfiguredKernelStub* operator->() const { return this; }
};
class KernelData {
 public:
  CUfunction function() const { return nullptr; }
  ConfiguredKernelStub configure(int, int, unsigned int = 0, CUstream_st* = nullptr) const { return {}; }
  const LoadedProgramDataStub& program() const {
    static LoadedProgramDataStub s;
    return s;
  }
};
class LinkedProgramData {};
}  // namespace jitify2
