// Oracle scaffolding (NOT product code): definitions for the ten symbols the reference
// library leaves unresolved when JitifyCache.cu, the three XML TUs and the generated
// version.cpp are left out of an offline build.
#include <stdexcept>
namespace tinyxml2 { class XMLDocument {}; class XMLNode {}; class XMLElement {}; }
#include "flamegpu/version.h"
#include "flamegpu/detail/JitifyCache.h"
#include "flamegpu/io/XMLStateReader.h"
#include "flamegpu/io/XMLStateWriter.h"
#include "flamegpu/io/XMLLogger.h"
#include "jitify/jitify.hpp"
namespace flamegpu {
const char VERSION_BUILDMETADATA[] = {"oracle"};
const char VERSION_STRING[] = {"2.0.0-rc.5"};
const char VERSION_FULL[] = {"2.0.0-rc.5+oracle"};
const char TELEMETRY_RANDOM_ID[] = {"oracle"};
namespace detail {
JitifyCache& JitifyCache::getInstance() { throw std::runtime_error("RTC disabled in oracle build"); }
void JitifyCache::clearDiskCache() {}
std::unique_ptr<jitify2::KernelData> JitifyCache::loadKernel(const std::string&, const std::vector<std::string>&, const std::string&, const std::string&) { throw std::runtime_error("RTC disabled in oracle build"); }
}  // namespace detail
namespace io {
void XMLStateReader::parse(const std::string&, const std::shared_ptr<const ModelData>&, Verbosity) { throw std::runtime_error("XML disabled in oracle build"); }
XMLStateWriter::XMLStateWriter() { throw std::runtime_error("XML disabled in oracle build"); }
XMLLogger::XMLLogger(const std::string&, bool, bool) { throw std::runtime_error("XML disabled in oracle build"); }
void XMLStateWriter::beginWrite(const std::string&, bool) {}
void XMLStateWriter::endWrite() {}
void XMLStateWriter::writeConfig(const Simulation*) {}
void XMLStateWriter::writeStats(unsigned int) {}
void XMLStateWriter::writeEnvironment(const std::shared_ptr<const detail::EnvironmentManager>&) {}
void XMLStateWriter::writeMacroEnvironment(const std::shared_ptr<const detail::CUDAMacroEnvironment>&, std::initializer_list<std::string>) {}
void XMLStateWriter::writeAgents(const util::StringPairUnorderedMap<std::shared_ptr<const AgentVector>>&) {}
void XMLLogger::log(const RunLog&, const RunPlan&, bool, bool, bool, bool) const {}
void XMLLogger::log(const RunLog&, bool, bool, bool, bool, bool) const {}
}  // namespace io
}  // namespace flamegpu
