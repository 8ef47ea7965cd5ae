// flamegpu/model/ModelDescription.h -- host-side model description for the hot-path API layer.
//
// Source compatible with the subset of the reference's description API that the spatial hot path
// needs (include/flamegpu/model/{ModelDescription,AgentDescription,AgentFunctionDescription,
// LayerDescription,EnvironmentDescription}.h and the MessageSpatial2D/3D::Description classes):
// same class names, method names and argument meaning, so the reference's examples
// (examples/cpp/circles_spatial3D, boids_spatial3D) compile against either tree.
// Everything here is plain host metadata; no device work.  The reference's model layer itself is
// out of scope -- this is the smallest description surface that lets the
// per-step path run, not a re-implementation of it (no sub-models, dependency graph, RTC, I/O).
#ifndef FGB_INCLUDE_FLAMEGPU_MODEL_MODELDESCRIPTION_H_
#define FGB_INCLUDE_FLAMEGPU_MODEL_MODELDESCRIPTION_H_

#include <array>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <typeindex>
#include <vector>

#include "flamegpu/defines.h"
#include "flamegpu/runtime/AgentFunction.cuh"

namespace flamegpu {

namespace exception {
struct FLAMEGPUException : public std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define FGB_DEF_EXC(name)                       \
  struct name : public FLAMEGPUException {      \
    using FLAMEGPUException::FLAMEGPUException; \
  }
FGB_DEF_EXC(InvalidVarName);
FGB_DEF_EXC(InvalidVarType);
FGB_DEF_EXC(UnsupportedVarType);
FGB_DEF_EXC(InvalidAgentName);
FGB_DEF_EXC(InvalidAgentFunc);
FGB_DEF_EXC(InvalidAgentVar);
FGB_DEF_EXC(InvalidStateName);
FGB_DEF_EXC(InvalidMessageName);
FGB_DEF_EXC(InvalidMessageVar);
FGB_DEF_EXC(InvalidMessageType);
FGB_DEF_EXC(InvalidMessage);
FGB_DEF_EXC(ReservedName);
FGB_DEF_EXC(InvalidArgument);
FGB_DEF_EXC(InvalidEnvProperty);
FGB_DEF_EXC(InvalidLayerMember);
FGB_DEF_EXC(InvalidCudaAgent);
FGB_DEF_EXC(InvalidPopulationData);
FGB_DEF_EXC(OutOfBoundsException);
FGB_DEF_EXC(CUDAError);
FGB_DEF_EXC(UnsupportedFeature);
#undef FGB_DEF_EXC
}  // namespace exception

// One variable of an agent or message.  The reference keeps these in a std::map keyed by name
// (model/Variable.h:128), so every per-variable loop runs in alphabetical order; same here.
struct Variable {
  std::type_index type = std::type_index(typeid(void));
  size_t type_size = 0;
  unsigned int elements = 1;
  std::vector<char> default_value;
  size_t bytes() const { return type_size * elements; }
};
typedef std::map<std::string, Variable> VariableMap;

template <typename T>
inline Variable make_variable(unsigned int elements, const T *defaults) {
  Variable v;
  v.type = std::type_index(typeid(T));
  v.type_size = sizeof(T);
  v.elements = elements;
  v.default_value.resize(sizeof(T) * elements);
  if (defaults) std::memcpy(v.default_value.data(), defaults, sizeof(T) * elements);
  else std::memset(v.default_value.data(), 0, sizeof(T) * elements);
  return v;
}

struct AgentData;
struct ModelData;
class HostAPI;
typedef void (*HostFunctionPointer)(HostAPI *);
typedef bool (*HostConditionPointer)(HostAPI *);

enum class MessageKind { BruteForce, Spatial2D, Spatial3D, Bucket };

struct MessageData {
  std::string name;
  MessageKind kind = MessageKind::BruteForce;
  VariableMap variables;
  float radius = 0.f;
  float min[3] = {0.f, 0.f, 0.f};
  float max[3] = {0.f, 0.f, 0.f};
  bool min_set[3] = {false, false, false}, max_set[3] = {false, false, false};
  bool persistent = false;
  // bucket messaging (reference MessageBucket::Data, MessageBucketHost.h:125-130): upper == INT_MAX means "not set"
  int bucket_lower = 0;
  int bucket_upper = std::numeric_limits<int>::max();
  int dims() const { return kind == MessageKind::Spatial3D ? 3 : (kind == MessageKind::Spatial2D ? 2 : 0); }
};

struct AgentFunctionData {
  std::string name;
  AgentFunctionWrapper *func = nullptr;
  AgentFunctionWrapper *func_filtered = nullptr;  // b200: radius-filtered iterator variant (== func for non-spatial input)
  AgentFunctionConditionWrapper *condition = nullptr;
  std::type_index in_type = std::type_index(typeid(void));
  std::type_index out_type = std::type_index(typeid(void));
  std::string message_input, message_output;
  bool message_output_optional = false;
  // b200 extension: the function ignores input messages beyond the list's radius of its search origin (it tests the
  // distance itself), so the spatial iterator may skip them (radius-filtered lock-step walk, MessageSpatial3D.cuh)
  bool radius_filtered_input = false;
  std::string agent_output, agent_output_state;
  bool has_agent_death = false;
  std::string initial_state = DEFAULT_STATE, end_state = DEFAULT_STATE;
  std::weak_ptr<AgentData> parent;
};

struct AgentData {
  std::string name;
  VariableMap variables;
  std::set<std::string> states;
  std::string initial_state = DEFAULT_STATE;
  bool default_state_only = true;
  std::map<std::string, std::shared_ptr<AgentFunctionData>> functions;
  unsigned int sort_period = 1;  // reference model/AgentData.cpp:17
};

struct EnvProperty {
  std::type_index type = std::type_index(typeid(void));
  size_t type_size = 0;
  unsigned int elements = 1;
  bool is_const = false;
  std::vector<char> data;
};

struct LayerData {
  std::string name;
  std::vector<std::shared_ptr<AgentFunctionData>> functions;
  std::vector<HostFunctionPointer> host_functions;
};

struct ModelData {
  std::string name;
  std::map<std::string, std::shared_ptr<AgentData>> agents;
  std::map<std::string, std::shared_ptr<MessageData>> messages;
  std::map<std::string, EnvProperty> environment;
  std::vector<std::shared_ptr<LayerData>> layers;
  std::vector<HostFunctionPointer> init_functions, step_functions, exit_functions;
  std::vector<HostConditionPointer> exit_conditions;
};

// ---------------------------------------------------------------------------------------------
// messages
// ---------------------------------------------------------------------------------------------
namespace detail {
class MessageDescriptionBase {
 public:
  explicit MessageDescriptionBase(std::shared_ptr<MessageData> d) : data(std::move(d)) {}
  std::string getName() const { return data->name; }
  template <typename T>
  void newVariable(const std::string &name) {
    add<T>(name, 1);
  }
  template <typename T, flamegpu::size_type N>
  void newVariable(const std::string &name) {
    add<T>(name, N);
  }
  bool hasVariable(const std::string &name) const { return data->variables.count(name) != 0; }
  void setPersistent(bool p) { data->persistent = p; }
  bool getPersistent() const { return data->persistent; }
  std::shared_ptr<MessageData> data;

 protected:
  template <typename T>
  void add(const std::string &name, unsigned int n) {
    if (!name.empty() && name[0] == '_') throw exception::ReservedName("message variable names may not begin with '_'");
    if (name.empty()) throw exception::InvalidMessageVar("message variable names may not be empty");
    if (data->variables.count(name)) throw exception::InvalidMessageVar("message '" + data->name + "' already has variable '" + name + "'");
    data->variables.emplace(name, make_variable<T>(n, nullptr));
  }
};
}  // namespace detail

class MessageBruteForce::Description : public detail::MessageDescriptionBase {
 public:
  using detail::MessageDescriptionBase::MessageDescriptionBase;
  static MessageKind kind() { return MessageKind::BruteForce; }
};

// reference src/flamegpu/runtime/messaging/MessageSpatial2D.cu:148-320 (setters and validation)
class MessageSpatial2D::Description : public detail::MessageDescriptionBase {
 public:
  explicit Description(std::shared_ptr<MessageData> d) : detail::MessageDescriptionBase(std::move(d)) {
    if (!data->variables.count("x")) {
      data->variables.emplace("x", make_variable<float>(1, nullptr));
      data->variables.emplace("y", make_variable<float>(1, nullptr));
    }
  }
  static MessageKind kind() { return MessageKind::Spatial2D; }
  void setRadius(float r) {
    if (!(r > 0.f)) throw exception::InvalidArgument("Spatial messaging radius must be a positive value");
    data->radius = r;
  }
  void setMinX(float v) { set_min(0, v); }
  void setMinY(float v) { set_min(1, v); }
  void setMin(float x, float y) { set_min(0, x); set_min(1, y); }
  void setMaxX(float v) { set_max(0, v); }
  void setMaxY(float v) { set_max(1, v); }
  void setMax(float x, float y) { set_max(0, x); set_max(1, y); }
  float getRadius() const { return data->radius; }
  float getMinX() const { return data->min[0]; }
  float getMinY() const { return data->min[1]; }
  float getMaxX() const { return data->max[0]; }
  float getMaxY() const { return data->max[1]; }

 protected:
  void set_min(int a, float v) {
    if (data->max_set[a] && v >= data->max[a]) throw exception::InvalidArgument("Spatial messaging min bound must be lower than max bound");
    data->min[a] = v;
    data->min_set[a] = true;
  }
  void set_max(int a, float v) {
    if (data->min_set[a] && v <= data->min[a]) throw exception::InvalidArgument("Spatial messaging max bound must be greater than min bound");
    data->max[a] = v;
    data->max_set[a] = true;
  }
};

// reference src/flamegpu/runtime/messaging/MessageSpatial3D.cu:175-284
class MessageSpatial3D::Description : public MessageSpatial2D::Description {
 public:
  explicit Description(std::shared_ptr<MessageData> d) : MessageSpatial2D::Description(std::move(d)) {
    if (!data->variables.count("z")) data->variables.emplace("z", make_variable<float>(1, nullptr));
  }
  static MessageKind kind() { return MessageKind::Spatial3D; }
  void setMinZ(float v) { set_min(2, v); }
  void setMaxZ(float v) { set_max(2, v); }
  void setMin(float x, float y, float z) { set_min(0, x); set_min(1, y); set_min(2, z); }
  void setMax(float x, float y, float z) { set_max(0, x); set_max(1, y); set_max(2, z); }
  float getMinZ() const { return data->min[2]; }
  float getMaxZ() const { return data->max[2]; }
};

// reference src/flamegpu/runtime/messaging/MessageBucket.cu:196-224 (setters, validation, the "_key" variable)
class MessageBucket::Description : public detail::MessageDescriptionBase {
 public:
  explicit Description(std::shared_ptr<MessageData> d) : detail::MessageDescriptionBase(std::move(d)) {
    if (!data->variables.count("_key")) data->variables.emplace("_key", make_variable<IntT>(1, nullptr));
  }
  static MessageKind kind() { return MessageKind::Bucket; }
  void setLowerBound(IntT min) {
    if (data->bucket_upper != std::numeric_limits<IntT>::max() && min >= data->bucket_upper)
      throw exception::InvalidArgument("Bucket messaging minimum bound must be lower than upper bound");
    data->bucket_lower = min;
  }
  void setUpperBound(IntT max) {
    if (max <= data->bucket_lower) throw exception::InvalidArgument("Bucket messaging upperBound bound must be greater than lower bound");
    data->bucket_upper = max;
  }
  void setBounds(IntT min, IntT max) {
    if (max <= min) throw exception::InvalidArgument("Bucket messaging upperBound bound must be greater than lower bound");
    data->bucket_lower = min;
    data->bucket_upper = max;
  }
  IntT getLowerBound() const { return data->bucket_lower; }
  IntT getUpperBound() const { return data->bucket_upper; }
};

// ---------------------------------------------------------------------------------------------
// agents and agent functions
// ---------------------------------------------------------------------------------------------
class AgentDescription;

class AgentFunctionDescription {
 public:
  AgentFunctionDescription(std::shared_ptr<AgentFunctionData> f, std::shared_ptr<ModelData> m) : function(std::move(f)), model(std::move(m)) {}
  std::string getName() const { return function->name; }
  void setInitialState(const std::string &s) { check_state(s); function->initial_state = s; }
  void setEndState(const std::string &s) { check_state(s); function->end_state = s; }
  void setMessageInput(const std::string &message_name) {
    auto it = lookup_message(message_name);
    check_type(it, function->in_type, "input");
    if (function->message_output == message_name) throw exception::InvalidMessageName("message '" + message_name + "' is already the output of this function");
    function->message_input = message_name;
  }
  void setMessageInput(const char *message_name) { setMessageInput(std::string(message_name)); }
  template <typename Desc>
  void setMessageInput(const Desc &d) { setMessageInput(d.getName()); }
  void setMessageOutput(const std::string &message_name) {
    auto it = lookup_message(message_name);
    check_type(it, function->out_type, "output");
    if (function->message_input == message_name) throw exception::InvalidMessageName("message '" + message_name + "' is already the input of this function");
    function->message_output = message_name;
  }
  void setMessageOutput(const char *message_name) { setMessageOutput(std::string(message_name)); }
  template <typename Desc>
  void setMessageOutput(const Desc &d) { setMessageOutput(d.getName()); }
  void setMessageOutputOptional(bool optional) { function->message_output_optional = optional; }
  // b200 extension (no reference counterpart).  Declares the radius contract of this function's spatial input: every
  // message within the radius of the search origin is presented exactly once, in the reference's relative order;
  // messages farther away MAY be skipped and far-away padding messages may be presented.  Safe for any function
  // that tests the distance itself and neither breaks out of the loop nor counts rejected messages (Circles, Boids).
  // CUDAConfig().spatialIterationMode = 0 overrides it (strict reference visit order for every function).
  void setMessageInputRadiusFiltered(bool filtered) { function->radius_filtered_input = filtered; }
  bool getMessageInputRadiusFiltered() const { return function->radius_filtered_input; }
  void setAgentOutput(const std::string &agent_name, const std::string &state = DEFAULT_STATE) {
    auto it = model->agents.find(agent_name);
    if (it == model->agents.end()) throw exception::InvalidAgentName("agent '" + agent_name + "' was not found");
    if (!it->second->states.count(state)) throw exception::InvalidStateName("agent '" + agent_name + "' has no state '" + state + "'");
    function->agent_output = agent_name;
    function->agent_output_state = state;
  }
  void setAgentOutput(const char *agent_name, const std::string &state = DEFAULT_STATE) { setAgentOutput(std::string(agent_name), state); }
  inline void setAgentOutput(const AgentDescription &agent, const std::string &state = DEFAULT_STATE);
  void setAllowAgentDeath(bool has_death) { function->has_agent_death = has_death; }
  template <typename Cdn>
  void setFunctionCondition(Cdn) { function->condition = Cdn::fnPtr(); }
  bool getAllowAgentDeath() const { return function->has_agent_death; }
  bool hasMessageInput() const { return !function->message_input.empty(); }
  bool hasMessageOutput() const { return !function->message_output.empty(); }
  bool getMessageOutputOptional() const { return function->message_output_optional; }
  bool hasAgentOutput() const { return !function->agent_output.empty(); }
  std::string getInitialState() const { return function->initial_state; }
  std::string getEndState() const { return function->end_state; }
  std::shared_ptr<AgentFunctionData> function;

 private:
  std::map<std::string, std::shared_ptr<MessageData>>::iterator lookup_message(const std::string &n) {
    auto it = model->messages.find(n);
    if (it == model->messages.end()) throw exception::InvalidMessageName("message '" + n + "' was not found in the model");
    return it;
  }
  void check_type(std::map<std::string, std::shared_ptr<MessageData>>::iterator it, const std::type_index &fn_type, const char *what) {
    const MessageKind k = it->second->kind;
    const std::type_index want = k == MessageKind::Spatial3D ? std::type_index(typeid(MessageSpatial3D))
                               : (k == MessageKind::Spatial2D ? std::type_index(typeid(MessageSpatial2D))
                                  : (k == MessageKind::Bucket ? std::type_index(typeid(MessageBucket))
                                                              : std::type_index(typeid(MessageBruteForce))));
    if (want != fn_type)
      throw exception::InvalidMessageType(std::string("message ") + what + " type of function '" + function->name + "' does not match message '" + it->first + "'");
  }
  void check_state(const std::string &s) {
    auto p = function->parent.lock();
    if (p && !p->states.count(s)) throw exception::InvalidStateName("agent '" + p->name + "' has no state '" + s + "'");
  }
  std::shared_ptr<ModelData> model;
};

class AgentDescription {
 public:
  AgentDescription(std::shared_ptr<AgentData> a, std::shared_ptr<ModelData> m) : agent(std::move(a)), model(std::move(m)) {}
  std::string getName() const { return agent->name; }
  void newState(const std::string &state) {
    if (agent->default_state_only) {  // reference AgentDescription.cpp: the first user state replaces "default"
      agent->default_state_only = false;
      agent->states.clear();
      agent->initial_state = state;
    }
    if (!agent->states.insert(state).second) throw exception::InvalidStateName("agent '" + agent->name + "' already has state '" + state + "'");
  }
  void setInitialState(const std::string &state) {
    if (!agent->states.count(state)) throw exception::InvalidStateName("agent '" + agent->name + "' has no state '" + state + "'");
    agent->initial_state = state;
  }
  template <typename T>
  void newVariable(const std::string &name, T default_value = T{}) { add<T>(name, 1, &default_value); }
  template <typename T, flamegpu::size_type N>
  void newVariable(const std::string &name, const std::array<T, N> &default_value = {}) { add<T>(name, N, default_value.data()); }
  template <typename AgentFunction>
  AgentFunctionDescription newFunction(const std::string &function_name, AgentFunction) {
    if (agent->functions.count(function_name)) throw exception::InvalidAgentFunc("agent '" + agent->name + "' already has function '" + function_name + "'");
    auto f = std::make_shared<AgentFunctionData>();
    f->name = function_name;
    f->func = AgentFunction::fnPtr();
    f->func_filtered = AgentFunction::fnPtrFiltered();
    f->in_type = AgentFunction::inType();
    f->out_type = AgentFunction::outType();
    f->initial_state = agent->initial_state;
    f->end_state = agent->initial_state;
    f->parent = agent;
    agent->functions.emplace(function_name, f);
    return AgentFunctionDescription(f, model);
  }
  AgentFunctionDescription getFunction(const std::string &function_name) {
    auto it = agent->functions.find(function_name);
    if (it == agent->functions.end()) throw exception::InvalidAgentFunc("agent '" + agent->name + "' has no function '" + function_name + "'");
    return AgentFunctionDescription(it->second, model);
  }
  AgentFunctionDescription Function(const std::string &function_name) { return getFunction(function_name); }
  void setSortPeriod(unsigned int p) { agent->sort_period = p; }  // reference AgentDescription.cpp:163
  bool hasVariable(const std::string &n) const { return agent->variables.count(n) != 0; }
  bool hasState(const std::string &s) const { return agent->states.count(s) != 0; }
  std::string getInitialState() const { return agent->initial_state; }
  unsigned int getVariablesCount() const { return static_cast<unsigned int>(agent->variables.size()); }
  std::shared_ptr<AgentData> agent;

 private:
  template <typename T>
  void add(const std::string &name, unsigned int n, const T *def) {
    if (name.empty() || name[0] == '_') throw exception::InvalidAgentVar("agent variable names may not be empty or begin with '_' ('" + name + "')");
    if (agent->variables.count(name)) throw exception::InvalidAgentVar("agent '" + agent->name + "' already has variable '" + name + "'");
    agent->variables.emplace(name, make_variable<T>(n, def));
  }
  std::shared_ptr<ModelData> model;
};

inline void AgentFunctionDescription::setAgentOutput(const AgentDescription &agent, const std::string &state) {
  setAgentOutput(agent.getName(), state);
}

// ---------------------------------------------------------------------------------------------
// environment, layers, model
// ---------------------------------------------------------------------------------------------
class EnvironmentDescription {
 public:
  explicit EnvironmentDescription(std::shared_ptr<ModelData> m) : model(std::move(m)) {}
  template <typename T>
  void newProperty(const std::string &name, T value, bool is_const = false) { add<T>(name, 1, &value, is_const); }
  template <typename T, flamegpu::size_type N>
  void newProperty(const std::string &name, const std::array<T, N> &value, bool is_const = false) { add<T>(name, N, value.data(), is_const); }
  template <typename T>
  T getProperty(const std::string &name) const {
    const EnvProperty &p = find<T>(name);
    T v;
    std::memcpy(&v, p.data.data(), sizeof(T));
    return v;
  }
  template <typename T>
  T setProperty(const std::string &name, T value) {
    EnvProperty &p = const_cast<EnvProperty &>(find<T>(name));
    T old;
    std::memcpy(&old, p.data.data(), sizeof(T));
    std::memcpy(p.data.data(), &value, sizeof(T));
    return old;
  }

 private:
  template <typename T>
  void add(const std::string &name, unsigned int n, const T *v, bool is_const) {
    if (name.empty() || name[0] == '_') throw exception::InvalidEnvProperty("environment property names may not be empty or begin with '_'");
    if (model->environment.count(name)) throw exception::InvalidEnvProperty("environment already has property '" + name + "'");
    EnvProperty p;
    p.type = std::type_index(typeid(T));
    p.type_size = sizeof(T);
    p.elements = n;
    p.is_const = is_const;
    p.data.resize(sizeof(T) * n);
    std::memcpy(p.data.data(), v, sizeof(T) * n);
    model->environment.emplace(name, std::move(p));
  }
  template <typename T>
  const EnvProperty &find(const std::string &name) const {
    auto it = model->environment.find(name);
    if (it == model->environment.end()) throw exception::InvalidEnvProperty("environment has no property '" + name + "'");
    if (it->second.type != std::type_index(typeid(T))) throw exception::InvalidEnvProperty("environment property '" + name + "' has a different type");
    return it->second;
  }
  std::shared_ptr<ModelData> model;
};

class LayerDescription {
 public:
  LayerDescription(std::shared_ptr<LayerData> l, std::shared_ptr<ModelData> m) : layer(std::move(l)), model(std::move(m)) {}
  // by function object, as the examples do (reference model/LayerDescription.h:85)
  template <typename AgentFunction>
  void addAgentFunction(AgentFunction) {
    AgentFunctionWrapper *want = AgentFunction::fnPtr();
    std::shared_ptr<AgentFunctionData> found;
    for (auto &a : model->agents)
      for (auto &f : a.second->functions)
        if (f.second->func == want) {
          if (found) throw exception::InvalidAgentFunc("agent function is attached to more than one agent; add it by (agent, function) name");
          found = f.second;
        }
    if (!found) throw exception::InvalidAgentFunc("agent function was not found in any agent of the model");
    add(found);
  }
  void addAgentFunction(const AgentFunctionDescription &f) { add(f.function); }
  void addAgentFunction(const std::string &agent_name, const std::string &function_name) {
    auto a = model->agents.find(agent_name);
    if (a == model->agents.end()) throw exception::InvalidAgentName("agent '" + agent_name + "' was not found");
    auto f = a->second->functions.find(function_name);
    if (f == a->second->functions.end()) throw exception::InvalidAgentFunc("agent '" + agent_name + "' has no function '" + function_name + "'");
    add(f->second);
  }
  void addHostFunction(HostFunctionPointer fn) {
    if (!layer->functions.empty()) throw exception::InvalidLayerMember("a layer cannot hold both agent and host functions");
    layer->host_functions.push_back(fn);
  }
  std::shared_ptr<LayerData> layer;

 private:
  // reference LayerDescription.cpp:100-190: functions of one layer run concurrently, so they may not share an agent
  // state, read a message list another one writes, nor birth into a state another one executes on
  void add(const std::shared_ptr<AgentFunctionData> &f) {
    if (!layer->host_functions.empty()) throw exception::InvalidLayerMember("a layer cannot hold both agent and host functions");
    for (auto &g : layer->functions) check_pair(*f, *g);
    layer->functions.push_back(f);
  }

 public:
  // Throws if f and g cannot run in the same layer.  Also used by CUDASimulation::initialise() to re-validate every
  // layer, because setInitialState / setEndState / setAgentOutput may be called after the function joined its layer.
  static void check_pair(const AgentFunctionData &f, const AgentFunctionData &g) {
    if (&g == &f) throw exception::InvalidAgentFunc("function '" + f.name + "' is already in this layer");
    auto fp = f.parent.lock(), gp = g.parent.lock();
    if (gp == fp && (g.initial_state == f.initial_state || g.end_state == f.end_state || g.initial_state == f.end_state ||
                     g.end_state == f.initial_state))
      throw exception::InvalidAgentFunc("two functions of one layer share an agent state");
    // births (reference LayerDescription.cpp:138-165): the state a function births into may not be the initial state
    // of another member, in either direction
    if (!f.agent_output.empty() && gp && gp->name == f.agent_output && g.initial_state == f.agent_output_state)
      throw exception::InvalidLayerMember("function '" + f.name + "' births into the state function '" + g.name + "' executes on");
    if (!g.agent_output.empty() && fp && fp->name == g.agent_output && f.initial_state == g.agent_output_state)
      throw exception::InvalidLayerMember("function '" + g.name + "' births into the state function '" + f.name + "' executes on");
    if ((!f.message_output.empty() && (f.message_output == g.message_input || f.message_output == g.message_output)) ||
        (!g.message_output.empty() && g.message_output == f.message_input))
      throw exception::InvalidLayerMember("functions of one layer may not write a message list another one uses");
  }

 private:
  std::shared_ptr<ModelData> model;
};

class ModelDescription {
 public:
  explicit ModelDescription(const std::string &model_name) : model(std::make_shared<ModelData>()) { model->name = model_name; }
  AgentDescription newAgent(const std::string &agent_name) {
    if (model->agents.count(agent_name)) throw exception::InvalidAgentName("model already has agent '" + agent_name + "'");
    auto a = std::make_shared<AgentData>();
    a->name = agent_name;
    a->states.insert(DEFAULT_STATE);
    a->variables.emplace(ID_VARIABLE_NAME, make_variable<id_t>(1, nullptr));  // reference AgentData.cpp:20
    model->agents.emplace(agent_name, a);
    return AgentDescription(a, model);
  }
  AgentDescription Agent(const std::string &agent_name) {
    auto it = model->agents.find(agent_name);
    if (it == model->agents.end()) throw exception::InvalidAgentName("agent '" + agent_name + "' was not found");
    return AgentDescription(it->second, model);
  }
  template <typename MessageType>
  typename MessageType::Description newMessage(const std::string &message_name) {
    if (model->messages.count(message_name)) throw exception::InvalidMessageName("model already has message '" + message_name + "'");
    auto m = std::make_shared<MessageData>();
    m->name = message_name;
    m->kind = MessageType::Description::kind();
    model->messages.emplace(message_name, m);
    return typename MessageType::Description(m);
  }
  MessageBruteForce::Description newMessage(const std::string &message_name) { return newMessage<MessageBruteForce>(message_name); }
  template <typename MessageType>
  typename MessageType::Description Message(const std::string &message_name) {
    auto it = model->messages.find(message_name);
    if (it == model->messages.end()) throw exception::InvalidMessageName("message '" + message_name + "' was not found");
    if (it->second->kind != MessageType::Description::kind()) throw exception::InvalidMessageType("message '" + message_name + "' has a different type");
    return typename MessageType::Description(it->second);
  }
  EnvironmentDescription Environment() { return EnvironmentDescription(model); }
  LayerDescription newLayer(const std::string &name = "") {
    auto l = std::make_shared<LayerData>();
    l->name = name;
    model->layers.push_back(l);
    return LayerDescription(l, model);
  }
  void addInitFunction(HostFunctionPointer f) { model->init_functions.push_back(f); }
  void addStepFunction(HostFunctionPointer f) { model->step_functions.push_back(f); }
  void addExitFunction(HostFunctionPointer f) { model->exit_functions.push_back(f); }
  void addExitCondition(HostConditionPointer f) { model->exit_conditions.push_back(f); }
  std::string getName() const { return model->name; }
  std::shared_ptr<ModelData> model;
};

}  // namespace flamegpu

#define FLAMEGPU_INIT_FUNCTION(funcName) \
  void funcName##_impl(flamegpu::HostAPI *FLAMEGPU); \
  flamegpu::HostFunctionPointer funcName = funcName##_impl; \
  void funcName##_impl(flamegpu::HostAPI *FLAMEGPU)
#define FLAMEGPU_STEP_FUNCTION(funcName) FLAMEGPU_INIT_FUNCTION(funcName)
#define FLAMEGPU_EXIT_FUNCTION(funcName) FLAMEGPU_INIT_FUNCTION(funcName)
#define FLAMEGPU_HOST_FUNCTION(funcName) FLAMEGPU_INIT_FUNCTION(funcName)
#define FLAMEGPU_EXIT_CONDITION(funcName) \
  bool funcName##_impl(flamegpu::HostAPI *FLAMEGPU); \
  flamegpu::HostConditionPointer funcName = funcName##_impl; \
  bool funcName##_impl(flamegpu::HostAPI *FLAMEGPU)

#endif  // FGB_INCLUDE_FLAMEGPU_MODEL_MODELDESCRIPTION_H_
