// flamegpu/simulation/AgentVector.h -- host-side population container (SoA), the exchange
// format of CUDASimulation::setPopulationData / getPopulationData.  API subset of the reference's
// include/flamegpu/simulation/AgentVector.h + AgentVector_Agent.h.
#ifndef FGB_INCLUDE_FLAMEGPU_SIMULATION_AGENTVECTOR_H_
#define FGB_INCLUDE_FLAMEGPU_SIMULATION_AGENTVECTOR_H_

#include <array>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "flamegpu/model/ModelDescription.h"

namespace flamegpu {

class AgentVector {
 public:
  class Agent {
   public:
    Agent(AgentVector *p, size_type i) : parent(p), index(i) {}
    template <typename T>
    T getVariable(const std::string &name) const {
      const auto &c = parent->column<T>(name, 1);
      T v;
      std::memcpy(&v, c.data() + static_cast<size_t>(index) * sizeof(T), sizeof(T));
      return v;
    }
    template <typename T, size_type N>
    std::array<T, N> getVariable(const std::string &name) const {
      const auto &c = parent->column<T>(name, N);
      std::array<T, N> v;
      std::memcpy(v.data(), c.data() + static_cast<size_t>(index) * sizeof(T) * N, sizeof(T) * N);
      return v;
    }
    template <typename T>
    void setVariable(const std::string &name, T value) {
      if (name.size() && name[0] == '_') throw exception::InvalidAgentVar("internal variables cannot be set");
      auto &c = parent->column<T>(name, 1);
      std::memcpy(c.data() + static_cast<size_t>(index) * sizeof(T), &value, sizeof(T));
    }
    template <typename T, size_type N>
    void setVariable(const std::string &name, const std::array<T, N> &value) {
      auto &c = parent->column<T>(name, N);
      std::memcpy(c.data() + static_cast<size_t>(index) * sizeof(T) * N, value.data(), sizeof(T) * N);
    }
    template <typename T, size_type N>
    void setVariable(const std::string &name, size_type element, T value) {
      auto &c = parent->column<T>(name, N);
      if (element >= N) throw exception::OutOfBoundsException("array index out of bounds");
      std::memcpy(c.data() + (static_cast<size_t>(index) * N + element) * sizeof(T), &value, sizeof(T));
    }
    id_t getID() const { return getVariable<id_t>(ID_VARIABLE_NAME); }
    size_type getIndex() const { return index; }

   private:
    AgentVector *parent;
    size_type index;
  };
  class iterator {
   public:
    iterator(AgentVector *p, size_type i) : parent(p), index(i) {}
    Agent operator*() const { return Agent(parent, index); }
    iterator &operator++() {
      ++index;
      return *this;
    }
    bool operator!=(const iterator &o) const { return index != o.index; }

   private:
    AgentVector *parent;
    size_type index;
  };

  explicit AgentVector(const AgentDescription &desc, size_type count = 0) : agent(desc.agent), count_(0) {
    for (const auto &v : agent->variables) columns[v.first];
    resize(count);
  }
  size_type size() const { return count_; }
  bool empty() const { return count_ == 0; }
  void resize(size_type n) {
    for (const auto &v : agent->variables) {
      auto &c = columns[v.first];
      const size_t b = v.second.bytes();
      c.resize(static_cast<size_t>(n) * b);
      for (size_type i = count_; i < n; ++i) std::memcpy(c.data() + static_cast<size_t>(i) * b, v.second.default_value.data(), b);
    }
    count_ = n;
  }
  void clear() { resize(0); }
  void push_back() { resize(count_ + 1); }
  Agent operator[](size_type i) {
    if (i >= count_) throw exception::OutOfBoundsException("AgentVector index out of bounds");
    return Agent(this, i);
  }
  Agent at(size_type i) { return (*this)[i]; }
  Agent back() { return (*this)[count_ - 1]; }
  iterator begin() { return iterator(this, 0); }
  iterator end() { return iterator(this, count_); }
  const std::shared_ptr<AgentData> &getAgentData() const { return agent; }
  // raw SoA column (bytes), used by the simulation for bulk copies
  std::vector<char> &raw(const std::string &name) { return columns.at(name); }
  const std::vector<char> &raw(const std::string &name) const { return columns.at(name); }

 private:
  template <typename T>
  std::vector<char> &column(const std::string &name, unsigned int elements) {
    auto v = agent->variables.find(name);
    if (v == agent->variables.end()) throw exception::InvalidAgentVar("agent '" + agent->name + "' has no variable '" + name + "'");
    if (v->second.type != std::type_index(typeid(T))) throw exception::InvalidVarType("variable '" + name + "' accessed with the wrong type");
    if (v->second.elements != elements) throw exception::InvalidVarType("variable '" + name + "' accessed with the wrong array length");
    return columns[name];
  }
  std::shared_ptr<AgentData> agent;
  size_type count_;
  std::map<std::string, std::vector<char>> columns;
};

}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_SIMULATION_AGENTVECTOR_H_
