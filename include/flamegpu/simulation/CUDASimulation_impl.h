// flamegpu/simulation/CUDASimulation_impl.h -- inline implementation of the step scheduler
// declared in CUDASimulation.h.  Included at the end of that header; do not include directly.
#ifndef FGB_INCLUDE_FLAMEGPU_SIMULATION_CUDASIMULATION_IMPL_H_
#define FGB_INCLUDE_FLAMEGPU_SIMULATION_CUDASIMULATION_IMPL_H_

namespace flamegpu {

inline void CUDASimulation::parse_args(int argc, const char **argv) {
  // the reference's argv parser (Simulation.cu:233-316): -s steps, -r seed, -d device, -t timing
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&](void) -> std::string { return i + 1 < argc ? std::string(argv[++i]) : std::string(); };
    if (a == "-s" || a == "--steps") config.steps = static_cast<unsigned int>(std::strtoul(next().c_str(), nullptr, 10));
    else if (a == "-r" || a == "--random") config.random_seed = std::strtoull(next().c_str(), nullptr, 10);
    else if (a == "-d" || a == "--device") cuda_config.device_id = std::atoi(next().c_str());
    else if (a == "-t" || a == "--timing") config.timing = true;
    else if (a == "-i" || a == "--in") config.input_file = next();
    else if (a == "-q" || a == "--quiet") config.verbosity = 0;
    else if (a == "-v" || a == "--verbose") config.verbosity = 2;
  }
}

inline void CUDASimulation::initialise() {
  if (initialised) return;
  FGB_CUDA_THROW(cudaSetDevice(cuda_config.device_id));
  FGB_ABI_THROW(fgb_ctx_create(cuda_config.device_id, &ctx));
  FGB_CUDA_THROW(cudaStreamCreateWithFlags(&main_stream, cudaStreamNonBlocking));
  FGB_CUDA_THROW(cudaEventCreateWithFlags(&fork_event, cudaEventDisableTiming));
  FGB_CUDA_THROW(cudaEventCreateWithFlags(&index_fork, cudaEventDisableTiming));
  FGB_CUDA_THROW(cudaEventCreateWithFlags(&index_done, cudaEventDisableTiming));
  FGB_CUDA_THROW(cudaStreamCreateWithFlags(&index_stream, cudaStreamNonBlocking));
  FGB_CUDA_THROW(cudaMalloc(&d_reduce_out, 8));
  FGB_CUDA_THROW(cudaMalloc(&d_reduce_epoch, 8));
  FGB_CUDA_THROW(cudaMemset(d_reduce_epoch, 0, 8));
  FGB_CUDA_THROW(cudaMalloc(&d_ctrl, kCtrlWords * 4));
  FGB_CUDA_THROW(cudaMemset(d_ctrl, 0, kCtrlWords * 4));
  next_slot = 1;
  slab_tmp_slot = alloc_slot();
  alloc_slot();  // two consecutive words for slabPack's counts

  // agents whose functions touch spatial messages carry the auto-sort key variable
  // (reference src/flamegpu/model/AgentFunctionDescription.cpp:534-535)
  for (auto &ap : model->agents)
    for (auto &fp : ap.second->functions) {
      auto spatial = [&](const std::string &m) {
        auto it = model->messages.find(m);
        return it != model->messages.end() && it->second->dims() > 0;
      };
      if (spatial(fp.second->message_input) || spatial(fp.second->message_output))
        if (!ap.second->variables.count("_auto_sort_bin_index"))
          ap.second->variables.emplace("_auto_sort_bin_index", make_variable<unsigned int>(1, nullptr));
    }

  std::vector<unsigned int> zero_slots;
  for (auto &mp : model->messages) {
    detail::CUDAMessage &m = messages[mp.first];
    m.desc = mp.second;
    m.list.init(mp.second->variables, true);
    m.list.gen = &alloc_gen;
    m.list.count_slot = alloc_slot();
    m.keyed_slot = alloc_slot();
    if (!mp.second->persistent) zero_slots.push_back(m.list.count_slot);
    if (mp.second->dims() > 0) {
      m.list.pad_slot = true;  // padding message of the radius-filtered iterator
      if (!(mp.second->radius > 0.f)) throw exception::InvalidMessageType("spatial message '" + mp.first + "' has no radius");
      auto w = windows.find(mp.first);
      if (w != windows.end()) {
        m.win_begin = w->second.first;
        m.win_count = w->second.second;
      }
      FGB_ABI_THROW(fgb_spatial_create_window(ctx, mp.second->dims(), mp.second->min, mp.second->max, mp.second->radius, m.win_begin,
                                              m.win_count, &m.spatial));
      FGB_ABI_THROW(fgb_spatial_get_window(m.spatial, &m.win_begin, &m.win_count));
      unsigned int bins = 0;
      FGB_ABI_THROW(fgb_spatial_get_metadata(m.spatial, &m.md, &bins));
    } else if (mp.second->kind == MessageKind::Bucket) {
      // reference MessageBucket::Data validation (CUDASimulation construction throws InvalidMessage when the
      // upper bound was never set, test_bucket.cu:38-47)
      if (mp.second->bucket_upper == std::numeric_limits<int>::max())
        throw exception::InvalidMessage("bucket message '" + mp.first + "' has no upper bound");
      FGB_ABI_THROW(fgb_bucket_create(ctx, mp.second->bucket_lower, mp.second->bucket_upper, &m.spatial));
      m.bucket = true;
      unsigned int bins = 0;
      FGB_ABI_THROW(fgb_spatial_get_metadata(m.spatial, &m.md, &bins));
    }
  }
  for (auto &ap : model->agents) {
    detail::CUDAAgent &a = agents[ap.first];
    a.desc = ap.second;
    a.next_id_slot = alloc_slot();
    a.host_next_id = 1;
    write_slot(a.next_id_slot, 1u);
    for (const auto &s : ap.second->states) {
      detail::DevList &l = a.states[s];
      l.init(ap.second->variables, true);
      l.gen = &alloc_gen;
      l.count_slot = alloc_slot();
      l.perm_limit_slot = alloc_slot();
    }
  }
  n_zero_slots = static_cast<unsigned int>(zero_slots.size());
  if (n_zero_slots) {
    FGB_CUDA_THROW(cudaMalloc(&d_zero_slots, n_zero_slots * 4));
    FGB_CUDA_THROW(cudaMemcpy(d_zero_slots, zero_slots.data(), n_zero_slots * 4, cudaMemcpyHostToDevice));
  }

  size_t max_width = 1;
  for (auto &lp : model->layers) {
    // functions may have changed states / outputs after they joined the layer: validate again (reference
    // LayerDescription.cpp:100-190 validates at add time only, ModelData::validate at construction)
    bool shared_birth_target = false;
    for (size_t i = 0; i < lp->functions.size(); ++i)
      for (size_t j = i + 1; j < lp->functions.size(); ++j) {
        LayerDescription::check_pair(*lp->functions[i], *lp->functions[j]);
        const AgentFunctionData &fi = *lp->functions[i], &fj = *lp->functions[j];
        if (!fi.agent_output.empty() && fi.agent_output == fj.agent_output && fi.agent_output_state == fj.agent_output_state)
          shared_birth_target = true;
      }
    // two members appending newborns to the same list would race on its count word: such a layer runs serially
    layer_serial.push_back(shared_birth_target);
    layers.emplace_back();
    if (!lp->host_functions.empty()) model_has_host_layers = true;
    for (auto &fp : lp->functions) {
      layers.back().emplace_back();
      detail::FunctionRT &f = layers.back().back();
      f.fn = fp;
      auto parent = fp->parent.lock();
      f.agent = &agent_rt(parent->name);
      if (!fp->message_input.empty()) f.msg_in = &messages.at(fp->message_input);
      if (!fp->message_output.empty()) f.msg_out = &messages.at(fp->message_output);
      f.tmp_slot = alloc_slot();
      f.failed_slot = alloc_slot();
      f.active_slot = alloc_slot();
      for (detail::DevFlags *fl : {&f.death_flag, &f.msg_flag, &f.birth_flag, &f.cond_flag, &f.move_flag, &f.exec_perm}) fl->gen = &alloc_gen;
      f.scratch_new.gen = &alloc_gen;
      if (!fp->agent_output.empty()) {
        model_has_births = true;
        f.out_agent = &agent_rt(fp->agent_output);
        f.scratch_new.init(f.out_agent->desc->variables, false);
        size_t total = 0;
        for (const auto &v : f.out_agent->desc->variables) {
          f.default_offsets.push_back(total);
          total += (v.second.bytes() + 15) & ~static_cast<size_t>(15);
        }
        std::vector<char> packed(std::max<size_t>(total, 16), 0);
        size_t k = 0;
        for (const auto &v : f.out_agent->desc->variables) std::memcpy(packed.data() + f.default_offsets[k++], v.second.default_value.data(), v.second.bytes());
        FGB_CUDA_THROW(cudaMalloc(&f.d_defaults, packed.size()));
        FGB_CUDA_THROW(cudaMemcpy(f.d_defaults, packed.data(), packed.size(), cudaMemcpyHostToDevice));
      }
      // auto-sort trigger (reference CUDASimulation.cu:410-460), scalar float x,y(,z) only
      if (f.msg_in && f.msg_in->desc->dims() > 0) {
        auto is_f = [&](const char *n) {
          auto it = parent->variables.find(n);
          return it != parent->variables.end() && it->second.type == std::type_index(typeid(float)) && it->second.elements == 1;
        };
        const int d = f.msg_in->desc->dims();
        if (is_f("x") && is_f("y") && (d == 2 || is_f("z"))) {
          f.sortable = true;
          f.sort_dims = d;
          const MessageData &md = *f.msg_in->desc;
          FGB_ABI_THROW(fgb_spatial_create_window(ctx, d, md.min, md.max, md.radius, f.msg_in->win_begin, f.msg_in->win_count,
                                                  &f.exec_binner));
        }
      }
    }
    max_width = std::max(max_width, lp->functions.size());
  }
  if (cuda_config.inLayerConcurrency && max_width > 1) {
    side_streams.resize(max_width);
    join_events.resize(max_width);
    for (size_t i = 0; i < max_width; ++i) {
      FGB_CUDA_THROW(cudaStreamCreateWithFlags(&side_streams[i], cudaStreamNonBlocking));
      FGB_CUDA_THROW(cudaEventCreateWithFlags(&join_events[i], cudaEventDisableTiming));
    }
  }

  // environment: packed property buffer + name-hash table (reference EnvironmentManager)
  size_t off = 0;
  {
    if (model->environment.size() > static_cast<size_t>(b200::kMaxEnvProps)) throw exception::UnsupportedFeature("more than b200::kMaxEnvProps environment properties");
    std::vector<uint32_t> h, slot(model->environment.size());
    for (const auto &p : model->environment) h.push_back(detail::name_hash_rt(p.first));
    std::memset(&env_table, 0, sizeof(env_table));
    if (!detail::build_perfect_table(env_table, h.data(), static_cast<uint32_t>(h.size()), slot.data()))
      throw exception::UnsupportedFeature("could not build a collision-free environment table");
    size_t k = 0;
    env_offsets.clear();
    for (const auto &p : model->environment) {
      off = (off + 15) & ~static_cast<size_t>(15);
      env_table.offset[slot[k++]] = static_cast<uint32_t>(off);
      env_offsets.push_back(off);
      off += p.second.data.size();
    }
  }
  env_host.assign(std::max<size_t>(off, 16), 0);
  FGB_CUDA_THROW(cudaMalloc(&d_env, env_host.size()));
  env_table.buffer = d_env;
  env_dirty = true;
  for (auto &f : slab_flags) f.gen = &alloc_gen;
  if (slab.enabled) slab_setup();
  initialised = true;
}

inline void CUDASimulation::destroy() {
  if (!initialised) return;
  cudaDeviceSynchronize();
  for (auto &g : graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  graphs.clear();
  for (auto &e : step_events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  step_events.clear();
  for (cudaEvent_t e : event_pool) cudaEventDestroy(e);
  event_pool.clear();
  for (auto &l : layers)
    for (auto &f : l) {
      f.scratch_new.release();
      f.death_flag.release();
      f.msg_flag.release();
      f.birth_flag.release();
      f.cond_flag.release();
      f.move_flag.release();
      if (f.d_defaults) cudaFree(f.d_defaults);
      f.exec_perm.release();
      if (f.exec_binner) fgb_spatial_destroy(f.exec_binner);
    }
  for (auto &a : agents)
    for (auto &s : a.second.states) s.second.release();
  for (auto &m : messages) {
    m.second.list.release();
    if (m.second.spatial) fgb_spatial_destroy(m.second.spatial);
  }
  for (auto &f : slab_flags) f.release();
  if (slab.enabled) {
    for (int r = 0; r < static_cast<int>(slab.peer.size()); ++r)
      if (r != slab.rank && slab.peer[r]) cudaIpcCloseMemHandle(slab.peer[r]);
    slab.peer.clear();
    if (slab.arena) cudaFree(slab.arena);
    slab.arena = nullptr;
  }
  for (int i = 0; i < 2; ++i) {
    if (h_ctrl_pinned[i]) cudaFreeHost(h_ctrl_pinned[i]);
    if (ctrl_events[i]) cudaEventDestroy(ctrl_events[i]);
  }
  if (stream_copy) cudaStreamDestroy(stream_copy);
  if (stream_chunk_done) cudaEventDestroy(stream_chunk_done);
  stream_copy = nullptr;
  stream_chunk_done = nullptr;
  if (d_new_aos) cudaFree(d_new_aos);
  d_new_aos = nullptr;
  new_aos_bytes = 0;
  if (h_words) cudaFreeHost(h_words);
  h_words = nullptr;
  if (d_env) cudaFree(d_env);
  if (d_zero_slots) cudaFree(d_zero_slots);
  if (d_ctrl) cudaFree(d_ctrl);
  for (auto s : side_streams) cudaStreamDestroy(s);
  for (auto e : join_events) cudaEventDestroy(e);
  if (fork_event) cudaEventDestroy(fork_event);
  if (index_fork) cudaEventDestroy(index_fork);
  if (index_done) cudaEventDestroy(index_done);
  if (index_stream) cudaStreamDestroy(index_stream);
  if (d_reduce_out) cudaFree(d_reduce_out);
  d_reduce_out = nullptr;
  if (h_prefetch) cudaFreeHost(h_prefetch);
  h_prefetch = nullptr;
  d_prefetch = nullptr;
  if (d_prefetch_vals) cudaFree(d_prefetch_vals);
  d_prefetch_vals = nullptr;
  if (d_reduce_epoch) cudaFree(d_reduce_epoch);
  d_reduce_epoch = nullptr;
  if (d_user_reduce) cudaFree(d_user_reduce);
  d_user_reduce = nullptr;
  if (d_hist_out) cudaFree(d_hist_out);
  d_hist_out = nullptr;
  hist_cap = 0;
  fork_event = index_fork = index_done = nullptr;
  index_stream = nullptr;
  if (main_stream) cudaStreamDestroy(main_stream);
  if (ctx) fgb_ctx_destroy(ctx);
  ctx = nullptr;
  initialised = false;
}

inline void CUDASimulation::upload_environment() {
  size_t k = 0;
  for (const auto &p : model->environment) std::memcpy(env_host.data() + env_offsets[k++], p.second.data.data(), p.second.data.size());
  FGB_CUDA_THROW(cudaMemcpyAsync(d_env, env_host.data(), env_host.size(), cudaMemcpyHostToDevice, main_stream));
  FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
  env_dirty = false;
}

inline void CUDASimulation::setPopulationData(AgentVector &pop, const std::string &state) {
  initialise();
  FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
  const std::string &an = pop.getAgentData()->name;
  detail::CUDAAgent &a = agent_rt(an);
  detail::DevList &l = state_list(an, state);
  const unsigned int n = pop.size();
  // ids: agents still at ID_NOT_SET get the next free ids in order (reference assignAgentIDs,
  // CUDAFatAgent.cu:311-391, does this lazily at the first step)
  {
    std::vector<char> &ids = pop.raw(ID_VARIABLE_NAME);
    id_t *p = reinterpret_cast<id_t *>(ids.data());
    id_t mx = 0;
    for (unsigned int i = 0; i < n; ++i) mx = std::max(mx, p[i]);
    a.host_next_id = std::max(a.host_next_id, mx + 1);
    for (unsigned int i = 0; i < n; ++i)
      if (p[i] == ID_NOT_SET) p[i] = a.host_next_id++;
    write_slot(a.next_id_slot, a.host_next_id);
  }
  l.bound = model_has_births ? quantise(n) : n;
  l.touch();
  a.recompute_pop_bound();
  l.reserve(std::max(l.bound, 1u), 0);
  for (size_t v = 0; v < l.names.size(); ++v) {
    const size_t b = l.meta[v].bytes();
    std::vector<char> tmp;
    const char *src = nullptr;
    try {
      src = pop.raw(l.names[v]).data();
      if (pop.raw(l.names[v]).size() != static_cast<size_t>(n) * b) src = nullptr;
    } catch (const std::out_of_range &) {
      src = nullptr;
    }
    if (!src) {  // variable added after the AgentVector was made (e.g. _auto_sort_bin_index): defaults
      tmp.resize(static_cast<size_t>(n) * b);
      for (unsigned int i = 0; i < n; ++i) std::memcpy(tmp.data() + static_cast<size_t>(i) * b, l.meta[v].default_value.data(), b);
      src = tmp.data();
    }
    if (n) FGB_CUDA_THROW(cudaMemcpy(l.data[v], src, static_cast<size_t>(n) * b, cudaMemcpyHostToDevice));
  }
  write_slot(l.count_slot, n);
}

inline unsigned int CUDASimulation::getAgentCount(const std::string &agent_name, const std::string &state) {
  initialise();
  return read_slot(state_list(agent_name, state).count_slot);
}

inline void CUDASimulation::getPopulationData(AgentVector &pop, const std::string &state) {
  initialise();
  const std::string &an = pop.getAgentData()->name;
  detail::DevList &l = state_list(an, state);
  const unsigned int n = read_slot(l.count_slot);  // synchronises the step stream
  pop.resize(n);
  for (size_t v = 0; v < l.names.size(); ++v) {
    const size_t b = l.meta[v].bytes();
    std::vector<char> *dst = nullptr;
    try {
      dst = &pop.raw(l.names[v]);
    } catch (const std::out_of_range &) {
      continue;
    }
    if (dst->size() != static_cast<size_t>(n) * b) continue;
    if (n) FGB_CUDA_THROW(cudaMemcpy(dst->data(), l.data[v], static_cast<size_t>(n) * b, cudaMemcpyDeviceToHost));
  }
}

inline void CUDASimulation::setPopulationDataSoA(const std::string &agent_name, const std::string &state, unsigned int n,
                                                 unsigned int nvars, const char *const *names, const void *const *host_ptrs) {
  initialise();
  detail::CUDAAgent &a = agent_rt(agent_name);
  detail::DevList &l = state_list(agent_name, state);
  const unsigned int bound = model_has_births ? quantise(n) : n;
  if (std::max(bound, 1u) > l.capacity) {
    FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
    l.reserve(std::max(bound, 1u), 0);
  }
  l.bound = bound;
  l.touch();
  a.recompute_pop_bound();
  for (size_t v = 0; v < l.names.size(); ++v) {
    const size_t b = l.meta[v].bytes();
    const void *src = nullptr;
    for (unsigned int k = 0; k < nvars; ++k)
      if (l.names[v] == names[k]) src = host_ptrs[k];
    if (!n) continue;
    if (src) {
      FGB_CUDA_THROW(cudaMemcpyAsync(l.data[v], src, static_cast<size_t>(n) * b, cudaMemcpyHostToDevice, main_stream));
    } else if (l.names[v] == ID_VARIABLE_NAME) {
      detail::k_iota<<<(n + 255) / 256, 256, 0, main_stream>>>(reinterpret_cast<unsigned int *>(l.data[v]), n, 1u);
      ++own_launches;
    } else {
      // defaults: all-zero defaults are a memset, anything else a broadcast of the default value
      bool zero = true;
      for (char c : l.meta[v].default_value) zero = zero && c == 0;
      if (zero) {
        FGB_CUDA_THROW(cudaMemsetAsync(l.data[v], 0, static_cast<size_t>(n) * b, main_stream));
      } else {
        std::vector<char> tmp(static_cast<size_t>(n) * b);
        for (unsigned int i = 0; i < n; ++i) std::memcpy(tmp.data() + static_cast<size_t>(i) * b, l.meta[v].default_value.data(), b);
        FGB_CUDA_THROW(cudaMemcpyAsync(l.data[v], tmp.data(), tmp.size(), cudaMemcpyHostToDevice, main_stream));
        FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
      }
    }
  }
  a.host_next_id = n + 1;
  // control words travel through pinned staging (two slots per call parity), so the call returns without draining the stream
  if (!h_words) FGB_CUDA_THROW(cudaMallocHost(&h_words, 16 * sizeof(unsigned int)));
  static_assert(sizeof(unsigned int) == 4, "");
  unsigned int *w = h_words + 2 * (soa_uploads++ & 7u);
  w[0] = n;
  w[1] = n + 1;
  FGB_CUDA_THROW(cudaMemcpyAsync(d_ctrl + l.count_slot, &w[0], 4, cudaMemcpyHostToDevice, main_stream));
  FGB_CUDA_THROW(cudaMemcpyAsync(d_ctrl + a.next_id_slot, &w[1], 4, cudaMemcpyHostToDevice, main_stream));
}

inline void CUDASimulation::streamPopulationDataSoA(const std::string &agent_name, const std::string &state, unsigned int nvars,
                                                    const char *const *names, void *const *host_ptrs, unsigned int chunks) {
  initialise();
  detail::DevList &l = state_list(agent_name, state);
  streamed = Streamed{};
  streamed.list = &l;
  for (unsigned int k = 0; k < nvars; ++k) {
    const int v = l.index_of(names[k]);
    if (v < 0) throw exception::InvalidAgentVar(std::string("agent has no variable '") + names[k] + "'");
    streamed.vars.push_back(v);
    streamed.host.push_back(host_ptrs[k]);
  }
  streamed.chunks = std::max(1u, chunks);
  if (!stream_copy) {
    FGB_CUDA_THROW(cudaStreamCreateWithFlags(&stream_copy, cudaStreamNonBlocking));
    FGB_CUDA_THROW(cudaEventCreateWithFlags(&stream_chunk_done, cudaEventDisableTiming));
  }
  streamed.stream = stream_copy;
  streamed.chunk_done = stream_chunk_done;
  // the step's last function that executes on the list; chunkable only if the list's membership and order are final once it ran
  for (auto &layer : layers)
    for (auto &f : layer) {
      const AgentFunctionData &fn = *f.fn;
      const bool touches = &f.agent->states.at(fn.initial_state) == &l || &f.agent->states.at(fn.end_state) == &l ||
                           (f.out_agent && &f.out_agent->states.at(fn.agent_output_state) == &l);
      if (!touches) continue;
      const bool simple = &f.agent->states.at(fn.initial_state) == &l && fn.initial_state == fn.end_state && !fn.has_agent_death &&
                          !fn.condition && !f.out_agent;
      streamed.function = simple ? &f : nullptr;
    }
  streamed.armed = true;
}

inline unsigned int CUDASimulation::finishStreamedPopulation() {
  if (!streamed.list) throw exception::InvalidArgument("finishStreamedPopulation without streamPopulationDataSoA");
  detail::DevList &l = *streamed.list;
  if (!streamed.chunked) {  // fallback: the copies could not ride on a chunked launch
    const unsigned int n = read_slot(l.count_slot);
    for (size_t k = 0; k < streamed.vars.size(); ++k)
      if (n) FGB_CUDA_THROW(cudaMemcpyAsync(streamed.host[k], l.data[streamed.vars[k]], static_cast<size_t>(n) * l.meta[streamed.vars[k]].bytes(),
                                            cudaMemcpyDeviceToHost, main_stream));
    FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
    streamed = Streamed{};
    return n;
  }
  FGB_CUDA_THROW(cudaStreamSynchronize(streamed.stream));
  FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
  const unsigned int n = streamed.n;
  streamed = Streamed{};
  return n;
}

inline unsigned int CUDASimulation::getPopulationDataSoA(const std::string &agent_name, const std::string &state, unsigned int nvars,
                                                         const char *const *names, void *const *host_ptrs, unsigned int capacity) {
  initialise();
  detail::DevList &l = state_list(agent_name, state);
  const unsigned int n = read_slot(l.count_slot);
  if (n > capacity) throw exception::OutOfBoundsException("getPopulationDataSoA: caller buffers too small");
  for (unsigned int k = 0; k < nvars; ++k) {
    const int v = l.index_of(names[k]);
    if (v < 0) throw exception::InvalidAgentVar(std::string("agent has no variable '") + names[k] + "'");
    if (n) FGB_CUDA_THROW(cudaMemcpyAsync(host_ptrs[k], l.data[v], static_cast<size_t>(n) * l.meta[v].bytes(), cudaMemcpyDeviceToHost, main_stream));
  }
  FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
  return n;
}

inline void *CUDASimulation::getAgentVariableDevicePtr(const std::string &agent_name, const std::string &state, const std::string &var) {
  initialise();
  detail::DevList &l = state_list(agent_name, state);
  const int i = l.index_of(var);
  if (i < 0) throw exception::InvalidAgentVar("agent '" + agent_name + "' has no variable '" + var + "'");
  return l.data[i];
}
inline void *CUDASimulation::getMessageVariableDevicePtr(const std::string &message_name, const std::string &var) {
  initialise();
  detail::CUDAMessage &m = messages.at(message_name);
  const int i = m.list.index_of(var);
  if (i < 0) throw exception::InvalidMessageVar("message '" + message_name + "' has no variable '" + var + "'");
  return m.list.data[i];
}
inline unsigned int CUDASimulation::getMessageCount(const std::string &message_name) {
  initialise();
  return read_slot(messages.at(message_name).list.count_slot);
}
inline fgb_spatial *CUDASimulation::getSpatialHandler(const std::string &message_name) {
  initialise();
  return messages.at(message_name).spatial;
}

// ---- host-side state that a recorded step mutates (pointer parity, bounds, list flags) -----------
inline std::vector<unsigned long long> CUDASimulation::snapshot_host_state() const {
  std::vector<unsigned long long> s;
  auto put_list = [&](const detail::DevList &l) {
    s.push_back(l.bound);
    s.push_back(l.capacity);
    for (size_t v = 0; v < l.data.size(); ++v) {
      s.push_back(reinterpret_cast<unsigned long long>(l.data[v]));
      s.push_back(reinterpret_cast<unsigned long long>(l.swap[v]));
    }
    s.push_back(l.perm_valid() ? (reinterpret_cast<unsigned long long>(l.cached_perm) | (l.perm_partial ? 1ull : 0ull) | (l.cached_perm_global ? 2ull : 0ull)) : 0ull);
  };
  for (const auto &a : agents) {
    s.push_back(a.second.pop_bound);
    for (const auto &st : a.second.states) put_list(st.second);
  }
  for (const auto &m : messages) {
    put_list(m.second.list);
    s.push_back((m.second.pbm_dirty ? 1ull : 0ull) | (m.second.truncate ? 2ull : 0ull) | (m.second.keyed_by_writer ? 4ull : 0ull) |
                (m.second.appended_after_keyed ? 8ull : 0ull) | (m.second.hist_dirty ? 16ull : 0ull) |
                (m.second.written_bin_ordered ? 32ull : 0ull));
  }
  return s;
}
inline void CUDASimulation::restore_host_state(const std::vector<unsigned long long> &s) {
  size_t k = 0;
  auto get_list = [&](detail::DevList &l) {
    l.bound = static_cast<unsigned int>(s[k++]);
    l.capacity = static_cast<unsigned int>(s[k++]);
    for (size_t v = 0; v < l.data.size(); ++v) {
      l.data[v] = reinterpret_cast<char *>(s[k++]);
      l.swap[v] = reinterpret_cast<char *>(s[k++]);
    }
    l.cached_perm = reinterpret_cast<const unsigned int *>(s[k] & ~3ull);
    l.perm_partial = (s[k] & 1ull) != 0;
    l.cached_perm_global = (s[k++] & 2ull) != 0;
    l.cached_perm_version = l.cached_perm ? l.order_version : 0ull;
  };
  for (auto &a : agents) {
    a.second.pop_bound = static_cast<unsigned int>(s[k++]);
    for (auto &st : a.second.states) get_list(st.second);
  }
  for (auto &m : messages) {
    get_list(m.second.list);
    const unsigned long long f = s[k++];
    m.second.pbm_dirty = (f & 1ull) != 0;
    m.second.truncate = (f & 2ull) != 0;
    m.second.keyed_by_writer = (f & 4ull) != 0;
    m.second.appended_after_keyed = (f & 8ull) != 0;
    m.second.written_bin_ordered = (f & 32ull) != 0;
    m.second.hist_dirty = (f & 16ull) != 0;
  }
}
inline std::vector<unsigned long long> CUDASimulation::graph_key() const {
  std::vector<unsigned long long> k = snapshot_host_state();
  // which functions sort this step (sort period), and the options that change the recorded work
  unsigned long long sort_bits = 0, bit = 0;
  for (const auto &l : layers)
    for (const auto &f : l) {
      const unsigned int p = f.agent->desc->sort_period;
      if (f.sortable && p != 0 && step_count % p == 0) sort_bits |= 1ull << (bit & 63);
      ++bit;
    }
  k.push_back(sort_bits);
  k.push_back(prefetch.size());  // reductions recorded behind the step (record_prefetched_reductions)
  k.push_back((cuda_config.stableMessageOrder ? 1ull : 0ull) | (cuda_config.trueSpatialSortKey ? 2ull : 0ull) |
              (cuda_config.binOrderExecution ? 4ull : 0ull) | (cuda_config.overlapIndexBuild ? 8ull : 0ull) |
              (cuda_config.tileLocalExecOrder ? 16ull : 0ull) | (cuda_config.fusedIndexBuild ? 32ull : 0ull) |
              (cuda_config.binOrderedOutput ? 64ull : 0ull) | (static_cast<unsigned long long>(cuda_config.spatialIterationMode + 1) << 8));
  return k;
}

// Geometry handed to the sort-key kernel, as spatialSortAgent_async computes it (reference
// CUDASimulation.cu:480-506,571).  REFERENCE QUIRK kept on purpose (agent order is a parity gate):
// MessageSpatial3D::Data derives from MessageSpatial2D::Data, so the reference's dynamic_cast at
// CUDASimulation.cu:487 always takes the 2D branch; for a 3D list envMin.z = envMax.z = 0, hence
// envWidth.z = 0 and gridDim.z = 1, while the kernel still reads z: the z term degenerates to
// floorf(((z-0)/0)*1) = +-inf/NaN -> saturated int and only x,y order the agents (on max_bit =
// floor(log2(gx*gy))+1 bits).  CUDAConfig().trueSpatialSortKey = true uses the intended 3D key.
inline int CUDASimulation::sort_geometry(const detail::FunctionRT &f, float mn[3], float width[3], unsigned int gd[3]) const {
  const fgb_spatial_metadata &md = f.msg_in->md;
  for (int a = 0; a < 3; ++a) {
    mn[a] = 0.f;
    width[a] = 0.f;
    gd[a] = 1u;
  }
  const int used = (f.sort_dims == 3 && !cuda_config.trueSpatialSortKey) ? 2 : f.sort_dims;
  for (int a = 0; a < used; ++a) {
    mn[a] = md.min[a];
    width[a] = md.environment_width[a];
    gd[a] = width[a] ? static_cast<unsigned int>(ceilf(width[a] / md.radius)) : 1u;
  }
  return static_cast<int>(std::floor(std::log2(static_cast<double>(gd[0]) * gd[1] * gd[2]))) + 1;
}

// Reserve capacities for everything the coming step can produce (same bound arithmetic as
// run_function, no launches).  Allocation is illegal during stream capture, so it happens here.
inline void CUDASimulation::plan_step() {
  // Reservations only grow and depend on nothing but the lists' bounds (and the model): when the bounds are what they were
  // when the last plan was made, every buffer is still large enough and the walk below is skipped (host time between two
  // steps is GPU idle time whenever a step function makes the host wait for the step).
  {
    std::vector<unsigned int> sig;
    sig.reserve(16);
    for (const auto &a : agents) {
      sig.push_back(a.second.pop_bound);
      for (const auto &st : a.second.states) sig.push_back(st.second.bound);
    }
    for (const auto &m : messages) sig.push_back(m.second.list.bound);
    sig.push_back(cuda_config.binOrderedOutput ? 1u : 0u);
    if (sig == last_plan_sig) return;
    last_plan_sig.swap(sig);
  }
  std::map<const detail::DevList *, unsigned int> b;  // running bounds
  auto bound_of = [&](const detail::DevList &l) -> unsigned int & {
    auto it = b.find(&l);
    if (it == b.end()) it = b.emplace(&l, l.bound).first;
    return it->second;
  };
  std::map<const detail::CUDAMessage *, bool> trunc;
  for (auto &m : messages) trunc[&m.second] = true;
  std::map<const detail::CUDAAgent *, unsigned int> pop;  // running population bounds (births raise them)
  for (auto &a : agents) pop[&a.second] = a.second.pop_bound;
  for (auto &layer : layers)
    for (auto &f : layer) {
      detail::DevList &L = f.agent->states.at(f.fn->initial_state);
      const unsigned int n = bound_of(L);
      if (n == 0) continue;
      if (f.fn->has_agent_death || f.fn->condition) f.death_flag.reserve(n);
      if (f.fn->condition) {
        f.cond_flag.reserve(n);
        f.move_flag.reserve(n);
      }
      if (f.exec_binner) {
        f.exec_perm.reserve(n);
        FGB_ABI_THROW(fgb_spatial_reserve(f.exec_binner, n));
      }
      // a spatial writer may need its own tile-local order (run_function step 3: the reader's order was a global one)
      if (f.msg_out && f.msg_out->spatial && !f.msg_out->bucket && !f.exec_binner && cuda_config.binOrderedOutput) f.exec_perm.reserve(n);
      if (f.msg_out) {
        detail::DevList &O = f.msg_out->list;
        unsigned int &ob = bound_of(O);
        ob = (trunc[f.msg_out] ? 0u : ob) + n;
        trunc[f.msg_out] = false;
        O.reserve(ob, O.capacity);
        if (f.fn->message_output_optional) f.msg_flag.reserve(n);
        if (f.msg_out->spatial) FGB_ABI_THROW(fgb_spatial_reserve(f.msg_out->spatial, ob));
      }
      if (f.fn->initial_state != f.fn->end_state) {
        detail::DevList &E = f.agent->states.at(f.fn->end_state);
        unsigned int &eb = bound_of(E);
        eb = std::min(eb + n, std::max(pop[f.agent], 1u));
        E.reserve(eb, E.capacity);
      }
      if (f.out_agent) {
        f.scratch_new.reserve(n, 0);
        f.birth_flag.reserve(n);
        detail::DevList &T = f.out_agent->states.at(f.fn->agent_output_state);
        unsigned int &tb = bound_of(T);
        if (&T == &L && f.fn->initial_state != f.fn->end_state) tb = n;
        else tb += n;
        pop[f.out_agent] += n;
        T.reserve(tb, T.capacity);
      }
      if (f.fn->initial_state != f.fn->end_state && !f.fn->condition &&
          !(f.out_agent && &f.out_agent->states.at(f.fn->agent_output_state) == &L))
        bound_of(L) = 0;
      if (f.sortable) {
        float mn[3], width[3];
        unsigned int gd[3];
        const int max_bit = sort_geometry(f, mn, width, gd);
        for (unsigned int sid = 0; sid < std::max<size_t>(1, side_streams.size()); ++sid)
          FGB_ABI_THROW(fgb_ctx_reserve(ctx, sid, n, max_bit));
      }
      for (unsigned int sid = 0; sid < std::max<size_t>(1, side_streams.size()); ++sid) FGB_ABI_THROW(fgb_ctx_reserve(ctx, sid, std::max(n, 1u), 0));
    }
  if (slab.enabled) {
    unsigned int most = 1;
    for (SlabList &S : slab.lists) {
      // room for what the neighbours may append (halo ghosts / arriving agents)
      const unsigned int b = S.is_message ? bound_of(*S.list) : S.list->bound;
      S.list->reserve(b + 2u * S.capacity, S.list->capacity);
      most = std::max(most, b + 2u * S.capacity);
      if (S.is_message) {
        for (auto &m : messages)
          if (&m.second.list == S.list && m.second.spatial) FGB_ABI_THROW(fgb_spatial_reserve(m.second.spatial, b + 2u * S.capacity));
      }
    }
    for (auto &f : slab_flags) f.reserve(most);
    FGB_ABI_THROW(fgb_ctx_reserve(ctx, kSlabScratchSlot, most, 0));
    FGB_ABI_THROW(fgb_slab_reserve(ctx, kSlabScratchSlot, std::max(std::max(slab.mig_cap, slab.halo_cap), 1u)));
  }
  // reserving may have raised list bounds' capacity only; bounds themselves are untouched
}

inline void CUDASimulation::run_function(detail::FunctionRT &f, cudaStream_t st, unsigned int sid) {
  const AgentFunctionData &fn = *f.fn;
  detail::DevList &L = f.agent->states.at(fn.initial_state);
  const unsigned int n = L.bound;
  if (n == 0) return;  // nothing can be in this state list
  unsigned int *d_n = slot_ptr(L.count_slot);
  const bool same_state = fn.initial_state == fn.end_state;

  // 1. automatic spatial sort of the executing agents (reference CUDASimulation.cu:463-573)
  const unsigned int period = f.agent->desc->sort_period;
  // after the auto sort the list is in the sort key's order: grouping inside tiles is enough for step 2b
  // ... unless the grid is deep along the slowest axis.  The sorted list is in (x,y)-column order (sort_geometry), so the
  // agents of a 2048-agent tile span EVERY plane of a few columns, and the ~110 tiles in flight (148 SMs) together touch
  // planes x 9 strips x 3 bins of messages each: at 256 planes that is ~100 MB, the whole L2, and `move` ran 3.5x slower
  // than at 32 planes (profiles/r02_move_depth.jsonl).  The global bin order keeps the blocks in flight inside ~3 planes.
  bool tile_local_fits = true;
  if (f.msg_in && f.msg_in->spatial && !f.msg_in->bucket && f.sort_dims == 3 && !cuda_config.trueSpatialSortKey) {
    size_t msg_bytes = 0;
    for (const auto &m : f.msg_in->list.meta) msg_bytes += m.bytes();
    const double per_bin = L.bound > 0 ? std::max(1.0, static_cast<double>(f.msg_in->list.bound) / std::max(1u, f.msg_in->spatial ? fgb_spatial_bin_count(f.msg_in->spatial) : 1u)) : 1.0;
    const int planes = f.msg_in->win_count > 0 ? f.msg_in->win_count : static_cast<int>(f.msg_in->md.grid_dim[f.msg_in->desc->dims() - 1]);
    const double footprint = 110.0 * planes * 27.0 * per_bin * static_cast<double>(msg_bytes);
    tile_local_fits = footprint < 32.0 * 1024 * 1024;  // a quarter of the 126 MB L2
  }
  const bool sorted_now = f.sortable && period != 0 && step_count % period == 0 && cuda_config.tileLocalExecOrder && tile_local_fits;
  if (f.sortable && period != 0 && step_count % period == 0) {
    float mn[3], width[3];
    unsigned int gd[3];
    const int max_bit = sort_geometry(f, mn, width, gd);
    prof_begin("agent_sort", st);
    const int ix = L.index_of("x"), iy = L.index_of("y"), iz = f.sort_dims == 3 ? L.index_of("z") : -1;
    const int ik = L.index_of("_auto_sort_bin_index");
    unsigned int *keys = reinterpret_cast<unsigned int *>(L.data[ik]);
    std::vector<fgb_var> vars = L.vars(true);
    FGB_ABI_THROW(fgb_sort_spatial(ctx, sid, reinterpret_cast<const float *>(L.data[ix]), reinterpret_cast<const float *>(L.data[iy]),
                                   iz >= 0 ? reinterpret_cast<const float *>(L.data[iz]) : nullptr, mn, width, gd, max_bit, n, d_n, keys,
                                   vars.data(), static_cast<unsigned int>(vars.size()), nullptr, st));
    L.swap_buffers();
    prof_end(st);
  }

  // 1b. function condition (reference CUDASimulation.cu:686-837, CUDAFatAgent.cu:186-236): evaluate it for every
  // agent, then partition the list stably into [failed | passed]; only the passed part executes
  const bool conditional = fn.condition != nullptr;
  unsigned int *d_failed = slot_ptr(f.failed_slot);
  unsigned int *d_active = slot_ptr(f.active_slot);
  if (conditional) {
    prof_begin("condition:" + fn.name, st);
    f.cond_flag.reserve(n);
    f.move_flag.reserve(n);
    f.death_flag.reserve(n);
    detail::FunctionArgs c;
    std::memset(&c, 0, sizeof(c));
    c.d_count = d_n;
    c.bound = n;
    L.fill_table(c.agent);
    c.death_flag = f.cond_flag.p;
    c.d_step = slot_ptr(kStepSlot);
    c.env = env_table;
    void *cargs[] = {&c};
    FGB_CUDA_THROW(cudaLaunchKernel(reinterpret_cast<const void *>(fn.condition), dim3((n + 127) / 128), dim3(128), cargs, 0, st));
    ++own_launches;
    std::vector<fgb_var> vars = L.vars(true);
    const unsigned int nv = static_cast<unsigned int>(vars.size());
    FGB_ABI_THROW(fgb_compact(ctx, sid, f.cond_flag.p, /*invert=*/1, n, d_n, 0, 0, nullptr, vars.data(), nv, d_failed, nullptr, st));
    FGB_ABI_THROW(fgb_compact(ctx, sid, f.cond_flag.p, /*invert=*/0, n, d_n, 0, 0, d_failed, vars.data(), nv, nullptr, nullptr, st));
    L.swap_buffers();
    detail::k_sub_word<<<1, 1, 0, st>>>(d_active, d_n, d_failed);
    ++own_launches;
    prof_end(st);
  }

  // 2. index of the input list, built lazily before its first reader (reference :864); record_layers has
  // normally issued it already (on the index stream, overlapping the sort above)
  if (f.msg_in && f.msg_in->spatial && f.msg_in->pbm_dirty) build_input_index(*f.msg_in, st);

  // 2b. execution order: bin the executing agents on the input list's grid (b200 extension)
  const bool bin_order = f.exec_binner && cuda_config.binOrderExecution && !conditional;
  if (bin_order) {
    prof_begin("exec_order", st);
    f.exec_perm.reserve(n);
    const int ix = L.index_of("x"), iy = L.index_of("y"), iz = f.sort_dims == 3 ? L.index_of("z") : -1;
    FGB_ABI_THROW(fgb_bin_permutation(f.exec_binner, n, d_n, reinterpret_cast<const float *>(L.data[ix]),
                                      reinterpret_cast<const float *>(L.data[iy]),
                                      iz >= 0 ? reinterpret_cast<const float *>(L.data[iz]) : nullptr, f.exec_perm.p,
                                      sorted_now ? FGB_BUILD_TILE_LOCAL : FGB_BUILD_DEFAULT, st));
    L.cached_perm = f.exec_perm.p;  // an output function of this list may reuse it while the list is unchanged
    L.cached_perm_version = L.order_version;
    L.perm_partial = false;
    L.cached_perm_global = !sorted_now;
    prof_end(st);
  }

  // 3. kernel arguments
  detail::FunctionArgs a;
  std::memset(&a, 0, sizeof(a));
  a.d_count = d_n;
  a.exec_perm = bin_order ? f.exec_perm.p : nullptr;
  a.d_agent_offset = conditional ? d_failed : nullptr;
  // the executing agents: all of them, or the part of the list that passed the condition
  unsigned int *d_exec = conditional ? d_active : d_n;
  a.bound = n;
  L.fill_table(a.agent);
  if (f.msg_in) {
    detail::CUDAMessage &M = *f.msg_in;
    M.list.fill_table(a.msg_in);
    a.d_msg_in_count = slot_ptr(M.list.count_slot);
    if (M.bucket) {
      // bucket lists: {min, max exclusive} travel in grid_dim[0..1] (MessageBucket.cuh reads them there)
      a.in_meta.grid_dim[0] = M.desc->bucket_lower;
      a.in_meta.grid_dim[1] = M.desc->bucket_upper + 1;
      a.in_meta.pbm = M.md.PBM;
    } else if (M.spatial) {
      for (int k = 0; k < 3; ++k) {
        a.in_meta.min[k] = M.md.min[k];
        a.in_meta.max[k] = M.md.max[k];
        a.in_meta.env_width[k] = M.md.environment_width[k];
        a.in_meta.grid_dim[k] = static_cast<int>(M.md.grid_dim[k]);
      }
      a.in_meta.win_begin = M.win_begin;
      a.in_meta.win_count = M.win_count;
      a.in_meta.radius = M.md.radius;
      a.in_meta.iter_mode = filtered_iteration(f) ? 1 : 0;
      a.in_meta.radius2_eps = M.md.radius * M.md.radius * 1.00001f;
      a.in_meta.pad_index = M.list.capacity;
      a.in_meta.wrap_compatible = M.md.wrap_compatible ? 1 : 0;
      a.in_meta.pbm = M.md.PBM;
    }
  }
  if (f.msg_out) {
    detail::CUDAMessage &MO = *f.msg_out;
    detail::DevList &O = MO.list;
    O.reserve((MO.truncate ? 0u : O.bound) + n, O.capacity);
    O.fill_table(a.msg_out, /*use_swap=*/true);  // functions write the swap list at their thread index
    const bool plain = !fn.message_output_optional && MO.truncate;  // every executing agent writes one message into an emptied list
    if (fn.message_output_optional) {
      f.msg_flag.reserve(n);
      a.msg_out_flag = f.msg_flag.p;
    } else if (MO.truncate) {
      a.d_msg_out_count = slot_ptr(O.count_slot);  // published by the function kernel itself (step 5)
    }
    if (plain && MO.spatial && !MO.bucket) {
      // bin-ordered output: thread t writes slot t for agent perm[t], perm = bin order of the list's last spatial reader.
      // Positions have moved a little since, so the list arrives NEARLY bin-grouped: every tile takes the direct scatter.
      // (Not with stableMessageOrder: the permutation's order inside a bin is atomic-arrival order, and the stable
      // build orders a bin by message slot.)
      if (cuda_config.binOrderedOutput && !cuda_config.stableMessageOrder && !bin_order && !conditional && !f.out_agent && L.perm_valid()) {
        const int ix = L.index_of("x"), iy = L.index_of("y"), iz = MO.desc->dims() == 3 ? L.index_of("z") : -1;
        if (!L.cached_perm_global) {
          a.exec_perm = L.cached_perm;
          a.d_perm_limit = L.perm_partial ? slot_ptr(L.perm_limit_slot) : nullptr;
          a.slot_by_thread = 1u;
        } else if (n > 0 && ix >= 0 && iy >= 0 && (MO.desc->dims() == 2 || iz >= 0)) {
          // the reader ran in GLOBAL bin order (deep grids, or a step without the auto sort): reading the agents through
          // that permutation is a random gather over the whole list (16.8 M agents: 983 us against 221 us).  Grouping inside
          // 2048-agent tiles is all the scatter needs to find runs, so the writer gets its own tile-local order
          // (one k_group_tile pass over the positions; the heuristic assumes the model's location variables are x, y, z).
          f.exec_perm.reserve(n);
          FGB_ABI_THROW(fgb_bin_permutation(MO.spatial, n, d_n, reinterpret_cast<const float *>(L.data[ix]),
                                            reinterpret_cast<const float *>(L.data[iy]),
                                            iz >= 0 ? reinterpret_cast<const float *>(L.data[iz]) : nullptr, f.exec_perm.p,
                                            FGB_BUILD_TILE_LOCAL, st));
          a.exec_perm = f.exec_perm.p;
          a.slot_by_thread = 1u;
        }
      }
      if (cuda_config.fusedIndexBuild) {
        if (MO.hist_dirty)  // keyed by an earlier writer but never built: start from a clean histogram
          FGB_ABI_THROW(fgb_spatial_clear_histogram(MO.spatial, st));
        unsigned int *keys = nullptr, *hist = nullptr;
        FGB_ABI_THROW(fgb_spatial_writer_args(MO.spatial, n, &keys, &hist));
        a.out_keys = keys;
        a.out_hist = hist;
        for (int k = 0; k < 3; ++k) {
          a.out_min[k] = MO.md.min[k];
          a.out_grid_dim[k] = static_cast<int>(MO.md.grid_dim[k]);
        }
        a.out_radius = MO.md.radius;
        a.out_win_begin = MO.win_begin;
        a.out_win_count = MO.win_count;
        a.d_keyed = slot_ptr(MO.keyed_slot);
      }
    }
  }
  if (f.out_agent) {
    f.scratch_new.reserve(n, 0);
    f.birth_flag.reserve(n);
    f.scratch_new.fill_table(a.agent_out);
    a.agent_out_nvars = static_cast<uint32_t>(f.default_offsets.size());
    for (size_t v = 0; v < f.default_offsets.size(); ++v) {
      a.agent_out_slot[v] = f.scratch_new.slots[v];
      a.agent_out_len[v] = static_cast<uint32_t>(f.scratch_new.meta[v].bytes());
      a.agent_out_defaults[v] = f.d_defaults + f.default_offsets[v];
    }
    a.agent_out_flag = f.birth_flag.p;
    a.next_id = slot_ptr(f.out_agent->next_id_slot);
  }
  if (fn.has_agent_death) {
    f.death_flag.reserve(n);
    a.death_flag = f.death_flag.p;
  }
  a.d_step = slot_ptr(kStepSlot);
  a.env = env_table;

  // 4. the agent function itself
  {
    void *kargs[] = {&a};
    const unsigned int bs = static_cast<unsigned int>(cuda_config.agentFunctionBlockSize > 0 ? cuda_config.agentFunctionBlockSize : f.block_size);
    if (index_pending && f.msg_in && f.msg_in->spatial) FGB_CUDA_THROW(cudaStreamWaitEvent(st, index_done, 0));
    prof_begin("function:" + fn.name, st);
    // radius-filtered iterator: kFilterQueueWords words of chunk queue per thread (FunctionArgs.h)
    const bool filtered = filtered_iteration(f);
    const size_t smem = filtered ? sizeof(uint32_t) * detail::kFilterQueueWords * bs : 0;
    const void *kernel = reinterpret_cast<const void *>(filtered ? fn.func_filtered : fn.func);
    // thread range == agent range only without a permutation or with the tile-local one (groups stay inside 2048-agent tiles)
    const bool chunkable = streamed.armed && streamed.function == &f && n > 0 && (!a.exec_perm || (bin_order && sorted_now));
    if (chunkable) {
      // streamed download: the function runs in thread-range chunks, every finished chunk of the requested variables is
      // copied to the host on the copy stream while the next chunk computes
      const unsigned int per = ((n + streamed.chunks - 1) / streamed.chunks + 2047u) & ~2047u;
      for (unsigned int c0 = 0; c0 < n; c0 += per) {
        const unsigned int c1 = std::min(n, c0 + per);
        a.first_thread = c0;
        a.last_thread = c1;
        FGB_CUDA_THROW(cudaLaunchKernel(kernel, dim3((c1 - c0 + bs - 1) / bs), dim3(bs), kargs, smem, st));
        ++own_launches;
        FGB_CUDA_THROW(cudaEventRecord(streamed.chunk_done, st));
        FGB_CUDA_THROW(cudaStreamWaitEvent(streamed.stream, streamed.chunk_done, 0));
        for (size_t k = 0; k < streamed.vars.size(); ++k) {
          const size_t b = L.meta[streamed.vars[k]].bytes();
          FGB_CUDA_THROW(cudaMemcpyAsync(static_cast<char *>(streamed.host[k]) + static_cast<size_t>(c0) * b, L.data[streamed.vars[k]] + static_cast<size_t>(c0) * b,
                                         static_cast<size_t>(c1 - c0) * b, cudaMemcpyDeviceToHost, streamed.stream));
        }
      }
      streamed.chunked = true;
      streamed.n = n;
    } else {
      FGB_CUDA_THROW(cudaLaunchKernel(kernel, dim3((n + bs - 1) / bs), dim3(bs), kargs, smem, st));
      ++own_launches;
    }
    prof_end(st);
  }

  // 5. output message list (reference CUDAMessage::swap, CUDAMessage.cu:171-208)
  prof_begin("post:" + fn.name, st);
  if (f.msg_out) {
    detail::CUDAMessage &O = *f.msg_out;
    unsigned int *d_mc = slot_ptr(O.list.count_slot);
    if (fn.message_output_optional) {
      std::vector<fgb_var> vars = O.list.vars(/*from_data=*/false);  // swap (written) -> data (read list)
      FGB_ABI_THROW(fgb_compact(ctx, sid, f.msg_flag.p, 0, n, d_exec, 0, 0, O.truncate ? nullptr : d_mc, vars.data(),
                                static_cast<unsigned int>(vars.size()), nullptr, d_mc, st));
      O.list.bound = (O.truncate ? 0u : O.list.bound) + n;
    } else if (O.truncate) {
      O.list.swap_buffers();  // the count word was written by the function kernel (FunctionArgs::d_msg_out_count)
      O.list.bound = n;
      if (a.out_keys) {
        O.keyed_by_writer = true;
        O.appended_after_keyed = false;
        O.hist_dirty = true;
      }
      O.written_bin_ordered = a.slot_by_thread != 0;
    } else {
      std::vector<fgb_var> vars = O.list.vars(false);
      FGB_ABI_THROW(fgb_compact(ctx, sid, nullptr, 0, n, d_exec, n, 0, d_mc, vars.data(), static_cast<unsigned int>(vars.size()), nullptr,
                                d_mc, st));
      O.list.bound += n;
    }
    if (fn.message_output_optional || !O.truncate) O.written_bin_ordered = false;
    if (!a.out_keys) {
      if (O.truncate) O.keyed_by_writer = false;      // a plain (unfused) rewrite of the list
      else if (O.keyed_by_writer) O.appended_after_keyed = true;  // the appended items are keyed by the build
    }
    O.truncate = false;
    O.pbm_dirty = true;  // reference CUDASimulation.cu:1057-1059
  }

  // 6. death (+ state transition) : reference processDeath then transitionState (:1064-1067)
  unsigned int *d_tmp = slot_ptr(f.tmp_slot);
  const bool births = f.out_agent != nullptr;
  detail::DevList *T = births ? &f.out_agent->states.at(fn.agent_output_state) : nullptr;
  if (conditional && fn.has_agent_death) {
    // disabled agents at the front are kept unconditionally (reference scatter_all_count, CUDAScatter.cu:82-83)
    detail::k_fill_front<<<(n + 255) / 256, 256, 0, st>>>(f.death_flag.p, d_failed, n, 1u);
    ++own_launches;
  }
  if (same_state) {
    if (fn.has_agent_death) {
      std::vector<fgb_var> vars = L.vars(true);
      FGB_ABI_THROW(fgb_compact(ctx, sid, f.death_flag.p, 0, n, d_n, 0, 0, nullptr, vars.data(), static_cast<unsigned int>(vars.size()),
                                births ? d_tmp : nullptr, births ? nullptr : d_n, st));
      L.swap_buffers();
    }
  } else {
    detail::DevList &E = f.agent->states.at(fn.end_state);
    E.reserve(std::min(E.bound + n, std::max(f.agent->pop_bound, 1u)), E.capacity);
    std::vector<fgb_var> vars(L.names.size());
    for (size_t v = 0; v < vars.size(); ++v) {
      vars[v].type_len = L.meta[v].bytes();
      vars[v].in = L.data[v];
      vars[v].out = E.data[v];
    }
    unsigned int *d_ec = slot_ptr(E.count_slot);
    if (conditional) {
      // only the executing survivors change state; the disabled front stays in the initial state
      detail::k_move_flags<<<(n + 255) / 256, 256, 0, st>>>(f.move_flag.p, fn.has_agent_death ? f.death_flag.p : nullptr, d_failed, d_n, n);
      ++own_launches;
      FGB_ABI_THROW(fgb_compact(ctx, sid, f.move_flag.p, 0, n, d_n, 0, 0, d_ec, vars.data(), static_cast<unsigned int>(vars.size()),
                                nullptr, d_ec, st));
    } else {
      FGB_ABI_THROW(fgb_compact(ctx, sid, fn.has_agent_death ? f.death_flag.p : nullptr, 0, n, d_n, fn.has_agent_death ? 0u : n, 0, d_ec,
                                vars.data(), static_cast<unsigned int>(vars.size()), nullptr, d_ec, st));
    }
    E.bound = std::min(E.bound + n, std::max(f.agent->pop_bound, 1u));  // a transition creates no agents
    E.touch();
    L.touch();
  }

  // 7. births appended after the survivors, in parent order (reference CUDAAgentStateList.cu:189-250)
  bool births_into_vacated = false;
  if (births) {
    f.out_agent->pop_bound += n;
    T->reserve(T->bound + n, T->capacity);
    std::vector<fgb_var> vars(f.scratch_new.names.size());
    for (size_t v = 0; v < vars.size(); ++v) {
      vars[v].type_len = f.scratch_new.meta[v].bytes();
      vars[v].in = f.scratch_new.data[v];
      vars[v].out = T->data[v];
    }
    const unsigned int nv = static_cast<unsigned int>(vars.size());
    unsigned int *d_tc = slot_ptr(T->count_slot);
    if (T == &L && same_state) {
      FGB_ABI_THROW(fgb_compact(ctx, sid, f.birth_flag.p, 0, n, d_exec, 0, 0, fn.has_agent_death ? d_tmp : d_n, vars.data(), nv, nullptr, d_n, st));
      L.bound += n;
      L.touch();
    } else if (T == &L) {  // the initial state was vacated by the transition: children start at 0
      FGB_ABI_THROW(fgb_compact(ctx, sid, f.birth_flag.p, 0, n, d_exec, 0, 0, nullptr, vars.data(), nv, nullptr, d_n, st));
      births_into_vacated = true;
      L.bound = n;
      L.touch();
    } else {
      FGB_ABI_THROW(fgb_compact(ctx, sid, f.birth_flag.p, 0, n, d_exec, 0, 0, d_tc, vars.data(), nv, nullptr, d_tc, st));
      T->bound += n;
      T->touch();
      if (same_state && fn.has_agent_death) {
        detail::k_copy_word<<<1, 1, 0, st>>>(d_n, d_tmp);
        ++own_launches;
      }
    }
  }
  if (!same_state && !births_into_vacated) {
    if (conditional) {
      detail::k_copy_word<<<1, 1, 0, st>>>(d_n, d_failed);  // the agents that failed the condition stay
      ++own_launches;
    } else {
      FGB_CUDA_THROW(cudaMemsetAsync(d_n, 0, 4, st));
      L.bound = 0;
    }
  }
  prof_end(st);
}

// which iterator variant a function runs with (CUDAConfig().spatialIterationMode)
inline bool CUDASimulation::filtered_iteration(const detail::FunctionRT &f) const {
  if (!(f.msg_in && f.msg_in->spatial && !f.msg_in->bucket && f.fn->func_filtered)) return false;
  if (f.msg_in->list.capacity >= (1u << 30)) return false;  // the filtered walk addresses locations with 32-bit byte offsets
  const int mode = cuda_config.spatialIterationMode;
  return mode > 0 || (mode < 0 && f.fn->radius_filtered_input);
}

inline void CUDASimulation::build_input_index(detail::CUDAMessage &M, cudaStream_t st) {
  if (M.bucket) {
    prof_begin("build_index", st);
    std::vector<fgb_var> vars = M.list.vars(true);
    const int ik = M.list.index_of("_key");
    // also for an empty list: the PBM must read all-zero (reference MessageBucket.cu:108-112)
    FGB_ABI_THROW(fgb_build_index_keys(M.spatial, M.list.bound, slot_ptr(M.list.count_slot), reinterpret_cast<const int *>(M.list.data[ik]),
                                       vars.data(), static_cast<unsigned int>(vars.size()),
                                       cuda_config.stableMessageOrder ? FGB_BUILD_STABLE : FGB_BUILD_DEFAULT, st));
    if (M.list.bound > 0) M.list.swap_buffers();
    prof_end(st);
    M.pbm_dirty = false;
    return;
  }
  if (slab.enabled && slab.connected && !slab.lists.empty() && slab.lists[0].list == &M.list) {
    // halo: the messages of the boundary planes z0 and z1-1 go to rank-1 / rank+1, theirs are appended as ghosts
    prof_begin("slab_halo", st);
    slab_exchange(slab.lists[0], slab.z0 + 1, slab.z1 - 1, /*remove=*/false, st);
    if (M.keyed_by_writer) M.appended_after_keyed = true;  // the ghosts are keyed by the build
    prof_end(st);
  }
  if (M.list.bound > 0) {
    prof_begin("build_index", st);
    std::vector<fgb_var> vars = M.list.vars(true);
    const int ix = M.list.index_of("x"), iy = M.list.index_of("y"), iz = M.desc->dims() == 3 ? M.list.index_of("z") : -1;
    unsigned int flags = cuda_config.stableMessageOrder ? FGB_BUILD_STABLE : FGB_BUILD_DEFAULT;
    if (M.keyed_by_writer) flags |= FGB_BUILD_KEYS_READY;
    // written in bin order: the few tiles that are not grouped (fast movers, appended ghosts) are scattered inside the scan +
    // scatter launch and the worklist launch is saved
    if (M.keyed_by_writer && M.written_bin_ordered && !cuda_config.stableMessageOrder) flags |= FGB_BUILD_EXPECT_GROUPED;
    FGB_ABI_THROW(fgb_build_index_ex(M.spatial, M.list.bound, slot_ptr(M.list.count_slot), reinterpret_cast<const float *>(M.list.data[ix]),
                                     reinterpret_cast<const float *>(M.list.data[iy]),
                                     iz >= 0 ? reinterpret_cast<const float *>(M.list.data[iz]) : nullptr, vars.data(),
                                     static_cast<unsigned int>(vars.size()), flags,
                                     (M.keyed_by_writer && M.appended_after_keyed) ? slot_ptr(M.keyed_slot) : nullptr, nullptr, st));
    M.list.swap_buffers();
    prof_end(st);
  }
  M.keyed_by_writer = false;  // the sorted list has new slots; the histogram was consumed (and re-zeroed) by the scan
  M.written_bin_ordered = false;
  M.appended_after_keyed = false;
  M.hist_dirty = false;
  M.pbm_dirty = false;
}

inline void CUDASimulation::record_layers(cudaStream_t main, size_t first, size_t last) {
  for (size_t li = first; li < last && li < layers.size(); ++li) {
    auto &layer = layers[li];
    // The PBMs this layer reads do not depend on the layer's own agent lists: build them on the index stream
    // while the functions' streams sort / partition their agents; each reader waits for index_done before its
    // kernel.  (Also orders the build before EVERY reader when several functions of a layer share a list.)
    bool forked = false;
    for (auto &f : layer) {
      if (!(f.msg_in && f.msg_in->spatial && f.msg_in->pbm_dirty)) continue;
      const bool overlap = cuda_config.overlapIndexBuild && !cuda_config.profile;
      if (overlap && !forked) {
        FGB_CUDA_THROW(cudaEventRecord(index_fork, main));
        FGB_CUDA_THROW(cudaStreamWaitEvent(index_stream, index_fork, 0));
        forked = true;
      }
      build_input_index(*f.msg_in, overlap ? index_stream : main);
    }
    if (forked) {
      FGB_CUDA_THROW(cudaEventRecord(index_done, index_stream));
      index_pending = true;
    }
    const bool fork = cuda_config.inLayerConcurrency && layer.size() > 1 && !side_streams.empty() && !layer_serial[li];
    if (fork) FGB_CUDA_THROW(cudaEventRecord(fork_event, main));
    for (size_t i = 0; i < layer.size(); ++i) {
      cudaStream_t st = fork ? side_streams[i] : main;
      if (fork) FGB_CUDA_THROW(cudaStreamWaitEvent(st, fork_event, 0));
      run_function(layer[i], st, fork ? static_cast<unsigned int>(i) : 0u);
      if (fork) {
        FGB_CUDA_THROW(cudaEventRecord(join_events[i], st));
        FGB_CUDA_THROW(cudaStreamWaitEvent(main, join_events[i], 0));
      }
    }
    if (index_pending) {  // every reader has waited already; this joins the index stream back for everything else
      FGB_CUDA_THROW(cudaStreamWaitEvent(main, index_done, 0));
      index_pending = false;
    }
    // host-function layers run between graphs (eager mode only)
    if (!model->layers[li]->host_functions.empty()) {
      FGB_CUDA_THROW(cudaStreamSynchronize(main));
      for (auto hf : model->layers[li]->host_functions) hf(&host_api);
      flush_host_agents();
      if (env_dirty) upload_environment();
    }
  }
}

inline void CUDASimulation::record_end_of_step(cudaStream_t main) {
  // end of step on the device: ++step counter, non-persistent lists emptied (reference :619-625)
  if (slab.enabled) {
    // migration: agents whose position left the slab go to the neighbour that owns their plane
    prof_begin("slab_migrate", main);
    for (size_t k = 1; k < slab.lists.size(); ++k) {
      SlabList &S = slab.lists[k];
      slab_exchange(S, slab.z0, slab.z1, /*remove=*/true, main);
      FGB_ABI_THROW(fgb_slab_check_bound(ctx, slot_ptr(S.list->count_slot), S.list->bound, slot_ptr(slab.err_slot), main));
    }
    prof_end(main);
  }
  detail::k_end_of_step<<<1, 32, 0, main>>>(d_ctrl, kStepSlot, d_zero_slots, n_zero_slots, slab.enabled ? slab.epoch_slot : 0u);
  ++own_launches;
  for (auto &m : messages)
    if (!m.second.desc->persistent) {
      m.second.truncate = true;
      m.second.pbm_dirty = true;
      m.second.keyed_by_writer = false;  // (hist_dirty stays: an unbuilt histogram is cleared by the next fused writer)
      m.second.appended_after_keyed = false;
    }
}

inline void CUDASimulation::record_step(cudaStream_t main) {
  for (auto &m : messages) m.second.truncate = true;  // reference CUDASimulation.cu:599-601
  record_layers(main, 0, layers.size());
  record_end_of_step(main);
  record_prefetched_reductions(main);
}

// The reductions the model's step functions are known to ask for, recorded behind the step (so inside its graph); their
// 8-byte results go straight to mapped pinned host memory, followed by the step's epoch.
inline void CUDASimulation::record_prefetched_reductions(cudaStream_t st) {
  last_step_prefetched = 0;
  if (prefetch.empty() || model->step_functions.empty()) return;
  if (!h_prefetch) return;  // allocated when the first reduction is learned (never during capture)
  for (const PrefetchedReduction &r : prefetch) {
    detail::DevList &l = state_list(r.agent, r.state);
    const int i = l.index_of(r.variable);
    unsigned long long *d_val = d_prefetch_vals + last_step_prefetched;
    FGB_ABI_THROW(fgb_reduce(ctx, 0, r.op, r.dtype, l.data[i], l.bound, slot_ptr(l.count_slot), d_val, st));
    if (slab.enabled) {  // every rank holds a part of the population: fold the per-rank results (same sequence on every rank)
      // sum<T> accumulates in the 8-byte type of T's kind, min / max keep T (fgb_reduce)
      const int rdtype = r.op == FGB_REDUCE_SUM ? (r.dtype == FGB_F32 || r.dtype == FGB_F64 ? FGB_F64 : (r.dtype == FGB_I32 || r.dtype == FGB_I64 ? FGB_I64 : FGB_U64))
                                                : r.dtype;
      std::vector<void *> boxes(slab.world);
      for (int k = 0; k < slab.world; ++k) boxes[k] = slab.peer[k] + slab.mail_off;
      FGB_ABI_THROW(fgb_slab_allreduce(ctx, r.op, rdtype, d_val, boxes.data(), slab.rank, slab.world, d_reduce_epoch, slot_ptr(slab.err_slot),
                                       cuda_config.slabTimeoutMs, st));
    }
    ++last_step_prefetched;
  }
  detail::k_publish_reductions<<<1, 32, 0, st>>>(d_prefetch, d_prefetch_vals, last_step_prefetched, d_ctrl, kStepSlot);
  ++own_launches;
}

// true: *out holds the result that the step which just ran has left for (agent, state, variable, op, dtype).  false: not
// recorded with that step; the request is remembered so that the following steps record it.
inline bool CUDASimulation::prefetched_result(const std::string &agent, const std::string &state, const std::string &variable, int op, int dtype,
                                              void *out, size_t bytes) {
  size_t k = 0;
  for (; k < prefetch.size(); ++k) {
    const PrefetchedReduction &r = prefetch[k];
    if (r.op == op && r.dtype == dtype && r.variable == variable && r.agent == agent && r.state == state) break;
  }
  if (k == prefetch.size()) {
    if (prefetch.size() < kMaxPrefetch) {
      if (!h_prefetch) {
        FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
        FGB_CUDA_THROW(cudaHostAlloc(&h_prefetch, (1 + kMaxPrefetch) * sizeof(unsigned long long), cudaHostAllocMapped));
        std::memset(h_prefetch, 0xFF, (1 + kMaxPrefetch) * sizeof(unsigned long long));
        FGB_CUDA_THROW(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_prefetch), h_prefetch, 0));
        FGB_CUDA_THROW(cudaMalloc(&d_prefetch_vals, kMaxPrefetch * sizeof(unsigned long long)));
      }
      prefetch.push_back(PrefetchedReduction{agent, state, variable, op, dtype});
    }
    return false;
  }
  if (k >= last_step_prefetched) return false;
  // the step (graph) that just ran publishes its epoch behind the results: spin on the host word; should the epoch not
  // show up (it always does unless the step failed), drain the stream and report the error through the normal path
  volatile unsigned long long *epoch = h_prefetch;
  for (unsigned long long spins = 0; *epoch != step_count; ++spins) {
    if ((spins & 0xFFFFFull) == 0xFFFFFull) {
      FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
      if (*epoch != step_count) return false;
    }
  }
  std::memcpy(out, const_cast<const unsigned long long *>(h_prefetch) + 1 + k, bytes);
  return true;
}

// ---- phase-wise execution for the multi-GPU slab driver (eager; exchanges happen between phases) ----
inline void CUDASimulation::runLayers(unsigned int first, unsigned int last) {
  initialise();
  if (env_dirty) upload_environment();
  if (first == 0) {
    plan_step();
    for (auto &m : messages) m.second.truncate = true;
  }
  record_layers(main_stream, first, last);
}
inline void CUDASimulation::endStep() {
  record_end_of_step(main_stream);
  ++step_count;
  if (model_has_births) refresh_bounds();
}

inline void CUDASimulation::endStepPipelined() {
  record_end_of_step(main_stream);
  ++step_count;
  if (!h_ctrl_pinned[0]) {
    for (int i = 0; i < 2; ++i) {
      FGB_CUDA_THROW(cudaMallocHost(&h_ctrl_pinned[i], kCtrlWords * 4));
      FGB_CUDA_THROW(cudaEventCreateWithFlags(&ctrl_events[i], cudaEventDisableTiming));
    }
  }
  const int cur = static_cast<int>(pipelined_steps & 1ull), prev = cur ^ 1;
  FGB_CUDA_THROW(cudaMemcpyAsync(h_ctrl_pinned[cur], d_ctrl, next_slot * 4, cudaMemcpyDeviceToHost, main_stream));
  FGB_CUDA_THROW(cudaEventRecord(ctrl_events[cur], main_stream));
  if (pipelined_steps > 0) {
    // counts as of the end of the PREVIOUS step (complete long ago: the device is busy with this step)
    FGB_CUDA_THROW(cudaEventSynchronize(ctrl_events[prev]));
    const unsigned int *h = h_ctrl_pinned[prev];
    for (auto &a : agents)
      for (auto &s : a.second.states) {
        detail::DevList &l = s.second;
        // everything appended since that snapshot is covered by appended_this_step (this step's appends)
        l.bound = std::min(l.capacity, h[l.count_slot] + l.appended_this_step);
        l.bound = std::max(l.bound, 1u);
      }
  }
  for (auto &a : agents) {
    for (auto &s : a.second.states) s.second.appended_this_step = 0;
    a.second.recompute_pop_bound();
  }
  for (auto &m : messages) m.second.list.appended_this_step = 0;
  ++pipelined_steps;
}

inline std::vector<std::pair<std::string, size_t>> CUDASimulation::listLayout(bool is_message, const std::string &name) {
  initialise();
  detail::DevList &l = is_message ? messages.at(name).list : state_list(name, agent_rt(name).desc->initial_state);
  std::vector<std::pair<std::string, size_t>> out;
  for (size_t v = 0; v < l.names.size(); ++v) out.emplace_back(l.names[v], l.meta[v].bytes());
  return out;
}

inline void CUDASimulation::slabPack(bool is_message, const std::string &name, const std::string &state,
                                     const std::string &geometry_message, int lo, int hi, void *const *dst_lo, void *const *dst_hi,
                                     unsigned int capacity, bool remove, unsigned int *d_count_lo, unsigned int *d_count_hi) {
  initialise();
  cudaStream_t xs = exchange_active ? index_stream : main_stream;
  const unsigned int xslot = exchange_active ? 1u : 0u;  // own look-back scratch: the main stream may be sorting
  detail::DevList &l = is_message ? messages.at(name).list : state_list(name, state);
  const detail::CUDAMessage &G = messages.at(geometry_message);
  const int slow = G.desc->dims() - 1;
  const char *axis = slow == 2 ? "z" : "y";
  const int ip = l.index_of(axis);
  if (ip < 0) throw exception::InvalidAgentVar(std::string("list has no position variable '") + axis + "'");
  const unsigned int n = l.bound;
  FGB_CUDA_THROW(cudaMemsetAsync(d_count_lo, 0, 4, xs));
  FGB_CUDA_THROW(cudaMemsetAsync(d_count_hi, 0, 4, xs));
  if (n == 0) return;
  unsigned int *d_n = slot_ptr(l.count_slot);
  for (auto &f : slab_flags) f.reserve(n);
  FGB_ABI_THROW(fgb_ctx_reserve(ctx, xslot, n, 0));
  FGB_ABI_THROW(fgb_plane_flags(ctx, reinterpret_cast<const float *>(l.data[ip]), n, d_n, G.md.min[slow], G.md.radius,
                                static_cast<int>(G.md.grid_dim[slow]), lo, hi, slab_flags[0].p, slab_flags[1].p, slab_flags[2].p,
                                xs));
  for (int side = 0; side < 2; ++side) {
    void *const *dst = side == 0 ? dst_lo : dst_hi;
    if (!dst) continue;
    // the destination buffers hold `capacity` items: writes are clamped, the caller checks the count
    std::vector<fgb_var> vars(l.names.size());
    for (size_t v = 0; v < vars.size(); ++v) {
      vars[v].type_len = l.meta[v].bytes();
      vars[v].in = l.data[v];
      vars[v].out = dst[v];
    }
    FGB_ABI_THROW(fgb_compact_limited(ctx, xslot, slab_flags[side == 0 ? 0 : 2].p, 0, n, d_n, 0, 0, nullptr, capacity, vars.data(),
                                      static_cast<unsigned int>(vars.size()), side == 0 ? d_count_lo : d_count_hi, nullptr, xs));
  }
  if (remove) {
    std::vector<fgb_var> vars = l.vars(true);
    FGB_ABI_THROW(fgb_compact(ctx, xslot, slab_flags[1].p, 0, n, d_n, 0, 0, nullptr, vars.data(), static_cast<unsigned int>(vars.size()),
                              nullptr, d_n, xs));
    l.swap_buffers();
  }
}

inline void CUDASimulation::listAppend(bool is_message, const std::string &name, const std::string &state, unsigned int n_max,
                                       const unsigned int *d_n_src, const void *const *src) {
  initialise();
  if (n_max == 0) return;
  cudaStream_t xs = exchange_active ? index_stream : main_stream;
  const unsigned int xslot = exchange_active ? 1u : 0u;
  detail::DevList &l = is_message ? messages.at(name).list : state_list(name, state);
  unsigned int *d_n = slot_ptr(l.count_slot);
  if (l.bound + n_max > l.capacity) {
    FGB_CUDA_THROW(cudaDeviceSynchronize());
    l.reserve(l.bound + n_max, l.capacity);
  }
  std::vector<fgb_var> vars(l.names.size());
  for (size_t v = 0; v < vars.size(); ++v) {
    vars[v].type_len = l.meta[v].bytes();
    vars[v].in = src[v];
    vars[v].out = l.data[v];
  }
  // copy-all compaction: keep_front == n_max keeps every item below the device count; offset and the new
  // size are the list's own device count word
  FGB_ABI_THROW(fgb_ctx_reserve(ctx, xslot, n_max, 0));
  FGB_ABI_THROW(fgb_compact(ctx, xslot, nullptr, 0, n_max, d_n_src, n_max, 0, d_n, vars.data(), static_cast<unsigned int>(vars.size()),
                            nullptr, d_n, xs));
  l.bound += n_max;
  l.touch();
  l.appended_this_step += n_max;
  if (!is_message) agent_rt(name).pop_bound += n_max;
  if (is_message) messages.at(name).pbm_dirty = true;
}

inline void CUDASimulation::beginMessageExchange() {
  initialise();
  FGB_CUDA_THROW(cudaEventRecord(index_fork, main_stream));
  FGB_CUDA_THROW(cudaStreamWaitEvent(index_stream, index_fork, 0));
  exchange_active = true;
}

inline void CUDASimulation::endMessageExchange(const std::string &message) {
  detail::CUDAMessage &M = messages.at(message);
  if (M.spatial && M.pbm_dirty) build_input_index(M, index_stream);
  FGB_CUDA_THROW(cudaEventRecord(index_done, index_stream));
  index_pending = true;  // consumed (and cleared) by the next record_layers
  exchange_active = false;
}

// ================================================================================================================
// multi-GPU: z-slab decomposition behind step()
// ================================================================================================================
inline void CUDASimulation::configureSlabs(int rank, int world, const std::string &message, unsigned int halo_capacity,
                                           unsigned int migrate_capacity) {
  if (initialised) throw exception::InvalidArgument("configureSlabs must be called before the simulation is initialised");
  if (world < 1 || rank < 0 || rank >= world || world > 32) throw exception::InvalidArgument("configureSlabs: bad rank / world");
  auto it = model->messages.find(message);
  if (it == model->messages.end() || it->second->dims() == 0) throw exception::InvalidMessageName("configureSlabs: '" + message + "' is not a spatial message list");
  const MessageData &md = *it->second;
  const int slow = md.dims() - 1;
  const int planes = static_cast<int>(std::ceil((md.max[slow] - md.min[slow]) / md.radius));
  if (world > planes) throw exception::InvalidArgument("configureSlabs: more ranks than bin planes");
  slab.enabled = world > 1;
  slab.rank = rank;
  slab.world = world;
  slab.message = message;
  slab.planes = planes;
  slab.z0 = static_cast<int>((static_cast<long long>(planes) * rank) / world);         // contiguous, as even as possible
  slab.z1 = static_cast<int>((static_cast<long long>(planes) * (rank + 1)) / world);
  slab.halo_cap = halo_capacity;
  slab.mig_cap = std::min(migrate_capacity, 131072u);  // fgb_slab_migrate_out pairs holes in a 262144-bit shared-memory bitmap
  if (slab.enabled) {
    const int w0 = std::max(slab.z0 - 1, 0), w1 = std::min(slab.z1 + 1, planes);  // own planes + one ghost plane per side
    windows[message] = std::make_pair(w0, w1 - w0);
  }
}

inline void CUDASimulation::slab_setup() {
  detail::CUDAMessage &M = messages.at(slab.message);
  const char *axis = M.desc->dims() == 3 ? "z" : "y";
  auto add = [&](bool is_message, detail::DevList &l, detail::CUDAAgent *agent, unsigned int cap) {
    const int pv = l.index_of(axis);
    if (pv < 0 || l.meta[pv].type != std::type_index(typeid(float)) || l.meta[pv].elements != 1) return;
    SlabList S;
    S.is_message = is_message;
    S.list = &l;
    S.agent = agent;
    S.pos_var = pv;
    S.capacity = cap;
    slab.lists.push_back(S);
  };
  add(true, M.list, nullptr, slab.halo_cap);
  if (slab.lists.empty()) throw exception::InvalidMessageVar("slab message has no float position variable");
  for (auto &a : agents)
    for (auto &st : a.second.states) add(false, st.second, &a.second, slab.mig_cap);
  size_t off = 0;
  auto align = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  for (SlabList &S : slab.lists)
    for (int side = 0; side < 2; ++side) {
      SlabStaging &g = S.st[side];
      for (size_t v = 0; v < S.list->names.size(); ++v) {
        g.var_off.push_back(off);
        off = align(off + static_cast<size_t>(S.capacity) * S.list->meta[v].bytes());
      }
      g.count_off = off;
      g.flag_off = off + 8;
      off = align(off + 16);
    }
  slab.mail_off = off;
  off = align(off + static_cast<size_t>(2) * slab.world * 16);
  slab.arena_bytes = off;
  FGB_CUDA_THROW(cudaMalloc(&slab.arena, slab.arena_bytes));
  FGB_CUDA_THROW(cudaMemset(slab.arena, 0, slab.arena_bytes));
  slab.peer.assign(slab.world, nullptr);
  slab.peer[slab.rank] = slab.arena;
  slab.epoch_slot = alloc_slot();
  slab.err_slot = alloc_slot();
}

inline void CUDASimulation::slabExportHandle(void *out) {
  initialise();
  if (!slab.enabled) throw exception::InvalidArgument("slab decomposition is not configured");
  FGB_CUDA_THROW(cudaDeviceSynchronize());  // the arena is zeroed before any peer can see it
  cudaIpcMemHandle_t h;
  FGB_CUDA_THROW(cudaIpcGetMemHandle(&h, slab.arena));
  std::memcpy(out, &h, sizeof(h));
}

inline void CUDASimulation::slabConnect(const void *all_handles) {
  initialise();
  if (!slab.enabled) throw exception::InvalidArgument("slab decomposition is not configured");
  const char *p = static_cast<const char *>(all_handles);
  for (int r = 0; r < slab.world; ++r) {
    if (r == slab.rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, p + static_cast<size_t>(r) * sizeof(h), sizeof(h));
    void *mapped = nullptr;
    FGB_CUDA_THROW(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
    slab.peer[r] = static_cast<char *>(mapped);
  }
  slab.connected = true;
}

inline unsigned int CUDASimulation::slabError() {
  initialise();
  return slab.enabled ? read_slot(slab.err_slot) : 0u;
}

// One exchange of a list with both neighbours: items below plane `lo_plane` go to rank-1, items at or above `hi_plane`
// to rank+1 (the compaction writes them -- and their count -- into the neighbour's staging buffer), optionally removed
// here; then the neighbours' items are appended.  Every rank runs the same sequence every step (also with an empty
// list: the neighbours wait for the flag).
inline void CUDASimulation::slab_exchange(SlabList &S, int lo_plane, int hi_plane, bool remove, cudaStream_t st) {
  detail::DevList &l = *S.list;
  const detail::CUDAMessage &G = messages.at(slab.message);
  const int slow = G.desc->dims() - 1;
  const unsigned int n = l.bound;
  unsigned int *d_n = slot_ptr(l.count_slot);
  const unsigned int *d_epoch = slot_ptr(slab.epoch_slot);
  unsigned int *d_err = slot_ptr(slab.err_slot);
  const bool has[2] = {slab.rank > 0, slab.rank < slab.world - 1};
  const unsigned int nv = static_cast<unsigned int>(l.names.size());
  // halo: a slab of >= 2 planes sends every boundary message to ONE side, so the messages are selected with one read of the
  // position column and gathered by index (fgb_slab_pack_planes); a one-plane slab keeps the flag + compaction form
  const bool gather_halo = !remove && slab.z1 - slab.z0 >= 2;
  if (n > 0 && !remove && !gather_halo) {
    for (auto &f : slab_flags) f.reserve(n);
    FGB_ABI_THROW(fgb_plane_flags(ctx, reinterpret_cast<const float *>(l.data[S.pos_var]), n, d_n, G.md.min[slow], G.md.radius,
                                  static_cast<int>(G.md.grid_dim[slow]), lo_plane, hi_plane, slab_flags[0].p, nullptr, slab_flags[2].p, st));
  }
  unsigned long long *peer_flag[2] = {nullptr, nullptr};
  // migration edits the agent list in place: a valid tile-local execution order survives below min(count before, after)
  const bool keep_perm = remove && cuda_config.binOrderedOutput && l.perm_valid() && l.perm_limit_slot != 0;
  if (keep_perm) {
    detail::k_perm_limit<<<1, 1, 0, st>>>(slot_ptr(l.perm_limit_slot), d_n, l.perm_partial ? 1 : 0, 0);
    ++own_launches;
  }
  if (remove) {
    // migration: only a few agents leave per step, so the list is not rewritten: the leavers are gathered by index into the
    // neighbours' staging buffers and their holes are filled from the tail (fgb_slab_migrate_out)
    std::vector<void *> peer_cols[2];
    unsigned int *peer_count[2] = {nullptr, nullptr};
    for (int side = 0; side < 2; ++side) {
      if (!has[side]) continue;
      char *base = slab.peer[slab.rank + (side == 0 ? -1 : 1)];
      const SlabStaging &g = S.st[side == 0 ? 1 : 0];
      for (unsigned int v = 0; v < nv; ++v) peer_cols[side].push_back(base + g.var_off[v]);
      peer_count[side] = reinterpret_cast<unsigned int *>(base + g.count_off);
      peer_flag[side] = reinterpret_cast<unsigned long long *>(base + g.flag_off);
    }
    std::vector<fgb_var> cols(nv);
    for (unsigned int v = 0; v < nv; ++v) {
      cols[v].type_len = l.meta[v].bytes();
      cols[v].in = l.data[v];
      cols[v].out = l.data[v];
    }
    FGB_ABI_THROW(fgb_slab_migrate_out(ctx, kSlabScratchSlot, reinterpret_cast<const float *>(l.data[S.pos_var]), n, d_n, G.md.min[slow], G.md.radius,
                                       static_cast<int>(G.md.grid_dim[slow]), lo_plane, hi_plane, S.capacity, cols.data(), nv,
                                       has[0] ? peer_cols[0].data() : nullptr, has[1] ? peer_cols[1].data() : nullptr, peer_count[0], peer_count[1],
                                       d_n, d_err, st));
  }
  if (gather_halo) {
    std::vector<void *> peer_cols[2];
    unsigned int *peer_count[2] = {nullptr, nullptr};
    for (int side = 0; side < 2; ++side) {
      if (!has[side]) continue;
      // I am the neighbour's OTHER side: what I send down arrives in rank-1's "from rank+1" buffer and vice versa
      char *base = slab.peer[slab.rank + (side == 0 ? -1 : 1)];
      const SlabStaging &g = S.st[side == 0 ? 1 : 0];
      for (unsigned int v = 0; v < nv; ++v) peer_cols[side].push_back(base + g.var_off[v]);
      peer_count[side] = reinterpret_cast<unsigned int *>(base + g.count_off);
      peer_flag[side] = reinterpret_cast<unsigned long long *>(base + g.flag_off);
    }
    std::vector<fgb_var> cols(nv);
    for (unsigned int v = 0; v < nv; ++v) {
      cols[v].type_len = l.meta[v].bytes();
      cols[v].in = l.data[v];
      cols[v].out = l.data[v];
    }
    // without a neighbour on a side nothing is selected for it: the plane range is opened up (planes are clamped to the grid)
    FGB_ABI_THROW(fgb_slab_pack_planes(ctx, kSlabScratchSlot, reinterpret_cast<const float *>(l.data[S.pos_var]), n, d_n, G.md.min[slow], G.md.radius,
                                       static_cast<int>(G.md.grid_dim[slow]), has[0] ? lo_plane : 0, has[1] ? hi_plane : static_cast<int>(G.md.grid_dim[slow]),
                                       S.capacity, cols.data(), nv, has[0] ? peer_cols[0].data() : nullptr, has[1] ? peer_cols[1].data() : nullptr,
                                       peer_count[0], peer_count[1], st));
  }
  for (int side = 0; side < 2 && !remove && !gather_halo; ++side) {
    if (!has[side]) continue;
    // I am the neighbour's OTHER side: what I send down arrives in rank-1's "from rank+1" buffer and vice versa
    char *base = slab.peer[slab.rank + (side == 0 ? -1 : 1)];
    const SlabStaging &g = S.st[side == 0 ? 1 : 0];
    std::vector<fgb_var> vars(nv);
    for (unsigned int v = 0; v < nv; ++v) {
      vars[v].type_len = l.meta[v].bytes();
      vars[v].in = l.data[v];
      vars[v].out = base + g.var_off[v];
    }
    FGB_ABI_THROW(fgb_compact_limited(ctx, kSlabScratchSlot, slab_flags[side == 0 ? 0 : 2].p, 0, n, d_n, 0, 0, nullptr, S.capacity, vars.data(), nv,
                                      reinterpret_cast<unsigned int *>(base + g.count_off), nullptr, st));
    peer_flag[side] = reinterpret_cast<unsigned long long *>(base + g.flag_off);
  }
  FGB_ABI_THROW(fgb_slab_signal(ctx, peer_flag[0], peer_flag[1], d_epoch, st));
  const SlabStaging &from_lo = S.st[0], &from_hi = S.st[1];
  FGB_ABI_THROW(fgb_slab_wait(ctx, has[0] ? reinterpret_cast<const unsigned long long *>(slab.arena + from_lo.flag_off) : nullptr,
                              has[1] ? reinterpret_cast<const unsigned long long *>(slab.arena + from_hi.flag_off) : nullptr,
                              has[0] ? reinterpret_cast<const unsigned int *>(slab.arena + from_lo.count_off) : nullptr,
                              has[1] ? reinterpret_cast<const unsigned int *>(slab.arena + from_hi.count_off) : nullptr, S.capacity, d_epoch, d_err,
                              cuda_config.slabTimeoutMs, st));
  for (int side = 0; side < 2; ++side) {
    if (!has[side]) continue;
    const SlabStaging &g = S.st[side];
    std::vector<fgb_var> vars(nv);
    for (unsigned int v = 0; v < nv; ++v) {
      vars[v].type_len = l.meta[v].bytes();
      vars[v].in = slab.arena + g.var_off[v];
      vars[v].out = l.data[v];
    }
    // copy-all append: keep_front == capacity keeps every item below the received count (clamped to the capacity)
    FGB_ABI_THROW(fgb_compact(ctx, kSlabScratchSlot, nullptr, 0, S.capacity, reinterpret_cast<const unsigned int *>(slab.arena + g.count_off),
                              S.capacity, 0, d_n, vars.data(), nv, nullptr, d_n, st));
  }
  if (S.is_message) {
    l.bound += (has[0] ? S.capacity : 0u) + (has[1] ? S.capacity : 0u);
    for (auto &m : messages)
      if (&m.second.list == &l) m.second.pbm_dirty = true;
  }
  // agent lists keep their (sticky) launch bound: fgb_slab_check_bound raises the error word if the population outgrows
  // it before the next slab_refresh_bounds()
  l.touch();
  if (keep_perm) {
    detail::k_perm_limit<<<1, 1, 0, st>>>(slot_ptr(l.perm_limit_slot), d_n, 1, 1);
    ++own_launches;
    l.cached_perm_version = l.order_version;
    l.perm_partial = true;
  }
}

// Re-read the list counts (one small copy + sync every slabRefreshPeriod steps) and re-centre the sticky launch bounds
// of the decomposed agent lists: generous enough to absorb the net inflow until the next refresh, and unchanged as
// long as they still fit, so that the captured step graphs are reused.
inline void CUDASimulation::slab_refresh_bounds() {
  std::vector<unsigned int> h(kCtrlWords);
  FGB_CUDA_THROW(cudaMemcpyAsync(h.data(), d_ctrl, next_slot * 4, cudaMemcpyDeviceToHost, main_stream));
  FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
  if (h[slab.err_slot]) throw exception::CUDAError("slab exchange failed on the device: error bits " + std::to_string(h[slab.err_slot]) +
                                                   " (1 neighbour timeout, 2 staging overflow, 4 list outgrew its launch bound)");
  const unsigned int period = std::max(1u, cuda_config.slabRefreshPeriod);
  for (size_t k = 1; k < slab.lists.size(); ++k) {
    detail::DevList &l = *slab.lists[k].list;
    const unsigned int count = h[l.count_slot], cap = slab.lists[k].capacity;
    const unsigned int want = quantise(count + 8u * cap + count / 32u);
    const bool too_small = static_cast<unsigned long long>(count) + 2ull * cap * period > l.bound;
    const bool too_big = l.bound > want + want / 2u;
    if (too_small || too_big) l.bound = want;
  }
  for (auto &a : agents) a.second.recompute_pop_bound();
}

inline void CUDASimulation::slab_allreduce(void *d_value, int dtype, int op) {
  std::vector<void *> boxes(slab.world);
  for (int r = 0; r < slab.world; ++r) boxes[r] = slab.peer[r] + slab.mail_off;
  FGB_ABI_THROW(fgb_slab_allreduce(ctx, op, dtype, d_value, boxes.data(), slab.rank, slab.world, d_reduce_epoch, slot_ptr(slab.err_slot),
                                   cuda_config.slabTimeoutMs, main_stream));
}

inline void CUDASimulation::refresh_bounds() {
  std::vector<unsigned int> h(kCtrlWords);
  FGB_CUDA_THROW(cudaMemcpyAsync(h.data(), d_ctrl, next_slot * 4, cudaMemcpyDeviceToHost, main_stream));
  FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
  for (auto &a : agents) {
    a.second.host_next_id = h[a.second.next_id_slot];
    for (auto &s : a.second.states) s.second.bound = quantise(h[s.second.count_slot]);
    a.second.recompute_pop_bound();
  }
}

inline bool CUDASimulation::step() {
  initialise();
  if (env_dirty) upload_environment();
  if (slab.enabled) {
    if (!slab.connected) throw exception::InvalidArgument("slab decomposition: slabConnect() must be called before the first step");
    // the launch bounds are re-centred from the device counts every slabRefreshPeriod steps -- and after each of the first
    // steps, when the initial transient (the first migrations) settles, so that the one re-capture it causes happens early
    if (step_count < 4 || step_count % std::max(1u, cuda_config.slabRefreshPeriod) == 0) slab_refresh_bounds();
  }
  plan_step();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (config.timing) {
    // events come from a pool and go back to it when their times are harvested (every 256 steps at the latest), so a
    // long run neither creates two events per step nor keeps one pair per step alive
    if (step_events.size() >= 256) harvest_step_events(/*only_finished=*/true);
    e0 = take_event();
    e1 = take_event();
    FGB_CUDA_THROW(cudaEventRecord(e0, main_stream));
  }
  if (cuda_config.useCUDAGraphs && !model_has_host_layers && !cuda_config.profile && !streamed.armed) {
    // the key covers list pointers, bounds and capacities; grow-only scratch (scan flags, exec_perm, the kernel
    // library's per-stream scratch and per-list index buffers) is covered by the allocation generation: any
    // reallocation since the capture invalidates every cached graph (they hold the freed pointers)
    const unsigned long long gen = alloc_gen + fgb_alloc_generation(ctx);
    if (gen != graphs_generation) {
      for (auto &g : graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
      graphs.clear();
      graphs_generation = gen;
    }
    const std::vector<unsigned long long> key = graph_key();
    GraphEntry *hit = nullptr;
    for (auto &g : graphs)
      if (g.key == key) {
        hit = &g;
        break;
      }
    if (!hit) {
      const unsigned long long l0 = getLaunchCount();
      cudaGraph_t graph = nullptr;
      FGB_CUDA_THROW(cudaStreamBeginCapture(main_stream, cudaStreamCaptureModeThreadLocal));
      try {
        record_step(main_stream);
      } catch (...) {
        cudaStreamEndCapture(main_stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      FGB_CUDA_THROW(cudaStreamEndCapture(main_stream, &graph));
      last_graph_width = graph_width(graph);
      GraphEntry g;
      g.key = key;
      FGB_CUDA_THROW(cudaGraphInstantiate(&g.exec, graph, 0));
      cudaGraphDestroy(graph);
      g.launches = getLaunchCount() - l0;
      g.prefetched = last_step_prefetched;
      g.post_state = snapshot_host_state();
      if (graphs.size() >= 64) {  // bounded cache
        cudaGraphExecDestroy(graphs.front().exec);
        graphs.erase(graphs.begin());
      }
      graphs.push_back(std::move(g));
      hit = &graphs.back();
    } else {
      restore_host_state(hit->post_state);
      own_launches += hit->launches;
      last_step_prefetched = hit->prefetched;
    }
    FGB_CUDA_THROW(cudaGraphLaunch(hit->exec, main_stream));
  } else {
    record_step(main_stream);
  }
  streamed.armed = false;  // a streamed download covers exactly one step
  ++step_count;
  if (!model->step_functions.empty()) {
    // step functions see the finished step (reference CUDASimulation.cu:603-617 runs them inside step(), inside its
    // per-step timer).  Everything HostAPI reads from the device goes through main_stream (reductions, counts, agent
    // data) and synchronises that stream itself when the value is handed to the host, so the step is NOT drained
    // first: a reduction's kernel is queued right behind the step's graph and the GPU never idles for a launch latency
    in_step_function = true;
    try {
      for (auto sf : model->step_functions) sf(&host_api);
    } catch (...) {
      in_step_function = false;
      throw;
    }
    in_step_function = false;
    flush_host_agents();
  }
  if (config.timing) {
    FGB_CUDA_THROW(cudaEventRecord(e1, main_stream));
    step_events.emplace_back(e0, e1);
  }
  if (model_has_births) refresh_bounds();
  bool go_on = true;
  for (auto ec : model->exit_conditions) {
    FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
    if (ec(&host_api)) go_on = false;
  }
  if (!model->exit_conditions.empty()) {
    if (go_on) {
      flush_host_agents();
    } else {  // reference test_host_agent_creation.cu:24-29: agents made alongside an EXIT verdict are not created
      for (auto &kv : host_new_agents) {
        kv.second.count = 0;
        kv.second.data.clear();
      }
    }
  }
  return go_on;
}

// Widest level of a captured graph: nodes are levelled by their longest path from a root; the width is the largest
// number of KERNEL nodes on one level.  1 = a chain; functions of one layer on their own streams show up as > 1.
inline unsigned int CUDASimulation::graph_width(cudaGraph_t graph) {
  size_t nn = 0, ne = 0;
  if (cudaGraphGetNodes(graph, nullptr, &nn) != cudaSuccess || nn == 0) return 0;
  std::vector<cudaGraphNode_t> nodes(nn);
  cudaGraphGetNodes(graph, nodes.data(), &nn);
  cudaGraphGetEdges(graph, nullptr, nullptr, &ne);
  std::vector<cudaGraphNode_t> from(ne), to(ne);
  if (ne) cudaGraphGetEdges(graph, from.data(), to.data(), &ne);
  std::map<cudaGraphNode_t, size_t> id;
  for (size_t i = 0; i < nn; ++i) id[nodes[i]] = i;
  std::vector<unsigned int> depth(nn, 0);
  for (size_t pass = 0; pass < nn; ++pass) {  // longest-path relaxation (graphs of a step have tens of nodes)
    bool changed = false;
    for (size_t e = 0; e < ne; ++e) {
      const size_t a = id[from[e]], b = id[to[e]];
      if (depth[b] < depth[a] + 1) {
        depth[b] = depth[a] + 1;
        changed = true;
      }
    }
    if (!changed) break;
  }
  std::map<unsigned int, unsigned int> per_level;
  unsigned int width = 0;
  for (size_t i = 0; i < nn; ++i) {
    cudaGraphNodeType t;
    if (cudaGraphNodeGetType(nodes[i], &t) != cudaSuccess || t != cudaGraphNodeTypeKernel) continue;
    width = std::max(width, ++per_level[depth[i]]);
  }
  return width;
}

inline void CUDASimulation::simulate() {
  initialise();
  const auto t0 = std::chrono::steady_clock::now();
  for (auto f : model->init_functions) f(&host_api);
  flush_host_agents();
  for (unsigned int i = 0; config.steps == 0 || i < config.steps; ++i)
    if (!step()) break;
  FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
  for (auto f : model->exit_functions) f(&host_api);
  elapsed_simulation = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (config.timing) std::printf("Total Processing time: %.3f ms\n", elapsed_simulation * 1e3);
}

inline cudaEvent_t CUDASimulation::take_event() {
  if (!event_pool.empty()) {
    cudaEvent_t e = event_pool.back();
    event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  FGB_CUDA_THROW(cudaEventCreate(&e));
  return e;
}

// convert recorded (start, end) pairs into seconds, oldest first, and recycle the events; only_finished stops at the
// first pair whose end event has not completed yet (never blocks)
inline void CUDASimulation::harvest_step_events(bool only_finished) {
  size_t done = 0;
  for (auto &e : step_events) {
    if (only_finished && cudaEventQuery(e.second) != cudaSuccess) break;
    float ms = 0.f;
    FGB_CUDA_THROW(cudaEventElapsedTime(&ms, e.first, e.second));
    step_seconds.push_back(ms * 1e-3);
    event_pool.push_back(e.first);
    event_pool.push_back(e.second);
    ++done;
  }
  step_events.erase(step_events.begin(), step_events.begin() + static_cast<std::ptrdiff_t>(done));
}

inline std::vector<double> CUDASimulation::getElapsedTimeSteps() {
  if (initialised) FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
  harvest_step_events(/*only_finished=*/false);
  return step_seconds;
}

inline std::map<std::string, std::pair<double, unsigned int>> CUDASimulation::getProfile() {
  std::map<std::string, std::pair<double, unsigned int>> out;
  if (initialised) FGB_CUDA_THROW(cudaDeviceSynchronize());
  for (auto &r : prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      out[r.name].first += ms;
      out[r.name].second += 1;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  prof.clear();
  return out;
}

// ---- host agent creation ---------------------------------------------------------------------
inline detail::HostNewAgents &CUDASimulation::host_new_buffer(const std::string &agent_name, const std::string &state) {
  initialise();
  auto key = std::make_pair(agent_name, state);
  auto it = host_new_agents.find(key);
  if (it != host_new_agents.end()) return it->second;
  detail::DevList &l = state_list(agent_name, state);
  detail::HostNewAgents b;
  size_t off = 0;
  for (size_t v = 0; v < l.names.size(); ++v) {
    const size_t bytes = l.meta[v].bytes();
    const size_t align = std::min<size_t>(16, bytes & (~bytes + 1));  // largest power of two dividing the size, at most 16
    off = (off + align - 1) / align * align;
    b.names.push_back(l.names[v]);
    b.meta.push_back(l.meta[v]);
    b.offset.push_back(off);
    off += bytes;
  }
  b.agent_size = (off + 15) & ~static_cast<size_t>(15);
  b.defaults.assign(b.agent_size, 0);
  for (size_t v = 0; v < b.names.size(); ++v)
    if (!b.meta[v].default_value.empty()) std::memcpy(b.defaults.data() + b.offset[v], b.meta[v].default_value.data(), b.meta[v].bytes());
  return host_new_agents.emplace(key, std::move(b)).first->second;
}

inline void CUDASimulation::flush_host_agents() {
  for (auto &kv : host_new_agents) {
    detail::HostNewAgents &b = kv.second;
    if (b.count == 0) continue;
    detail::CUDAAgent &a = agent_rt(kv.first.first);
    detail::DevList &l = state_list(kv.first.first, kv.first.second);
    FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));
    // ids: the next free ones of the agent type (device counter: device births may have advanced it)
    id_t next = read_slot(a.next_id_slot);
    const int iv = l.index_of(ID_VARIABLE_NAME);
    for (unsigned int i = 0; i < b.count; ++i) *reinterpret_cast<id_t *>(b.data.data() + static_cast<size_t>(i) * b.agent_size + b.offset[iv]) = next++;
    a.host_next_id = next;
    write_slot(a.next_id_slot, next);
    const unsigned int have = read_slot(l.count_slot);
    l.reserve(have + b.count, have);
    const size_t bytes = static_cast<size_t>(b.count) * b.agent_size;
    if (bytes > new_aos_bytes) {
      if (d_new_aos) cudaFree(d_new_aos);
      FGB_CUDA_THROW(cudaMalloc(&d_new_aos, bytes + bytes / 2));
      new_aos_bytes = bytes + bytes / 2;
    }
    FGB_CUDA_THROW(cudaMemcpyAsync(d_new_aos, b.data.data(), bytes, cudaMemcpyHostToDevice, main_stream));
    std::vector<fgb_var> vars(l.names.size());
    for (size_t v = 0; v < vars.size(); ++v) {
      vars[v].type_len = l.meta[v].bytes();
      vars[v].in = d_new_aos + b.offset[v];
      vars[v].out = l.data[v];
    }
    FGB_ABI_THROW(fgb_scatter_new_agents(ctx, d_new_aos, static_cast<unsigned int>(b.agent_size), vars.data(), static_cast<unsigned int>(vars.size()),
                                         b.count, have, nullptr, main_stream));
    FGB_CUDA_THROW(cudaStreamSynchronize(main_stream));  // host creation is not a hot path: the staging vector is reused below
    write_slot(l.count_slot, have + b.count);
    l.bound = std::max(l.bound, model_has_births ? quantise(have + b.count) : have + b.count);
    l.touch();
    a.recompute_pop_bound();
    b.count = 0;
    b.data.clear();
  }
}

inline HostNewAgentAPI HostAgentAPI::newAgent() {
  detail::HostNewAgents &b = sim->host_new_buffer(agent, state);
  b.data.insert(b.data.end(), b.defaults.begin(), b.defaults.end());
  return HostNewAgentAPI(&b, b.count++);
}

// ---- minimal HostAPI -------------------------------------------------------------------------
inline unsigned int HostAPI::getStepCounter() const { return sim->getStepCounter(); }
inline unsigned int HostAgentAPI::count() {
  const unsigned int local = sim->getAgentCount(agent, state);
  if (!sim->slab.enabled) return local;
  unsigned long long v = local;  // global population: all-reduce of the per-slab counts
  FGB_CUDA_THROW(cudaMemcpyAsync(sim->d_reduce_out, &v, 8, cudaMemcpyHostToDevice, sim->main_stream));
  sim->slab_allreduce(sim->d_reduce_out, FGB_U64, FGB_REDUCE_SUM);
  FGB_CUDA_THROW(cudaMemcpyAsync(&v, sim->d_reduce_out, 8, cudaMemcpyDeviceToHost, sim->main_stream));
  FGB_CUDA_THROW(cudaStreamSynchronize(sim->main_stream));
  return static_cast<unsigned int>(v);
}
namespace detail {
template <typename T> struct reduce_dtype;
template <> struct reduce_dtype<float> { static constexpr int value = FGB_F32; using sum_t = double; };
template <> struct reduce_dtype<double> { static constexpr int value = FGB_F64; using sum_t = double; };
template <> struct reduce_dtype<int> { static constexpr int value = FGB_I32; using sum_t = long long; };
template <> struct reduce_dtype<unsigned int> { static constexpr int value = FGB_U32; using sum_t = unsigned long long; };
template <> struct reduce_dtype<long long> { static constexpr int value = FGB_I64; using sum_t = long long; };
template <> struct reduce_dtype<unsigned long long> { static constexpr int value = FGB_U64; using sum_t = unsigned long long; };
}  // namespace detail
// under a slab decomposition every rank holds a part of the population: fold the per-rank results (8 bytes each)
#define FGB_SLAB_FOLD(R, op)                                                                                  \
  if (sim->slab.enabled) sim->slab_allreduce(sim->d_reduce_out, detail::reduce_dtype<R>::value, (op))

// one reduction kernel over the state list's variable (device-resident count), then 8 bytes to the host
template <typename T, typename R>
inline R HostAgentAPI::reduce(const std::string &variable, int op) {
  sim->initialise();
  detail::DevList &l = sim->state_list(agent, state);
  const int i = l.index_of(variable);
  if (i < 0) throw exception::InvalidAgentVar("agent '" + agent + "' has no variable '" + variable + "'");
  if (l.meta[i].type != std::type_index(typeid(T)) || l.meta[i].elements != 1) throw exception::InvalidVarType("wrong type for '" + variable + "'");
  R r{};
  static_assert(sizeof(R) <= 8, "reduction results travel in an 8-byte slot");
  // recorded behind the step that just ran (learned from an earlier step's step function)?  then the value is already in
  // mapped host memory; the request is valid only for the list as the step left it
  if (sim->in_step_function && sim->prefetched_result(agent, state, variable, op, detail::reduce_dtype<T>::value, &r, sizeof(R))) return r;
  FGB_ABI_THROW(fgb_reduce(sim->ctx, 0, op, detail::reduce_dtype<T>::value, l.data[i], l.bound, sim->slot_ptr(l.count_slot),
                           sim->d_reduce_out, sim->main_stream));
  FGB_SLAB_FOLD(R, op);
  FGB_CUDA_THROW(cudaMemcpyAsync(&r, sim->d_reduce_out, sizeof(R), cudaMemcpyDeviceToHost, sim->main_stream));
  FGB_CUDA_THROW(cudaStreamSynchronize(sim->main_stream));
  return r;
}
template <typename T, typename R>
inline R HostAgentAPI::transform_reduce(const std::string &variable, int transform, const void *param) {
  sim->initialise();
  detail::DevList &l = sim->state_list(agent, state);
  const int i = l.index_of(variable);
  if (i < 0) throw exception::InvalidAgentVar("agent '" + agent + "' has no variable '" + variable + "'");
  if (l.meta[i].type != std::type_index(typeid(T)) || l.meta[i].elements != 1) throw exception::InvalidVarType("wrong type for '" + variable + "'");
  FGB_ABI_THROW(fgb_transform_reduce(sim->ctx, 0, transform, detail::reduce_dtype<T>::value, l.data[i], l.bound,
                                     sim->slot_ptr(l.count_slot), param, sim->d_reduce_out, sim->main_stream));
  FGB_SLAB_FOLD(R, FGB_REDUCE_SUM);
  R r{};
  FGB_CUDA_THROW(cudaMemcpyAsync(&r, sim->d_reduce_out, sizeof(R), cudaMemcpyDeviceToHost, sim->main_stream));
  FGB_CUDA_THROW(cudaStreamSynchronize(sim->main_stream));
  return r;
}
template <typename T>
inline unsigned int HostAgentAPI::count(const std::string &variable, T value) {
  return static_cast<unsigned int>(transform_reduce<T, unsigned long long>(variable, FGB_TRANSFORM_COUNT_EQUAL, &value));
}
template <typename T>
inline std::pair<double, double> HostAgentAPI::meanStandardDeviation(const std::string &variable) {
  const unsigned int n = count();
  if (n == 0) return std::make_pair(0.0, 0.0);
  const double mean = static_cast<double>(reduce<T, typename detail::reduce_dtype<T>::sum_t>(variable, FGB_REDUCE_SUM)) / static_cast<double>(n);
  const double ss = transform_reduce<T, double>(variable, FGB_TRANSFORM_SUM_SQ_DEV, &mean);
  return std::make_pair(mean, std::sqrt(ss / static_cast<double>(n)));
}
template <typename InT, typename OutT>
inline std::vector<OutT> HostAgentAPI::histogramEven(const std::string &variable, unsigned int histogramBins, InT lowerBound, InT upperBound) {
  sim->initialise();
  if (!(lowerBound < upperBound)) throw exception::InvalidArgument("lowerBound must be lower than upperBound in HostAgentAPI::histogramEven()");
  if (sim->slab.enabled) throw exception::UnsupportedFeature("histogramEven under a slab decomposition");
  detail::DevList &l = sim->state_list(agent, state);
  const int i = l.index_of(variable);
  if (i < 0) throw exception::InvalidAgentVar("agent '" + agent + "' has no variable '" + variable + "'");
  if (l.meta[i].elements != 1) throw exception::UnsupportedVarType("HostAgentAPI::histogramEven() does not support agent array variables");
  if (l.meta[i].type != std::type_index(typeid(InT))) throw exception::InvalidVarType("wrong type for '" + variable + "'");
  if (histogramBins > sim->hist_cap) {
    if (sim->d_hist_out) cudaFree(sim->d_hist_out);
    FGB_CUDA_THROW(cudaMalloc(&sim->d_hist_out, static_cast<size_t>(histogramBins) * 4));
    sim->hist_cap = histogramBins;
  }
  FGB_ABI_THROW(fgb_histogram_even(sim->ctx, detail::reduce_dtype<InT>::value, l.data[i], l.bound, sim->slot_ptr(l.count_slot), histogramBins,
                                   static_cast<double>(lowerBound), static_cast<double>(upperBound), sim->d_hist_out, sim->main_stream));
  std::vector<unsigned int> h(histogramBins);
  FGB_CUDA_THROW(cudaMemcpyAsync(h.data(), sim->d_hist_out, static_cast<size_t>(histogramBins) * 4, cudaMemcpyDeviceToHost, sim->main_stream));
  FGB_CUDA_THROW(cudaStreamSynchronize(sim->main_stream));
  return std::vector<OutT>(h.begin(), h.end());
}
#if defined(__CUDACC__)
namespace detail {
template <typename InT>
struct identity_transform {  // reduce() == transformReduce() with the identity
  template <typename A, typename B>
  struct unary_function {
    __host__ __device__ B operator()(const A &a) const { return static_cast<B>(a); }
  };
};
}  // namespace detail
template <typename InT, typename OutT, typename transformOperatorT, typename reductionOperatorT>
inline OutT HostAgentAPI::transformReduce(const std::string &variable, transformOperatorT, reductionOperatorT, OutT init) {
  static_assert(sizeof(OutT) <= 8, "reduction results are at most 8 bytes");
  sim->initialise();
  if (sim->slab.enabled) throw exception::UnsupportedFeature("user-functor reductions under a slab decomposition");
  detail::DevList &l = sim->state_list(agent, state);
  const int i = l.index_of(variable);
  if (i < 0) throw exception::InvalidAgentVar("agent '" + agent + "' has no variable '" + variable + "'");
  if (l.meta[i].elements != 1) throw exception::UnsupportedVarType("HostAgentAPI::transformReduce() does not support agent array variables");
  if (l.meta[i].type != std::type_index(typeid(InT))) throw exception::InvalidVarType("wrong type for '" + variable + "'");
  if (!sim->d_user_reduce) {
    const size_t bytes = static_cast<size_t>(detail::kUserRedBlocks) * 8 + static_cast<size_t>(detail::kUserRedBlocks) * 4 + 16;
    FGB_CUDA_THROW(cudaMalloc(&sim->d_user_reduce, bytes));
    FGB_CUDA_THROW(cudaMemset(sim->d_user_reduce, 0, bytes));
  }
  OutT *partial = static_cast<OutT *>(sim->d_user_reduce);
  // layout: kUserRedBlocks partials of sizeof(OutT) <= 8 bytes, kUserRedBlocks "has data" words right behind the partials
  // (as the kernel addresses them), the arrival counter at the very end
  unsigned int *done = reinterpret_cast<unsigned int *>(static_cast<char *>(sim->d_user_reduce) + static_cast<size_t>(detail::kUserRedBlocks) * 12 + 8);
  unsigned int blocks = (l.bound + 2047u) / 2048u;
  blocks = std::max(1u, std::min(blocks, detail::kUserRedBlocks));
  using Tr = typename transformOperatorT::template unary_function<InT, OutT>;
  using Rd = typename reductionOperatorT::template binary_function<OutT>;
  detail::k_user_transform_reduce<InT, OutT, Tr, Rd><<<blocks, 256, 0, sim->main_stream>>>(
      reinterpret_cast<const InT *>(l.data[i]), l.bound, sim->slot_ptr(l.count_slot), init, partial, done, static_cast<OutT *>(sim->d_reduce_out));
  FGB_CUDA_THROW(cudaPeekAtLastError());
  ++sim->own_launches;
  OutT r{};
  FGB_CUDA_THROW(cudaMemcpyAsync(&r, sim->d_reduce_out, sizeof(OutT), cudaMemcpyDeviceToHost, sim->main_stream));
  FGB_CUDA_THROW(cudaStreamSynchronize(sim->main_stream));
  return r;
}
template <typename InT, typename reductionOperatorT>
inline InT HostAgentAPI::reduce(const std::string &variable, reductionOperatorT r, InT init) {
  return transformReduce<InT, InT, detail::identity_transform<InT>, reductionOperatorT>(variable, detail::identity_transform<InT>(), r, init);
}
#endif
template <typename T>
inline T HostAgentAPI::sum(const std::string &variable) {
  return static_cast<T>(reduce<T, typename detail::reduce_dtype<T>::sum_t>(variable, FGB_REDUCE_SUM));
}
template <typename T>
inline T HostAgentAPI::min(const std::string &variable) {
  return count() ? reduce<T, T>(variable, FGB_REDUCE_MIN) : T{};
}
template <typename T>
inline T HostAgentAPI::max(const std::string &variable) {
  return count() ? reduce<T, T>(variable, FGB_REDUCE_MAX) : T{};
}
template <typename T>
inline T HostEnvironment::getProperty(const std::string &name) { return sim->getEnvironmentProperty<T>(name); }
template <typename T>
inline T HostEnvironment::setProperty(const std::string &name, T value) { return sim->setEnvironmentProperty<T>(name, value); }

}  // namespace flamegpu

#endif  // FGB_INCLUDE_FLAMEGPU_SIMULATION_CUDASIMULATION_IMPL_H_
