// engine.cu -- PipelineOp: FilterExec / ProjectionExec / AggregateExec (and fused chains of them)
// executed by the tile-pipeline kernel.  Mirrors DataFusion's operator contract as Sail uses it
// (SURVEY.md section 8b): streaming operators emit one output batch per input batch; the aggregate
// consumes its whole input and emits at end of stream; NULL predicate rows are dropped; schemas are
// fixed at construction.
#include "engine.hpp"

#include <algorithm>

namespace sg {

static uint64_t now_ns() {
  return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static uint64_t next_pow2(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return p; }

// ------------------------------------------------------------------------------------------------
// spec parsing
// ------------------------------------------------------------------------------------------------
StageSpec parse_stage(const Json& j, const Schema& in, Schema* out) {
  StageSpec st;
  const std::string op = j.at("op").as_str();
  if (op == "filter") {
    st.kind = StageSpec::Filter;
    st.predicate = parse_expr(j.at("predicate"), in);
    SG_CHECK(st.predicate->type.id == TypeId::Bool, SAILGPU_ERR_INVALID, "filter predicate must be boolean");
    const Json* p = j.find("projection");
    if (p && !p->is_null()) {
      st.has_projection = true;
      for (auto& x : p->a) {
        int i = (int)x.as_int();
        SG_CHECK(i >= 0 && i < (int)in.size(), SAILGPU_ERR_INVALID, "filter projection index out of range");
        st.projection.push_back(i);
        out->push_back(in[(size_t)i]);
      }
    } else *out = in;
  } else if (op == "projection") {
    st.kind = StageSpec::Projection;
    for (auto& x : j.at("exprs").a) {
      ExprPtr e = parse_expr(x.at("expr"), in);
      st.exprs.push_back(e);
      st.names.push_back(x.at("name").as_str());
      out->push_back({st.names.back(), e->type, e->nullable});
    }
  } else if (op == "aggregate") {
    st.kind = StageSpec::Aggregate;
    const Json* md = j.find("mode");
    st.mode = md ? md->as_str() : "single";
    SG_CHECK(st.mode == "single" || st.mode == "partial" || st.mode == "final" || st.mode == "final_partitioned", SAILGPU_ERR_INVALID,
             "aggregate mode '" + st.mode + "'");
    const bool merging = st.mode == "final" || st.mode == "final_partitioned";
    for (auto& g : j.at("group_by").a) {
      ExprPtr e = parse_expr(g.at("expr"), in);
      st.group_exprs.push_back(e);
      st.group_names.push_back(g.at("name").as_str());
      out->push_back({st.group_names.back(), e->type, e->nullable});
    }
    size_t state_col = st.group_exprs.size();
    for (auto& a : j.at("aggs").a) {
      StageSpec::Agg ag;
      ag.fn = a.at("fn").as_str();
      ag.name = a.at("name").as_str();
      const Json* it = a.find("input_type");
      if (it && !it->is_null()) ag.input_type = parse_type(it->as_str());
      const Json* args = a.find("args");
      if (!merging && args && args->kind == Json::Arr && !args->a.empty()) { ag.arg = parse_expr(args->a[0], in); ag.has_arg = true; ag.input_type = ag.arg->type; }
      SG_CHECK(ag.fn == "count" || ag.has_arg || merging, SAILGPU_ERR_INVALID, "aggregate '" + ag.fn + "' needs an argument");
      SG_CHECK(ag.fn == "count" || ag.input_type.id != TypeId::Null, SAILGPU_ERR_INVALID, "aggregate '" + ag.name + "' needs input_type in final mode");
      AggTypes at = agg_types(ag.fn, ag.fn == "count" ? T(TypeId::Int64) : ag.input_type);
      if (merging) {
        for (size_t k = 0; k < at.state.size(); ++k) {
          SG_CHECK(state_col + k < in.size(), SAILGPU_ERR_INVALID, "final aggregate: input has too few state columns");
          const DataType& have = in[state_col + k].type;
          SG_CHECK(have == at.state[k] || (have.id == TypeId::Int64 && at.state[k].id == TypeId::UInt64) || (have.id == TypeId::UInt64 && at.state[k].id == TypeId::Int64),
                   SAILGPU_ERR_INVALID, "final aggregate: state column " + std::to_string(state_col + k) + " is " + have.str() + ", expected " + at.state[k].str());
        }
      }
      state_col += at.state.size();
      if (st.mode == "partial") {
        if (ag.fn == "avg") { out->push_back({ag.name + "[count]", at.state[0], false}); out->push_back({ag.name + "[sum]", at.state[1], true}); }
        else out->push_back({ag.name + "[" + ag.fn + "]", at.state[0], ag.fn != "count"});
      } else {
        out->push_back({ag.name, at.final_type, ag.fn != "count"});
      }
      st.aggs.push_back(ag);
    }
  } else {
    fail(SAILGPU_ERR_UNSUPPORTED, "operator '" + op + "' cannot be part of a fused pipeline");
  }
  return st;
}

void check_device_error(Ctx* ctx, uint32_t* dev_flag) {
  uint32_t f = 0;
  SG_CUDA(cudaMemcpyAsync(&f, dev_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!f) return;
  SG_CUDA(cudaMemsetAsync(dev_flag, 0, 4, ctx->stream));
  if (f & ERR_DIV_ZERO) fail(SAILGPU_ERR_ARITHMETIC, "Divide by zero");
  if (f & ERR_OVERFLOW) fail(SAILGPU_ERR_ARITHMETIC, "Arithmetic overflow");
  if (f & ERR_TABLE_FULL) fail(SAILGPU_ERR_CUDA, "hash table overflow");
  fail(SAILGPU_ERR_UNSUPPORTED, "unsupported value encountered on device");
}

// ------------------------------------------------------------------------------------------------
// launching one compiled pipeline over one batch
// ------------------------------------------------------------------------------------------------
struct DevProgram {
  BufPtr literals;
  std::vector<uint64_t> literal_ptrs;
};

static bool timing_enabled() { static const bool on = getenv("SAILGPU_TIMING") != nullptr; return on; }

struct PipelineRunner {
  Ctx* ctx;
  Schema in_schema;
  std::vector<StageSpec> stages;
  std::map<std::vector<bool>, std::shared_ptr<CompiledPipeline>> cache;
  std::map<const CompiledPipeline*, DevProgram> programs;
  DevScalars scal;
  int hot_wanted = 8;
  std::function<void(PipelineCompiler&, CompiledPipeline&)> custom_sink;   // build / partition sinks
  std::function<void(PipelineCompiler&, CompiledPipeline&)> pre_stages;    // probe ops injected before the stages

  void init(Ctx* c, const Schema& in) {
    ctx = c; in_schema = in;
    scal.buf = dev_alloc_zero(ctx, 256);
  }

  std::shared_ptr<CompiledPipeline> compiled_for(const DevBatch& b) {
    std::vector<bool> sig;
    for (auto& c : b.cols) sig.push_back((bool)c.validity);
    auto it = cache.find(sig);
    if (it != cache.end()) return it->second;
    auto cp = std::make_shared<CompiledPipeline>();
    PipelineCompiler pc(in_schema, sig);
    if (pre_stages) pre_stages(pc, *cp);
    bool agg = false;
    for (auto& st : stages) {
      SG_CHECK(!agg, SAILGPU_ERR_INVALID, "aggregate must be the last stage of a pipeline");
      if (st.kind == StageSpec::Filter) {
        pc.add_filter(st.predicate);
        if (st.has_projection) pc.set_projection(st.projection);
      } else if (st.kind == StageSpec::Projection) {
        pc.set_exprs(st.exprs);
      } else {
        pc.finish_aggregate(*cp, st);
        agg = true;
      }
    }
    if (!agg) { if (custom_sink) custom_sink(pc, *cp); else pc.finish_store_or_compact(*cp); }
    int hot = 0;
    if (agg) hot = cp->agg.n_keys == 0 ? 1 : hot_wanted;
    pc.finalize(*cp, ctx, hot);
    cache[sig] = cp;
    return cp;
  }

  DevProgram& program_for(const std::shared_ptr<CompiledPipeline>& cp) {
    auto it = programs.find(cp.get());
    if (it == programs.end()) {
      DevProgram dp;
      size_t lit_bytes = 0;
      for (auto& s : cp->literals) lit_bytes += (s.size() + 15) & ~(size_t)15;
      dp.literals = dev_alloc(ctx, lit_bytes + 16);
      std::vector<uint8_t> blob(lit_bytes + 16, 0);
      size_t off = 0;
      for (auto& s : cp->literals) {
        memcpy(blob.data() + off, s.data(), s.size());
        dp.literal_ptrs.push_back(reinterpret_cast<uint64_t>(dp.literals->ptr) + off);
        off += (s.size() + 15) & ~(size_t)15;
      }
      if (!cp->literals.empty()) {
        SG_CUDA(cudaMemcpyAsync(dp.literals->ptr, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
        SG_CUDA(cudaStreamSynchronize(ctx->stream));   // host blob goes out of scope
      }
      for (auto& fx : cp->literal_fixups) cp->prog[(size_t)fx.first].imm1 = dp.literal_ptrs[(size_t)fx.second];
      it = programs.emplace(cp.get(), std::move(dp)).first;
    }
    return it->second;
  }

  // fills inputs + common fields; caller fills sink buffers; then launch()
  void prepare(PipelineParams& P, const CompiledPipeline& cp, const DevBatch& b, int64_t row0, int64_t nrows) {
    memset(&P, 0, sizeof(P));
    P.n_rows = nrows;
    P.tile_rows = cp.rpt * NT;
    P.n_inputs = (int)cp.inputs.size();
    P.n_inst = (int)cp.prog.size();
    P.sink = cp.sink;
    P.arena_bytes = cp.arena_bytes;
    P.mask_slot = cp.mask_slot;
    P.error_flag = scal.error();
    P.n_probes = cp.n_probes;
    bool tma = getenv("SAILGPU_NO_TMA") == nullptr;
    SG_CHECK(row0 % 1024 == 0, SAILGPU_ERR_INVALID, "chunk offset must be a multiple of 1024 rows");
    for (size_t i = 0; i < cp.inputs.size(); ++i) {
      const InputReg& r = cp.inputs[i];
      const DevColumn& c = b.cols[(size_t)r.col];
      const uint8_t* base = static_cast<const uint8_t*>(r.validity ? c.validity->ptr : c.data->ptr);
      base += r.width ? row0 * r.width : row0 / 8;
      P.in[i].data = base;
      P.in[i].slot = (uint32_t)r.slot;
      P.in[i].width = r.width;
      P.in[i].tma_ok = (reinterpret_cast<uint64_t>(base) & 15) == 0 ? 1 : 0;
      tma &= P.in[i].tma_ok != 0;
    }
    P.use_tma = tma ? 1 : 0;
  }

  static uint32_t rs(uint32_t off, uint32_t add) { return off == NO_SLOT ? NO_SLOT : (off & 0x7FFFFFFFu) + ((off >> 31) ? add : 0u); }
  static void rs_key(KeyDesc& k, uint32_t add) { k.slot = rs(k.slot, add); k.valid_slot = rs(k.valid_slot, add); }

  // Builds the kernel argument block: program, descriptors and parameters resolved for both stages.
  void launch(PipelineParams& P, const std::shared_ptr<CompiledPipeline>& cp, const PipelineAux* aux_host, Metrics& m) {
    program_for(cp);
    const int64_t n_tiles = (P.n_rows + P.tile_rows - 1) / P.tile_rows;
    if (n_tiles == 0) return;
    P.prog = nullptr;
    auto K = std::make_unique<KernelArgs>();
    memset(K.get(), 0, sizeof(KernelArgs));
    for (int st = 0; st < 2; ++st) {
      const uint32_t add = (uint32_t)st * cp->stage_bytes;
      PipelineParams& Q = K->P[st];
      Q = P;
      Q.mask_slot = rs(P.mask_slot, add);
      for (int i = 0; i < Q.n_inputs; ++i) Q.in[i].slot = rs(P.in[i].slot, add);
      for (int j = 0; j < Q.n_out; ++j) { Q.out[j].slot = rs(P.out[j].slot, add); Q.out[j].valid_slot = rs(P.out[j].valid_slot, add); }
      for (size_t i = 0; i < cp->prog.size(); ++i) {
        VmInst I = cp->prog[i];
        I.dst = rs(I.dst, add); I.a = rs(I.a, add); I.b = rs(I.b, add); I.c = rs(I.c, add);
        K->prog[st][i] = I;
      }
      if (aux_host) {
        PipelineAux& A = K->aux[st];
        A = *aux_host;
        for (int i = 0; i < A.agg.n_keys; ++i) rs_key(A.agg.keys[i], add);
        for (int w = 0; w < A.agg.key_words; ++w) { A.agg.kwords[w].slot = rs(A.agg.kwords[w].slot, add); A.agg.kwords[w].valid_slot = rs(A.agg.kwords[w].valid_slot, add); }
        for (int j = 0; j < A.agg.n_accs; ++j) { A.agg.accs[j].value_slot = rs(A.agg.accs[j].value_slot, add); A.agg.accs[j].valid_slot = rs(A.agg.accs[j].valid_slot, add); }
        for (int i = 0; i < A.build.n_keys; ++i) rs_key(A.build.keys[i], add);
        for (int i = 0; i < A.part.n_keys; ++i) rs_key(A.part.keys[i], add);
        A.part.pid_slot = rs(A.part.pid_slot, add);
        for (int q = 0; q < MAX_PROBES; ++q) {
          for (int i = 0; i < A.probe[q].n_keys; ++i) rs_key(A.probe[q].keys[i], add);
          A.probe[q].rowid_slot = rs(A.probe[q].rowid_slot, add);
          A.probe[q].match_slot = rs(A.probe[q].match_slot, add);
        }
      }
    }
    int per_sm = std::max(1, (int)(ctx->max_smem / (cp->smem_bytes + 1024)));
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * per_sm);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool timed = timing_enabled();
    if (timed) { SG_CUDA(cudaEventCreate(&e0)); SG_CUDA(cudaEventCreate(&e1)); SG_CUDA(cudaEventRecord(e0, ctx->stream)); }
    SG_CUDA(launch_pipeline(*K, cp->rpt, cp->n_stages, cp->smem_bytes, grid, ctx->stream));
    if (timed) { SG_CUDA(cudaEventRecord(e1, ctx->stream)); m.pending.emplace_back(e0, e1); }
    m.kernel_launches++;
    m.pipeline_launches++;
  }
};

// ------------------------------------------------------------------------------------------------
// PipelineOp
// ------------------------------------------------------------------------------------------------
struct AggTable {
  BufPtr table, state;
  uint64_t capacity = 0;
  uint64_t rows_bound = 0;     // upper bound on the number of groups (rows fed so far)
};

struct PipelineOp : Op {
  PipelineRunner run;
  bool has_agg = false;
  StageSpec* agg_stage = nullptr;
  AggTable tab;
  std::shared_ptr<CompiledPipeline> agg_cp;   // all launches of an aggregate must share one layout
  std::deque<BatchPtr> ready;
  std::vector<BufPtr> kept_heaps;
  bool input_done = false, emitted = false;

  void collect_heaps(const DevBatch& b) {
    for (auto& c : b.cols)
      if (c.type.is_string()) for (auto& h : c.heaps) kept_heaps.push_back(h);
  }

  void push(int input, const BatchPtr& b) override {
    SG_CHECK(input == 0, SAILGPU_ERR_INVALID, "operator has one input");
    SG_CHECK(!input_done, SAILGPU_ERR_STATE, "push after finish_input");
    const uint64_t t0 = now_ns();
    m.input_rows += (uint64_t)b->rows; m.input_batches++;
    if (has_agg) { collect_heaps(*b); push_agg(b); }
    else ready.push_back(run_stream(b));
    m.elapsed_compute_ns += now_ns() - t0;
  }
  void finish(int input) override {
    SG_CHECK(input == 0, SAILGPU_ERR_INVALID, "operator has one input");
    input_done = true;
  }
  bool pull(BatchPtr* out) override {
    *out = nullptr;
    if (has_agg) {
      if (!input_done) return true;
      if (emitted) return false;
      const uint64_t t0 = now_ns();
      *out = extract_agg();
      m.elapsed_compute_ns += now_ns() - t0;
      emitted = true;
      m.output_rows += (uint64_t)(*out)->rows; m.output_batches++;
      return false;
    }
    if (!ready.empty()) {
      *out = ready.front(); ready.pop_front();
      m.output_rows += (uint64_t)(*out)->rows; m.output_batches++;
      return !(input_done && ready.empty());
    }
    return !input_done;
  }

  // ---- streaming: filter / projection ----------------------------------------------------------
  BatchPtr run_stream(const BatchPtr& b) {
    auto cp = run.compiled_for(*b);
    auto out = std::make_shared<DevBatch>();
    const int64_t n = b->rows;
    PipelineParams P;
    run.prepare(P, *cp, *b, 0, n);
    const bool compact = cp->sink == SINK_COMPACT;
    P.n_out = (int)cp->outs.size();
    SG_CHECK(P.n_out <= MAX_OUTPUTS, SAILGPU_ERR_UNSUPPORTED, "more than " + std::to_string(MAX_OUTPUTS) + " output columns");
    std::vector<BufPtr> valid_tmp((size_t)P.n_out), bool_tmp((size_t)P.n_out);
    std::vector<BufPtr> all_heaps;
    for (auto& c : b->cols) if (c.type.is_string()) for (auto& h : c.heaps) all_heaps.push_back(h);
    for (int j = 0; j < P.n_out; ++j) {
      OutputCol o = cp->outs[(size_t)j];
      DevColumn c; c.type = cp->out_types[(size_t)j]; c.arrow_is_utf8 = c.type.id == TypeId::Utf8;
      if (o.width) { c.data = dev_alloc(ctx, (size_t)n * o.width); o.data = static_cast<uint8_t*>(c.data->ptr); }
      else if (compact) { bool_tmp[(size_t)j] = dev_alloc(ctx, (size_t)n); o.data = static_cast<uint8_t*>(bool_tmp[(size_t)j]->ptr); }
      else { c.data = dev_alloc_zero(ctx, (size_t)((n + 31) / 32 * 4)); o.data = static_cast<uint8_t*>(c.data->ptr); }
      if (o.valid_slot != NO_SLOT) {
        if (compact) { valid_tmp[(size_t)j] = dev_alloc(ctx, (size_t)n); o.valid_bytes = static_cast<uint8_t*>(valid_tmp[(size_t)j]->ptr); }
        else { c.validity = dev_alloc_zero(ctx, (size_t)((n + 31) / 32 * 4)); o.valid_bytes = static_cast<uint8_t*>(c.validity->ptr); c.null_count = -1; }
      }
      if (c.type.is_string()) c.heaps = all_heaps;
      P.out[j] = o;
      out->cols.push_back(c);
    }
    BufPtr status;
    if (compact) {
      const int64_t n_tiles = (n + P.tile_rows - 1) / P.tile_rows;
      status = dev_alloc_zero(ctx, (size_t)(n_tiles + 1) * 8);
      P.tile_status = static_cast<unsigned long long*>(status->ptr);
      P.ticket = run.scal.ticket();
      P.out_count = run.scal.out_count();
      SG_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(run.scal.buf->ptr) + 8, 0, 16, ctx->stream));
    }
    run.launch(P, cp, nullptr, m);
    int64_t out_rows = n;
    if (compact) {
      unsigned long long cnt = 0;
      SG_CUDA(cudaMemcpyAsync(&cnt, run.scal.out_count(), 8, cudaMemcpyDeviceToHost, ctx->stream));
      check_device_error(ctx, run.scal.error());   // synchronises
      out_rows = (int64_t)cnt;
      for (int j = 0; j < P.n_out; ++j) {
        DevColumn& c = out->cols[(size_t)j];
        if (bool_tmp[(size_t)j]) {
          c.data = dev_alloc_zero(ctx, (size_t)((out_rows + 31) / 32 * 4));
          SG_CUDA(launch_pack_bytes(static_cast<const uint8_t*>(bool_tmp[(size_t)j]->ptr), static_cast<uint32_t*>(c.data->ptr), out_rows, nullptr, ctx->stream));
        }
        if (valid_tmp[(size_t)j]) {
          c.validity = dev_alloc_zero(ctx, (size_t)((out_rows + 31) / 32 * 4));
          SG_CUDA(launch_pack_bytes(static_cast<const uint8_t*>(valid_tmp[(size_t)j]->ptr), static_cast<uint32_t*>(c.validity->ptr), out_rows, nullptr, ctx->stream));
          c.null_count = -1;
        }
      }
      if (!bool_tmp.empty() || !valid_tmp.empty()) SG_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
      check_device_error(ctx, run.scal.error());
    }
    out->rows = out_rows;
    for (auto& c : out->cols) c.length = out_rows;
    return out;
  }

  // ---- aggregate ------------------------------------------------------------------------------
  static constexpr uint64_t MAX_CAPACITY = 1ull << 27;

  uint64_t read_n_groups() {
    unsigned long long g = 0;
    SG_CUDA(cudaMemcpyAsync(&g, run.scal.n_groups(), 8, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    return g;
  }

  void fill_table(AggParams& A) {
    A.table = static_cast<uint8_t*>(tab.table->ptr);
    A.state = static_cast<uint32_t*>(tab.state->ptr);
    A.capacity_mask = tab.capacity - 1;
    A.n_groups = run.scal.n_groups();
  }

  // make room for `rows` more input rows (each may be a new group)
  void ensure_capacity(const CompiledPipeline& cp, uint64_t rows) {
    const AggParams& A0 = cp.agg;
    if (A0.n_keys == 0) rows = 1;
    uint64_t need = next_pow2(std::max<uint64_t>(1024, 2 * (tab.rows_bound + rows)));
    if (need <= tab.capacity) { tab.rows_bound += rows; return; }
    uint64_t groups = 0;
    if (tab.capacity) {
      groups = read_n_groups();
      tab.rows_bound = groups;
      need = next_pow2(std::max<uint64_t>(1024, 2 * (groups + rows)));
      if (need <= tab.capacity) { tab.rows_bound += rows; return; }
    }
    SG_CHECK(need <= MAX_CAPACITY, SAILGPU_ERR_UNSUPPORTED, "aggregate needs more than 2^27 group slots in one step");
    AggTable old = tab;
    tab.capacity = need;
    tab.table = dev_alloc(ctx, (size_t)need * A0.entry_words * 8);
    tab.state = dev_alloc_zero(ctx, (size_t)need * 4);
    if (old.capacity && groups) {
      AggParams A = A0;
      fill_table(A);
      SG_CUDA(cudaMemsetAsync(run.scal.n_groups(), 0, 8, ctx->stream));
      SG_CUDA(launch_agg_rehash(A, static_cast<const uint8_t*>(old.table->ptr), static_cast<const uint32_t*>(old.state->ptr), old.capacity, run.scal.error(), ctx->stream));
      m.kernel_launches++;
    }
    tab.rows_bound = groups + rows;
  }

  void push_agg(const BatchPtr& b) {
    auto cp = run.compiled_for(*b);
    if (!agg_cp) agg_cp = cp;
    SG_CHECK(cp->agg.entry_words == agg_cp->agg.entry_words && cp->agg.key_words == agg_cp->agg.key_words, SAILGPU_ERR_UNSUPPORTED,
             "aggregate input batches differ in which key columns carry validity buffers");
    const int64_t n = b->rows;
    int64_t done = 0;
    while (done < n) {
      int64_t chunk = n - done;
      if (cp->agg.n_keys > 0) {
        const int64_t max_chunk = (int64_t)(MAX_CAPACITY / 2) - (int64_t)std::min<uint64_t>(tab.rows_bound, MAX_CAPACITY / 4);
        if (chunk > max_chunk) {
          if (tab.capacity) tab.rows_bound = read_n_groups();
          chunk = std::min<int64_t>(chunk, (int64_t)(MAX_CAPACITY / 2) - (int64_t)tab.rows_bound);
          chunk &= ~(int64_t)1023;
          SG_CHECK(chunk > 0, SAILGPU_ERR_UNSUPPORTED, "aggregate exceeds 2^26 groups");
        }
      }
      ensure_capacity(*cp, (uint64_t)chunk);
      PipelineParams P;
      run.prepare(P, *cp, *b, done, chunk);
      PipelineAux aux;
      memset(&aux, 0, sizeof(aux));
      aux.agg = cp->agg;
      fill_table(aux.agg);
      run.launch(P, cp, &aux, m);
      done += chunk;
    }
  }

  BatchPtr extract_agg() {
    std::shared_ptr<CompiledPipeline> cp = agg_cp;
    if (!cp) {   // no input at all: compile against an all-valid signature to learn the output layout
      DevBatch dummy;
      for (auto& f : run.in_schema) { DevColumn c; c.type = f.type; dummy.cols.push_back(c); }
      cp = run.compiled_for(dummy);
    }
    const AggParams& A0 = cp->agg;
    uint64_t groups = 0;
    if (tab.capacity) { check_device_error(ctx, run.scal.error()); groups = read_n_groups(); }
    const bool synth = A0.n_keys == 0 && groups == 0;   // global aggregate over zero rows: one row of NULLs / zero counts
    const int64_t rows = synth ? 1 : (int64_t)groups;
    auto out = std::make_shared<DevBatch>();
    out->rows = rows;
    AggExtractParams X;
    memset(&X, 0, sizeof(X));
    X.n_cols = (int)cp->agg_outs.size();
    std::vector<BufPtr> vbytes((size_t)X.n_cols);
    for (int i = 0; i < X.n_cols; ++i) {
      const AggOutSpec& s = cp->agg_outs[(size_t)i];
      AggOutCol& o = X.cols[i];
      o.kind = s.kind; o.a = s.a; o.b = s.b; o.nullable = s.nullable ? 1 : 0;
      o.width = s.type.is_string() ? 16 : s.type.arrow_width();
      SG_CHECK(s.type.id != TypeId::Bool, SAILGPU_ERR_UNSUPPORTED, "boolean group keys are not supported yet");
      if (s.kind == 0) { o.key_word = s.b; o.src_words = A0.keys[s.a].width == 16 ? 2 : 1; }
      else if (s.kind == 1) { const int op = A0.accs[s.a].op; o.src_words = (op == ACC_SUM_I128 || op == ACC_MIN_I128 || op == ACC_MAX_I128) ? 2 : 1; }
      else {
        o.is_float = s.type.is_float() ? 1 : 0;
        if (!o.is_float) { i128 mul = pow10_i128(s.type.scale - s.in_type.scale); o.scale_mul_lo = (uint64_t)(u128)mul; o.scale_mul_hi = (uint64_t)((u128)mul >> 64); }
      }
      DevColumn c; c.type = s.type; c.length = rows; c.arrow_is_utf8 = s.type.id == TypeId::Utf8;
      c.data = dev_alloc_zero(ctx, (size_t)rows * o.width);
      o.data = static_cast<uint8_t*>(c.data->ptr);
      if (s.nullable) { vbytes[(size_t)i] = dev_alloc_zero(ctx, (size_t)rows + 4); o.valid_bytes = static_cast<uint8_t*>(vbytes[(size_t)i]->ptr); }
      if (c.type.is_string()) c.heaps = kept_heaps;
      out->cols.push_back(c);
    }
    if (!synth && rows > 0) {
      AggParams A = A0;
      fill_table(A);
      SG_CUDA(cudaMemsetAsync(run.scal.cursor(), 0, 8, ctx->stream));
      SG_CUDA(launch_agg_extract(A, X, run.scal.cursor(), run.scal.error(), ctx->stream));
      m.kernel_launches++;
    } else if (synth) {
      // counts are 0 (valid); every other aggregate is NULL -> validity bytes stay 0, count columns get no validity
      for (int i = 0; i < X.n_cols; ++i) {
        const AggOutSpec& s = cp->agg_outs[(size_t)i];
        const bool is_count = s.kind == 1 && (A0.accs[s.a].op == ACC_COUNT || (A0.accs[s.a].op == ACC_SUM_I64 && !A0.accs[s.a].track_seen));
        if (vbytes[(size_t)i] && is_count) SG_CUDA(cudaMemsetAsync(vbytes[(size_t)i]->ptr, 1, 1, ctx->stream));
      }
    }
    SG_CUDA(cudaMemsetAsync(run.scal.nulls(0), 0, 8 * 20, ctx->stream));
    SG_CHECK(X.n_cols <= 20 || true, SAILGPU_ERR_UNSUPPORTED, "");
    std::vector<unsigned long long> nulls((size_t)X.n_cols, 0);
    BufPtr nullctr = dev_alloc_zero(ctx, (size_t)X.n_cols * 8 + 8);
    for (int i = 0; i < X.n_cols; ++i) {
      if (!vbytes[(size_t)i] || rows == 0) continue;
      DevColumn& c = out->cols[(size_t)i];
      c.validity = dev_alloc_zero(ctx, (size_t)((rows + 31) / 32 * 4));
      SG_CUDA(launch_pack_bytes(static_cast<const uint8_t*>(vbytes[(size_t)i]->ptr), static_cast<uint32_t*>(c.validity->ptr), rows,
                                static_cast<unsigned long long*>(nullctr->ptr) + i, ctx->stream));
    }
    SG_CUDA(cudaMemcpyAsync(nulls.data(), nullctr->ptr, (size_t)X.n_cols * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < X.n_cols; ++i) {
      DevColumn& c = out->cols[(size_t)i];
      if (c.validity) { c.null_count = (int64_t)nulls[(size_t)i]; if (c.null_count == 0) c.validity = nullptr; }
    }
    return out;
  }
};

std::unique_ptr<Op> make_join_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);
std::unique_ptr<Op> make_sort_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);
std::unique_ptr<Op> make_repartition_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);

std::unique_ptr<Op> make_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs, int partition) {
  (void)partition;
  const std::string kind = spec.at("op").as_str();
  if (kind == "hash_join") return make_join_op(ctx, spec, inputs);
  if (kind == "sort") return make_sort_op(ctx, spec, inputs);
  if (kind == "repartition") return make_repartition_op(ctx, spec, inputs);
  SG_CHECK(inputs.size() == 1, SAILGPU_ERR_INVALID, "operator '" + kind + "' takes exactly one input");
  auto op = std::make_unique<PipelineOp>();
  op->ctx = ctx; op->kind = kind; op->in_schemas = inputs;
  op->run.init(ctx, inputs[0]);
  std::vector<const Json*> stage_specs;
  if (kind == "pipeline") for (auto& s : spec.at("stages").a) stage_specs.push_back(&s);
  else stage_specs.push_back(&spec);
  SG_CHECK(!stage_specs.empty(), SAILGPU_ERR_INVALID, "empty pipeline");
  Schema cur = inputs[0];
  for (const Json* s : stage_specs) {
    Schema next;
    op->run.stages.push_back(parse_stage(*s, cur, &next));
    cur = next;
  }
  op->has_agg = op->run.stages.back().kind == StageSpec::Aggregate;
  for (size_t i = 0; i + 1 < op->run.stages.size(); ++i)
    SG_CHECK(op->run.stages[i].kind != StageSpec::Aggregate, SAILGPU_ERR_INVALID, "aggregate must be the last stage of a pipeline");
  op->out_schema = cur;
  return op;
}

}  // namespace sg
