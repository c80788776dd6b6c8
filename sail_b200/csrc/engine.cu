// engine.cu -- PipelineOp: FilterExec / ProjectionExec / AggregateExec (and fused chains of them)
// executed by the tile-pipeline kernel.  Mirrors DataFusion's operator contract as Sail uses it
// (SURVEY.md section 8b): streaming operators emit one output batch per input batch; the aggregate
// consumes its whole input and emits at end of stream; NULL predicate rows are dropped; schemas are
// fixed at construction.
#include "runner.hpp"

namespace sg {


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
// PipelineOp
// ------------------------------------------------------------------------------------------------
struct AggTable {
  BufPtr table, state, occ;
  uint64_t capacity = 0;
  bool direct = false;        // direct-key protocol (vm.h AggParams::direct_key): `occ` is built on demand
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
    else if (const std::vector<int>* cols = rename_only()) {
      // a ProjectionExec of plain column references (the `expr=[_0@0 as #9, _3@1 as #12]` nodes over every scan in the reference's
      // plans) moves no data: the output batch shares the input's buffers, as DataFusion's Arc-cloned arrays do
      auto out = std::make_shared<DevBatch>();
      out->rows = b->rows;
      for (int c : *cols) out->cols.push_back(b->cols[(size_t)c]);
      ready.push_back(out);
    }
    else ready.push_back(run_stream(b));
    m.elapsed_compute_ns += now_ns() - t0;
  }
  // the input column of every output when the whole pipeline is projections of bare column references, else null
  int rename_state = -1;      // -1: not analysed yet, 0: no, 1: yes
  std::vector<int> rename_cols;
  const std::vector<int>* rename_only() {
    if (rename_state < 0) {
      rename_state = 0;
      std::vector<int> cur;
      bool ok = !run.stages.empty();
      for (size_t si = 0; ok && si < run.stages.size(); ++si) {
        const StageSpec& st = run.stages[si];
        if (st.kind != StageSpec::Projection) { ok = false; break; }
        std::vector<int> next;
        for (auto& e : st.exprs) {
          if (e->kind != Expr::Col) { ok = false; break; }
          next.push_back(si == 0 ? e->col : cur[(size_t)e->col]);
        }
        cur.swap(next);
      }
      if (ok) { rename_cols = cur; rename_state = 1; }
    }
    return rename_state == 1 ? &rename_cols : nullptr;
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

  // ---- two-pass filter ---------------------------------------------------------------------------
  // An order-preserving single-pass compaction chains every tile to all earlier ones (decoupled look-back), and since a
  // tile only knows its count after evaluating the predicate the chain runs at the pace of the slowest resident tile
  // (profiles/: ~45 % of the FilterExec kernel's stall samples).  Large batches therefore take two passes: pass 1
  // reads only the predicate's columns and stores the mask (1 bit / row); the per-tile popcounts are scanned; pass 2
  // reads the projected columns + mask and stores every tile at its known offset, tiles in any order.
  static constexpr int64_t TWO_PASS_MIN_ROWS = 1 << 18;
  PipelineRunner mask_run, sel_run;
  bool two_pass_ready = false;
  std::string share_key;
  void setup_two_pass() {
    const Schema& in = run.in_schema;
    mask_run.init(ctx, in);
    if (!share_key.empty()) mask_run.share(share_key + "#mask");
    StageSpec f; f.kind = StageSpec::Filter; f.predicate = run.stages[0].predicate;
    mask_run.stages.push_back(f);
    mask_run.custom_sink = [](PipelineCompiler& pc, CompiledPipeline& cp) { pc.finish_mask_store(cp); };
    Schema in2 = in;
    Field mf; mf.name = "__mask"; mf.type.id = TypeId::Bool; mf.nullable = false;
    in2.push_back(mf);
    sel_run.init(ctx, in2);
    if (!share_key.empty()) sel_run.share(share_key + "#select");
    sel_run.stages = run.stages;
    StageSpec& s0 = sel_run.stages[0];
    auto me = std::make_shared<Expr>(); me->kind = Expr::Col; me->col = (int)in.size(); me->type.id = TypeId::Bool; me->nullable = false;
    s0.predicate = me;
    if (!s0.has_projection) { s0.has_projection = true; s0.projection.clear(); for (size_t i = 0; i < in.size(); ++i) s0.projection.push_back((int)i); }
    two_pass_ready = true;
  }
  bool two_pass_applies(const BatchPtr& b) const {
    const char* mn = getenv("SAILGPU_TWO_PASS_MIN");     // rows from which a filter runs in two passes (tests lower it; "off" disables)
    if (mn && !strcmp(mn, "off")) return false;
    const int64_t min_rows = mn && *mn ? atoll(mn) : TWO_PASS_MIN_ROWS;
    return !has_agg && !run.stages.empty() && run.stages[0].kind == StageSpec::Filter && b->rows >= min_rows;
  }
  BatchPtr run_two_pass(const BatchPtr& b) {
    if (!two_pass_ready) setup_two_pass();
    BatchPtr mb = run_streaming(mask_run, ctx, b, m, nullptr, {});
    auto b2 = std::make_shared<DevBatch>(*b);
    b2->cols.push_back(mb->cols[0]);
    auto cp2 = sel_run.compiled_for(*b2);
    const int tile_rows = cp2->rpt * NT;
    const int64_t n = b->rows, n_tiles = (n + tile_rows - 1) / tile_rows;
    BufPtr counts = dev_alloc_zero(ctx, (size_t)(n_tiles + 1) * 4), offs = dev_alloc(ctx, (size_t)(n_tiles + 1) * 8), scratch = dev_alloc(ctx, 1026 * 8);
    SG_CUDA(launch_tile_popcount(static_cast<const uint32_t*>(mb->cols[0].data->ptr), n, tile_rows, n_tiles, static_cast<uint32_t*>(counts->ptr), ctx->stream));
    SG_CUDA(launch_exclusive_scan_u32(static_cast<const uint32_t*>(counts->ptr), n_tiles + 1, static_cast<uint64_t*>(offs->ptr), static_cast<uint64_t*>(scratch->ptr), ctx->stream));
    unsigned long long total = 0;
    SG_CUDA(cudaMemcpyAsync(&total, static_cast<const uint64_t*>(offs->ptr) + n_tiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    m.kernel_launches += 4;
    return run_streaming(sel_run, ctx, b2, m, nullptr, {}, static_cast<const unsigned long long*>(offs->ptr), (int64_t)total);
  }

  BatchPtr run_stream(const BatchPtr& b) {
    if (two_pass_applies(b)) return run_two_pass(b);
    return run_streaming(run, ctx, b, m, nullptr, {});
  }

  // ---- aggregate ------------------------------------------------------------------------------
  // The group table is sized for the groups the operator expects, not for its input rows.  Every launch carries a group
  // limit (half the capacity minus what tiles in flight could still add); CTAs that see the table above it stop taking
  // tiles and append the ones they still owned to a deferred list.  The host reads (groups, deferred) back at the next
  // synchronisation point it needs anyway (next push / finish), and if tiles were handed back it grows the table by the
  // observed groups-per-row ratio, rehashes, and re-launches over the list -- with the pipeline variant compiled for
  // the global table alone when the input turned out to have many groups.
  static constexpr uint64_t MIN_CAPACITY = 1ull << 22, MAX_CAPACITY = 1ull << 28;
  static constexpr uint64_t CARD_MANY_GROUPS = 256, HOT_GROUP_LIMIT = 1 << 16;
  static uint64_t min_capacity() { const char* v = getenv("SAILGPU_AGG_MIN_CAPACITY"); return v && *v ? next_pow2((uint64_t)atoll(v)) : MIN_CAPACITY; }
  bool use_cold = false;
  int64_t rows_in_table = 0;
  int64_t known_groups = -1;      // group count read (and error flag checked) by the last resolve_pending(), -1 = stale
  struct Pending { BatchPtr batch; std::shared_ptr<CompiledPipeline> cp; BufPtr deferred; int64_t n_tiles = 0; bool active = false; } pend;

  unsigned long long* n_deferred_ptr() { return reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(run.scal.buf->ptr) + 40); }

  uint64_t read_n_groups() {
    run.ensure_scratch();
    unsigned long long g = 0;
    SG_CUDA(cudaMemcpyAsync(&g, run.scal.n_groups(), 8, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    return g;
  }

  void fill_table(AggParams& A) {
    A.table = static_cast<uint8_t*>(tab.table->ptr);
    A.state = static_cast<uint32_t*>(tab.state->ptr);
    A.occ = static_cast<uint32_t*>(tab.occ->ptr);
    A.capacity_mask = tab.capacity - 1;
    A.n_groups = run.scal.n_groups();
    A.direct_key = tab.direct ? 1 : 0;
  }
  // One never-null 8-byte key word: the table CAN take the direct-key protocol.  Opt-in (SAILGPU_DIRECT_KEY=1): measured on
  // GROUP BY l_orderkey over SF10 (60 M rows -> 15 M groups) the insert kernel itself did not get faster without the fence and the
  // counter round trip (6.25 vs 6.5 ms: it is bound by the per-row accumulator atomics that need their old value for the 128-bit
  // carry), while initialising and scanning every ENTRY of the 64 M-slot table cost 4.3 ms more than the 4-byte state words of
  // the general protocol (11.5 vs 8.4 ms per aggregation, profiles/README.md).
  static bool direct_eligible(const AggParams& A) {
    const char* e = getenv("SAILGPU_DIRECT_KEY");
    return A.n_keys == 1 && A.key_words == 1 && !A.has_null_word && e && *e && atoi(e) != 0;
  }
  // direct-key tables keep no list of occupied entries while they are filled: build it (extraction / re-hash / migration read it)
  void build_occ(const AggTable& t, const AggParams& layout) {
    if (!t.direct) return;
    AggParams A = layout;
    A.table = static_cast<uint8_t*>(t.table->ptr); A.occ = static_cast<uint32_t*>(t.occ->ptr); A.capacity_mask = t.capacity - 1;
    BufPtr counter = dev_alloc_zero(ctx, 8);
    SG_CUDA(launch_agg_build_occ(A, static_cast<unsigned long long*>(counter->ptr), ctx->stream));
    m.kernel_launches++;
  }

  // (re)allocates the table with `cap` slots and moves the `groups` existing entries over
  void alloc_table(const AggParams& A0, uint64_t cap, uint64_t groups) {
    SG_CHECK(cap <= MAX_CAPACITY, SAILGPU_ERR_UNSUPPORTED, "aggregate needs more than 2^28 group slots");
    AggTable old = tab;
    tab.capacity = cap;
    tab.direct = direct_eligible(A0);
    tab.table = dev_alloc(ctx, (size_t)(cap + 1) * A0.entry_words * 8);
    tab.state = tab.direct ? dev_alloc(ctx, 4) : dev_alloc_zero(ctx, (size_t)cap * 4);
    tab.occ = dev_alloc(ctx, (size_t)(cap + 1) * 4);
    if (tab.direct) { run.ensure_scratch(); AggParams A = A0; fill_table(A); SG_CUDA(launch_agg_init_direct(A, ctx->stream)); m.kernel_launches++; }
    if (old.capacity && groups) {
      build_occ(old, A0);
      AggParams A = A0;
      fill_table(A);
      SG_CUDA(cudaMemsetAsync(run.scal.n_groups(), 0, 8, ctx->stream));
      SG_CUDA(launch_agg_rehash(A, static_cast<const uint8_t*>(old.table->ptr), static_cast<const uint32_t*>(old.occ->ptr), groups, run.scal.error(), ctx->stream));
      m.kernel_launches++;
    }
  }

  // one launch over all tiles of `b` (list == null) or over the listed tiles; tiles handed back land in `deferred_out`
  void launch_agg(const std::shared_ptr<CompiledPipeline>& cp, const DevBatch& b, const BufPtr& list, int64_t n_list, const BufPtr& deferred_out) {
    PipelineParams P;
    run.prepare(P, *cp, b, 0, b.rows);
    if (list) { P.tile_list = static_cast<const uint32_t*>(list->ptr); P.n_list = n_list; }
    PipelineAux aux;
    memset(&aux, 0, sizeof(aux));
    aux.agg = cp->agg;
    fill_table(aux.agg);
    if (deferred_out) {
      SG_CUDA(cudaMemsetAsync(n_deferred_ptr(), 0, 8, ctx->stream));
      aux.agg.deferred = static_cast<uint32_t*>(deferred_out->ptr);
      aux.agg.n_deferred = n_deferred_ptr();
      const char* fl = getenv("SAILGPU_AGG_FIRST_LIMIT");      // tests: force an early hand-back on the first pass
      aux.agg.group_limit = (!list && fl && *fl) ? (unsigned long long)atoll(fl) : ~0ull;   // ~0: launch() derives it from the grid
      // the dictionary variant is only worth running while there are few groups: it hands back early, and the
      // re-launch (many-groups variant, table sized by the observed ratio) takes over
      run.group_limit_cap = (cp->cold_variant || list) ? ~0ull : HOT_GROUP_LIMIT;      // re-launches over a list always make progress
    }
    run.launch(P, cp, &aux, m);
  }

  // ---- validity signatures ------------------------------------------------------------------------
  // The entry layout (null-mask word, seen bits, one counter per nullable avg/count argument) follows from which input
  // columns carry validity buffers, and Arrow batches of one stream differ in that (a producer drops the bitmap of a
  // batch without nulls).  The operator therefore keeps the union of the signatures it has seen: a batch that lacks a
  // bitmap of the union gets an all-ones one, and when a batch widens the union the table is carried over to the wider
  // layout (migrate_layout).  DataFusion's accumulators take `null_count == 0` fast paths per batch the same way.
  std::vector<bool> sticky_sig;
  BufPtr ones; size_t ones_bytes = 0;

  BatchPtr with_union_signature(const BatchPtr& b, bool* widened) {
    *widened = false;
    if (sticky_sig.empty()) sticky_sig.assign(b->cols.size(), false);
    bool same = true;
    for (size_t i = 0; i < b->cols.size(); ++i) {
      const bool has = (bool)b->cols[i].validity;
      if (has && !sticky_sig[i]) { sticky_sig[i] = true; *widened = true; }
      same &= has == sticky_sig[i];
    }
    if (same) return b;
    const size_t need = (size_t)((b->rows + 31) / 32 * 4);
    if (!ones || ones_bytes < need) {
      ones_bytes = std::max<size_t>(need, 1 << 16);
      ones = dev_alloc(ctx, ones_bytes);
      SG_CUDA(cudaMemsetAsync(ones->ptr, 0xFF, ones_bytes, ctx->stream));
    }
    auto bb = std::make_shared<DevBatch>(*b);
    for (size_t i = 0; i < bb->cols.size(); ++i)
      if (sticky_sig[i] && !bb->cols[i].validity) { bb->cols[i].validity = ones; bb->cols[i].null_count = 0; }
    return bb;
  }

  static bool same_layout(const CompiledPipeline& a, const CompiledPipeline& b) {
    const AggParams &x = a.agg, &y = b.agg;
    if (x.entry_words != y.entry_words || x.key_words != y.key_words || x.has_null_word != y.has_null_word || x.n_accs != y.n_accs) return false;
    for (int j = 0; j < x.n_accs; ++j)
      if (x.accs[j].op != y.accs[j].op || x.accs[j].word != y.accs[j].word || x.accs[j].track_seen != y.accs[j].track_seen) return false;
    return a.acc_ident == b.acc_ident;
  }

  // the table built under `from`'s layout continues under `to`'s (a superset: more seen bits, a null-mask word, separate
  // counters where one counter served several aggregates)
  void migrate_layout(const std::shared_ptr<CompiledPipeline>& from, const std::shared_ptr<CompiledPipeline>& to) {
    Trace tr(ctx, "agg.migrate");
    const AggParams &O = from->agg, &N = to->agg;
    AggMigrateMap M;
    memset(&M, 0, sizeof(M));
    M.old_entry_words = (int32_t)O.entry_words; M.old_key_words = O.key_words;
    M.key_shift = (N.has_null_word && !O.has_null_word) ? 1 : 0;
    SG_CHECK(N.key_words == O.key_words + M.key_shift && N.has_null_word >= O.has_null_word, SAILGPU_ERR_UNSUPPORTED, "aggregate layouts of two batches cannot be reconciled (group keys)");
    std::vector<int> src((size_t)N.n_accs, -1);
    for (auto& kv : to->acc_ident) {
      auto it = from->acc_ident.find(kv.first);
      SG_CHECK(it != from->acc_ident.end(), SAILGPU_ERR_UNSUPPORTED, "aggregate layouts of two batches cannot be reconciled (" + kv.first + ")");
      src[(size_t)kv.second] = it->second;
    }
    for (int j = 0; j < N.n_accs; ++j) {
      SG_CHECK(src[(size_t)j] >= 0 && O.accs[src[(size_t)j]].op == N.accs[j].op, SAILGPU_ERR_UNSUPPORTED, "aggregate layouts of two batches cannot be reconciled (accumulator kinds)");
      M.acc_src_word[j] = (int16_t)O.accs[src[(size_t)j]].word;
      M.seen_src[j] = N.accs[j].track_seen ? (O.accs[src[(size_t)j]].track_seen ? (int8_t)src[(size_t)j] : (int8_t)-1) : (int8_t)-2;
    }
    check_device_error(ctx, run.scal.error());
    const uint64_t groups = read_n_groups();
    AggTable old = tab;
    build_occ(old, O);
    tab.direct = direct_eligible(N);
    tab.table = dev_alloc(ctx, (size_t)(tab.capacity + 1) * N.entry_words * 8);
    tab.state = tab.direct ? dev_alloc(ctx, 4) : dev_alloc_zero(ctx, (size_t)tab.capacity * 4);
    tab.occ = dev_alloc(ctx, (size_t)(tab.capacity + 1) * 4);
    run.ensure_scratch();
    AggParams A = N;
    fill_table(A);
    if (tab.direct) { SG_CUDA(launch_agg_init_direct(A, ctx->stream)); m.kernel_launches++; }
    SG_CUDA(cudaMemsetAsync(run.scal.n_groups(), 0, 8, ctx->stream));
    SG_CUDA(launch_agg_migrate(A, M, static_cast<const uint8_t*>(old.table->ptr), static_cast<const uint32_t*>(old.occ->ptr), groups, run.scal.error(), ctx->stream));
    m.kernel_launches++;
    known_groups = -1;
  }

  void push_agg(const BatchPtr& b0) {
    Trace tr(ctx, "agg.push");
    resolve_pending();
    if (b0->rows == 0) return;
    bool widened = false;
    const BatchPtr b = with_union_signature(b0, &widened);
    auto cp = run.compiled_for(*b, use_cold);
    if (agg_cp && tab.capacity && !same_layout(*agg_cp, *cp)) {
      SG_CHECK(widened, SAILGPU_ERR_STATE, "aggregate layout changed without a new validity buffer");
      migrate_layout(agg_cp, cp);
    }
    agg_cp = cp;
    const bool grouped = cp->agg.n_keys > 0;
    // first table: 4 M slots for real inputs, but a final aggregate over a few partial rows gets a few KB (a hand-back
    // grows it if later batches are bigger)
    if (!tab.capacity) alloc_table(cp->agg, grouped ? std::min<uint64_t>(min_capacity(), next_pow2(8 * (uint64_t)b->rows + 2048)) : 1024, 0);
    const int64_t tile_rows = (int64_t)cp->rpt * NT;
    const int64_t n_tiles = (b->rows + tile_rows - 1) / tile_rows;
    BufPtr deferred = grouped ? dev_alloc(ctx, (size_t)n_tiles * 4) : nullptr;
    launch_agg(cp, *b, nullptr, 0, deferred);
    rows_in_table += b->rows;
    known_groups = -1;
    if (grouped) { pend.batch = b; pend.cp = cp; pend.deferred = deferred; pend.n_tiles = n_tiles; pend.active = true; }
  }

  // reads back what the last launch handed back and, until nothing is left, grows the table and re-launches over it
  void resolve_pending() {
    if (!pend.active) return;
    Trace tr(ctx, "agg.resolve");
    uint64_t prev_def = 0;
    for (;;) {
      unsigned long long gd[3] = {0, 0, 0};      // n_groups @24, cursor @32, n_deferred @40
      SG_CUDA(cudaMemcpyAsync(gd, run.scal.n_groups(), 24, cudaMemcpyDeviceToHost, ctx->stream));
      check_device_error(ctx, run.scal.error());   // synchronises
      const uint64_t groups = gd[0], n_def = gd[2];
      if (n_def == 0) { known_groups = (int64_t)groups; break; }      // extract_agg() right after needs no second read-back
      const int64_t tile_rows = (int64_t)pend.cp->rpt * NT;
      const double rows_def = (double)std::min<int64_t>((int64_t)n_def * tile_rows, pend.batch->rows);
      const double rows_done = std::max(1.0, (double)rows_in_table - rows_def);
      // groups still to come, by the ratio seen so far (every row a new group when nothing was processed yet)
      const double ratio = groups ? std::min(1.0, (double)groups / rows_done) : 1.0;
      const double est = ratio * rows_def * 1.25 + 1024.0;
      uint64_t cap = next_pow2((uint64_t)(2.0 * ((double)groups + est)) + 2 * MIN_CAPACITY / 2);
      // a hand-back does not always mean a full table: the dictionary variant gives up at HOT_GROUP_LIMIT groups whatever the
      // capacity.  The table only grows when the estimate asks for it, or when a re-launch at this capacity made no progress.
      const bool stuck = prev_def != 0 && n_def >= prev_def;
      if (cap <= tab.capacity && stuck) cap = tab.capacity * 2;
      prev_def = n_def;
      if (cap > tab.capacity) alloc_table(pend.cp->agg, cap, groups);
      std::shared_ptr<CompiledPipeline> cp = pend.cp;
      if (!use_cold && groups > CARD_MANY_GROUPS && getenv("SAILGPU_NO_COLD") == nullptr) {
        auto cold = run.compiled_for(*pend.batch, true);
        if (cold->agg.entry_words == cp->agg.entry_words) {
          use_cold = true;                                    // later batches start on the many-groups variant
          if (cold->rpt == cp->rpt) cp = cold;                // same tile size: this batch's deferred list carries over
        }
      }
      BufPtr next_deferred = dev_alloc(ctx, (size_t)n_def * 4);
      launch_agg(cp, *pend.batch, pend.deferred, (int64_t)n_def, next_deferred);
      pend.cp = cp; pend.deferred = next_deferred;
    }
    pend = Pending();
  }

  BatchPtr extract_agg() {
    resolve_pending();
    Trace tr(ctx, "agg.extract");
    std::shared_ptr<CompiledPipeline> cp = agg_cp;
    if (!cp) {   // no input at all: compile against an all-valid signature to learn the output layout
      DevBatch dummy;
      for (auto& f : run.in_schema) { DevColumn c; c.type = f.type; dummy.cols.push_back(c); }
      cp = run.compiled_for(dummy);
    }
    const AggParams& A0 = cp->agg;
    run.ensure_scratch();
    uint64_t groups = 0;
    if (tab.capacity) {
      if (known_groups >= 0) groups = (uint64_t)known_groups;
      else { check_device_error(ctx, run.scal.error()); groups = read_n_groups(); }
    }
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
      build_occ(tab, A0);
      AggParams A = A0;
      fill_table(A);
      SG_CUDA(launch_agg_extract(A, X, groups, run.scal.error(), ctx->stream));
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

// Plan-time specialisation (sailgpu_jit_precompile): compiles the pipeline of `spec` for a batch whose columns carry
// validity buffers where `validity_mask` has a bit set, writes the specialised kernel's source to *source and, with
// `compile`, its cubin into the kernel cache -- all without a device (NVRTC cross-compiles for sm_100a).
size_t pipeline_precompile(const Json& spec, const std::vector<Schema>& inputs, uint64_t validity_mask, bool cold, bool compile, std::string* source) {
  static Ctx plan_ctx;      // B200 geometry (148 SMs, 227 KB of shared memory per CTA); never touches a device
  std::unique_ptr<Op> op = make_op(&plan_ctx, spec, inputs, 0);
  PipelineOp* p = dynamic_cast<PipelineOp*>(op.get());
  SG_CHECK(p != nullptr, SAILGPU_ERR_UNSUPPORTED, "only filter / projection / aggregate pipelines are specialised");
  DevBatch shape;
  for (size_t i = 0; i < inputs[0].size(); ++i) {
    DevColumn c; c.type = inputs[0][i].type;
    if (i < 64 && ((validity_mask >> i) & 1)) c.validity = std::make_shared<DevBuf>();
    shape.cols.push_back(c);
  }
  auto cp = p->run.compiled_for(shape, cold);
  std::string why;
  SG_CHECK(jit_supported(*cp, &why), SAILGPU_ERR_UNSUPPORTED, "kernel specialiser: " + why);
  JitPlan plan;
  SG_CHECK(jit_plan(*cp, plan_ctx.max_smem, &plan), SAILGPU_ERR_UNSUPPORTED, "kernel specialiser: pipeline does not fit in shared memory");
  const std::string src = jit_generate(*cp, plan);
  if (source) *source = src;
  return compile ? jit_precompile_to_cache(src) : 0;
}

// Plan-time limits of a filter / projection / aggregate pipeline (sailgpu_spec_validate): the tile program is compiled against the
// B200 geometry for a batch without validity buffers, so that the data-independent limits -- "more than 6 group keys", "group key
// wider than 64 bytes", "does not fit in shared memory" ... -- are answered while planning, never after the first batch arrived.
// (Validity buffers add one column buffer each; a batch whose nullable columns push a pipeline over the 20-buffer limit is still
// reported at push time as SAILGPU_ERR_UNSUPPORTED.)  Touches no device.
static thread_local bool g_in_static_check = false;
std::unique_ptr<Op> make_wide_agg_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs, const Schema& out_schema);
void pipeline_static_check(const Json& spec, const std::vector<Schema>& inputs) {
  static Ctx plan_ctx;
  std::unique_ptr<Op> op = make_op(&plan_ctx, spec, inputs, 0);
  PipelineOp* p = dynamic_cast<PipelineOp*>(op.get());
  if (p == nullptr || p->rename_only()) return;       // a projection that only picks / renames columns passes buffers on: nothing to compile
  DevBatch shape;
  for (auto& f : inputs[0]) { DevColumn c; c.type = f.type; shape.cols.push_back(c); }
  p->run.compiled_for(shape, false);
}

std::unique_ptr<Op> make_join_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);
std::unique_ptr<Op> make_sort_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);
std::unique_ptr<Op> make_merge_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);
std::unique_ptr<Op> make_nlj_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);
std::unique_ptr<Op> make_repartition_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);
std::unique_ptr<Op> make_chain_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);
std::unique_ptr<Op> make_exchange_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs);

std::unique_ptr<Op> make_op(Ctx* ctx, const Json& spec, const std::vector<Schema>& inputs, int partition) {
  (void)partition;
  const std::string kind = spec.at("op").as_str();
  if (kind == "hash_join") return make_join_op(ctx, spec, inputs);
  if (kind == "nested_loop_join") return make_nlj_op(ctx, spec, inputs);
  if (kind == "sort") return make_sort_op(ctx, spec, inputs);
  if (kind == "sort_preserving_merge") return make_merge_op(ctx, spec, inputs);
  if (kind == "repartition") return make_repartition_op(ctx, spec, inputs);
  if (kind == "chain") return make_chain_op(ctx, spec, inputs);
  if (kind == "exchange") return make_exchange_op(ctx, spec, inputs);
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
  {   // operators created from the same spec over the same schema share their compiled pipelines (and specialised kernels)
    std::string key = "pipeline|";
    json_dump(spec, &key);
    for (auto& f : inputs[0]) { key += '|'; key += f.type.str(); key += f.nullable ? '?' : '!'; }
    op->share_key = key;
    op->run.share(key);
  }
  op->has_agg = op->run.stages.back().kind == StageSpec::Aggregate;
  for (size_t i = 0; i + 1 < op->run.stages.size(); ++i)
    SG_CHECK(op->run.stages[i].kind != StageSpec::Aggregate, SAILGPU_ERR_INVALID, "aggregate must be the last stage of a pipeline");
  op->out_schema = cur;
  // a group key the hash table cannot pack (more than 6 keys / 64 bytes): grouping by sorting instead (ops_more.cu WideAggOp)
  if (kind == "aggregate" && !g_in_static_check) {
    bool wide = false;
    g_in_static_check = true;
    try { pipeline_static_check(spec, inputs); }
    catch (const Error& e) { wide = e.code == SAILGPU_ERR_UNSUPPORTED && std::string(e.what()).find("group key") != std::string::npos; }
    g_in_static_check = false;
    if (wide) return make_wide_agg_op(ctx, spec, inputs, cur);
  }
  return op;
}

}  // namespace sg
