//! The execution-time rewrite: DataFusion nodes -> `GpuExec`.  NOT COMPILED here.
//!
//! Call sites (SURVEY.md 8b "Where it is plugged in", option (i)):
//!   crates/sail-execution/src/job_runner.rs:63        `let plan = sail_gpu::rewrite_for_gpu(plan, &opts)?;` before `trace_execution_plan`
//!   crates/sail-execution/src/task_runner/core.rs:110 same, on cluster workers
//! It runs after Sail's physical optimizer rules (crates/sail-physical-optimizer/src/lib.rs:38-70), so EXPLAIN snapshots and
//! `RemoteExecutionCodec` never see a `GpuExec`.
use std::ffi::CString;
use std::sync::Arc;

use arrow::ffi::FFI_ArrowSchema;
use datafusion::physical_plan::aggregates::{AggregateExec, AggregateMode};
use datafusion::physical_plan::filter::FilterExec;
use datafusion::physical_plan::projection::ProjectionExec;
use datafusion::physical_plan::ExecutionPlan;
use datafusion_common::tree_node::{Transformed, TreeNode};
use datafusion_common::Result;
use serde_json::{json, Value};

use crate::config::GpuOptions;
use crate::exec::GpuExec;
use crate::{ffi, spec};

/// Asks the library whether it executes `spec` over these inputs and whether its output schema is the node's
/// (`sailgpu_spec_validate`: spec parser + DataFusion/arrow type inference, no device).  Unsupported -> keep the CPU node.
fn accepted(spec_text: &str, plan: &Arc<dyn ExecutionPlan>, children: &[Arc<dyn ExecutionPlan>]) -> bool {
    let schemas: Vec<FFI_ArrowSchema> = match children.iter().map(|c| FFI_ArrowSchema::try_from(c.schema().as_ref())).collect() { Ok(v) => v, Err(_) => return false };
    let ptrs: Vec<*const FFI_ArrowSchema> = schemas.iter().map(|s| s as *const _).collect();
    let c = match CString::new(spec_text) { Ok(c) => c, Err(_) => return false };
    let mut out = FFI_ArrowSchema::empty();
    let mut err = vec![0i8; 512];
    let rc = unsafe { ffi::sailgpu_spec_validate(c.as_ptr(), spec_text.len(), ptrs.as_ptr(), ptrs.len() as i32, &mut out, err.as_mut_ptr(), err.len()) };
    if rc != ffi::SAILGPU_OK {
        log::debug!("GPU rewrite: {} stays on the CPU: {}", plan.name(), unsafe { std::ffi::CStr::from_ptr(err.as_ptr()) }.to_string_lossy());
        return false;
    }
    arrow::datatypes::Schema::try_from(&out).map(|s| s.fields() == plan.schema().fields()).unwrap_or(false)
}

fn is_streaming_stage(p: &Arc<dyn ExecutionPlan>) -> bool {
    let a = p.as_any();
    a.is::<FilterExec>() || a.is::<ProjectionExec>()
        || a.downcast_ref::<AggregateExec>().map(|x| matches!(x.mode(), AggregateMode::Partial | AggregateMode::Single | AggregateMode::SinglePartitioned)).unwrap_or(false)
}

/// Bottom-up: every supported node becomes a `GpuExec`; a Filter/Projection/Aggregate(Partial|Single) whose only child is
/// a streaming `GpuExec` is FUSED into that child's spec (`{"op":"pipeline","stages":[..]}`: one kernel, one pass over HBM).
/// All inputs of a hash-partitioned consumer are repartitioned by the same engine: a `RepartitionExec Hash` is only
/// rewritten when its consumer's other inputs are rewritten as well (the partition hash of libsailgpu is its own, not
/// DataFusion's `REPARTITION_RANDOM_STATE` ahash) -- `consistent_partitioning` below undoes the rest.
pub fn rewrite_for_gpu(plan: Arc<dyn ExecutionPlan>, opts: &GpuOptions) -> Result<Arc<dyn ExecutionPlan>> {
    if !opts.enabled { return Ok(plan); }
    let rewritten = plan.transform_up(|node| {
        let Some(mut sp) = spec::of_plan(&node) else { return Ok(Transformed::no(node)) };
        let mut children: Vec<Arc<dyn ExecutionPlan>> = node.children().into_iter().cloned().collect();
        if is_streaming_stage(&node) && children.len() == 1 {
            if let Some(child) = children[0].as_any().downcast_ref::<GpuExec>() {
                let cs: Value = serde_json::from_str(child.spec()).unwrap_or(Value::Null);
                let fusable = matches!(cs["op"].as_str(), Some("filter") | Some("projection") | Some("pipeline"))
                    && !(cs["op"] == "pipeline" && cs["stages"].as_array().and_then(|s| s.last()).map(|l| l["op"] == "aggregate").unwrap_or(false));
                if fusable {
                    let mut stages = if cs["op"] == "pipeline" { cs["stages"].as_array().cloned().unwrap_or_default() } else { vec![cs] };
                    stages.push(sp);
                    sp = json!({"op": "pipeline", "stages": stages});
                    children = child.children().into_iter().cloned().collect();
                }
            }
        }
        let text = sp.to_string();
        if !accepted(&text, &node, &children) { return Ok(Transformed::no(node)); }
        if opts.jit_at_plan_time { precompile(&text, &children); }
        Ok(Transformed::yes(Arc::new(GpuExec::new(text, &node, children, opts.clone())) as Arc<dyn ExecutionPlan>))
    })?.data;
    consistent_partitioning(rewritten)
}

/// Plan-time kernel specialisation (include/sailgpu.h: sailgpu_jit_precompile): the cubin is in the cache before the first batch.
fn precompile(spec_text: &str, children: &[Arc<dyn ExecutionPlan>]) {
    let schemas: Vec<FFI_ArrowSchema> = match children.iter().map(|c| FFI_ArrowSchema::try_from(c.schema().as_ref())).collect() { Ok(v) => v, Err(_) => return };
    let ptrs: Vec<*const FFI_ArrowSchema> = schemas.iter().map(|s| s as *const _).collect();
    if let Ok(c) = CString::new(spec_text) {
        let mut buf = vec![0i8; 256];
        // validity mask 0: batches without validity buffers (the common case for scans of NOT NULL data); other signatures compile on first use
        unsafe { ffi::sailgpu_jit_precompile(c.as_ptr(), spec_text.len(), ptrs.as_ptr(), ptrs.len() as i32, 0, ffi::SAILGPU_JIT_COMPILE, buf.as_mut_ptr(), buf.len()) };
    }
}

/// Which engine computes the hash partitioning an input of a multi-input operator arrives in.
#[derive(Clone, Copy, PartialEq, Eq)]
enum Exchange { Gpu, DataFusion }

/// The nearest hash exchange below `p`, looking through single-child nodes (filters, projections, partial aggregates, sorts keep
/// the partitioning of their input); `None`: the input is not hash-partitioned by an exchange in this plan (e.g. a CollectLeft
/// build side or a scan that is already partitioned).
fn exchange_below(p: &Arc<dyn ExecutionPlan>) -> Option<Exchange> {
    if let Some(g) = p.as_any().downcast_ref::<GpuExec>() {
        if g.op_kind() == "repartition" { return Some(Exchange::Gpu); }
    } else if let Some(r) = p.as_any().downcast_ref::<datafusion::physical_plan::repartition::RepartitionExec>() {
        if matches!(r.partitioning(), datafusion::physical_plan::Partitioning::Hash(..)) { return Some(Exchange::DataFusion); }
    }
    let children = p.children();
    if children.len() == 1 { exchange_below(children[0]) } else { None }
}

/// Turns the GPU hash exchange below `p` (if any, same search as `exchange_below`) back into the DataFusion node it replaced,
/// keeping whatever sits under it.
fn restore_exchange(p: &Arc<dyn ExecutionPlan>) -> Result<Arc<dyn ExecutionPlan>> {
    if let Some(g) = p.as_any().downcast_ref::<GpuExec>() {
        if g.op_kind() == "repartition" {
            let kept: Vec<Arc<dyn ExecutionPlan>> = g.children().into_iter().cloned().collect();
            return g.original().clone().with_new_children(kept);
        }
    }
    let children = p.children();
    if children.len() != 1 { return Ok(p.clone()); }
    let restored = restore_exchange(children[0])?;
    p.clone().with_new_children(vec![restored])
}

/// A GPU `repartition` under a consumer whose sibling input is partitioned by DataFusion (or the reverse) would break
/// co-partitioning: libsailgpu's partition hash is its own, not DataFusion's `REPARTITION_RANDOM_STATE` ahash.  For every node
/// with two or more inputs the exchanges below them must come from ONE engine; where they are mixed, the GPU exchanges are
/// turned back into the DataFusion nodes they replaced (the operators above and below stay on the GPU).
fn consistent_partitioning(plan: Arc<dyn ExecutionPlan>) -> Result<Arc<dyn ExecutionPlan>> {
    Ok(plan.transform_down(|node| {
        let children: Vec<Arc<dyn ExecutionPlan>> = node.children().into_iter().cloned().collect();
        if children.len() < 2 { return Ok(Transformed::no(node)); }
        let kinds: Vec<Option<Exchange>> = children.iter().map(exchange_below).collect();
        let mixed = kinds.contains(&Some(Exchange::Gpu)) && kinds.contains(&Some(Exchange::DataFusion));
        if !mixed { return Ok(Transformed::no(node)); }
        let restored = children.iter().map(restore_exchange).collect::<Result<Vec<_>>>()?;
        Ok(Transformed::yes(node.with_new_children(restored)?))
    })?.data)
}
