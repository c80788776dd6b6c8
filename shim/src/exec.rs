//! `GpuExec`: a DataFusion `ExecutionPlan` node executed by libsailgpu.  NOT COMPILED here.
//!
//! Trait surface as Sail's own nodes implement it (crates/sail-physical-plan/src/range.rs:80-150,
//! crates/sail-execution/src/plan/shuffle_write.rs:115-207).  One `sailgpu_op` per (node, partition); the stream is
//! pull based like every `SendableRecordBatchStream`, dropping it cancels the operator
//! (cf. `RoundRobinReceiverStream::drop`, repartition.rs:104-119).
use std::any::Any;
use std::ffi::CString;
use std::fmt;
use std::sync::Arc;

use arrow::array::{Array, RecordBatch, StructArray};
use arrow::compute::concat_batches;
use arrow::datatypes::SchemaRef;
use arrow::ffi::{from_ffi, to_ffi, FFI_ArrowArray, FFI_ArrowSchema};
use datafusion::execution::{SendableRecordBatchStream, TaskContext};
use datafusion::physical_plan::metrics::{ExecutionPlanMetricsSet, MetricsSet};
use datafusion::physical_plan::stream::RecordBatchStreamAdapter;
use datafusion::physical_plan::{DisplayAs, DisplayFormatType, ExecutionPlan, PlanProperties};
use datafusion_common::{internal_err, DataFusionError, Result};
use futures::StreamExt;

use crate::config::GpuOptions;
use crate::ffi::{self, SailGpuCtx, SailGpuOp};

/// A context per (device, partition slot); contexts are cheap and share nothing.
pub struct GpuContext(pub *mut SailGpuCtx);
unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}
impl GpuContext {
    pub fn create(device: i32) -> Result<Arc<Self>> {
        let mut p = std::ptr::null_mut();
        ffi::check(std::ptr::null(), unsafe { ffi::sailgpu_ctx_create(device, &mut p) })?;
        Ok(Arc::new(Self(p)))
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) { unsafe { ffi::sailgpu_ctx_destroy(self.0) } }
}

#[derive(Debug)]
pub struct GpuExec {
    /// JSON operator spec (include/sailgpu.h); a fused chain is `{"op":"pipeline","stages":[..]}`
    spec: String,
    /// display name of the DataFusion node(s) this replaces
    replaces: String,
    children: Vec<Arc<dyn ExecutionPlan>>,
    schema: SchemaRef,
    properties: Arc<PlanProperties>,
    options: GpuOptions,
    metrics: ExecutionPlanMetricsSet,
    /// the DataFusion node this one replaced (kept so that the rewrite can take a replacement back: rewrite.rs consistent_partitioning)
    original: Arc<dyn ExecutionPlan>,
}

impl GpuExec {
    /// `template`: the DataFusion node being replaced -- its `PlanProperties` (partitioning, ordering, emission) carry over
    /// unchanged except where INTEGRATION.md section 2 says otherwise (hash_join reports no maintained input order).
    pub fn new(spec: String, template: &Arc<dyn ExecutionPlan>, children: Vec<Arc<dyn ExecutionPlan>>, options: GpuOptions) -> Self {
        Self { spec, replaces: template.name().to_string(), children, schema: template.schema(), properties: template.properties().clone(),
               options, metrics: ExecutionPlanMetricsSet::new(), original: template.clone() }
    }
    pub fn spec(&self) -> &str { &self.spec }
    pub fn original(&self) -> &Arc<dyn ExecutionPlan> { &self.original }
    /// `"filter"`, `"hash_join"`, `"repartition"`, ... (the spec's `op`)
    pub fn op_kind(&self) -> String {
        serde_json::from_str::<serde_json::Value>(&self.spec).ok().and_then(|v| v["op"].as_str().map(str::to_string)).unwrap_or_default()
    }
}

impl DisplayAs for GpuExec {
    fn fmt_as(&self, _t: DisplayFormatType, f: &mut fmt::Formatter) -> fmt::Result {
        write!(f, "GpuExec: replaces={}, spec={}", self.replaces, self.spec)
    }
}

impl ExecutionPlan for GpuExec {
    fn name(&self) -> &'static str { "GpuExec" }
    fn as_any(&self) -> &dyn Any { self }
    fn properties(&self) -> &Arc<PlanProperties> { &self.properties }
    fn children(&self) -> Vec<&Arc<dyn ExecutionPlan>> { self.children.iter().collect() }
    fn maintains_input_order(&self) -> Vec<bool> {
        // filter / projection / pipeline-without-aggregate keep row order; joins, aggregates, sorts and repartitions do not
        let streaming = self.spec.starts_with("{\"op\":\"filter\"") || self.spec.starts_with("{\"op\":\"projection\"");
        vec![streaming; self.children.len()]
    }
    fn with_new_children(self: Arc<Self>, children: Vec<Arc<dyn ExecutionPlan>>) -> Result<Arc<dyn ExecutionPlan>> {
        if children.len() != self.children.len() { return internal_err!("GpuExec: wrong number of children"); }
        Ok(Arc::new(Self { spec: self.spec.clone(), replaces: self.replaces.clone(), children, schema: self.schema.clone(),
                           properties: self.properties.clone(), options: self.options.clone(), metrics: ExecutionPlanMetricsSet::new(),
                           original: self.original.clone() }))
    }
    fn metrics(&self) -> Option<MetricsSet> { Some(self.metrics.clone_inner()) }

    fn execute(&self, partition: usize, context: Arc<TaskContext>) -> Result<SendableRecordBatchStream> {
        // `execute` is synchronous and must not block (SURVEY.md 8b "Threading"): create child streams only.  A child that is a
        // GpuExec itself is NOT executed as a RecordBatch stream: it is run inside this node's blocking task and hands its output
        // over as device handles (sailgpu_op_pull_device_handle -> sailgpu_op_push_device), so the rows never leave HBM.
        let inputs: Vec<Input> = self.children.iter().map(|c| match c.as_any().downcast_ref::<GpuExec>() {
            Some(g) => Ok(Input::Gpu(g.clone_node(), context.clone())),
            None => c.execute(partition, context.clone()).map(Input::Host),
        }).collect::<Result<Vec<_>>>()?;
        let node = self.clone_node();
        let schema = self.schema.clone();
        let stream = futures::stream::once(async move {
            // CUDA synchronisation points block the calling thread: keep them off the tokio workers
            tokio::task::spawn_blocking(move || {
                let ctx = GpuContext::create(node.options.device_for(partition))?;
                let op = node.run(&ctx, partition, inputs)?;
                pull_host(&op, &node.schema)
            })
            .await
            .map_err(|e| DataFusionError::Execution(format!("GpuExec task: {e}")))?
        })
        .flat_map(|r| match r {
            Ok(batches) => futures::stream::iter(batches.into_iter().map(Ok)).boxed(),
            Err(e) => futures::stream::iter(vec![Err(e)]).boxed(),
        });
        Ok(Box::pin(RecordBatchStreamAdapter::new(schema, stream)))
    }
}

/// What feeds one input of a GpuExec
enum Input {
    Host(SendableRecordBatchStream),
    /// a child GpuExec (and the task context its own host-side children are executed with)
    Gpu(GpuExec, Arc<TaskContext>),
}

struct OpHandle { raw: *mut SailGpuOp, out_schema: FFI_ArrowSchema }
unsafe impl Send for OpHandle {}
impl Drop for OpHandle {
    fn drop(&mut self) { unsafe { ffi::sailgpu_op_destroy(self.raw) } } // idempotent cancel
}

impl GpuExec {
    fn clone_node(&self) -> GpuExec {
        GpuExec { spec: self.spec.clone(), replaces: self.replaces.clone(), children: self.children.clone(), schema: self.schema.clone(),
                  properties: self.properties.clone(), options: self.options.clone(), metrics: ExecutionPlanMetricsSet::new(),
                  original: self.original.clone() }
    }

    /// Creates the operator in `ctx`, feeds every input (input 0 first: the build side of a join) and returns it ready to be pulled.
    /// Runs on a blocking thread.  GPU children run recursively in the SAME context: one stream, so a handle needs no event.
    fn run(&self, ctx: &Arc<GpuContext>, partition: usize, inputs: Vec<Input>) -> Result<OpHandle> {
        let in_schemas: Vec<SchemaRef> = self.children.iter().map(|c| c.schema()).collect();
        let ffi_schemas: Vec<FFI_ArrowSchema> = in_schemas.iter().map(|s| FFI_ArrowSchema::try_from(s.as_ref())).collect::<std::result::Result<_, _>>()?;
        let ptrs: Vec<*const FFI_ArrowSchema> = ffi_schemas.iter().map(|s| s as *const _).collect();
        let cspec = CString::new(self.spec.as_str()).map_err(|e| DataFusionError::Plan(e.to_string()))?;
        let mut out_c = FFI_ArrowSchema::empty();
        let mut raw = std::ptr::null_mut();
        ffi::check(std::ptr::null(), unsafe {
            ffi::sailgpu_op_create(ctx.0, cspec.as_ptr(), self.spec.len(), ptrs.as_ptr(), ptrs.len() as i32, partition as i32, &mut raw, &mut out_c)
        })?;
        let op = OpHandle { raw, out_schema: out_c };
        let rt = tokio::runtime::Handle::current();
        for (i, input) in inputs.into_iter().enumerate() {
            match input {
                Input::Gpu(child, task_ctx) => {
                    // the child's own inputs: host streams for DataFusion children, recursion for GPU children
                    let grand: Vec<Input> = child.children.iter().map(|c| match c.as_any().downcast_ref::<GpuExec>() {
                        Some(g) => Ok(Input::Gpu(g.clone_node(), task_ctx.clone())),
                        None => c.execute(partition, task_ctx.clone()).map(Input::Host),
                    }).collect::<Result<Vec<_>>>()?;
                    let child_op = child.run(ctx, partition, grand)?;
                    loop {
                        let mut dev = ffi::FFI_ArrowDeviceArray::empty();
                        let mut more = 0i32;
                        ffi::check(child_op.raw, unsafe { ffi::sailgpu_op_pull_device_handle(child_op.raw, &mut dev, &mut more) })?;
                        if dev.array_length() > 0 {
                            ffi::check(op.raw, unsafe { ffi::sailgpu_op_push_device(op.raw, i as i32, &mut dev) })?; // takes ownership
                        } else {
                            dev.release();
                        }
                        if more == 0 { break; }
                    }
                }
                Input::Host(mut stream) => {
                    // DataFusion streams 8192-row batches (application.yaml:247-251); a launch wants millions: coalesce before the copy
                    let mut pending: Vec<RecordBatch> = vec![];
                    let mut rows = 0usize;
                    let flush = |pending: &mut Vec<RecordBatch>| -> Result<()> {
                        if pending.is_empty() { return Ok(()); }
                        let batch = concat_batches(&pending[0].schema(), pending.iter())?;
                        pending.clear();
                        let (mut arr, _schema) = to_ffi(&StructArray::from(batch).to_data())?;
                        ffi::check(op.raw, unsafe { ffi::sailgpu_op_push(op.raw, i as i32, &mut arr as *mut FFI_ArrowArray) }) // takes ownership
                    };
                    while let Some(b) = rt.block_on(stream.next()) {
                        let b = b?;
                        rows += b.num_rows();
                        pending.push(b);
                        if rows >= self.options.coalesce_rows { flush(&mut pending)?; rows = 0; }
                    }
                    flush(&mut pending)?;
                }
            }
            ffi::check(op.raw, unsafe { ffi::sailgpu_op_finish_input(op.raw, i as i32) })?;
        }
        Ok(op)
    }
}

/// pull until `has_more == 0`, importing every batch into host Arrow memory (the node's parent is a DataFusion operator)
fn pull_host(op: &OpHandle, out_schema: &SchemaRef) -> Result<Vec<RecordBatch>> {
    let mut out = vec![];
    loop {
        let mut arr = FFI_ArrowArray::empty();
        let mut more = 0i32;
        ffi::check(op.raw, unsafe { ffi::sailgpu_op_pull(op.raw, &mut arr, &mut more) })?;
        let data = unsafe { from_ffi(arr, &op.out_schema) }?;
        let batch = RecordBatch::from(StructArray::from(data)).with_schema(out_schema.clone())?;
        if batch.num_rows() > 0 || more == 0 { out.push(batch); }
        if more == 0 { break; }
    }
    Ok(out)
}
