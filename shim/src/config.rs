//! `execution.gpu.*` -- the keys to add to crates/sail-common/src/config/application.yaml (next to
//! `execution.batch_size`, application.yaml:247-251).  NOT COMPILED here.
//!
//! ```yaml
//! - key: execution.gpu.enabled          # bool,  default false: rewrite eligible physical operators to GpuExec
//! - key: execution.gpu.devices          # list,  default [0]:  CUDA ordinals; partition p runs on devices[p % len]
//! - key: execution.gpu.min_rows         # int,   default 1048576: operators whose input statistics are smaller stay on the CPU
//! - key: execution.gpu.coalesce_rows    # int,   default 4194304: child batches are concatenated up to this many rows per push
//! - key: execution.gpu.jit_at_plan_time # bool,  default true: sailgpu_jit_precompile while the job is being planned
//! ```
#[derive(Debug, Clone)]
pub struct GpuOptions {
    pub enabled: bool,
    pub devices: Vec<i32>,
    pub min_rows: usize,
    pub coalesce_rows: usize,
    pub jit_at_plan_time: bool,
}

impl Default for GpuOptions {
    fn default() -> Self {
        Self { enabled: false, devices: vec![0], min_rows: 1 << 20, coalesce_rows: 4 << 20, jit_at_plan_time: true }
    }
}

impl GpuOptions {
    /// one context (= one CUDA stream, allocation cache and staging pool) per partition slot: calls on one context are
    /// serialised inside the library, contexts run concurrently (include/sailgpu.h, "Threading")
    pub fn device_for(&self, partition: usize) -> i32 {
        self.devices[partition % self.devices.len()]
    }
}
