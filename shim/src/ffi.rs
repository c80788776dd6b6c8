//! Raw bindings of include/sailgpu.h (C ABI: plain pointers, sizes, Arrow C Data / C Device Data structs).
//! NOT COMPILED here.  Each function cites the reference interface it stands in for in include/sailgpu.h.
use std::ffi::{c_char, c_void, CStr};

use arrow::ffi::{FFI_ArrowArray, FFI_ArrowSchema};
use datafusion_common::{DataFusionError, Result};

#[repr(C)]
pub struct SailGpuCtx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct SailGpuOp {
    _private: [u8; 0],
}

/// Arrow C Device Data Interface (`ArrowDeviceArray`); arrow-rs 58 has no binding yet.
#[repr(C)]
pub struct FFI_ArrowDeviceArray {
    pub array: FFI_ArrowArray,
    pub device_id: i64,
    pub device_type: i32, // ARROW_DEVICE_CUDA = 2
    pub sync_event: *mut c_void,
    pub reserved: [i64; 3],
}

impl FFI_ArrowDeviceArray {
    pub fn empty() -> Self {
        Self { array: FFI_ArrowArray::empty(), device_id: -1, device_type: 0, sync_event: std::ptr::null_mut(), reserved: [0; 3] }
    }
    /// rows of the batch (the only field of a HANDLE a consumer outside the library may look at)
    pub fn array_length(&self) -> usize { self.array.len() }
    /// drops the batch: `FFI_ArrowArray`'s own Drop calls the release callback the library installed
    pub fn release(self) { drop(self) }
}

pub const SAILGPU_OK: i32 = 0;
pub const SAILGPU_ERR_INVALID: i32 = 1;
pub const SAILGPU_ERR_UNSUPPORTED: i32 = 2;
pub const SAILGPU_ERR_ARITHMETIC: i32 = 4;
pub const SAILGPU_ERR_NO_DEVICE: i32 = 5;
pub const SAILGPU_JIT_COLD_VARIANT: i32 = 1;
pub const SAILGPU_JIT_COMPILE: i32 = 2;

#[link(name = "sailgpu")]
extern "C" {
    pub fn sailgpu_version() -> u32;
    pub fn sailgpu_ctx_create(device: i32, out: *mut *mut SailGpuCtx) -> i32;
    pub fn sailgpu_ctx_destroy(ctx: *mut SailGpuCtx);
    pub fn sailgpu_ctx_last_error(ctx: *const SailGpuCtx) -> *const c_char;
    pub fn sailgpu_ctx_synchronize(ctx: *mut SailGpuCtx) -> i32;
    pub fn sailgpu_comm_unique_id(out128: *mut u8) -> i32;
    pub fn sailgpu_ctx_comm_init(ctx: *mut SailGpuCtx, unique_id128: *const u8, rank: i32, world_size: i32) -> i32;
    pub fn sailgpu_spec_validate(
        spec: *const c_char, spec_len: usize, input_schemas: *const *const FFI_ArrowSchema, n_inputs: i32,
        out_schema: *mut FFI_ArrowSchema, err_buf: *mut c_char, err_cap: usize,
    ) -> i32;
    pub fn sailgpu_jit_precompile(
        spec: *const c_char, spec_len: usize, input_schemas: *const *const FFI_ArrowSchema, n_inputs: i32,
        validity_mask: u64, flags: i32, buf: *mut c_char, cap: usize,
    ) -> i64;
    pub fn sailgpu_op_create(
        ctx: *mut SailGpuCtx, spec: *const c_char, spec_len: usize, input_schemas: *const *const FFI_ArrowSchema,
        n_inputs: i32, partition: i32, out: *mut *mut SailGpuOp, out_schema: *mut FFI_ArrowSchema,
    ) -> i32;
    pub fn sailgpu_op_push(op: *mut SailGpuOp, input_idx: i32, batch: *mut FFI_ArrowArray) -> i32;
    pub fn sailgpu_op_push_device(op: *mut SailGpuOp, input_idx: i32, batch: *mut FFI_ArrowDeviceArray) -> i32;
    pub fn sailgpu_op_finish_input(op: *mut SailGpuOp, input_idx: i32) -> i32;
    pub fn sailgpu_op_pull(op: *mut SailGpuOp, out: *mut FFI_ArrowArray, has_more: *mut i32) -> i32;
    pub fn sailgpu_op_pull_device_handle(op: *mut SailGpuOp, out: *mut FFI_ArrowDeviceArray, has_more: *mut i32) -> i32;
    pub fn sailgpu_op_pull_device(op: *mut SailGpuOp, out: *mut FFI_ArrowDeviceArray, has_more: *mut i32) -> i32;
    pub fn sailgpu_op_pull_partition(op: *mut SailGpuOp, part: i32, out: *mut FFI_ArrowDeviceArray, has_more: *mut i32) -> i32;
    pub fn sailgpu_op_metrics(op: *mut SailGpuOp, json_buf: *mut c_char, cap: usize) -> i64;
    pub fn sailgpu_last_error(op: *const SailGpuOp) -> *const c_char;
    pub fn sailgpu_op_destroy(op: *mut SailGpuOp);
    /// result sink: one self-contained Arrow IPC stream per batch (what `to_arrow_batch` writes, executor.rs:320-330)
    pub fn sailgpu_ipc_stream(schema: *const FFI_ArrowSchema, batch: *const FFI_ArrowArray, data: *mut *mut u8, len: *mut usize) -> i32;
    pub fn sailgpu_op_pull_ipc(op: *mut SailGpuOp, data: *mut *mut u8, len: *mut usize, rows: *mut i64, has_more: *mut i32) -> i32;
    pub fn sailgpu_ipc_last_error() -> *const c_char;
    pub fn sailgpu_ipc_free(data: *mut u8);
    pub fn sailgpu_exchange(
        ctx: *mut SailGpuCtx, schema: *const FFI_ArrowSchema, send: *mut FFI_ArrowDeviceArray, n: i32,
        recv: *mut FFI_ArrowDeviceArray,
    ) -> i32;
}

/// Non-zero status -> the DataFusion error the CPU operator would have raised
/// (error variants Sail serialises across workers: crates/sail-execution/src/stream/error.rs:85-118).
pub fn check(op: *const SailGpuOp, rc: i32) -> Result<()> {
    if rc == SAILGPU_OK {
        return Ok(());
    }
    let msg = unsafe {
        let p = if op.is_null() { sailgpu_ctx_last_error(std::ptr::null()) } else { sailgpu_last_error(op) };
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    Err(match rc {
        SAILGPU_ERR_ARITHMETIC if msg.contains("Divide by zero") => DataFusionError::ArrowError(Box::new(arrow::error::ArrowError::DivideByZero), None),
        SAILGPU_ERR_ARITHMETIC => DataFusionError::ArrowError(Box::new(arrow::error::ArrowError::ArithmeticOverflow(msg)), None),
        SAILGPU_ERR_UNSUPPORTED => DataFusionError::NotImplemented(msg),
        SAILGPU_ERR_INVALID => DataFusionError::Plan(msg),
        _ => DataFusionError::Execution(format!("libsailgpu: {msg} (status {rc})")),
    })
}
