//! `extern "C"` declarations mirroring include/modchol_b200.h one to one.
use std::os::raw::{c_char, c_int, c_void};

pub const MODCHOL_GMW81: c_int = 0;
pub const MODCHOL_SE90: c_int = 1;
pub const MODCHOL_SE99: c_int = 2;
pub const MODCHOL_MEM_HOST: c_int = 0;
pub const MODCHOL_MEM_DEVICE: c_int = 1;
pub const MODCHOL_MODE_STRICT: c_int = 0;
pub const MODCHOL_MODE_FAST: c_int = 1;

extern "C" {
    pub fn modchol_factor_batched_f64(
        algo: c_int, mode: c_int, batch: i64, n: i32, a: *const f64, l: *mut f64, e: *mut f64,
        p: *mut u64, phase2_start: *mut i32, mem_kind: c_int, stream: *mut c_void,
    ) -> c_int;
    pub fn modchol_factor_batched_multi_gpu_f64(
        algo: c_int, mode: c_int, batch: i64, n: i32, a: *const f64, l: *mut f64, e: *mut f64,
        p: *mut u64, phase2_start: *mut i32, ngpus: c_int,
    ) -> c_int;
    pub fn modchol_gershgorin_circles_batched_f64(
        batch: i64, n: i32, a: *const f64, circles: *mut f64, mem_kind: c_int, stream: *mut c_void,
    ) -> c_int;
    pub fn modchol_verify_batched_f64(
        batch: i64, n: i32, a: *const f64, l: *const f64, e: *const f64, p: *const u64,
        resid: *mut f64, flags: *mut i32, mem_kind: c_int, stream: *mut c_void,
    ) -> c_int;
    pub fn modchol_solve_batched_f64(
        batch: i64, n: i32, nrhs: i32, l: *const f64, p: *const u64, b: *const f64, x: *mut f64,
        mem_kind: c_int, stream: *mut c_void,
    ) -> c_int;
    pub fn modchol_factor_solve_batched_f64(
        algo: c_int, mode: c_int, batch: i64, n: i32, nrhs: i32, a: *const f64, b: *const f64,
        x: *mut f64, l: *mut f64, e: *mut f64, p: *mut u64, phase2_start: *mut i32,
        mem_kind: c_int, stream: *mut c_void,
    ) -> c_int;
    pub fn modchol_assemble_qdqt_batched_f64(
        batch: i64, n: i32, nrefl: i32, v: *const f64, d: *const f64, a: *mut f64, mem_kind: c_int,
        stream: *mut c_void,
    ) -> c_int;
    pub fn modchol_fp64_peak_tflops(dfma_tflops: *mut f64, dmma_tflops: *mut f64) -> c_int;
    pub fn modchol_set_small_kernel_family(family: c_int) -> c_int;
    pub fn modchol_kernel_class(algo: c_int, mode: c_int, n: i32) -> *const c_char;
    pub fn modchol_launch_count() -> i64;
    pub fn modchol_last_error() -> *const c_char;
    pub fn modchol_version() -> c_int;
    pub fn modchol_device_count() -> c_int;
    pub fn modchol_shutdown();
}
