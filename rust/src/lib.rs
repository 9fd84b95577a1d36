//! Modified Cholesky decompositions on NVIDIA B200.
//!
//! Drop-in for the hot path of `argmin-rs/modcholesky` v0.2.0: the same three traits over
//! `ndarray::Array2<f64>`, the same `Decomposition { l, e, p }`, plus batched entry points over
//! `ndarray::Array3<f64>`.  All arithmetic runs in hand-written sm_100a CUDA kernels behind the
//! C-ABI of `include/modchol_b200.h`; there is no CPU fallback, so a call on a machine without
//! a CUDA device panics (the reference's own failure mode is a panic, too).
//!
//! ```ignore
//! use modcholesky::ModCholeskySE99;
//! let a = ndarray::arr2(&[[1.0, 1.0, 2.0], [1.0, 1.0, 3.0], [2.0, 3.0, 1.0]]);
//! let decomp = a.mod_cholesky_se99();          // identical call to the reference crate
//! println!("{:?} {:?} {:?}", decomp.l, decomp.e, decomp.p);
//! ```
mod ffi;
pub mod utils;

use ndarray::{Array1, Array2, Array3};
use std::ffi::CStr;
use std::os::raw::c_int;

/// Arithmetic mode of the kernels (see `include/modchol_b200.h`).  `Strict` reproduces every
/// rounding of the reference (bit-identical to its CPU arithmetic); `Fast` fuses `a - b*c`
/// and scales by a reciprocal square root, agreeing within 1e-12 / 1e-10.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Mode {
    Strict = 0,
    Fast = 1,
}

static MODE: std::sync::atomic::AtomicI32 = std::sync::atomic::AtomicI32::new(0);

/// Process-wide default mode used by the trait methods (which take no arguments, like the
/// reference's).
pub fn set_mode(m: Mode) {
    MODE.store(m as i32, std::sync::atomic::Ordering::Relaxed);
}
fn mode() -> c_int {
    MODE.load(std::sync::atomic::Ordering::Relaxed) as c_int
}

/// Same struct as the reference (lib.rs:109-119).
pub struct Decomposition<L, E, P> {
    pub l: L,
    pub e: E,
    pub p: P,
}

impl<L, E, P> Decomposition<L, E, P> {
    pub fn new(l: L, e: E, p: P) -> Self {
        Decomposition { l, e, p }
    }
}

pub(crate) fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(ffi::modchol_last_error()) }.to_string_lossy().into_owned();
        // the reference's traits are infallible; it panics on misuse, and so do we
        panic!("modchol_b200 error {}: {}", rc, msg);
    }
}

fn factor2(algo: c_int, a: &Array2<f64>) -> Decomposition<Array2<f64>, Array1<f64>, Array1<usize>> {
    assert!(a.is_square()); // se90.rs:41, se99.rs:38 (gmw81.rs:43 is a debug_assert)
    let n = a.nrows();
    let a_std = a.as_standard_layout(); // row-major contiguous, as the C-ABI expects
    let mut l = Array2::<f64>::zeros((n, n));
    let mut e = Array1::<f64>::zeros(n);
    let mut p = vec![0u64; n];
    check(unsafe {
        ffi::modchol_factor_batched_f64(
            algo, mode(), 1, n as i32, a_std.as_ptr(), l.as_mut_ptr(), e.as_mut_ptr(),
            p.as_mut_ptr(), std::ptr::null_mut(), ffi::MODCHOL_MEM_HOST, std::ptr::null_mut(),
        )
    });
    Decomposition::new(l, e, Array1::from_iter(p.into_iter().map(|x| x as usize)))
}

fn factor3(
    algo: c_int, a: &Array3<f64>, ngpus: i32,
) -> Decomposition<Array3<f64>, Array2<f64>, Array2<usize>> {
    let (b, n, n2) = a.dim();
    assert!(n == n2);
    let a_std = a.as_standard_layout();
    let mut l = Array3::<f64>::zeros((b, n, n));
    let mut e = Array2::<f64>::zeros((b, n));
    let mut p = vec![0u64; b * n];
    let rc = unsafe {
        if ngpus == 1 {
            ffi::modchol_factor_batched_f64(
                algo, mode(), b as i64, n as i32, a_std.as_ptr(), l.as_mut_ptr(), e.as_mut_ptr(),
                p.as_mut_ptr(), std::ptr::null_mut(), ffi::MODCHOL_MEM_HOST, std::ptr::null_mut(),
            )
        } else {
            ffi::modchol_factor_batched_multi_gpu_f64(
                algo, mode(), b as i64, n as i32, a_std.as_ptr(), l.as_mut_ptr(), e.as_mut_ptr(),
                p.as_mut_ptr(), std::ptr::null_mut(), ngpus,
            )
        }
    };
    check(rc);
    let p = Array2::from_shape_vec((b, n), p.into_iter().map(|x| x as usize).collect()).unwrap();
    Decomposition::new(l, e, p)
}

/// Gill, Murray and Wright (1981) -- gmw81.rs:24-32.
pub trait ModCholeskyGMW81<L, E, P>
where
    Self: Sized,
{
    fn mod_cholesky_gmw81(&self) -> Decomposition<L, E, P> {
        panic!("Not implemented!")
    }
}
/// Schnabel & Eskow (1990) -- se90.rs:24-32.
pub trait ModCholeskySE90<L, E, P>
where
    Self: Sized,
{
    fn mod_cholesky_se90(&self) -> Decomposition<L, E, P> {
        panic!("Not implemented!")
    }
}
/// Schnabel & Eskow (1999) -- se99.rs:21-29.
pub trait ModCholeskySE99<L, E, P>
where
    Self: Sized,
{
    fn mod_cholesky_se99(&self) -> Decomposition<L, E, P> {
        panic!("Not implemented!")
    }
}

impl ModCholeskyGMW81<Array2<f64>, Array1<f64>, Array1<usize>> for Array2<f64> {
    fn mod_cholesky_gmw81(&self) -> Decomposition<Array2<f64>, Array1<f64>, Array1<usize>> {
        factor2(ffi::MODCHOL_GMW81, self)
    }
}
impl ModCholeskySE90<Array2<f64>, Array1<f64>, Array1<usize>> for Array2<f64> {
    fn mod_cholesky_se90(&self) -> Decomposition<Array2<f64>, Array1<f64>, Array1<usize>> {
        factor2(ffi::MODCHOL_SE90, self)
    }
}
impl ModCholeskySE99<Array2<f64>, Array1<f64>, Array1<usize>> for Array2<f64> {
    fn mod_cholesky_se99(&self) -> Decomposition<Array2<f64>, Array1<f64>, Array1<usize>> {
        factor2(ffi::MODCHOL_SE99, self)
    }
}

/// New: the batched entry point over `Array3<f64>` of shape (batch, n, n).
pub trait ModCholeskyBatched {
    fn mod_cholesky_gmw81_batched(&self) -> Decomposition<Array3<f64>, Array2<f64>, Array2<usize>>;
    fn mod_cholesky_se90_batched(&self) -> Decomposition<Array3<f64>, Array2<f64>, Array2<usize>>;
    fn mod_cholesky_se99_batched(&self) -> Decomposition<Array3<f64>, Array2<f64>, Array2<usize>>;
    /// Same, with the batch split over `ngpus` devices of this box (0 = all).
    fn mod_cholesky_se99_batched_multi_gpu(
        &self, ngpus: i32,
    ) -> Decomposition<Array3<f64>, Array2<f64>, Array2<usize>>;
}
impl ModCholeskyBatched for Array3<f64> {
    fn mod_cholesky_gmw81_batched(&self) -> Decomposition<Array3<f64>, Array2<f64>, Array2<usize>> {
        factor3(ffi::MODCHOL_GMW81, self, 1)
    }
    fn mod_cholesky_se90_batched(&self) -> Decomposition<Array3<f64>, Array2<f64>, Array2<usize>> {
        factor3(ffi::MODCHOL_SE90, self, 1)
    }
    fn mod_cholesky_se99_batched(&self) -> Decomposition<Array3<f64>, Array2<f64>, Array2<usize>> {
        factor3(ffi::MODCHOL_SE99, self, 1)
    }
    fn mod_cholesky_se99_batched_multi_gpu(
        &self, ngpus: i32,
    ) -> Decomposition<Array3<f64>, Array2<f64>, Array2<usize>> {
        factor3(ffi::MODCHOL_SE99, self, ngpus)
    }
}

/// Gershgorin circles -- gershgorin.rs:15-42.
pub trait GershgorinCircles {
    fn gershgorin_circles(&self) -> Vec<(f64, f64)>;
}
impl GershgorinCircles for Array2<f64> {
    fn gershgorin_circles(&self) -> Vec<(f64, f64)> {
        debug_assert!(self.is_square());
        let n = self.nrows();
        let a_std = self.as_standard_layout();
        let mut out = vec![0.0f64; 2 * n];
        check(unsafe {
            ffi::modchol_gershgorin_circles_batched_f64(
                1, n as i32, a_std.as_ptr(), out.as_mut_ptr(), ffi::MODCHOL_MEM_HOST, std::ptr::null_mut(),
            )
        });
        (0..n).map(|i| (out[2 * i], out[2 * i + 1])).collect()
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use crate::utils::{diag_mat_from_arr, index_to_permutation_mat};

    fn max_abs(m: &Array2<f64>) -> f64 {
        m.iter().fold(0.0, |acc, x| acc.max(x.abs()))
    }

    /// Acceptance identity of every reference test (e.g. se99.rs:221-231):
    /// P A P^T + P E P^T = L L^T, on the reference's 3x3 example (lib.rs:32-34).
    #[test]
    fn identity_holds_for_all_three_algorithms() {
        let a = ndarray::arr2(&[[1.0, 1.0, 2.0], [1.0, 1.0, 3.0], [2.0, 3.0, 1.0]]);
        for d in [a.mod_cholesky_gmw81(), a.mod_cholesky_se90(), a.mod_cholesky_se99()] {
            let pm = index_to_permutation_mat(d.p.as_slice().unwrap());
            let em = diag_mat_from_arr(d.e.as_slice().unwrap());
            let lhs = pm.dot(&(&a + &em)).dot(&pm.t());
            assert!(max_abs(&(lhs - d.l.dot(&d.l.t()))) < 1e-12);
        }
    }
}
