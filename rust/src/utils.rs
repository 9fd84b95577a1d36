//! Host-side helpers with the reference's names (utils.rs:14-147), including the seeded
//! generators `random_householder` / `random_diagonal` (public through `pub mod utils`,
//! lib.rs:102), plus `assemble_qdqt_batched`, the device-side assembly of the bench inputs
//! (benches/bench.rs:48-53) over the C-ABI.  The Python mirror with a bit-exact restatement of
//! the generators' arithmetic lives in `modcholesky_b200/utils.py`.

/// utils.rs:14-27
pub fn eigenvalues_2x2(mat: &ndarray::ArrayView2<f64>) -> (f64, f64) {
    let (a, b, c, d) = (mat[(0, 0)], mat[(0, 1)], mat[(1, 0)], mat[(1, 1)]);
    let tmp = ((-(a + d) / 2.0).powi(2) - a * d + b * c).sqrt();
    let (l1, l2) = ((a + d) / 2.0 + tmp, (a + d) / 2.0 - tmp);
    if l1 > l2 { (l1, l2) } else { (l2, l1) }
}

/// utils.rs:30-38
pub fn swap_columns<T>(mat: &mut ndarray::Array2<T>, idx1: usize, idx2: usize) {
    for i in 0..mat.nrows() {
        mat.swap((i, idx1), (i, idx2));
    }
}

/// utils.rs:41-49
pub fn swap_rows<T>(mat: &mut ndarray::Array2<T>, idx1: usize, idx2: usize) {
    for i in 0..mat.ncols() {
        mat.swap((idx1, i), (idx2, i));
    }
}

/// utils.rs:52-70 -- first strict maximum
pub fn index_of_largest(c: &ndarray::ArrayView1<f64>) -> usize {
    let (mut max, mut max_idx) = (c[0], 0);
    for (i, &ci) in c.iter().enumerate() {
        if ci > max {
            max = ci;
            max_idx = i;
        }
    }
    max_idx
}

/// utils.rs:73-91
pub fn index_of_largest_abs(c: &ndarray::ArrayView1<f64>) -> usize {
    let (mut max, mut max_idx) = (c[0].abs(), 0);
    for (i, &ci) in c.iter().enumerate() {
        if ci.abs() > max {
            max = ci.abs();
            max_idx = i;
        }
    }
    max_idx
}

/// utils.rs:94-101
pub fn index_to_permutation_mat(idxs: &[usize]) -> ndarray::Array2<f64> {
    let n = idxs.len();
    let mut mat = ndarray::Array2::zeros((n, n));
    for i in 0..n {
        mat[(i, idxs[i])] = 1.0;
    }
    mat
}

/// utils.rs:104-113
pub fn diag_mat_from_arr(arr: &[f64]) -> ndarray::Array2<f64> {
    ndarray::Array2::from_diag(&ndarray::Array1::from(arr.to_vec()))
}

use ndarray_rand::rand_distr::Uniform;
use ndarray_rand::RandomExt;
use rand::prelude::*;

/// utils.rs:116-121 -- a random Householder matrix of dimension `dim` from `seed`.
pub fn random_householder(dim: usize, seed: u8) -> ndarray::Array2<f64> {
    let w = householder_vector(dim, seed).insert_axis(ndarray::Axis(1));
    let denom = w.fold(0.0, |acc, &x: &f64| acc + x.powi(2));
    ndarray::Array::eye(dim) - 2.0 / denom * w.dot(&w.t())
}

/// The vector `w` of `random_householder` (utils.rs:117-118): what `assemble_qdqt_batched` takes.
pub fn householder_vector(dim: usize, seed: u8) -> ndarray::Array1<f64> {
    let mut rng = rand_xorshift::XorShiftRng::from_seed([seed; 16]);
    ndarray::Array::random_using(dim, Uniform::new_inclusive(-1.0, 1.0), &mut rng)
}

/// utils.rs:126-147 -- a random diagonal matrix with eigenvalues in the given range and at
/// least `min_neg_eigenvalues` negative ones.
pub fn random_diagonal(
    dim: usize,
    (eig_range_min, eig_range_max): (f64, f64),
    min_neg_eigenvalues: usize,
    seed: u8,
) -> ndarray::Array2<f64> {
    let mut rng = rand_xorshift::XorShiftRng::from_seed([seed; 16]);
    let mut w = ndarray::Array::random_using(dim, Uniform::new_inclusive(eig_range_min, eig_range_max), &mut rng);
    let mut idxs: Vec<usize> = (0..dim).collect();
    for _ in 0..min_neg_eigenvalues {
        let idxidx: usize = (rng.gen::<f64>() * (idxs.len() - 1) as f64).floor() as usize;
        let idx = idxs.remove(idxidx);
        w[idx] = rng.gen::<f64>() - 1.0;
    }
    let mut out = ndarray::Array::eye(dim);
    out.diag_mut().assign(&w);
    out
}

/// `A_b = Q_b diag(d_b) Q_b^T`, `Q_b` the product of the Householder reflectors of `v[b, 0..]`,
/// assembled on the GPU (`modchol_assemble_qdqt_batched_f64`): the batched form of
/// benches/bench.rs:48-53.  `v`: batch x nrefl x n, `d`: batch x n.
pub fn assemble_qdqt_batched(v: &ndarray::Array3<f64>, d: &ndarray::Array2<f64>) -> ndarray::Array3<f64> {
    let (batch, nrefl, n) = v.dim();
    assert_eq!(d.dim(), (batch, n));
    let v = v.as_standard_layout();
    let d = d.as_standard_layout();
    let mut a = ndarray::Array3::<f64>::zeros((batch, n, n));
    let rc = unsafe {
        crate::ffi::modchol_assemble_qdqt_batched_f64(
            batch as i64, n as i32, nrefl as i32, v.as_ptr(), d.as_ptr(), a.as_mut_ptr(),
            crate::ffi::MODCHOL_MEM_HOST, std::ptr::null_mut())
    };
    crate::check(rc);
    a
}
