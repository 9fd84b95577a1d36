//! Host-side helpers with the reference's names (utils.rs:14-113).  The seeded generators
//! (`random_householder`, `random_diagonal`, utils.rs:116-147) are test-input synthesis and
//! stay in the reference crate; their bit-exact restatement lives in
//! `modcholesky_b200/utils.py`.

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
