//! matrices/s of the reference crate's own `mod_cholesky_{gmw81,se90,se99}` (gmw81.rs:39,
//! se90.rs:38, se99.rs:35) looped over a batch of symmetric-uniform matrices with rayon --
//! the CPU baseline BASELINE.json's north star names.  Prints one JSON line shaped like
//! `bench.py --impl reference`.
//!
//!     cargo run --release -- <gmw81|se90|se99> <n> <batch> [steps]
use modcholesky::{ModCholeskyGMW81, ModCholeskySE90, ModCholeskySE99};
use ndarray::Array2;
use rand::{Rng, SeedableRng};
use rayon::prelude::*;
use std::time::Instant;

/// Family F1 of SURVEY.md 8(d): symmetric, entries ~ U[-1, 1] (indefinite; SE90/SE99 enter
/// phase 2 at column 0).
fn sym_uniform(n: usize, seed: u64) -> Array2<f64> {
    let mut rng = rand_xorshift::XorShiftRng::seed_from_u64(seed);
    let mut a = Array2::<f64>::zeros((n, n));
    for i in 0..n {
        for j in 0..=i {
            let v: f64 = rng.gen_range(-1.0..=1.0);
            a[(i, j)] = v;
            a[(j, i)] = v;
        }
    }
    a
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let algo = args.get(1).map(String::as_str).unwrap_or("se99").to_string();
    let n: usize = args.get(2).and_then(|s| s.parse().ok()).unwrap_or(32);
    let batch: usize = args.get(3).and_then(|s| s.parse().ok()).unwrap_or(1 << 18);
    let steps: usize = args.get(4).and_then(|s| s.parse().ok()).unwrap_or(3);
    let mats: Vec<Array2<f64>> = (0..batch).into_par_iter().map(|i| sym_uniform(n, 0xC0FFEE + i as u64)).collect();
    let run = |m: &Array2<f64>| -> f64 {
        // the result is consumed so that the call cannot be optimised away
        match algo.as_str() {
            "gmw81" => m.mod_cholesky_gmw81().e[0],
            "se90" => m.mod_cholesky_se90().e[0],
            _ => m.mod_cholesky_se99().e[0],
        }
    };
    let warm: f64 = mats.par_iter().take(batch / 8 + 1).map(run).sum();
    let t0 = Instant::now();
    let mut sink = warm;
    for _ in 0..steps {
        sink += mats.par_iter().map(run).sum::<f64>();
    }
    let dt = t0.elapsed().as_secs_f64();
    let value = (steps * batch) as f64 / dt;
    println!(
        "{{\"impl\": \"reference\", \"metric\": \"matrices_per_sec\", \"value\": {:.1}, \"unit\": \"matrices/s\", \
         \"steps\": {}, \"ms_per_step\": {:.3}, \"higher_is_better\": true, \"dtype\": \"f64\", \"data\": \"synthetic\", \
         \"config\": {{\"workload\": \"{} n={}, {} symmetric-uniform matrices per step\"}}, \
         \"cpu_baseline\": {{\"value\": {:.1}, \"unit\": \"matrices/s\", \"cores\": {}, \"kind\": \"reference\", \
         \"sample\": \"argmin-rs/modcholesky 0.2.0 looped with rayon\"}}, \"sink\": {:e}}}",
        value, steps, 1e3 * dt / steps as f64, algo, n, batch, value, rayon::current_num_threads(), sink
    );
}
