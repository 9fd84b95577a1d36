// Links the prebuilt C-ABI library.  Point MODCHOL_B200_LIB_DIR at the directory holding
// libmodchol_b200.so (default: ../modcholesky_b200/csrc, where `python -m
// modcholesky_b200.build` leaves it).
fn main() {
    let dir = std::env::var("MODCHOL_B200_LIB_DIR")
        .unwrap_or_else(|_| format!("{}/../modcholesky_b200/csrc", env!("CARGO_MANIFEST_DIR")));
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=modchol_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=MODCHOL_B200_LIB_DIR");
}
