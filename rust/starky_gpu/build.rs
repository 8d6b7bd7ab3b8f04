// Links libstarkyb200.so (built by `make -C starky_bls12_381_b200/csrc`).  STARKYB200_LIB_DIR overrides the location.
fn main() {
    println!("cargo:rerun-if-env-changed=CARGO_FEATURE_GPU");
    if std::env::var("CARGO_FEATURE_GPU").is_err() { return; }   // pure-Rust build: nothing to link
    let dir = std::env::var("STARKYB200_LIB_DIR").unwrap_or_else(|_| {
        format!("{}/../../starky_bls12_381_b200", std::env::var("CARGO_MANIFEST_DIR").unwrap())
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=starkyb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=STARKYB200_LIB_DIR");
}
