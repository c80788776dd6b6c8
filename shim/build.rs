// NOT COMPILED here (no Rust toolchain in the build image).
fn main() {
    let dir = std::env::var("SAILGPU_LIB_DIR").unwrap_or_else(|_| "../sail_b200/_build".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sailgpu");
    println!("cargo:rerun-if-env-changed=SAILGPU_LIB_DIR");
}
