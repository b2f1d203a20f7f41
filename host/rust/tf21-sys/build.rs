// Links against the in-tree CUDA library.  TF21_LIB_DIR defaults to ../../../twenty-first_b200.
fn main() {
    let dir = std::env::var("TF21_LIB_DIR").unwrap_or_else(|_| {
        format!("{}/../../../twenty-first_b200", env!("CARGO_MANIFEST_DIR"))
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=tf21");
    println!("cargo:rerun-if-env-changed=TF21_LIB_DIR");
}
