// Links libwebsplat_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    let dir = std::env::var("WEBSPLAT_B200_LIB_DIR").unwrap_or_else(|_| "../../web-splat_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=websplat_b200");
    println!("cargo:rerun-if-env-changed=WEBSPLAT_B200_LIB_DIR");
}
