// Links libdust_b200.so (built by `make -C dust_b200/csrc`).  DUST_B200_LIB_DIR overrides the in-tree location.
fn main() {
    let dir = std::env::var("DUST_B200_LIB_DIR").unwrap_or_else(|_| {
        let root = std::path::Path::new(env!("CARGO_MANIFEST_DIR")).join("../../dust_b200/lib");
        root.to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=dust_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=DUST_B200_LIB_DIR");
}
