// bindings/rust/build.rs — link against the C-ABI shared library built by `make -C serenade_b200/csrc`.
// VMIS_B200_LIB_DIR points at the directory that holds libvmis_b200.so (default: ../../serenade_b200).
fn main() {
    let dir = std::env::var("VMIS_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../../serenade_b200", here)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=vmis_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=VMIS_B200_LIB_DIR");
}
