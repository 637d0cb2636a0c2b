// Links libacm.so.  ACM_LIB_DIR points at apex_camera_models_b200/lib of this repository.
fn main() {
    let dir = std::env::var("ACM_LIB_DIR").unwrap_or_else(|_| "../../apex_camera_models_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=acm");
    println!("cargo:rerun-if-env-changed=ACM_LIB_DIR");
}
