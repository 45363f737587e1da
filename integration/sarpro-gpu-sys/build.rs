// Links libsarpro_gpu.so (built by `python -m sarpro_b200.build` in the sarpro-b200 tree: nvcc, sm_100a, static cudart).
// SARPRO_GPU_LIB_DIR points at the directory that holds it.
fn main() {
    if let Ok(dir) = std::env::var("SARPRO_GPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=sarpro_gpu");
    println!("cargo:rerun-if-env-changed=SARPRO_GPU_LIB_DIR");
}
