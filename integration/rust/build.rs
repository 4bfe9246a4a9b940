// protocol_decoder/build.rs
// Links the prebuilt CUDA library (built by `make -C proof_protocol_decoder_b200/csrc`,
// nvcc -gencode arch=compute_100a,code=sm_100a) and the CUDA runtime.
fn main() {
    let dir = std::env::var("PPD_B200_LIB_DIR").expect("set PPD_B200_LIB_DIR to the directory holding libppd_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ppd_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-env-changed=PPD_B200_LIB_DIR");
}
