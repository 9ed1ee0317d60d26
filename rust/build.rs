// build.rs of dmmt-jpeg-encoder with the `cuda` feature: links the C-ABI library of the B200 encode path
// (include/dmmt_cuda.h; built by `python -m dmmt_jpeg_encoder_b200.build` into dmmt_jpeg_encoder_b200/lib/).
//
//   DMMT_CUDA_LIB_DIR  directory holding libdmmt_cuda.a (static, default) or libdmmt_cuda.so
//   DMMT_CUDA_SHARED=1 link libdmmt_cuda.so instead of the static archive
//   CUDA_HOME          CUDA toolkit root (default /usr/local/cuda): libcudart
//
// NOT BUILT IN THE DEVELOPMENT IMAGE OF THIS REPOSITORY (no rustc / cargo there); rust/check_layout.py keeps the
// declarations of src/image/writer/jpeg/cuda.rs in step with the header instead.
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=DMMT_CUDA_LIB_DIR");
    println!("cargo:rerun-if-env-changed=DMMT_CUDA_SHARED");
    println!("cargo:rerun-if-env-changed=CUDA_HOME");
    if env::var_os("CARGO_FEATURE_CUDA").is_none() {
        return; // the stock CPU encoder
    }
    let lib_dir = env::var("DMMT_CUDA_LIB_DIR")
        .expect("set DMMT_CUDA_LIB_DIR to the directory holding libdmmt_cuda.a (dmmt_jpeg_encoder_b200/lib)");
    let cuda_home = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".to_owned());
    println!("cargo:rustc-link-search=native={lib_dir}");
    if env::var("DMMT_CUDA_SHARED").map(|v| v == "1").unwrap_or(false) {
        println!("cargo:rustc-link-lib=dylib=dmmt_cuda");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{lib_dir}");
    } else {
        println!("cargo:rustc-link-lib=static=dmmt_cuda");
    }
    println!("cargo:rustc-link-search=native={cuda_home}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
