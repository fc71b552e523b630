// build.rs -- compiles the CUDA library with nvcc for sm_100a and links it (no CPU fallback exists).
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    // AKAZE_B200_CSRC = <this repository>/akaze-rust_b200/csrc (the overlay lives inside the reference's akaze/ directory)
    let csrc = env::var("AKAZE_B200_CSRC").map(PathBuf::from).unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../csrc"));
    let lib = out.join("libakaze_b200.so");
    // every source of akaze-rust_b200/build.py (tests/test_host_layer.py keeps the two lists equal)
    let sources = ["akaze_api.cu", "scale_space.cu", "detector.cu", "keypoints.cu", "matcher.cu", "matcher_tc.cu", "ransac.cu"];
    let mut cmd = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()));
    cmd.args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false",
               "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-ldl", "-o"]).arg(&lib);
    for s in &sources {
        cmd.arg(csrc.join(s));
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let status = cmd.status().expect("nvcc not found: the B200 engine cannot be built");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=akaze_b200");
}
