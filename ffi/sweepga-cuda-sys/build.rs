// build.rs — compiles the CUDA / C++ sources of libsweepga_b200 with nvcc for sm_100a (Blackwell B200) and links the
// result.  One architecture, no PTX fallback, no CPU path: on a machine without nvcc the build fails here, and on a
// machine without an sm_100 device swg_create() fails at run time with a message.
//
//   SWEEPGA_B200_SRC   directory holding filter_pipeline.cu, host_parsers.cpp, paf_io.cpp, multi_gpu.cpp and the .cuh/.h files
//                      (default: ../../sweepga_b200/csrc relative to this crate)
//   SWEEPGA_B200_INC   directory holding sweepga_b200.h (default: ../../include)
//   NVCC               the compiler (default: /usr/local/cuda/bin/nvcc)
use std::{env, path::PathBuf, process::Command};

fn main() {
    let here = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let src = env::var("SWEEPGA_B200_SRC").map(PathBuf::from).unwrap_or_else(|_| here.join("../../sweepga_b200/csrc"));
    let inc = env::var("SWEEPGA_B200_INC").map(PathBuf::from).unwrap_or_else(|_| here.join("../../include"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libsweepga_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let status = Command::new(&nvcc)
        .args([
            "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
            "-fmad=false", // parity-critical f64 expressions: no FMA contraction
            "--extended-lambda", "-Xcompiler", "-fPIC,-O3,-pthread", "-cudart", "static", "-shared",
        ])
        .arg(format!("-I{}", inc.display()))
        .arg(format!("-I{}", src.display()))
        .arg("-o")
        .arg(&lib)
        .args(["filter_pipeline.cu", "host_parsers.cpp", "paf_io.cpp", "multi_gpu.cpp"].iter().map(|f| src.join(f)))
        .args(["-lpthread", "-ldl", "-lrt", "-lz"])
        .status()
        .unwrap_or_else(|e| panic!("cannot run {nvcc}: {e}"));
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=sweepga_b200");
    println!("cargo:rerun-if-changed={}", src.display());
    println!("cargo:rerun-if-changed={}", inc.display());
    println!("cargo:rerun-if-env-changed=NVCC");
}
