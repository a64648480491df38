// Builds libpantax_gpu.so from the CUDA sources with nvcc (sm_100a) and links it.
// north_star: "a thin extern "C" layer built by build.rs and nvcc".  Mirrors pantax_b200/build.py.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("PANTAX_B200_ROOT").unwrap_or_else(|_| "../..".into()));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = root.join("pantax_b200/csrc");
    let lib = out.join("libpantax_gpu.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let status = Command::new(nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "-Xcompiler", "-fPIC", "--shared"])
        .arg(csrc.join("ptx_kernels.cu"))
        .arg(csrc.join("ptx_api.cu"))
        .arg("-o").arg(&lib)
        .args(["-lcudart", "-ldl"])
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=pantax_gpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    for f in ["ptx_kernels.cu", "ptx_api.cu", "ptx_core.cuh", "ptx_fast.cuh", "ptx_fxorder.h", "ptx_internal.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/pantax_gpu.h").display());
}
