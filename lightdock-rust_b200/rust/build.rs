// build.rs — add to the upstream crate root (Cargo.toml: `build = "build.rs"`, `links = "lightdock_b200"`).
// Compiles the CUDA library for sm_100a only (no multi-arch fat binary, no CPU fallback) and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from(env::var("LIGHTDOCK_B200_SRC").unwrap_or_else(|_| "lightdock-rust_b200".into()));
    let lib = out.join("liblightdock_b200.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17"])
        .args(["-Xcompiler", "-fPIC", "-shared"])
        .arg(format!("-I{}", src.join("../include").display()))
        .arg("-o").arg(&lib)
        .arg(src.join("csrc/ld_capi.cu")).arg(src.join("csrc/ld_probe.cu"))
        .status()
        .expect("nvcc not found: the B200 scoring path has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=lightdock_b200");
    println!("cargo:rerun-if-changed={}", src.join("csrc").display());
}
