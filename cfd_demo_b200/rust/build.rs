// build.rs — compiles the sm_100a kernels and links them into the crate (SOURCE ONLY: no Rust toolchain
// exists in the build environment of this repository, so this file has never been run).
// Attach as `build = "build.rs"` in the reference's Cargo.toml; `cc` is already in its Cargo.lock (1.2.13).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CFD_B200_ROOT").unwrap_or_else(|_| "../..".into()));
    let csrc = root.join("cfd_demo_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libcfd_b200.so");
    let status = Command::new("nvcc")
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false"])
        .args(["-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("cfd_model.cu"))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=cfd_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
