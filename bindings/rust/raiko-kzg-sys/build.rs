// Builds libraiko_kzg.so with nvcc for sm_100a only (no other backend, no CPU fallback).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    // csrc = <repo>/raiko_b200/csrc (set RAIKO_KZG_CSRC to override)
    let csrc = env::var("RAIKO_KZG_CSRC").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../../raiko_b200/csrc")
    });
    let lib = out.join("libraiko_kzg.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("kzg_ctx.cu"))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=raiko_kzg");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
