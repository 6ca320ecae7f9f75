// Builds libraiko_kzg.so with nvcc for sm_100a only (no other backend, no CPU fallback).
// The kernels live in five translation units beside the host file; ALL of them go on one nvcc
// command line (`-t 0` compiles them in parallel).  tests/test_rust_boundary.py extracts this
// exact argument list, runs it, dlopens the result and checks every rk_* symbol of src/lib.rs.
use std::{env, path::PathBuf, process::Command};

const NVCC_FLAGS: &[&str] = &[
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-t", "0",
];
const SOURCES: &[&str] = &[
    "kzg_ctx.cu", "tu_msm.cu", "tu_path.cu", "tu_table.cu", "tu_verify.cu", "tu_pairing.cu",
];

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    // csrc = <repo>/raiko_b200/csrc (set RAIKO_KZG_CSRC to override)
    let csrc = env::var("RAIKO_KZG_CSRC").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../../raiko_b200/csrc")
    });
    let lib = out.join("libraiko_kzg.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(NVCC_FLAGS)
        .arg("-o")
        .arg(&lib)
        .args(SOURCES.iter().map(|s| csrc.join(s)))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=raiko_kzg");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
