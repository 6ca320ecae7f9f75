"""The Rust boundary as shipped (bindings/rust/): the nvcc command line that build.rs runs must
produce a LOADABLE library exporting every symbol src/lib.rs binds, and src/lib.rs must bind
exactly what include/raiko_kzg.h declares.  No Rust toolchain exists in this image, so the test
replays build.rs's command itself (it is pure data: NVCC_FLAGS + SOURCES)."""
import ctypes
import os
import re
import shutil
import subprocess

import pytest

from test_abi import header_functions

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SYS = os.path.join(ROOT, "bindings", "rust", "raiko-kzg-sys")


def _rust_str_array(src: str, name: str):
    m = re.search(r"const %s: &\[&str\] = &\[(.*?)\];" % name, src, flags=re.S)
    assert m, "build.rs has no %s array" % name
    return re.findall(r'"([^"]*)"', m.group(1))


def rust_extern_functions():
    src = open(os.path.join(SYS, "src", "lib.rs")).read()
    return sorted(set(re.findall(r"pub fn (rk_[a-z0-9_]+)\s*\(", src)))


def test_sys_crate_binds_exactly_the_header():
    assert rust_extern_functions() == header_functions()


def test_build_rs_lists_every_translation_unit():
    src = open(os.path.join(SYS, "build.rs")).read()
    listed = set(_rust_str_array(src, "SOURCES"))
    present = {f for f in os.listdir(os.path.join(ROOT, "raiko_b200", "csrc")) if f.endswith(".cu")}
    assert listed == present, "build.rs SOURCES %s != csrc/*.cu %s" % (sorted(listed), sorted(present))
    flags = _rust_str_array(src, "NVCC_FLAGS")
    assert "arch=compute_100a,code=sm_100a" in flags and "-shared" in flags


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_build_rs_command_builds_a_loadable_library(tmp_path):
    """Round 1 shipped a build.rs that compiled kzg_ctx.cu alone: it linked, and then failed at
    dlopen with `undefined symbol: rk::launch_k_roots_export`.  This runs what build.rs runs."""
    src = open(os.path.join(SYS, "build.rs")).read()
    flags, sources = _rust_str_array(src, "NVCC_FLAGS"), _rust_str_array(src, "SOURCES")
    csrc = os.path.join(ROOT, "raiko_b200", "csrc")
    lib = str(tmp_path / "libraiko_kzg.so")
    cmd = ["nvcc"] + flags + ["-o", lib] + [os.path.join(csrc, s) for s in sources]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    so = ctypes.CDLL(lib)                     # RTLD_NOW: an undefined launch_* symbol fails here
    for name in rust_extern_functions():
        assert hasattr(so, name), "library built by build.rs does not export %s" % name
    so.rk_version.restype = ctypes.c_char_p
    assert b"sm_100a" in so.rk_version()


def test_lib_patch_keeps_the_reference_surface():
    """Every public item of lib/src/primitives/eip4844.rs (14-99, and the re-exports of :13 that
    core/src/preflight.rs:305-316 uses) is present in the replacement module."""
    src = open(os.path.join(ROOT, "bindings", "rust", "raiko-lib-patch", "eip4844_gpu.rs")).read()
    for item in ("KZG_SETTINGS_BIN", "KZG_SETTINGS", "VERSIONED_HASH_VERSION_KZG", "KzgGroup", "KzgField", "KzgCommitment",
                 "Eip4844Error", "get_evaluation_point", "proof_of_equivalence", "calc_kzg_proof", "calc_kzg_proof_with_point",
                 "calc_kzg_proof_commitment", "commitment_to_version_hash", "kzg_proof_to_bytes",
                 "KZGSettings", "deserialize_blob_rust", "blob_to_kzg_commitment_rust", "Blob",
                 "KzgSettings", "blob_to_kzg_commitment", "compute_kzg_proof", "kzg_to_versioned_hash"):
        assert re.search(r"\bpub (static|const|type|struct|enum|fn|use)\b[^;{]*\b%s\b" % item, src), item
    assert "impl Clone" in src or "derive(Clone)" in src      # KZG_SETTINGS.clone() at preflight.rs:311
