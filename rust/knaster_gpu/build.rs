// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo/rustc).  The text of INTEGRATION.md as files:
// the binding a knaster maintainer would add next to knaster_graph.  tests/test_host_plan.py checks that
// src/ffi.rs declares every entry point of include/knaster_gpu.h.
// knaster_gpu/build.rs -- builds the CUDA library with nvcc for sm_100a and links it
use std::{env, path::PathBuf, process::Command};
fn main() {
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("csrc"); // = knaster_b200/csrc
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let status = Command::new("make").arg("-C").arg(&csrc).arg(format!("OUT={}", out.display())).status().unwrap();
    assert!(status.success(), "nvcc build of libknaster_gpu.so failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=knaster_gpu");
    println!("cargo:rerun-if-changed=csrc");
}
