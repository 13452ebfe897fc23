// build.rs — links libfluxb200.so.  (UNCOMPILED here: no Rust toolchain in this image.)
//
// The library is built by nvcc for sm_100a, outside cargo:
//     python -c "import __graft_entry__ as g; g.build()"      ->  flux_b200/lib/libfluxb200.so
// FLUXB200_LIB_DIR names the directory that holds it (default: ../../flux_b200/lib relative to this crate).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("FLUXB200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../flux_b200/lib")
    });
    if !dir.join("libfluxb200.so").exists() {
        panic!("libfluxb200.so not found in {} — build it first (see this file's header) or set FLUXB200_LIB_DIR",
               dir.display());
    }
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=fluxb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=FLUXB200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/fluxb200.h");
}
