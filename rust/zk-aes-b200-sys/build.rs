// Locates libzkaes_b200.so.  ZKAES_B200_LIB_DIR names the directory that holds it; the default is this repository's in-tree
// build (aes_zero_knowledge_proof_circuit_b200/, produced by `make -C aes_zero_knowledge_proof_circuit_b200/csrc`).
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=ZKAES_B200_LIB_DIR");
    let dir = env::var("ZKAES_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").expect("CARGO_MANIFEST_DIR"));
        manifest.join("..").join("..").join("aes_zero_knowledge_proof_circuit_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=zkaes_b200");
    // let binaries find the library next to where it was linked from
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
