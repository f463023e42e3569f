// PRE-STAGED, UNCOMPILED.  Points rustc at the in-tree libzkmsm.so (built by `python -c "import __graft_entry__ as g;
// g.build()"` with nvcc for sm_100a).  ZKMSM_LIB_DIR overrides the search directory.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("ZKMSM_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
        manifest.join("..").join("..").join("zkvm_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=zkmsm");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=ZKMSM_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/zkmsm.h");
}
