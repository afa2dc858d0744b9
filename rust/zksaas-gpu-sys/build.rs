// Links libzksaas_gpu.so.  ZKSAAS_GPU_LIB_DIR points at the directory that holds it (in this repository:
// zk-saas_b200/, after `python -c 'import __graft_entry__ as g; g.build()'`).
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=ZKSAAS_GPU_LIB_DIR");
    if let Ok(dir) = env::var("ZKSAAS_GPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        // so that `cargo test` / the prover binary find the library without LD_LIBRARY_PATH
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=zksaas_gpu");
}
