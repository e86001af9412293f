// Links libzkgpu.so.  ZKGPU_LIB_DIR points at zkos-monorepo_b200/lib (where `make -C zkos-monorepo_b200/csrc` puts it).
fn main() {
    if let Ok(dir) = std::env::var("ZKGPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=zkgpu");
    println!("cargo:rerun-if-env-changed=ZKGPU_LIB_DIR");
}
