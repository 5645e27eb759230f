fn main() {
    // SFGPU_LIB_DIR = directory that holds libsfgpu.so (solverforge_b200/ in this repository)
    if let Ok(dir) = std::env::var("SFGPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=sfgpu");
    println!("cargo:rerun-if-env-changed=SFGPU_LIB_DIR");
}
