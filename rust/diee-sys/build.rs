// Links libdiee_cuda.so.  DIEE_LIB_DIR = the directory that holds it (die_e_b200/ of this repository after
// `python die_e_b200/build.py`); the library itself needs libcudart and, at run time only, libnccl.so.2 (dlopen).
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=DIEE_LIB_DIR");
    match env::var("DIEE_LIB_DIR") {
        Ok(dir) => {
            println!("cargo:rustc-link-search=native={}", dir);
            println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
        }
        Err(_) => println!("cargo:warning=DIEE_LIB_DIR is not set: libdiee_cuda.so must be on the linker's search path"),
    }
    println!("cargo:rustc-link-lib=dylib=diee_cuda");
}
