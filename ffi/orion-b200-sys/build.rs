// Links liborion_b200.so.  ORION_B200_LIB_DIR points at the directory holding the library
// (orion-sdr_b200/lib in this repository); no bindgen, the declarations in src/lib.rs are
// written by hand against include/orion_b200.h (ABI version 1).
fn main() {
    if let Ok(dir) = std::env::var("ORION_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=orion_b200");
    println!("cargo:rerun-if-env-changed=ORION_B200_LIB_DIR");
}
