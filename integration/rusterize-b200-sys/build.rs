// Links librz_b200.so.  RZ_B200_LIB_DIR = directory holding the library (rusterize_b200/ of this repository after
// `python -m rusterize_b200.build`); the library links cudart statically and needs only the NVIDIA driver.
fn main() {
    println!("cargo:rerun-if-env-changed=RZ_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("RZ_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=rz_b200");
}
