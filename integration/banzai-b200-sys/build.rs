// BANZAI_B200_LIB_DIR = directory that holds libbanzai_b200.so (banzai_b200/ in this repository)
fn main() {
    let dir = std::env::var("BANZAI_B200_LIB_DIR").expect("set BANZAI_B200_LIB_DIR to the directory of libbanzai_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=banzai_b200");
    println!("cargo:rerun-if-env-changed=BANZAI_B200_LIB_DIR");
}
