// build.rs for the shimmed crate: link liblmrs_b200.so (set LMRS_B200_LIB_DIR to lm.rs_b200/lmrs_b200).
fn main() {
    let dir = std::env::var("LMRS_B200_LIB_DIR").unwrap_or_else(|_| "../lmrs_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=lmrs_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
}
