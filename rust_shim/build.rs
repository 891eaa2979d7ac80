// Points rustc at the in-tree libb200tfhe.so (built by `make -C tfhe_rs_string_b200/csrc`).
fn main() {
    let dir = std::env::var("B200TFHE_LIB_DIR").unwrap_or_else(|_| "../tfhe_rs_string_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b200tfhe");
}
