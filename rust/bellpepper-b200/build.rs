// Link against the in-tree libbp_r1cs.so (built by `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    let dir = std::env::var("BP_R1CS_LIB_DIR").unwrap_or_else(|_| "../../bellpepper_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=bp_r1cs");
    println!("cargo:rerun-if-env-changed=BP_R1CS_LIB_DIR");
}
