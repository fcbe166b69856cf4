// Links librofl_b200.so.  ROFL_B200_DIR = checkout of the rofl_b200 repository (the library is built in-tree by
// `python -c 'import __graft_entry__ as g; g.build()'` or `make -C rofl-project-code_b200/csrc`).
fn main() {
    let dir = std::env::var("ROFL_B200_DIR").expect("set ROFL_B200_DIR to the rofl_b200 checkout");
    println!("cargo:rustc-link-search=native={}/rofl-project-code_b200", dir);
    println!("cargo:rustc-link-lib=dylib=rofl_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}/rofl-project-code_b200", dir);
    println!("cargo:rerun-if-env-changed=ROFL_B200_DIR");
}
