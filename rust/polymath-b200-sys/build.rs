// Builds libpolymath_b200.so for sm_100a with nvcc (BASELINE.json north_star: "a thin C-ABI FFI crate, built by a
// build.rs that invokes nvcc for sm_100a") and links it.  Mirrors build.py of the repository root: every .cu / .cpp
// under polymath_b200/csrc is one translation unit, compiled with the same flags, linked into one shared object.
//
// Environment:
//   POLYMATH_B200_LIB_DIR  directory of a prebuilt libpolymath_b200.so (skips nvcc; required with feature "prebuilt")
//   POLYMATH_B200_ROOT     repository root (default: two levels above this crate)
//   NVCC                   compiler (default: nvcc on PATH, else /usr/local/cuda/bin/nvcc)
use std::env;
use std::fs;
use std::path::{Path, PathBuf};
use std::process::Command;

fn collect(dir: &Path, out: &mut Vec<PathBuf>) {
    let mut entries: Vec<PathBuf> = fs::read_dir(dir)
        .unwrap_or_else(|e| panic!("cannot read {}: {e}", dir.display()))
        .map(|e| e.unwrap().path())
        .collect();
    entries.sort();
    for p in entries {
        if p.is_dir() {
            collect(&p, out);
        } else if matches!(p.extension().and_then(|e| e.to_str()), Some("cu") | Some("cpp")) {
            out.push(p);
        }
    }
}

fn watch_headers(dir: &Path) {
    for e in fs::read_dir(dir).unwrap() {
        let p = e.unwrap().path();
        if p.is_dir() {
            watch_headers(&p);
        } else if matches!(p.extension().and_then(|e| e.to_str()), Some("cuh") | Some("hpp") | Some("h")) {
            println!("cargo:rerun-if-changed={}", p.display());
        }
    }
}

fn main() {
    println!("cargo:rerun-if-env-changed=POLYMATH_B200_LIB_DIR");
    println!("cargo:rerun-if-env-changed=POLYMATH_B200_ROOT");
    println!("cargo:rerun-if-env-changed=NVCC");
    if let Ok(dir) = env::var("POLYMATH_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=polymath_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
        return;
    }
    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        panic!("feature `prebuilt` needs POLYMATH_B200_LIB_DIR");
    }
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let root = env::var("POLYMATH_B200_ROOT").map(PathBuf::from).unwrap_or_else(|_| manifest.join("../.."));
    let csrc = root.join("polymath_b200").join("csrc");
    let include = root.join("include");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| {
        if Command::new("nvcc").arg("--version").output().is_ok() { "nvcc".into() } else { "/usr/local/cuda/bin/nvcc".into() }
    });
    let mut sources = Vec::new();
    collect(&csrc, &mut sources);
    assert!(!sources.is_empty(), "no CUDA sources under {}", csrc.display());
    watch_headers(&csrc);
    println!("cargo:rerun-if-changed={}", include.join("polymath_b200.h").display());
    let mut objects = Vec::new();
    for src in &sources {
        println!("cargo:rerun-if-changed={}", src.display());
        let rel = src.strip_prefix(&csrc).unwrap().to_string_lossy().replace(['/', '\\'], "_");
        let obj = out.join(format!("{rel}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC,-O3", "--expt-relaxed-constexpr"])
            .arg("-I").arg(&include)
            .args(["-x", "cu", "-c"]).arg(src)
            .arg("-o").arg(&obj)
            .status()
            .unwrap_or_else(|e| panic!("cannot run {nvcc}: {e}"));
        assert!(status.success(), "nvcc failed for {}", src.display());
        objects.push(obj);
    }
    let lib = out.join("libpolymath_b200.so");
    let status = Command::new(&nvcc)
        .arg("-shared").arg("-o").arg(&lib)
        .args(&objects)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"])
        .status()
        .expect("link step");
    assert!(status.success(), "linking libpolymath_b200.so failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=polymath_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    println!("cargo:root={}", out.display());
}
