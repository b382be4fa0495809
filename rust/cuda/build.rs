// UNVERIFIED (never compiled here: the build image has no Rust toolchain).  Builds librtw_cuda.so from the CUDA sources of this
// repository with nvcc for sm_100a — the same five translation units, flags and link line as ray_tracing_weekend_b200/build.py
// (UNITS / COMMON / ARCH there; tests/test_abi_and_host.py::test_rust_build_recipe_matches_build_py keeps the two lists equal) — and
// links it.  RTW_CUDA_SRC points at the checkout of the CUDA backend (the directory that holds include/rtw.h).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let src = PathBuf::from(env::var("RTW_CUDA_SRC").expect("set RTW_CUDA_SRC to the CUDA backend checkout"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = src.join("ray_tracing_weekend_b200/csrc");
    let host = src.join("ray_tracing_weekend_b200/host");
    let arch = ["-gencode", "arch=compute_100a,code=sm_100a"];
    let common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"];
    // (unit, extra flags): the kernel units are compiled with -fmad=false — FP32 fused multiply-adds are explicit fmaf calls so that
    // every kernel rounds alike, and the f64 path must not contract at all (Rust never does)
    let units: [(PathBuf, &[&str]); 5] = [
        (csrc.join("kernels_f32.cu"), &["-fmad=false"]),
        (csrc.join("kernels_f64.cu"), &["-fmad=false"]),
        (csrc.join("bvh_device.cu"), &["-fmad=false"]),
        (csrc.join("capi.cu"), &[]),
        (host.join("rtw_host_capi.cpp"), &[]),
    ];
    let mut objs = vec![];
    for (unit, extra) in units.iter() {
        let obj = out.join(unit.file_stem().unwrap()).with_extension("o");
        let ok = Command::new("nvcc").args(arch).args(common).args(*extra).arg("-c").arg(unit).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {}", unit.display());
        objs.push(obj);
    }
    let lib = out.join("librtw_cuda.so");
    // cudart is linked statically; NCCL is NOT linked: the library dlopen()s libnccl.so.2 at first use (rtw_render_multi with
    // RTW_COLLECTIVE_NCCL, rtw_comm_*), hence -ldl
    let ok = Command::new("nvcc").args(arch).args(["-shared", "-cudart", "static", "-o"]).arg(&lib).args(&objs).arg("-ldl")
        .status().unwrap().success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rtw_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", host.display());
    println!("cargo:rerun-if-changed={}", src.join("include/rtw.h").display());
}
