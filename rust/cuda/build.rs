// UNVERIFIED (never compiled here).  Builds librtw_cuda.so from the CUDA sources of this repository with
// nvcc for sm_100a — the same three commands as ray_tracing_weekend_b200/build.py — and links it.
// RTW_CUDA_SRC points at the checkout of the CUDA backend (the directory that holds include/rtw.h).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let src = PathBuf::from(env::var("RTW_CUDA_SRC").expect("set RTW_CUDA_SRC to the CUDA backend checkout"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = src.join("ray_tracing_weekend_b200/csrc");
    let arch = ["-gencode", "arch=compute_100a,code=sm_100a"];
    let common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-fmad=false"];
    let mut objs = vec![];
    for unit in ["kernels_f32.cu", "kernels_f64.cu", "capi.cu"] {
        let obj = out.join(unit.replace(".cu", ".o"));
        let ok = Command::new("nvcc").args(arch).args(common).arg("-c").arg(csrc.join(unit)).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {unit}");
        objs.push(obj);
    }
    let lib = out.join("librtw_cuda.so");
    let ok = Command::new("nvcc").args(arch).args(["-shared", "-cudart", "static", "-o"]).arg(&lib).args(&objs)
        .status().unwrap().success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rtw_cuda");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", src.join("include/rtw.h").display());
}
