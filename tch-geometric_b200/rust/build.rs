// build.rs for the reference crate (Cargo.toml gains `build = "build.rs"`; the reference has none today).
// Compiles the CUDA sources for sm_100a with nvcc and links them statically into the cdylib.
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo/rustc); delivered as the glue a maintainer adds.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("TCHGEO_CSRC").unwrap_or_else(|_| "cuda/csrc".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = vec![];
    for src in ["capi.cu", "csx_build.cu", "csx_transform.cu", "gather.cu", "negative_sampling.cu", "neighbor_sampling.cu",
                "partitioned.cu", "partitioned_fixed.cu", "random_walk.cu", "relabel.cu", "transport.cu"] {
        let obj = out.join(src.replace(".cu", ".o"));
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                   "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(src)).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {src}");
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
        objs.push(obj);
    }
    // host half of the compact transport: plain C++ (AVX-512 paths are selected at run time)
    let hu = out.join("host_unpack.o");
    assert!(Command::new("g++").args(["-std=c++17", "-O3", "-fPIC", "-c"]).arg(csrc.join("host_unpack.cpp")).arg("-o").arg(&hu)
        .status().unwrap().success());
    objs.push(hu);
    let lib = out.join("libtchgeo_cuda.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=tchgeo_cuda");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
    // torch's current CUDA stream for the bindings (tch-rs does not expose it): one C++ file against the libtorch headers
    // the reference's torch-sys build already requires (LIBTORCH)
    let libtorch = env::var("LIBTORCH").expect("LIBTORCH must point at the libtorch the crate links against");
    let shim = out.join("torch_stream_shim.o");
    assert!(Command::new("g++").args(["-std=c++17", "-O2", "-fPIC", "-c"]).arg(csrc.join("torch_stream_shim.cpp"))
        .arg(format!("-I{libtorch}/include")).arg("-I/usr/local/cuda/include").arg("-o").arg(&shim).status().unwrap().success());
    assert!(Command::new("ar").arg("crs").arg(out.join("libtchgeo_shim.a")).arg(&shim).status().unwrap().success());
    println!("cargo:rustc-link-lib=static=tchgeo_shim");
    println!("cargo:rustc-link-lib=c10_cuda");
}
