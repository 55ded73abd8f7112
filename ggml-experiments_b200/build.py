"""In-tree build of the native libraries (nvcc, sm_100a only).

    python -m ggml_experiments_b200.build        # or: python ggml-experiments_b200/build.py

Outputs (git-ignored, shipped to the GPU box by gpurun):
    _build/libggml_b200.so      the drop-in ggml boundary + CUDA kernels   (include/ggml/ggml.h)
    _build/libmobilevit_b200.so the host MobileViT program + its C ABI     (include/mobilevit_b200.h)
    _build/mobilevit_cli        the reference's main() equivalent
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
INC = os.path.join(ROOT, "include")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-I" + INC,
          "-I" + os.path.join(HERE, "csrc")]

LIB_SOURCES = ["csrc/ggml_b200.cpp", "csrc/plan.cpp", "csrc/exec_exact.cu", "csrc/fuse.cpp", "csrc/fast_kernels.cu",
               "csrc/gemm_tcgen05.cu", "csrc/dwconv_tma.cu", "csrc/dwreduce.cu", "csrc/ir_fused.cu", "csrc/attention_tc.cu", "csrc/vit_stage.cu", "csrc/debug_api.cu"]
HOST_SOURCES = ["host/mobilevit.cpp", "host/gru.cpp"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _all_deps() -> list[str]:
    deps = []
    for base in (os.path.join(HERE, "csrc"), os.path.join(HERE, "host"), os.path.join(INC), os.path.join(INC, "ggml")):
        for fn in os.listdir(base):
            if fn.endswith((".h", ".cuh", ".cu", ".cpp")):
                deps.append(os.path.join(base, fn))
    return deps


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force: bool = False, verbose_ptxas: bool = False) -> dict:
    os.makedirs(OUT, exist_ok=True)
    deps = _all_deps()
    lib = os.path.join(OUT, "libggml_b200.so")
    host = os.path.join(OUT, "libmobilevit_b200.so")
    cli = os.path.join(OUT, "mobilevit_cli")
    srcs = [os.path.join(HERE, s) for s in LIB_SOURCES if os.path.exists(os.path.join(HERE, s))]
    if force or _newer(lib, deps):
        objs = []
        for s in srcs:  # one object per TU so an edit recompiles only that file
            o = os.path.join(OUT, os.path.basename(s) + ".o")
            if force or _newer(o, [s] + [d for d in deps if d.endswith((".h", ".cuh"))]):
                extra = ["-Xptxas", "-v"] if verbose_ptxas else []
                if os.environ.get("GGML_B200_IR_PROFILE"):  # clock64 phase profile inside the fused inverted-residual kernel
                    extra.append("-DGGML_B200_IR_PROFILE")
                if os.environ.get("GGML_B200_VIT_PROFILE"):
                    extra.append("-DGGML_B200_VIT_PROFILE")
                if os.environ.get("GGML_B200_GEMM_PROFILE"):
                    extra.append("-DGGML_B200_GEMM_PROFILE")
                if os.environ.get("GGML_B200_ATTN_PROFILE"):
                    extra.append("-DGGML_B200_ATTN_PROFILE")
                _run([NVCC, *ARCH, *COMMON, *extra, "-c", s, "-o", o])
            objs.append(o)
        _run([NVCC, *ARCH, "-shared", "-o", lib, *objs, "-lcudart"])
    if force or _newer(host, deps + [lib]):
        _run([NVCC, *ARCH, *COMMON, "-shared", "-o", host, *[os.path.join(HERE, s) for s in HOST_SOURCES],
              "-L" + OUT, "-lggml_b200", "-Xlinker", "-rpath,$ORIGIN"])
    main_src = os.path.join(HERE, "host", "main.cpp")
    if os.path.exists(main_src) and (force or _newer(cli, deps + [host])):
        _run([NVCC, *ARCH, *COMMON, "-o", cli, main_src, "-L" + OUT, "-lmobilevit_b200", "-lggml_b200",
              "-Xlinker", "-rpath,$ORIGIN"])
    out = {"lib": lib, "host": host, "cli": cli}
    out.update(build_reference_programs(force))
    return out


REFERENCE = "/root/reference"


def build_reference_programs(force: bool = False) -> dict:
    """Drop-in proof: compile the UNMODIFIED reference programs, from where they lie under /root/reference, against
    include/ and link them with libggml_b200.so instead of upstream ggml's objects (mobilevit/Makefile:7-20).
    Only possible in the dev container (the GPU box has no /root/reference; the built binaries travel with the repo)."""
    res = {}
    progs = {"ref_main_b200": (os.path.join(REFERENCE, "mobilevit", "main.cpp"), ["-I" + REFERENCE]),
             "ref_rnn_b200": (os.path.join(REFERENCE, "rnn_text_gen", "rnn_text_generation.cpp"), [])}
    for name, (src, extra) in progs.items():
        if not os.path.exists(src):
            continue
        exe = os.path.join(OUT, name)
        if force or _newer(exe, [src, os.path.join(OUT, "libggml_b200.so"), os.path.join(INC, "ggml", "ggml.h")]):
            _run(["g++", "-O2", "-std=c++17", "-w", "-I" + INC, *extra, src, "-o", exe, "-L" + OUT, "-lggml_b200",
                  "-Wl,-rpath,$ORIGIN"])
        res[name] = exe
    return res


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="--ptxas" in sys.argv)
