"""Build recipe for libopnet_b200.so: plain nvcc, sm_100a only, in-tree output.

    python -m objectpermanence_b200.build          # (re)build if sources are newer

No torch headers are involved: the library is a C-ABI shared object (include/opnet_b200.h)
and is loaded with ctypes (objectpermanence_b200/_lib.py).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libopnet_b200.so")
SOURCES = ["opn_api.cu", "opn_lstm.cu", "opn_lstm_mma.cu", "opn_lstm_tc.cu", "opn_attention_tc.cu", "opn_wgrad_tc.cu", "opn_opnet_fused.cu", "opn_opnet_l1head.cu", "opn_opnet_fused_bwd.cu", "opn_opnet_l1bwd.cu", "opn_gemm.cu", "opn_gemm_skinny.cu", "opn_gemm_proj.cu", "opn_gemm_tc.cu", "opn_pointwise.cu", "opn_head_loss.cu", "opn_train_eval.cu"]
HEADERS = [os.path.join(CSRC, "opn_common.cuh"), os.path.join(CSRC, "opn_lstm_common.cuh"), os.path.join(CSRC, "opn_mma_common.cuh"), os.path.join(CSRC, "opn_tc_common.cuh"), os.path.join(os.path.dirname(PKG), "include", "opnet_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libopnet_b200 needs the CUDA toolkit to build")
    return nvcc


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), lib_path: str = LIB_PATH) -> str:
    """defines / lib_path: development variants (e.g. -DOPN_LSTM_PHASES into lib/libopnet_b200_phases.so)."""
    if not force and not defines and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    tag = "" if lib_path == LIB_PATH else "." + os.path.basename(lib_path).split(".")[0]
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", tag + ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, proc in procs:
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("linking libopnet_b200.so failed")
    return lib_path


if __name__ == "__main__":
    if "--exp" in sys.argv:  # experiment variants: --exp NAME DEFINE[=VALUE] ...
        i = sys.argv.index("--exp")
        print(build(force=True, defines=sys.argv[i + 2:], lib_path=os.path.join(LIB_DIR, f"libopnet_b200_{sys.argv[i + 1]}.so")))
    elif "--phases" in sys.argv:  # development variant with per-phase cycle counters in the recurrence kernels
        print(build(force=True, defines=["OPN_LSTM_PHASES"], lib_path=os.path.join(LIB_DIR, "libopnet_b200_phases.so")))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
