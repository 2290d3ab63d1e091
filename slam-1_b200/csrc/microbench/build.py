"""Build the standalone micro-benchmarks (roofline calibration; not part of libslammatch.so).

    python slam-1_b200/csrc/microbench/build.py      -> slam-1_b200/csrc/microbench/bin/{pipe_rates,tc_probe}
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "bin")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"]


def main():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(BIN, exist_ok=True)
    for name in sorted(f[:-3] for f in os.listdir(HERE) if f.endswith(".cu")):
        src, out = os.path.join(HERE, name + ".cu"), os.path.join(BIN, name)
        deps = [src, os.path.join(HERE, "..", "tc_common.cuh")]
        if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
            continue
        subprocess.check_call([nvcc, *FLAGS, "-o", out, src])


if __name__ == "__main__":
    main()
