"""Build A/B variants of libhw1f.so into <pkg>/lib/variants/ (tools/ab_variants.sh measures them on the GPU box).
usage: python tools/build_variants.py name=-DHW1F_X=1,-DHW1F_Y=2 [name2=...]   (up to 4 nvcc runs in parallel)"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib

b = importlib.import_module("monte-carlo-simulation-of-hull-white-model-and-sensitivities-computation_b200.build")
VAR = os.path.join(b.LIB_DIR, "variants")


def one(spec):
    name, _, flags = spec.partition("=")
    out = os.path.join(VAR, name + ".so")
    cmd = [b.nvcc_path()] + b.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-Xptxas", "-v", "-o", out] + b.SOURCES
    res = subprocess.run(cmd, cwd=b.CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        return name, "FAILED\n" + res.stderr[-3000:]
    # registers / spills of the Q1 kernel
    lines = res.stderr.splitlines()
    info = ""
    for i, l in enumerate(lines):
        if "fast_kernelILi1ELi0ELi0ELi0ELi0ELi0" in l and "Compiling" in l:
            info = " ".join(x.strip() for x in lines[i + 1:i + 4])
            break
    return name, info


if __name__ == "__main__":
    os.makedirs(VAR, exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        for name, info in ex.map(one, sys.argv[1:]):
            print(name, "::", info)
