"""TEST INFRASTRUCTURE: builds tests/twin/libphysics_twin.so = csrc/physics.cuh compiled for the host (bit-exact CPU twin of the
device arithmetic).  nvcc drives g++ for the host pass; -ffp-contract=off keeps every rounding where the source puts it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "physics_twin.cu")
HDR = os.path.join(HERE, "..", "..", "reacherdistilation_b200", "csrc", "physics.cuh")
HDR2 = os.path.join(HERE, "..", "..", "reacherdistilation_b200", "csrc", "philox.cuh")
OUT = os.path.join(HERE, "libphysics_twin.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build(force=False):
    newest = max(os.path.getmtime(f) for f in (SRC, HDR, HDR2))
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        cmd = [NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-fopenmp,-fno-fast-math", "-o", OUT, SRC, "-lgomp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("twin build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return OUT


if __name__ == "__main__":
    print(build(force=True))
