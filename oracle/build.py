"""Build recipe for the C oracle (test infrastructure).  gcc only; output oracle/libreacher_oracle.so (git-ignored,
travels to the GPU box with the snapshot).  The reference itself is pure Python over TensorFlow-1.10 / gym / MuJoCo-1.50,
none of which is in /root/reference or installable here, so there is no oracle/_ref build -- see DESIGN.md."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "reacher_oracle.c")
OUT = os.path.join(HERE, "libreacher_oracle.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
