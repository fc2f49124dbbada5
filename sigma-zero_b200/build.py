"""Builds libszb200.so (hand-written CUDA for sm_100a) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libszb200.so")
SOURCES = ["engine.cu", "net.cu", "train.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unknown-pragmas",
    "--expt-relaxed-constexpr", "--expt-extended-lambda", "-Wno-deprecated-gpu-targets",
]
# engine.cu holds the bit-exact search arithmetic: no silent FMA contraction there (it also uses explicit
# round-to-nearest intrinsics); the network kernels want FFMA.
EXTRA_FLAGS = {"engine.cu": ["--fmad=false"]}


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, *EXTRA_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    tmp = OUT + ".tmp.%d" % os.getpid()
    subprocess.check_call(["nvcc", "-Wno-deprecated-gpu-targets", "-shared", "-o", tmp, *objs, "-lcudart"])
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
