"""Compiles the UNMODIFIED reference host (where it lies under /root/reference) against this backend
directory, exactly as `make KERNELS=cuda` would, into c_kernels/cuda/bin/ (git-ignored; travels to the GPU box):
    c_kernels/cuda/bin/tealeaf_cuda            host-driven plugin path (drivers/*.c call run_* per kernel)
    c_kernels/cuda/bin/tealeaf_cuda_resident   -DDIFFUSE_OVERLOAD: device-resident solver loop
Does nothing when /root/reference is absent (the GPU box uses the prebuilt files)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TL_REFERENCE", "/root/reference/TeaLeaf")
OUT = os.path.join(HERE, "bin")
CXX = "/usr/bin/g++"


def build():
    if not os.path.isdir(REF):
        return
    libdir = os.path.join(ROOT, "exploringsycl_b200")
    for name, opts in (("tealeaf_cuda", []), ("tealeaf_cuda_resident", ["-DDIFFUSE_OVERLOAD"])):
        objdir = os.path.join(OUT, "obj_" + name)
        os.makedirs(os.path.join(objdir, "drivers"), exist_ok=True)
        flags = ["-x", "c++", "-std=c++11", "-O2", "-w", "-I" + HERE, "-I" + REF, "-I" + os.path.join(ROOT, "include")] + opts
        srcs = [s for s in glob.glob(os.path.join(REF, "*.c")) + glob.glob(os.path.join(REF, "drivers", "*.c"))
                if os.path.basename(s) != "comms.c"]  # replaced by comms_b200.cpp
        objs = []
        procs = []
        for s in srcs:
            o = os.path.join(objdir, os.path.relpath(s, REF)[:-2] + ".o")
            procs.append(subprocess.Popen([CXX] + flags + ["-c", s, "-o", o]))
            objs.append(o)
        glue = ["kernel_interface.cpp", "comms_b200.cpp"] + (["diffuse_overload.cpp"] if opts else [])
        for g in glue:
            o = os.path.join(objdir, g[:-4] + ".o")
            procs.append(subprocess.Popen([CXX] + flags + ["-c", os.path.join(HERE, g), "-o", o]))
            objs.append(o)
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("drop-in compile failed")
        exe = os.path.join(OUT, name)
        subprocess.check_call([CXX, "-o", exe] + objs + ["-L" + libdir, "-ltealeaf_b200",
                                                          "-Wl,-rpath,$ORIGIN/../../../exploringsycl_b200", "-lm", "-lrt"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    build()
    print("built", OUT)
