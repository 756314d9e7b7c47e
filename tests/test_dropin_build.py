"""The drop-in directory builds with the reference's OWN build system: copy the reference TeaLeaf tree
to a scratch directory, drop c_kernels/cuda into it and run `make KERNELS=cuda` exactly as
INTEGRATION.md says (plain and -DDIFFUSE_OVERLOAD).  CPU-only check (compile + link + symbols); the
binaries are exercised on the GPU by tests/test_dropin_gpu.py.  Skipped where /root/reference is absent."""
import os
import shutil
import subprocess

import pytest

REF = "/root/reference/TeaLeaf"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


@pytest.mark.parametrize("options", ["", "-DDIFFUSE_OVERLOAD"])
def test_make_kernels_cuda(tmp_path, options):
    tree = tmp_path / "TeaLeaf"
    shutil.copytree(REF, tree)
    shutil.copytree(os.path.join(ROOT, "c_kernels", "cuda"), tree / "c_kernels" / "cuda")
    cmd = ["make", "KERNELS=cuda", "COMPILER=GNU", "TL_B200=" + ROOT]
    if options:
        cmd.append("OPTIONS=" + options)
    r = subprocess.run(cmd, cwd=tree, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    exe = tree / "tealeaf"
    assert exe.exists()
    syms = subprocess.run(["nm", "-C", str(exe)], capture_output=True, text=True).stdout
    for name in ("run_cg_calc_w(Chunk*, Settings*, double*)", "run_pack_or_unpack(", "run_field_summary(",
                 "sum_over_ranks(Settings*, double*)", "send_recv_message("):
        assert name in syms, name
    assert ("diffuse_overload(" in syms) == bool(options)
    assert "tl_run_cg_calc_w" in syms  # resolved from libtealeaf_b200.so
    # the reference's MPI comms.c was compiled (against the shim) but is not linked
    assert "MPI_Allreduce" not in syms
