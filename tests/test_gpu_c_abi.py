"""-m gpu: the drop-in boundary exercised from plain C (examples/c_abi_demo.c): no Python objects, no torch types —
cudaMalloc'd buffers, a caller-owned stream, pcv_attn_workspace_bytes / pcv_attn_fwd for the tcgen05 and the
CUDA-core kernels, checked inside the program against a double-precision restatement of modules.py:146-164."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def test_plain_c_caller(tmp_path):
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "no C compiler on the box"
    lib_dir = os.path.join(ROOT, "perceiver_io_b200", "lib")
    exe = str(tmp_path / "c_abi_demo")
    build = subprocess.run(
        [cc, "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
         os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L", lib_dir, "-lpcv_attn",
         "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{lib_dir}",
         f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"],
        capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "C_ABI_DEMO_OK" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]
    assert run.stdout.count(" ok") == 4, run.stdout
