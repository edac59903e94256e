"""The C-ABI library loads without a GPU and exports exactly what include/oak_b200.h declares;
compute calls fail loudly without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "oak_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(oak_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    from oak_b200 import _cabi

    lib = _cabi.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/oak_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == names, "ctypes signatures out of sync with the header"
    assert lib.oak_version() >= 100


def test_struct_layouts_match_header():
    from oak_b200 import _cabi

    assert ctypes.sizeof(_cabi.DimDesc) == 6 * 4 + 4 * 8 + 3 * 8
    assert ctypes.sizeof(_cabi.KernelDesc) == 4 * 4 + 2 * 8


def test_compute_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oak_b200 import _cabi
    from oak_b200.oak_kernel import OAKKernel
    from oak_b200.ortho_rbf_kernel import RBF

    assert _cabi.load().oak_device_count() == 0
    k = OAKKernel([RBF, RBF], num_dims=2, max_interaction_depth=2, constrain_orthogonal=True)
    with pytest.raises(_cabi.OakNativeError):
        k.K(np.zeros((4, 2)))
    with pytest.raises(_cabi.OakNativeError):
        _cabi.Spec([_cabi.DimSpec(_cabi.DIM_RBF, 0, measure=_cabi.MEASURE_GAUSSIAN, m1=1.0)], 1, [0.0, 1.0])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "orthogonal-additive-gaussian-processes_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_header_is_plain_c():
    """include/oak_b200.h is the boundary a maintainer binds from C / cgo / ctypes: it must parse as C99 on its own
    (no C++-isms, no torch or CUDA types in the signatures)."""
    import os
    import shutil
    import subprocess
    import tempfile

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "hdr.c")
        with open(src, "w") as f:
            f.write('#include "oak_b200.h"\nint main(void) { return (int)OAK_LINK_PROBIT - 1; }\n')
        out = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                              "-fsyntax-only", src], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "oak_b200.h")).read(), flags=re.S)
    assert "torch" not in text and "cudaStream_t" not in text and "at::" not in text  # outside the comments


def _build_c_program(tmpdir):
    """gcc -std=c99 tests/c/abi_link.c against include/oak_b200.h and the in-tree liboak_b200.so."""
    import os
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if gcc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("needs gcc and the CUDA runtime headers")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "orthogonal-additive-gaussian-processes_b200")
    exe = os.path.join(str(tmpdir), "abi_link")
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(root, "include"),
           "-I", os.path.join(cuda, "include"), os.path.join(root, "tests", "c", "abi_link.c"), "-o", exe,
           "-L", pkg, "-loak_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_c_program_links_the_abi(tmp_path):
    """A C99 translation unit compiles against the header, links the shared library and runs its housekeeping entry
    points (no device needed)."""
    import subprocess

    out = subprocess.run([_build_c_program(tmp_path), "--no-gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "oak_version" in out.stdout


@pytest.mark.gpu
def test_c_program_computes_a_gram_through_the_abi(tmp_path):
    """The same program end to end on the device: spec, prologue, Gram and K_diag against closed forms in C."""
    import subprocess

    out = subprocess.run([_build_c_program(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
