"""CPU tests of the C-ABI boundary and the host-side mirror of the reference surface (no GPU compute)."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from diffqcqp_b200 import _lib, build
    build.build()  # nvcc cross-compiles without a GPU
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "diffqcqp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from diffqcqp_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/diffqcqp_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms


def test_identification(lib):
    assert lib.dq_build_arch() == b"sm_100a"
    assert lib.dq_version() >= 100
    assert lib.dq_max_n() == 128
    assert lib.dq_error_string(0) == b"ok"
    for code in (1, 2, 3, 4):
        assert len(lib.dq_error_string(code)) > 3


def test_sass_is_sm100a_with_256bit_global_accesses():
    """The library holds sm_100a SASS only, and the kernels stream P / grad_P rows with the 256-bit global
    loads/stores that exist from sm_100 on (LDG.E...256 / STG.E...256) and vote-based control (VOTE)."""
    from diffqcqp_b200 import _lib
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    archs = set(l.split(".")[-2] for l in out.splitlines() if l.strip().endswith(".cubin"))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "admm_fwd_kernel" in sass and "qp_bwd_kernel" in sass and "qcqp_bwd_kernel" in sass
    assert ".256" in sass and "LDG" in sass and "STG" in sass and "VOTE" in sass and "DFMA" in sass


def test_argument_validation_without_gpu(lib):
    buf = np.zeros(64 * 64 + 64, dtype=np.float64)
    p = buf.ctypes.data
    DQ_OK, BAD, UNSUP, ALIGN = 0, 1, 2, 3
    # empty batch is a no-op success (no launch, so no device needed)
    assert lib.dq_qp_forward(p, p, None, p, None, 0, 8, 1e-7, 1e-7, 10, 1, None) == DQ_OK
    assert lib.dq_qcqp_forward(p, p, p, p, None, p, None, 0, 8, 1e-7, 1e-7, 10, 1, None) == DQ_OK
    assert lib.dq_qp_backward(p, p, p, p, p, p, 0, 8, None) == DQ_OK
    # bad arguments are rejected before anything touches CUDA
    assert lib.dq_qp_forward(None, p, None, p, None, 4, 8, 1e-7, 1e-7, 10, 1, None) == BAD
    assert lib.dq_qp_forward(p, p, None, p, None, -1, 8, 1e-7, 1e-7, 10, 1, None) == BAD
    assert lib.dq_qp_forward(p, p, None, p, None, 4, 0, 1e-7, 1e-7, 10, 1, None) == BAD
    assert lib.dq_qp_forward(p, p, None, p, None, 4, 129, 1e-7, 1e-7, 10, 1, None) == UNSUP
    assert lib.dq_qp_forward(p, p, p, p, None, 4, 33, 1e-7, 1e-7, 10, 3, None) == UNSUP  # warm-start extension: tile kernels only
    assert lib.dq_qp_forward(p + 4, p, None, p, None, 4, 8, 1e-7, 1e-7, 10, 1, None) == ALIGN
    assert lib.dq_qcqp_forward(p, p, p, p, None, p, None, 4, 7, 1e-7, 1e-7, 10, 1, None) == BAD  # odd N
    assert lib.dq_qcqp_forward(p, p, None, p, None, p, None, 4, 8, 1e-7, 1e-7, 10, 1, None) == BAD
    assert lib.dq_qp_backward(p, p, p, None, p, p, 4, 8, None) == BAD
    assert lib.dq_qcqp_backward(p, p, p, p, p, p, p, p, p, p, 4, 9, None) == BAD
    assert lib.dq_qp_solve_host(p, p, p, None, None, None, 4, 140, 1e-7, 1e-7, 10, -1) == UNSUP


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    """Without a device the product path must fail loudly, never compute on the CPU."""
    import qcqp
    from diffqcqp_b200 import DiffQCQPError
    P = torch.diag_embed(torch.rand(4, 8)); q = torch.rand(4, 8, 1)
    with pytest.raises(DiffQCQPError):
        qcqp.QPFn2.apply(P, q, torch.zeros_like(q), 1e-7, 100)
    buf = np.ones(4 * 64 + 64, dtype=np.float64)
    p = buf.ctypes.data
    assert lib.dq_qp_forward(p, p, None, p, None, 4, 8, 1e-7, 1e-7, 10, 1, None) == 4  # DQ_ERR_CUDA
    assert lib.dq_last_cuda_error() != 0


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "diffqcqp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("# no oracle", ""), f"{f} mentions the oracle"
    src = open(os.path.join(ROOT, "qcqp.py")).read()
    assert "oracle" not in src


def test_surface_mirrors_reference_signatures():
    import qcqp
    sig = inspect.signature(qcqp.QPFn2.forward)
    assert list(sig.parameters) == ["ctx", "P", "q", "warm_start", "eps", "max_iter", "mu_prox"]   # qcqp.py:24
    assert sig.parameters["mu_prox"].default == 1e-7
    sig = inspect.signature(qcqp.QCQPFn2.forward)
    assert list(sig.parameters) == ["ctx", "P", "q", "l_n", "mu", "warm_start", "eps", "max_iter", "mu_prox"]  # :144
    assert torch.get_default_dtype() == torch.float64                                               # qcqp.py:13


def test_shape_validation():
    from diffqcqp_b200.qcqp import _check_shapes
    P, q = torch.zeros(3, 8, 8), torch.zeros(3, 8, 1)
    assert _check_shapes(P, q) == (3, 8)
    with pytest.raises(ValueError):
        _check_shapes(torch.zeros(3, 8, 7), q)
    with pytest.raises(ValueError):
        _check_shapes(P, torch.zeros(3, 8))
    with pytest.raises(ValueError):
        _check_shapes(torch.zeros(3, 7, 7), torch.zeros(3, 7, 1), torch.zeros(3, 3, 1), torch.zeros(3, 3, 1))
    with pytest.raises(ValueError):
        _check_shapes(P, q, torch.zeros(3, 3, 1), torch.zeros(3, 4, 1))
    assert _check_shapes(P, q, torch.zeros(3, 4, 1), torch.zeros(3, 4, 1)) == (3, 8)


def test_algorithmic_bytes_match_survey():
    from diffqcqp_b200 import workloads as wl
    # SURVEY.md section 8d minus the dead warm_start read (8N) that the kernels never perform
    assert wl.qp_bytes(8) == 1984 - 64
    assert wl.qcqp_bytes(16) == 7424 - 128
    assert wl.qcqp_bytes(24) == 15744 - 192
    assert wl.qp_bytes(32) == 26368 - 256


def test_sass_has_no_per_lane_indexed_jumps():
    """No kernel contains an indexed jump (BRX / JMX).  nvcc 12.9 can fold an unrolled `if (j == lane_index) x[j] = ...`
    chain into a switch lowered to per-lane jump tables: every lane of the warp then runs its own path (measured: the
    N = 24 forward 60x slower with one such chain in the refactorisation).  common.cuh's sel() keeps those updates selects."""
    import shutil
    import subprocess
    from diffqcqp_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    fn, bad = None, {}
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split(":")[1].strip()
        elif " BRX " in line or " JMX " in line:
            bad[fn] = bad.get(fn, 0) + 1
    assert not bad, bad


def test_sass_of_the_persistent_forward_hot_loop():
    """Regression guard for the headline kernel (admm_fwd_diag8_kernel<PROX_NONNEG>): the code object holds it, it fits 72
    registers, its ADMM loops carry no divergence guards (BRA.DIV: the warp index is read through a shuffle so that
    the compiler knows the loop to be warp-uniform) and no local-memory traffic, and the straggler loop uses the uniform
    redux (CREDUX) for the tile maximum."""
    import re
    import shutil
    import subprocess
    from diffqcqp_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cuobjdump) or not os.path.exists(nvcc):
        pytest.skip("cuobjdump / nvcc not available")
    # the thresholds below describe what nvcc 12.9 generates; another compiler release may legitimately differ
    ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout
    if "release 12.9" not in ver:
        pytest.skip("SASS regression guard is pinned to nvcc 12.9")
    fun = "_ZN2dq21admm_fwd_diag8_kernelILi0EEEvNS_9FwdParamsE"
    res = subprocess.run([cuobjdump, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    m = re.search(fun + r".*?\n\s*REG:(\d+)", res, re.S)
    assert m and int(m.group(1)) <= 72, m and m.group(1)
    sass = subprocess.run([cuobjdump, "-sass", "-fun", fun, _lib.LIB_PATH], capture_output=True, text=True).stdout
    ins = [(int(mm.group(1), 16), mm.group(2)) for mm in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", sass)]
    assert len(ins) > 3000
    assert any("CREDUX" in t for _, t in ins)
    # the diagonal loops: backward branches whose body holds DSETP + VOTE and spans 400..900 instructions, ahead of the
    # inlined generic group routine's loops (the two at the lowest addresses: normal and straggler loop)
    loops = []
    for a, t in ins:
        mm = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if mm and int(mm.group(1), 16) < a and 400 <= (a - int(mm.group(1), 16)) // 16 <= 900:
            loops.append((int(mm.group(1), 16), a))
    loops.sort()
    assert len(loops) >= 2
    for lo, hi in loops[:2]:
        body = [t for a, t in ins if lo <= a <= hi]
        assert any("VOTE" in t for t in body) and any("DSETP" in t for t in body)
        assert not any("BRA.DIV" in t for t in body), "divergence guards inside the persistent forward's loop"
        assert not any(t.startswith(("LDL", "STL")) or " LDL" in t or " STL" in t for t in body), "spills inside the loop"


def test_sass_of_the_thread_per_problem_forward():
    """The round-2 headline kernel (admm_fwd_tpp8_kernel<PROX_NONNEG, 8>): present in the code object, within the register
    budget of three CTAs per SM, no spills, and its main loop carries the branch-free double-precision seeds
    (MUFU.RSQ64H / MUFU.RCP64H) of fast_sqrt / fast_rcp."""
    import re
    import shutil
    import subprocess
    from diffqcqp_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cuobjdump) or not os.path.exists(nvcc):
        pytest.skip("cuobjdump / nvcc not available")
    if "release 12.9" not in subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout:
        pytest.skip("SASS regression guard is pinned to nvcc 12.9")
    fun = "_ZN2dq20admm_fwd_tpp8_kernelILi0ELi8EEEvNS_9FwdParamsEi"
    res = subprocess.run([cuobjdump, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    m = re.search(fun + r".*?\n\s*REG:(\d+)\s+STACK:(\d+)", res, re.S)
    assert m, "kernel missing from the code object"
    assert int(m.group(1)) <= 168 and int(m.group(2)) == 0, (m.group(1), m.group(2))
    sass = subprocess.run([cuobjdump, "-sass", "-fun", fun, _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "MUFU.RSQ64H" in sass and "MUFU.RCP64H" in sass and "ATOMS" in sass
    assert not re.search(r"\b(LDL|STL)\b", sass), "local-memory traffic in the thread-per-problem forward"
