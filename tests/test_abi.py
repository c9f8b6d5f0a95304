"""The drop-in boundary without a GPU: the library loads, exports every symbol the headers
declare (C ABI) and the 17 C++-linkage entry points under exactly the mangled names the
reference's own HostCUDA.cu produces (HostCUDA.h:99-126, EwaldCUDA.h:59-62, CudaFunctions.h:7-8)."""
import os
import re
import subprocess
import sys

import pytest

from changa_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# nm -D of the reference's HostCUDA.cu compiled unmodified (oracle/_ref/libhostcuda_ref.so), host entry points
REFERENCE_MANGLED = [
    "_Z24allocatePinnedHostMemoryPPvm",
    "_Z20freePinnedHostMemoryPv",
    "_Z28DataManagerTransferLocalTreePvmS_mS_mPS_S0_S0_P11CUstream_stiS_",
    "_Z30DataManagerTransferRemoteChunkPvmS_mPS_S0_P11CUstream_stS_",
    "_Z24TransferParticleVarsBackP16VariablePartDatamPvP11CUstream_stS1_",
    "_Z34TreePieceCellListDataTransferLocalP12_CudaRequest",
    "_Z35TreePieceCellListDataTransferRemoteP12_CudaRequest",
    "_Z41TreePieceCellListDataTransferRemoteResumeP12_CudaRequest",
    "_Z34TreePiecePartListDataTransferLocalP12_CudaRequest",
    "_Z44TreePiecePartListDataTransferLocalSmallPhaseP12_CudaRequestP15CompactPartDatai",
    "_Z35TreePiecePartListDataTransferRemoteP12_CudaRequest",
    "_Z41TreePiecePartListDataTransferRemoteResumeP12_CudaRequest",
    "_Z26TreePieceDataTransferBasicP12_CudaRequestP11_CudaDevPtr",
    "_Z33TreePieceDataTransferBasicCleanupP11_CudaDevPtr",
    "_Z20EwaldHostMemorySetupP9EwaldDataiii",
    "_Z19EwaldHostMemoryFreeP9EwaldDatai",
    "_Z9EwaldHostP15CompactPartDataP16VariablePartDataP9EwaldDataP11CUstream_stPvii",
]


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {ln.split()[-1] for ln in out.splitlines() if ln.strip()}


@pytest.mark.parametrize("double", [False, True])
def test_library_loads_and_exports_the_c_abi(double):
    L = lib.load(double)
    missing = [s for s in lib.C_ABI_SYMBOLS if not hasattr(L, s)]
    assert not missing, missing
    assert L.cb200_real_bytes() == (8 if double else 4)
    assert L.cb200_abi_version() >= 1
    assert b"sm_100a" in L.cb200_build_info()


def test_header_declarations_are_all_exported():
    """every cb200_* function include/changa_b200_api.h declares is in C_ABI_SYMBOLS and in the .so"""
    hdr = open(os.path.join(ROOT, "include", "changa_b200_api.h")).read()
    declared = set(re.findall(r"\b(cb200_[A-Za-z0-9_]+)\s*\(", hdr)) - {"cb200_callback_fn"}
    syms = exported(lib.library_path(False))
    assert declared <= syms, declared - syms
    assert declared == set(lib.C_ABI_SYMBOLS), declared ^ set(lib.C_ABI_SYMBOLS)


def test_cxx_entry_points_have_the_reference_mangled_names():
    syms = exported(lib.library_path(False))
    missing = [m for m in REFERENCE_MANGLED if m not in syms]
    assert not missing, missing


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhostcuda_ref.so")),
                    reason="oracle/_ref not built")
def test_every_host_export_of_the_reference_build_is_matched():
    ref = exported(os.path.join(ROOT, "oracle", "_ref", "libhostcuda_ref.so"))
    ours = exported(lib.library_path(False))
    host = {s for s in ref if s.startswith("_Z") and "device_stub" not in s
            and not re.search(r"hapi|nodeGravityComputation|particleGravityComputation|EwaldKernel|ZeroVars|gpuLocalTreeWalk", s)}
    assert host == set(REFERENCE_MANGLED), host ^ set(REFERENCE_MANGLED)
    assert host <= ours


def test_partition_buckets_host_side():
    import numpy as np
    cost = np.random.default_rng(0).uniform(1, 10, 1000)
    for ranks in (1, 2, 4, 8):
        cuts = lib.partition_buckets(cost, ranks)
        assert cuts[0] == 0 and cuts[-1] == 1000 and np.all(np.diff(cuts) > 0)
        loads = np.add.reduceat(cost, cuts[:-1])
        assert loads.max() / loads.mean() < 1.02


LINK_TEST = os.path.join(ROOT, "oracle", "_ref", "abi_link_test")


@pytest.mark.skipif(not os.path.exists(LINK_TEST), reason="oracle/_ref/abi_link_test not built (needs /root/reference)")
def test_reference_headers_translation_unit_links_the_product_library():
    """SURVEY row f4 surrogate, CPU half: oracle/abi_link_test.cu includes the REFERENCE's HostCUDA.h /
    EwaldCUDA.h / cuda_typedef.h / CudaFunctions.h, static_asserts every field offset against
    include/changa_b200_types.h (so the binary existing means the layouts agree), and its undefined
    symbols -- the reference's mangled names -- resolve against libchanga_b200.so."""
    out = subprocess.run(["nm", "-D", "--undefined-only", LINK_TEST], capture_output=True, text=True, check=True).stdout
    wanted = {ln.split()[-1] for ln in out.splitlines() if ln.strip()} & set(REFERENCE_MANGLED)
    assert len(wanted) >= 7, wanted
    assert wanted <= exported(lib.library_path(False))
    ldd = subprocess.run(["ldd", LINK_TEST], capture_output=True, text=True).stdout
    assert "libchanga_b200.so" in ldd and "not found" not in ldd.split("libchanga_b200.so")[1].splitlines()[0]


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(LINK_TEST), reason="oracle/_ref/abi_link_test not built (needs /root/reference)")
def test_reference_headers_translation_unit_runs_requests_on_the_gpu():
    """GPU half: the same binary uploads a tree, issues a Local cell request, a Local particle request,
    EwaldHost and TransferParticleVarsBack through the reference-declared C++ functions, and checks the
    result against a double direct sum and the five completion callbacks."""
    r = subprocess.run([LINK_TEST], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "-> ok" in r.stdout, r.stdout


def test_every_environment_switch_is_documented():
    """INTEGRATION.md's table names every CB200_* variable the library reads"""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    read = set()
    for f in glob.glob(os.path.join(root, "changa_b200", "csrc", "*.cu*")) + glob.glob(os.path.join(root, "changa_b200", "csrc", "*.cpp")):
        read |= set(re.findall(r'getenv\("(CB200_[A-Z0-9_]+)"\)', open(f).read()))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    assert read and not [v for v in sorted(read) if v not in doc]


def test_the_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under changa_b200/ imports, links or loads it, importing the package
    does not pull it in, and bench.py reaches it only from its checker / CPU-baseline legs"""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "changa_b200")
    for f in glob.glob(os.path.join(pkg, "**", "*"), recursive=True):
        if os.path.isfile(f) and f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
            text = open(f, errors="replace").read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
            assert "liboracle" not in text and "oracle/_ref" not in text and "gravity_oracle" not in text, f
    code = ("import sys; import changa_b200, changa_b200.hostcuda, changa_b200.step, changa_b200.device_step; "
            "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; assert not bad, bad")
    subprocess.check_call([sys.executable, "-c", code], cwd=root)
    bench = open(os.path.join(root, "bench.py")).read()
    assert not re.search(r"^(from|import)\s+oracle\b", bench, re.M)  # only inside the functions of its CPU legs


def test_a_missing_library_or_device_fails_loudly():
    """no CPU fallback anywhere: without the built library the loader raises, and without a CUDA device the first
    entry point that needs one ends the process with a message (this container has no GPU; on a GPU box the second
    half is skipped)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CB200_LIB="/nonexistent/libchanga_b200.so")
    code = ("from changa_b200 import lib\n"
            "try:\n    lib.load()\nexcept lib.LibraryMissing as e:\n    print('missing:', e)\nelse:\n    raise SystemExit(3)\n")
    p = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True)
    assert p.returncode == 0 and "no CPU fallback" in p.stdout, (p.returncode, p.stdout, p.stderr)
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    code = "from changa_b200.hostcuda import HostCUDA\nHostCUDA(double=False, device=0)\nprint('survived')\n"
    p = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert p.returncode != 0 and "survived" not in p.stdout, (p.returncode, p.stdout)
    assert "CUDA" in p.stderr or "cuda" in p.stderr, p.stderr[-400:]
