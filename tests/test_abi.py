"""CPU checks of the drop-in boundary: every symbol include/impec2d_b200.h declares is exported by the built
library, the ctypes structs match the C structs, and the product loader refuses a CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import tempfile
import pytest
import impec2d_b200 as I
from impec2d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "impec2d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(impec_[a-z_0-9]+)\s*\(", src)))


def build_emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "impec2d_b200", "csrc"), "emu"])


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_cuda_library_exports_every_symbol():
    path = I.library_path(False)
    if not os.path.exists(path):
        pytest.skip("CUDA library not built yet (run __graft_entry__.build())")
    out = subprocess.check_output(["nm", "-D", "--defined-only", path]).decode()
    exported = set(re.findall(r" T (impec_[a-z_0-9]+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    # the product library must contain sm_100a code
    elf = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-lelf", path]).decode()
    assert "sm_100a" in elf


def test_struct_sizes_match_c():
    code = r'''
    #include <stdio.h>
    #include "include/impec2d_b200.h"
    #include "oracle/impec_oracle.h"
    int main(){ printf("%zu %zu %zu %zu %zu\n", sizeof(impec_params), sizeof(impec_diag), sizeof(impec_solve_info), sizeof(orc_params), sizeof(orc_state)); return 0; }
    '''
    import oracle as O
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(code)
        subprocess.check_call(["/usr/bin/gcc", "-I", ROOT, "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(I.Params), C.sizeof(I.Diag), C.sizeof(I.SolveInfo), C.sizeof(O.Params), C.sizeof(O.State)]


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never silently compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    path = I.library_path(False)
    if not os.path.exists(path):
        with pytest.raises(RuntimeError):
            I.load(False)
        return
    P = I.default_params(N1=16, N2=16)
    with pytest.raises(I.ImpecError):
        I.Context(P)


def test_emulated_library_is_flagged_and_never_substituted():
    build_emu()
    assert I.load(emulated=True).impec_is_emulated() == 1
    # the loader selects by file name and then verifies the flag: pointing it at the wrong file must raise
    saved = _lib.library_path
    try:
        _lib._CACHE.pop(False, None)
        _lib.library_path = lambda emulated=False: saved(True)      # pretend the product .so IS the emulator
        with pytest.raises(RuntimeError, match="emulated"):
            _lib.load(False)
    finally:
        _lib.library_path = saved
        _lib._CACHE.pop(False, None)


def test_argument_errors_are_reported_not_fatal():
    build_emu()
    P = I.default_params(emulated=True, N1=16, N2=16, g=3)
    with pytest.raises(I.ImpecError, match="num_ghostcells"):
        I.Context(P, emulated=True)
    P = I.default_params(emulated=True, N1=16, N2=16)
    ctx = I.Context(P, emulated=True)
    with pytest.raises(I.ImpecError, match="upload_q"):
        ctx.pressure_solve()
    with pytest.raises(I.ImpecError, match="unknown option"):
        ctx.set_option("nonsense", 1)
    # slab contexts: extents that cannot be cut / coarsened are refused at creation, before any communication
    with pytest.raises(I.ImpecError, match="cannot be cut into 2 slabs"):
        I.Context(I.default_params(emulated=True, N1=42, N2=32), rank=0, nranks=2, emulated=True)
    with pytest.raises(I.ImpecError, match="no coarser level"):
        I.Context(I.default_params(emulated=True, N1=48, N2=37), rank=1, nranks=2, emulated=True)


def test_cycle_shape_follows_the_rank_count():
    """The W range in effect (impec_mg_w_range): the default 2..4 everywhere except on 8 or more ranks with slabs of 10-50 M cells
    (16384^2 on 8 GPUs), where the replicated multigrid tail is entered half as often (2..3; measured in profiles/r2ad_*).
    An explicit mg_w_to always wins.  (Creating a context allocates but does not touch the slab: no compute here.)"""
    build_emu()
    for n, nranks, want in ((9088, 8, (2, 3)), (9088, 4, (2, 4)), (1024, 8, (2, 4)), (64, 1, (2, 4))):
        P = I.default_params(emulated=True, N1=n, N2=n)
        ctx = I.Context(P, rank=0, nranks=nranks, emulated=True)
        assert ctx.mg_w_range() == want, (n, nranks, ctx.mg_w_range())
        ctx.set_option("mg_w_to", 4)
        assert ctx.mg_w_range() == (2, 4)
        ctx.close()
    P = I.default_params(emulated=True, N1=9088, N2=9088, mg_w_to=5)         # not the default range: left alone
    ctx = I.Context(P, rank=0, nranks=8, emulated=True)
    assert ctx.mg_w_range() == (2, 5)
    ctx.close()


def test_bind_host_to_device_logic():
    """impec_bind_host_to_device (host placement of one-process-per-GPU runs): the device's local CPU list is intersected with the mask the
    process has; nothing changes when the list is unreadable, malformed or disjoint from the mask.  The emulated build has no PCI device
    and takes the list from IMPEC_EMU_LOCAL_CPULIST; run in a child process because the call changes the caller's affinity."""
    build_emu()
    have = sorted(os.sched_getaffinity(0))
    if len(have) < 2:
        pytest.skip("needs two allowed CPUs")
    code = ("import os, sys, json; sys.path.insert(0, %r); import impec2d_b200 as I; "
            "r = I.bind_host_to_device(0, emulated=True); print(json.dumps([r[0], r[1], sorted(os.sched_getaffinity(0))]))" % ROOT)

    def child(cpulist):
        env = dict(os.environ)
        env.pop("IMPEC_EMU_LOCAL_CPULIST", None)
        if cpulist is not None:
            env["IMPEC_EMU_LOCAL_CPULIST"] = cpulist
        import json
        import sys
        return json.loads(subprocess.check_output([sys.executable, "-c", code], env=env).decode().strip().splitlines()[-1])

    a, b = have[0], have[-1]
    assert child("%d,%d" % (a, b)) == [2, "%d,%d" % (a, b), [a, b]]
    assert child("%d-%d,100000-100003" % (a, a)) == [1, "%d-%d,100000-100003" % (a, a), [a]]          # CPUs beyond the mask are ignored
    for bad in (None, "", "x", "3-1", "100000-100003"):                                                  # no list / malformed / disjoint: unchanged
        assert child(bad) == [0, None, have], bad


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: no source file of the product (package, kernels, host layer, header) imports, includes, links or
    dlopens anything under it, and the built libraries neither depend on it nor carry its symbols."""
    pkg = os.path.join(ROOT, "impec2d_b200")
    offenders = []
    for base in (pkg, os.path.join(ROOT, "include")):
        for dp, _, fs in os.walk(base):
            for f in fs:
                if not f.endswith((".py", ".cu", ".inl", ".h", ".cpp", ".hpp", "Makefile")):
                    continue
                text = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"^\s*(import|from)\s+oracle\b|#\s*include\s*[\"<][^\"<>]*oracle|liboracle|impec_oracle|\borc_[a-z_]+\s*\(", text, flags=re.M):
                    offenders.append(os.path.relpath(os.path.join(dp, f), ROOT))
    assert not offenders, offenders
    for name in ("libimpec2d_b200.so", "libimpec2d_host.so"):
        path = os.path.join(pkg, name)
        if not os.path.exists(path):
            continue
        needed = subprocess.check_output(["readelf", "-d", path]).decode()
        assert "oracle" not in needed and "impec_ref" not in needed, needed
        syms = subprocess.check_output(["nm", "-D", path]).decode()
        assert not re.search(r"\borc_[a-z_]+", syms)
