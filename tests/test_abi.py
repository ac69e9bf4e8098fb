"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol include/swiftest_cuda.h
declares, refuses to run without a GPU (no fallback), and its GPU-free helpers agree with the host logic."""
import ctypes as C
import os
import subprocess

import pytest

from swiftest_b200 import _lib, shard


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "swiftest_cuda.h"\nint main(void){return SWCU_OK;}\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, "-c", str(src), "-o",
                           str(tmp_path / "t.o")])


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU creating a context must fail loudly (SWCU_ERR_NOGPU), never fall back."""
    L = _lib.load()
    h = C.c_void_p()
    rc = L.swcu_create(0, C.byref(h))
    if rc == 0:  # running on the GPU box
        L.swcu_destroy(h)
        pytest.skip("GPU present")
    assert rc == 5 and not h.value
    from swiftest_b200 import Context, SwcuError
    with pytest.raises(SwcuError):
        Context(0)


def test_product_package_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "swiftest_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "swiftest_oracle" not in text, f


@pytest.mark.parametrize("n,nranks", [(100000, 8), (12, 5), (3, 8), (0, 2), (1000001, 4)])
def test_partition_matches_c_abi(n, nranks):
    L = _lib.load()
    covered = 0
    for r in range(nranks):
        i0, i1 = C.c_int32(), C.c_int32()
        assert L.swcu_partition(n, nranks, r, C.byref(i0), C.byref(i1)) == 0
        assert (i0.value, i1.value) == shard.partition(n, nranks, r)
        assert i0.value == covered
        covered = i1.value
    assert covered == n
    sizes = [shard.partition(n, nranks, r)[1] - shard.partition(n, nranks, r)[0] for r in range(nranks)]
    assert max(sizes) - min(sizes) <= 1


def test_tp_block_partition_is_coarray_shape():
    # swiftest_coarray.f90:705-711: ceil(ntot/nimages) per image, last image takes the remainder
    assert [shard.tp_block_partition(10, 4, k) for k in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [shard.tp_block_partition(2, 4, k) for k in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]


def test_null_context_and_bad_arguments_are_rejected_without_a_gpu():
    """Argument validation happens before any CUDA call: status codes, never a crash, never a silent success."""
    L = _lib.load()
    assert L.swcu_last_error(None) == b"null context"
    assert L.swcu_destroy(None) == 2                       # SWCU_ERR_ARG
    assert L.swcu_encounter_fetch(None, 0, None, None, None) == 2
    assert L.swcu_encounter_stats(None, None, None) == 2
    assert L.swcu_body_count(None, 0, None, None, None) == 2
    assert L.swcu_create(0, None) == 2
    i0, i1 = C.c_int32(), C.c_int32()
    assert L.swcu_partition(10, 0, 0, C.byref(i0), C.byref(i1)) == 2
    assert L.swcu_partition(10, 4, 4, C.byref(i0), C.byref(i1)) == 2
    assert L.swcu_partition(-1, 4, 0, C.byref(i0), C.byref(i1)) == 2
    assert L.swcu_version() == 100
    assert L.swcu_launch_count(None) == 0


def test_python_harness_validates_shapes_before_calling_the_library():
    import numpy as np
    from swiftest_b200.context import Context, _vec, _vec3
    with pytest.raises(ValueError):
        _vec3(np.zeros((4, 2)))
    with pytest.raises(ValueError):
        _vec3(np.zeros((4, 3)), n=5)
    with pytest.raises(ValueError):
        _vec(np.zeros(3), n=4)
    with pytest.raises(ValueError):
        Context._inplace3(np.zeros((4, 3), dtype=np.float32), 4)
    with pytest.raises(ValueError):
        Context._inplace3(np.zeros((3, 4)).T, 4)  # not C-contiguous


def _c_prototypes():
    """name -> list of (is_pointer, text) for every function include/swiftest_cuda.h declares."""
    import re
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char \*)\s*(swcu_[a-z0-9_]+)\s*\((.*?)\)\s*;", text, flags=re.S):
        args = " ".join(m.group(2).split())
        params = [] if args == "void" else [a.strip() for a in args.split(",")]
        protos[m.group(1)] = [("*" in a, a) for a in params]
    return protos


def _fortran_interfaces():
    """bind(C) name -> (dummy argument list, set of dummies declared with the `value` attribute)."""
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fortran", "swiftest_cuda.f90")
    src = open(path).read()
    src = re.sub(r"&\s*\n\s*", " ", src)                      # join continuation lines
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)(.*?)end function", src,
                         flags=re.S | re.I):
        dummies = [a.strip() for a in m.group(2).split(",") if a.strip()]
        body = re.sub(r"!.*", "", m.group(4))
        by_value = set()
        for line in body.splitlines():
            if "::" in line and re.search(r"\bvalue\b", line.split("::")[0]):
                by_value.update(v.strip().split("(")[0] for v in line.split("::")[1].split(","))
        assert m.group(1) == m.group(3), (m.group(1), m.group(3))
        out[m.group(3)] = (dummies, by_value)
    return out


def test_fortran_interfaces_match_the_c_prototypes():
    """The Fortran binding module cannot be compiled here (no Fortran compiler), so guard it structurally: every
    bind(C) interface names a function the header declares, with the same number of arguments, and passes exactly the
    C scalars by value (type(c_ptr), value counts as a pointer-sized scalar for a C pointer parameter)."""
    protos = _c_prototypes()
    ifaces = _fortran_interfaces()
    assert sorted(ifaces) == sorted(protos), sorted(set(protos) ^ set(ifaces))   # the module binds the WHOLE header
    for name, (dummies, by_value) in ifaces.items():
        assert name in protos, f"{name}: not declared in swiftest_cuda.h"
        cparams = protos[name]
        assert len(dummies) == len(cparams), f"{name}: {len(dummies)} Fortran dummies vs {len(cparams)} C parameters"
        for dummy, (is_ptr, ctext) in zip(dummies, cparams):
            if not is_ptr:
                assert dummy in by_value, f"{name}: C scalar `{ctext}` must be passed by value (dummy {dummy})"
    # and the names the hot path itself needs are all bound
    for need in ("swcu_create", "swcu_kick_getacch_int_all_flat_pl", "swcu_kick_getacch_int_all_tri_pl",
                 "swcu_kick_getacch_int_all_tp", "swcu_drift_all", "swcu_encounter_check_all_sort_and_sweep_plpl",
                 "swcu_encounter_fetch", "swcu_body_sync", "swcu_helio_step_pl"):
        assert need in ifaces, need


def test_ctypes_signatures_match_the_c_prototypes():
    """Every function of the header has ctypes argtypes in swiftest_b200/_lib.py with the right arity, pointers where C
    has pointers, and the right scalar width (int32_t / int64_t / uint64_t / double / int)."""
    L = _lib.load()
    protos = _c_prototypes()
    scalar = {"int32_t": C.c_int32, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "double": C.c_double, "int": C.c_int}
    for name, params in protos.items():
        fn = getattr(L, name)
        at = fn.argtypes
        if name in ("swcu_version", "swcu_last_error", "swcu_launch_count") and at is None:
            continue
        assert at is not None, f"{name}: no argtypes"
        assert len(at) == len(params), f"{name}: {len(at)} ctypes args vs {len(params)} C parameters"
        for t, (is_ptr, text) in zip(at, params):
            if is_ptr:
                assert t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or issubclass(t, C._Pointer), (name, text)
            else:
                ctype = text.replace("const ", "").split()[0]
                assert C.sizeof(t) == C.sizeof(scalar[ctype]), f"{name}: `{text}` bound as {t}"


def test_cpp_host_mirror_cpu_checks():
    """The GPU-free parts of the compiled C++ host mirror: pair counts of pl%flatten, symba_pl%set_renc, and the
    no-fallback rule (cuda_context throws without an sm_100 GPU)."""
    exe = os.path.join(os.path.dirname(_lib.LIB_PATH), "host_cpu_check")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(_lib.LIB_PATH)), "csrc")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "HOST-CPU-CHECK-OK" in out.stdout, out.stdout + out.stderr


def test_partition_and_rebalance_properties_hypothesis():
    """Property tests of the host sharding logic: slices tile [0,n) in rank order with sizes differing by at most one;
    a rebalance plan moves every particle exactly once and keeps the global order."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 10 ** 6), st.integers(1, 16))
    def slices(n, nranks):
        cuts = [shard.partition(n, nranks, r) for r in range(nranks)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
        blocks = [shard.tp_block_partition(n, nranks, r) for r in range(nranks)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(0, 500), min_size=1, max_size=8))
    def plan(counts):
        moves = shard.rebalance_plan(counts)
        ntot, nimg = sum(counts), len(counts)
        taken = [[False] * c for c in counts]
        dest = {}
        for src, lo, hi, dst, dlo in moves:
            for t in range(hi - lo):
                assert not taken[src][lo + t]
                taken[src][lo + t] = True
                dest[(dst, dlo + t)] = (src, lo + t)
        assert all(all(row) for row in taken) and len(dest) == ntot
        order = [dest[k] for k in sorted(dest)]
        assert order == sorted(order)
        for k in range(nimg):
            a, b = shard.tp_block_partition(ntot, nimg, k)
            assert sum(1 for (d, _) in dest if d == k) == b - a
        if max(counts) - min(counts) < nimg:
            assert not shard.needs_rebalance(counts)

    slices()
    plan()


def test_use_cuda_patch_is_current_and_names_real_entry_points(tmp_path):
    """patches/swiftest_use_cuda.diff: every swcu_* function the Fortran blocks call is declared in the C header and has
    an interface in fortran/swiftest_cuda.f90; when the reference tree is present (this container only) the committed
    diff equals what patches/make_patch.py generates and applies to it."""
    import re
    import shutil
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    diff = open(os.path.join(ROOT, "patches", "swiftest_use_cuda.diff")).read()
    added = "\n".join(l[1:] for l in diff.splitlines() if l.startswith("+") and not l.startswith("+++"))
    called = set(re.findall(r"\b(swcu_[a-z0-9_]+)\s*\(", added)) - {"swcu_check", "swcu_ensure_ctx", "swcu_iplanet"}
    assert len(called) >= 10
    declared = set(_lib.declared_symbols())
    fortran = open(os.path.join(ROOT, "fortran", "swiftest_cuda.f90")).read()
    for name in called:
        assert name in declared, name
        assert f'name="{name}"' in fortran, name
    assert "subroutine swcu_ensure_ctx" in fortran and "subroutine swcu_check" in fortran
    assert added.count("#ifdef USE_CUDA") == added.count("#endif") >= 14
    ref = "/root/reference"
    if os.path.isdir(os.path.join(ref, "src")) and shutil.which("patch"):
        gen = subprocess.run([sys.executable, os.path.join(ROOT, "patches", "make_patch.py"), ref], capture_output=True, text=True,
                             cwd=str(tmp_path), env=dict(os.environ, SWCU_PATCH_OUT=str(tmp_path / "out.diff")))
        assert gen.returncode == 0, gen.stderr
        assert open(tmp_path / "out.diff").read() == diff, "patches/swiftest_use_cuda.diff is stale: rerun patches/make_patch.py"
        work = tmp_path / "tree"
        shutil.copytree(os.path.join(ref, "src"), work / "src")
        shutil.copy(os.path.join(ref, "CMakeLists.txt"), work / "CMakeLists.txt")
        ap = subprocess.run(["patch", "-p1", "--dry-run", "-i", os.path.join(ROOT, "patches", "swiftest_use_cuda.diff")],
                            capture_output=True, text=True, cwd=str(work))
        assert ap.returncode == 0, ap.stdout + ap.stderr
