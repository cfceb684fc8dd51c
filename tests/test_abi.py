"""CPU-side checks of the drop-in boundary: libtevo.so loads and exports every symbol include/tevo.h declares,
the ctypes prototypes cover them all, and the product path refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "tevo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tevo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(tevo):
    import __graft_entry__ as g
    if not os.path.exists(tevo.LIB_PATH):
        g.build()
    lib = ctypes.CDLL(tevo.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s


def test_ctypes_prototypes_cover_header(tevo):
    from timeevomps_jl_b200._lib import PROTOTYPES
    assert sorted(PROTOTYPES) == _header_symbols()


def test_no_torch_or_cpp_types_in_abi():
    src = open(os.path.join(ROOT, "include", "tevo.h")).read()
    assert "torch" not in src and "std::" not in src and "at::" not in src


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "timeevomps.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                assert "oracle" not in open(os.path.join(dp, f)).read().replace("# oracle-free", ""), f


def test_fails_loudly_without_gpu(tevo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tevo.TevoError) as ei:
        tevo.MPS.product([0, 0, 0, 0])
    assert "no CPU fallback" in str(ei.value)


def test_host_side_gate_construction_matches_oracle(tevo, oracle):
    """BondOperator/gates/exp/time_evo_gates are host code on both sides; the product uses its own expm."""
    import numpy as np
    N = 9
    for (hp, ho) in ((tevo.tfi_bondop(N, 1.0, 0.5), oracle.tfi_bondop(N, 1.0, 0.5)),
                     (tevo.xxz_bondop(N, 0.7, 0.3), oracle.xxz_bondop(N, 0.7, 0.3))):
        gp, go = tevo.gates(hp), ho.gates()
        for a, b in zip(gp, go):
            assert a.b == b.b and np.array_equal(a.G, b.G)
        for alg_p, alg_o in ((tevo.TEBD2(), oracle.TEBD2()), (tevo.TEBD4(), oracle.TEBD4()),
                             (tevo.TEBD2Sweep(), oracle.TEBD2Sweep())):
            for dt in (0.05, -0.1j):
                P = tevo.time_evo_gates(dt, gp, alg_p)
                O = oracle.time_evo_gates(dt, go, alg_o)
                for Lp, Lo in zip(P, O):
                    assert len(Lp) == len(Lo)
                    for layer_p, layer_o in zip(Lp, Lo):
                        assert [g.b for g in layer_p] == [g.b for g in layer_o]
                        for a, b in zip(layer_p, layer_o):
                            assert np.abs(a.G - b.G).max() < 1e-14


def test_callback_gating_host_logic(tevo):
    """test/test_callbacks.jl:21-24 gating without touching the device."""
    cb = tevo.LocalMeasurementCallback(["Sz", "Sx"], 10, 0.1)
    assert cb.callback_dt() == 0.1 and len(tevo.measurement_ts(cb)) == 0
    assert not cb.wants(0.03, True, 0, tevo.TEBD2())
    assert not cb.wants(0.1, False, 0, tevo.TEBD2())
    assert cb.wants(0.1, True, 2, tevo.TEBD2())
    assert not cb.wants(0.1, True, 1, tevo.TEBD2())      # even (1-based) bond filtered for TEBD (callback.jl:84)
    assert cb.wants(0.1, True, 1, tevo.TDVP2())
    nb = tevo.NoTEvoCallback()
    assert nb.apply(None, t=0, bond=1, sweepend=True) is None and nb.checkdone(None) is False
    assert nb.callback_dt() == 0
    with pytest.raises(ValueError):
        tevo.SpecCallback(0.1, 10, bonds=[0, 9])


def test_julia_shim_binds_only_declared_symbols():
    """Every `ccall((:tevo_..., LIB) ...)` in the (unexecuted) Julia shim names a function that include/tevo.h
    declares and the built library exports; the reference-facing entry points are all bound."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    jl = open(os.path.join(root, "timeevomps.jl_b200", "julia", "TevoB200.jl")).read()
    hdr = open(os.path.join(root, "include", "tevo.h")).read()
    declared = set(re.findall(r"\b(tevo_[a-z0-9_]+)\s*\(", hdr))
    bound = set(re.findall(r"\(:(tevo_[a-z0-9_]+),\s*LIB\)", jl))
    assert bound, "no ccall found"
    assert bound <= declared, sorted(bound - declared)
    for must in ("tevo_ctx_create", "tevo_mps_create", "tevo_mps_upload_site", "tevo_mps_download_site",
                 "tevo_apply_gate", "tevo_apply_layer", "tevo_measure_bond", "tevo_mps_orthogonalize",
                 "tevo_mps_reorthogonalize", "tevo_mps_inner", "tevo_mps_inner_mpo", "tevo_mpo_create",
                 "tevo_env_create", "tevo_env_position", "tevo_tdvp_twosite", "tevo_tdvp_onesite",
                 "tevo_mps_normalize_site", "tevo_mps_free", "tevo_env_free", "tevo_mpo_free", "tevo_ctx_destroy"):
        assert must in bound, must


def test_header_is_plain_c_and_links(tmp_path):
    """include/tevo.h compiles as C99 and as C++17 (no torch / CUDA types in the signatures), and a C program that
    references entry points links against the built library (no compute call: there is no GPU here)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "tevo.h"\n'
                   'int main(void) {\n'
                   '  tevo_trunc t; tevo_spec s; (void)t; (void)s;\n'
                   '  /* take addresses so that the linker must resolve the symbols */\n'
                   '  void* p[] = {(void*)tevo_ctx_create, (void*)tevo_apply_gate, (void*)tevo_apply_layer,\n'
                   '               (void*)tevo_tdvp_twosite, (void*)tevo_tdvp_onesite, (void*)tevo_mps_upload_site};\n'
                   '  printf("%d %s\\n", (int)(sizeof(p) / sizeof(p[0])), tevo_version());\n'
                   '  return 0;\n}\n')
    inc = os.path.join(root, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", str(src)], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)],
                   check=True)
    libdir = os.path.join(root, "timeevomps.jl_b200")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=gnu99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-ltevo",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "6" and "libtevo" in out[1:]


def test_julia_shim_only_binds_declared_symbols():
    """The Julia shim cannot be executed here (no Julia in the image); at least every symbol it `ccall`s must be
    declared in include/tevo.h and exported by the built library, and the two drop-in drivers must exist."""
    import re
    shim = open(os.path.join(ROOT, "timeevomps.jl_b200", "julia", "TevoB200.jl")).read()
    header = open(os.path.join(ROOT, "include", "tevo.h")).read()
    syms = set(re.findall(r"ccall\(\(:(tevo_\w+), LIB\)", shim))
    assert len(syms) >= 25
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "timeevomps.jl_b200", "libtevo.so"))
    for s in sorted(syms):
        assert re.search(r"\b%s\s*\(" % s, header), "%s is not declared in include/tevo.h" % s
        assert hasattr(lib, s), "%s is not exported by libtevo.so" % s
    assert re.search(r"^function tebd!\(psi::MPS, H::GateList, dt::Number, tf::Number, alg::TEBDalg=TEBD2\(\)", shim, re.M)
    assert re.search(r"^function tdvp!\(psi::MPS, H::MPO, dt::Number, tf::Number", shim, re.M)
