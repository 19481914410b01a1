"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/gvigp.h declares; argument
validation that needs no device work returns the documented error codes (no compute calls here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gvigp.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gvi_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for must in ("gvi_ard_gram_f64", "gvi_gp_factor_f64", "gvi_gp_predict_f64", "gvi_risk_f64",
                 "gvi_fused_step_f64", "gvi_cond_var_select_f64", "gvi_select_update_f64"):
        assert must in syms


def test_library_exports_every_header_symbol(lib):
    from gvi_gaussian_process_b200 import _lib

    syms = header_symbols()
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gvi_[a-z0-9_]+)", out))
    assert set(syms) <= exported, sorted(set(syms) - exported)
    for s in syms:
        assert getattr(lib, s) is not None


def test_abi_version_and_size_queries(lib):
    assert lib.gvi_abi_version() == 1
    assert lib.gvi_gp_factor_bytes(1024, 8) > 8 * 512 * 32 * 33
    assert lib.gvi_gp_factor_bytes(0, 8) == 0
    assert lib.gvi_gp_factor_workspace_bytes(1000) >= 2 * 1024 * 1024 * 8
    assert lib.gvi_select_record_len(8, 1024) == 2 + 8 + 1023
    assert lib.gvi_select_workspace_bytes(1000, 10, 3) >= 1000 * 9 * 8
    assert lib.gvi_select_workspace_bytes(1000, 1, 3) == 0  # fewer than 2 inducing points is invalid


def test_argument_validation_without_device(lib):
    # all rejected before any CUDA call
    assert lib.gvi_ard_gram_f64(None, -1, None, 4, 3, None, None, 3, None, 4, None, None) == -1
    assert b"invalid argument" in lib.gvi_last_error_string()
    assert lib.gvi_ard_gram_f64(None, 2, None, 4, 3, None, None, 2, None, 4, None, None) == -1  # ls_len not in {1, D}
    assert lib.gvi_cond_var_select_f64(None, 10, 3, 1, None, None, 3, 1e-12, None, None, None, None) == -1
    assert b"at least 2 inducing points" in lib.gvi_last_error_string()
    assert lib.gvi_risk_f64(99, 0.5, None, None, None, None, None, 10, None, None, None) == -1
    # empty inputs are a no-op, as in the reference (a 0 x m Gram)
    assert lib.gvi_ard_gram_f64(None, 0, None, 4, 3, None, None, 3, None, 4, None, None) == 0
    assert lib.gvi_gp_predict_f64(None, 4, 3, None, 0, None, None, None, None, None, None) == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gvi_gaussian_process_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def test_argument_validation_of_the_widened_entry_points(lib):
    """SURVEY.md section 8(f) entry points: bad arguments are rejected before any CUDA call; empty inputs are a no-op."""
    # class probabilities: 2 <= K <= 64, H >= 1, N >= 0
    assert lib.gvi_class_probabilities_f64(None, None, 1, 10, None, None, 50, 0.01, 1e-10, None, None) == -1
    assert b"2 <= K <= 64" in lib.gvi_last_error_string()
    assert lib.gvi_class_probabilities_f64(None, None, 65, 10, None, None, 50, 0.01, 1e-10, None, None) == -1
    assert lib.gvi_class_probabilities_f64(None, None, 4, 10, None, None, 0, 0.01, 1e-10, None, None) == -1
    assert lib.gvi_class_probabilities_f64(None, None, 4, 0, None, None, 50, 0.01, 1e-10, None, None) == 0
    assert lib.gvi_class_probabilities_grad_f64(None, None, 4, 0, None, None, 50, 0.01, 1e-10, None, None, None,
                                                None) == 0
    assert lib.gvi_class_probabilities_f64(None, None, 4, 10, None, None, 50, 0.01, 1e-10, None, None) == -1  # nulls
    # risks on probabilities
    assert lib.gvi_multinomial_risk_f64(None, None, None, 1, 10, 2, None, None, None) == -1
    assert lib.gvi_multinomial_risk_f64(None, None, None, 4, 10, 0, None, None, None) == -1  # power >= 1
    assert lib.gvi_multinomial_risk_grad_f64(None, None, None, 4, 10, 2, None, None, None) == -1
    assert lib.gvi_multinomial_risk_workspace_bytes(1000) > 0
    assert lib.gvi_risk_grad_f64(99, 0.5, None, None, None, None, None, 10, None, None, None, None) == -1
    # vector-Jacobian products
    assert lib.gvi_gp_predict_grad_f64(None, 4, 3, None, 0, None, 0, None, None, None) == -1  # N >= 1
    assert lib.gvi_weighted_gram_f64(None, 4, 40, None, 10, None, None, None, None) == -1  # D <= 32
    assert b"D <= 32" in lib.gvi_last_error_string()
    assert lib.gvi_exact_gp_predict_grad_f64(None, 4, 40, None, 10, None, None, None, None, None, None) == -1  # nulls
    assert lib.gvi_exact_gp_predict_grad_workspace_bytes(1024, 1024, 8) >= 4 * 1024 * 1024 * 8
    assert (lib.gvi_exact_gp_predict_grad_workspace_bytes(1024, 1024, 784)
            > lib.gvi_exact_gp_predict_grad_workspace_bytes(1024, 1024, 8))  # wide-input route
    assert lib.gvi_fused_step_grad_workspace_bytes(60000, 2048, 784) >= 3 * 2048 * 2048 * 8  # wide-input route
    # optimiser step
    assert lib.gvi_optimiser_step_f64(7, None, None, None, None, 10, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, None) == -1
    assert lib.gvi_optimiser_step_f64(0, None, None, None, None, 10, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, None) == -1
    assert lib.gvi_optimiser_step_f64(0, None, None, None, None, 0, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, None) == 0


def test_argument_validation_of_the_other_kernel_entry_points(lib):
    """Non-ARD kernels (caller-supplied Grams): bad arguments are rejected before any CUDA call; empty inputs are a no-op."""
    assert lib.gvi_dot_gram_f64(None, 0, None, 4, 3, None, None, 2.0, None, 4, None) == 0
    assert lib.gvi_dot_gram_f64(None, 2, None, 4, 3, None, None, 2.0, None, 3, None) == -1  # ld_out < m2
    assert lib.gvi_dot_gram_f64(None, 2, None, 4, 3, None, None, 2.0, None, 4, None) == -1  # nulls
    assert lib.gvi_kernel_pairs_f64(5, None, None, 10, 3, None, None, 3, 1.0, None, None) == -1  # unknown kind
    assert lib.gvi_kernel_pairs_f64(0, None, None, 0, 3, None, None, 3, 1.0, None, None) == 0
    assert lib.gvi_gp_factor_gram_f64(None, 4, 4, 1e-5, 0, None, None, None, None, None, None) == -1
    assert lib.gvi_gp_predict_gram_f64(None, 4, None, 4, None, 0, None, None, None, None, None, None) == 0
    assert lib.gvi_gp_predict_gram_f64(None, 4, None, 4, None, 10, None, None, None, None, None, None) == -1
    assert lib.gvi_select_init_gram_f64(None, 10, 0, 3, 4, None, 1e-12, None, None, None) == -1
    assert lib.gvi_select_update_gram_f64(None, 10, 0, 3, 4, 0, None, None, 1e-12, None, None, None, None) == -1
    from gvi_gaussian_process_b200 import _lib

    assert not hasattr(lib, "gvi_test_set_grad_chunk_waves"), "test hooks must not be exported by the product library"
    tlib = _lib.load_test()
    assert tlib.gvi_test_set_grad_chunk_waves(3) == 0 and tlib.gvi_test_set_grad_chunk_waves(0) == 3


def test_xla_ffi_adapter_is_consistent_and_gated():
    """The XLA-FFI adapter cannot be compiled here (no jaxlib): check that its handler symbols and the registration
    table agree, that every C-ABI function it forwards to is declared in the header, and that the JAX entry points fail
    with a clear message instead of falling back to anything."""
    from gvi_gaussian_process_b200 import build, jax_ffi

    src = open(build.FFI_SOURCE).read()
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((gvi_[a-z0-9_]+),", src))
    assert defined == set(jax_ffi.FFI_TARGETS.values())
    forwarded = set(re.findall(r"\b(gvi_[a-z0-9_]+_f64)\(", src))
    assert forwarded and forwarded <= set(header_symbols())
    try:
        import jax  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="needs jax/jaxlib"):
            jax_ffi.register()
        assert build.xla_ffi_include_dir() is None
        with pytest.raises(RuntimeError, match="cannot be built"):
            build.build_ffi()


def test_xla_ffi_adapter_type_checks_against_the_c_abi_and_its_bindings(tmp_path):
    """jaxlib's headers are absent, so the adapter is compiled (syntax / type check only) against tests/xla_ffi_stub - a
    stub of the public xla/ffi/api/ffi.h surface the adapter uses, whose binding builder accumulates the handler's
    parameter list.  This verifies every forwarded call against include/gvigp.h and every handler signature against its
    binding; a deliberately broken copy must be rejected (so the check has teeth)."""
    import shutil
    import subprocess

    from gvi_gaussian_process_b200 import build

    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stub = os.path.join(root, "tests", "xla_ffi_stub")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", stub, "-I", cuda_inc]
    ok = subprocess.run(cmd + [build.FFI_SOURCE], capture_output=True, text=True)
    assert ok.returncode == 0, ok.stderr
    src = open(build.FFI_SOURCE).read().replace("../../../include/gvigp.h", os.path.join(root, "include", "gvigp.h"))
    swapped = src.replace('.Attr<int64_t>("m_sel").Attr<double>("jitter")', '.Attr<double>("jitter").Attr<int64_t>("m_sel")')
    assert swapped != src
    bad = tmp_path / "swapped_attrs.cc"
    bad.write_text(swapped)
    res = subprocess.run(cmd + [str(bad)], capture_output=True, text=True)
    assert res.returncode != 0 and "does not match its binding" in res.stderr
    dropped = src.replace("(int)m_sel,", "", 1)  # one argument fewer than gvi_cond_var_select_f64 takes
    assert dropped != src
    bad = tmp_path / "dropped_argument.cc"
    bad.write_text(dropped)
    res = subprocess.run(cmd + [str(bad)], capture_output=True, text=True)
    assert res.returncode != 0 and "gvi_cond_var_select_f64" in res.stderr


def _ffi_bindings():
    """target symbol -> (n_args, n_rets, [(attr type, attr name), ...]) parsed from the adapter's binding chains."""
    from gvi_gaussian_process_b200 import build

    src = open(build.FFI_SOURCE).read()
    out = {}
    for sym, chain in re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((gvi_[a-z0-9_]+),\s*\w+,\s*(.*?)\);", src, flags=re.S):
        out[sym] = (len(re.findall(r"\.Arg<", chain)), len(re.findall(r"\.Ret<", chain)),
                    re.findall(r'\.Attr<(\w+)>\("(\w+)"\)', chain))
    return out


def test_jax_wrappers_match_the_adapter_bindings(monkeypatch):
    """jax is absent, so the wrappers of jax_ffi.py run against a recording stand-in for jax / jax.numpy / jax.ffi:
    every ffi_call must pass as many operands and results as the C++ binding declares, with the binding's attribute
    names, and attribute values of the Python types XLA decodes as the declared C++ types (float -> double,
    int -> int64_t, bool -> bool).  Result buffers must have the sizes the C ABI's *_bytes queries ask for."""
    import sys
    import types

    import numpy as np

    from gvi_gaussian_process_b200 import _lib, jax_ffi

    calls = []

    class Spec:
        def __init__(self, shape, dtype):
            self.shape, self.dtype = tuple(shape), dtype

    def ffi_call(target, result_specs):
        def run(*operands, **attrs):
            specs = result_specs if isinstance(result_specs, (tuple, list)) else (result_specs,)
            calls.append((target, operands, specs, attrs))
            outs = tuple(np.ones(s.shape, dtype=s.dtype) for s in specs)
            return outs if isinstance(result_specs, (tuple, list)) else outs[0]
        return run

    class CustomVjp:
        def __init__(self, fn):
            self.fn = fn
        def defvjp(self, fwd, bwd):
            self.fwd, self.bwd = fwd, bwd
        def __call__(self, *a):
            return self.fn(*a)

    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    for name in ("atleast_1d", "zeros", "float64", "int64", "uint8", "ndim", "size", "sum", "shape"):
        setattr(jnp, name, getattr(np, name))
    jax.numpy = jnp
    jax.ShapeDtypeStruct = Spec
    jax.custom_vjp = CustomVjp
    jax.ffi = types.SimpleNamespace(ffi_call=ffi_call)
    monkeypatch.setitem(sys.modules, "jax", jax)
    monkeypatch.setitem(sys.modules, "jax.numpy", jnp)

    lib = _lib.load()
    n, m, d = 40, 6, 3
    x, z, y = np.zeros((n, d)), np.zeros((m, d)), np.zeros(n)
    jax_ffi.ard_gram(0.1, np.zeros(d), x, z)
    jax_ffi.dot_gram(0.1, None, 1, x, z)
    jax_ffi.dot_gram(0.1, 0.2, 3, x, z)
    factor = jax_ffi.gp_factor(z, 0.1, np.zeros(d))
    jax_ffi.gp_factor(z, 0.1, np.zeros(d), 1e-5, True, log_noise=0.0, y_z=np.zeros(m))
    mean, var = jax_ffi.gp_predict(factor, m, x)
    jax_ffi.risk(2, 0.5, y, mean, var)
    loss = jax_ffi.fused_loss(m, 5)
    args = (0.1, np.zeros(d), np.zeros(n), z, factor, np.zeros(n), x, y)
    loss(*args)
    grads = loss.bwd(args, 1.0)
    assert len(grads) == len(args) and np.shape(grads[1]) == (d,) and np.shape(grads[2]) == (n,)
    idx, trace = jax_ffi.cond_var_select(x, m, 0.1, np.zeros(d))
    assert idx.shape == (m,) and idx.dtype == np.int64 and trace.shape == (m - 1,)

    bindings = _ffi_bindings()
    py_type = {"double": float, "int64_t": int, "bool": bool}
    seen = set()
    for target, operands, specs, attrs in calls:
        n_args, n_rets, attr_decl = bindings[jax_ffi.FFI_TARGETS[target]]
        seen.add(target)
        assert len(operands) == n_args, target
        assert len(specs) == n_rets, target
        assert list(attrs) == [name for _, name in attr_decl], target  # same names, same order
        for ctype, name in attr_decl:
            assert type(attrs[name]) is py_type[ctype], (target, name)
        for op in operands:
            assert np.asarray(op).dtype == np.float64, target  # every operand is an F64 buffer in the adapter
    assert seen == set(jax_ffi.FFI_TARGETS)
    # result buffers sized by the C ABI's own queries (f64 words; the selector workspace is bytes)
    sizes = {t: [int(np.prod(s.shape)) for s in specs] for t, _, specs, _ in calls}
    assert sizes["gvi_gp_factor"] == [lib.gvi_gp_factor_bytes(m, d) // 8, lib.gvi_gp_factor_workspace_bytes(m) // 8]
    assert sizes["gvi_gp_predict"] == [n, n, lib.gvi_gp_predict_workspace_bytes(m, d) // 8]
    assert sizes["gvi_fused_step"] == [3, lib.gvi_fused_step_workspace_bytes(n, m) // 8]
    assert sizes["gvi_fused_step_grad"] == [4 + d, n, lib.gvi_fused_step_grad_workspace_bytes(n, m, d) // 8]
    assert sizes["gvi_cond_var_select"] == [m, m - 1, lib.gvi_select_workspace_bytes(n, m, d)]


def test_header_is_plain_c_and_a_c_program_links_against_the_library(lib, tmp_path):
    """The boundary is a C ABI: include/gvigp.h compiles as C99 (-pedantic-errors: no C++ isms, no torch types), and a
    C program with no Python in the process links against libgvigp.so, reads the ABI version, runs a size query and
    gets the documented error code + message from an argument check that needs no device."""
    import shutil

    from gvi_gaussian_process_b200 import _lib

    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    src = tmp_path / "abi_probe.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "gvigp.h"\n'
        "int main(void) {\n"
        "  if (gvi_abi_version() != 1) return 1;\n"
        "  if (gvi_select_record_len(8, 1024) != 2 + 8 + 1023) return 2;\n"
        "  if (gvi_ard_gram_f64(NULL, -1, NULL, 4, 3, NULL, NULL, 3, NULL, 4, NULL, NULL) != GVI_EINVAL) return 3;\n"
        '  if (!strstr(gvi_last_error_string(), "invalid argument")) return 4;\n'
        "  if (gvi_ard_gram_f64(NULL, 0, NULL, 4, 3, NULL, NULL, 3, NULL, 4, NULL, NULL) != GVI_OK) return 5;\n"
        '  printf("abi %d ok\\n", gvi_abi_version());\n  return 0;\n}\n')
    exe = tmp_path / "abi_probe"
    libdir = os.path.dirname(_lib.LIB_PATH)
    libname = os.path.basename(_lib.LIB_PATH)
    build = subprocess.run([gcc, "-std=c99", "-pedantic-errors", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                            str(src), "-o", str(exe), "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert run.stdout.strip() == "abi 1 ok"


def test_ctypes_signatures_match_the_header_prototypes():
    """Argument count and C type of every prototype in include/gvigp.h against the ctypes table the host code calls
    through (a drifted table would pass garbage without any error: ctypes cannot see the real prototype)."""
    from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_uint32, c_void_p

    from gvi_gaussian_process_b200 import _lib

    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(gvi_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text)
    assert sorted(p[1] for p in protos) == header_symbols()

    def ctype(t):
        t = t.strip()
        if "*" in t:
            return c_char_p if t.replace(" ", "") == "constchar*" else c_void_p
        return {"int": c_int, "int64_t": c_int64, "double": c_double, "size_t": c_size_t, "uint32_t": c_uint32}[t.replace("const", "").strip()]

    for ret, name, args in protos:
        args = [] if args.strip() == "void" else [a.strip() for a in args.split(",")]
        expected = [ctype(re.match(r"(.*?)[A-Za-z_][A-Za-z0-9_]*$", a).group(1)) for a in args]
        res, sig = _lib.SIGNATURES[name]
        assert res == ctype(ret), name
        assert sig == expected, (name, len(sig), len(expected))
