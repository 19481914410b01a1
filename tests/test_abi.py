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
    assert lib.gvi_test_set_grad_chunk_waves(3) == 0 and lib.gvi_test_set_grad_chunk_waves(0) == 3


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
