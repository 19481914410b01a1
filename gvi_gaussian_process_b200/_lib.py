"""ctypes binding of libgvigp.so (the C ABI declared in include/gvigp.h).

There is NO fallback: if the CUDA library is missing or a call fails, an exception is raised.
torch is used only for device memory, streams and (elsewhere) torch.distributed.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p

import torch  # noqa: F401  (loads libcudart so the .so shares torch's CUDA runtime)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GVI_LIB_PATH") or os.path.join(_HERE, "libgvigp.so")  # override: A/B builds when profiling

# symbol -> (restype, argtypes); must list every symbol of include/gvigp.h (tests/test_abi.py checks this)
SIGNATURES = {
    "gvi_abi_version": (c_int, []),
    "gvi_last_error_string": (c_char_p, []),
    "gvi_launch_count": (c_int64, []),
    "gvi_ard_gram_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "gvi_ard_gram_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                 c_int64, c_void_p, c_void_p]),
    "gvi_ard_gram_diag_f64": (c_int, [c_int64, c_void_p, c_void_p, c_void_p]),
    "gvi_add_diagonal_regulariser_f64": (c_int, [c_void_p, c_int, c_int64, c_double, c_int, c_void_p]),
    "gvi_chol_factor_f64": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "gvi_gp_factor_bytes": (c_size_t, [c_int, c_int]),
    "gvi_gp_factor_workspace_bytes": (c_size_t, [c_int]),
    "gvi_gp_factor_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_double, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_gp_factor_frozen_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                         c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_gp_factor_set_extra_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "gvi_gp_predict_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gvi_gp_predict_workspace_bytes_n": (c_size_t, [c_int, c_int, c_int64]),
    "gvi_gp_predict_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "gvi_risk_workspace_bytes": (c_size_t, [c_int64]),
    "gvi_risk_f64": (c_int, [c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                             c_void_p, c_void_p]),
    "gvi_risk_grad_f64": (c_int, [c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "gvi_fused_step_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "gvi_fused_step_workspace_bytes_d": (c_size_t, [c_int64, c_int, c_int]),
    "gvi_fused_step_f64": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int64, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "gvi_fused_step_grad_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "gvi_fused_step_grad_f64": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_int64, c_int, c_double, c_int, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_svgp_step_grad_f64": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int64, c_int, c_double, c_void_p, c_int64, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_gp_predict_grad_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p,
                                        c_void_p, c_void_p]),
    "gvi_exact_gp_predict_grad_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "gvi_exact_gp_predict_grad_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p]),
    "gvi_class_probabilities_f64": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_double,
                                            c_double, c_void_p, c_void_p]),
    "gvi_class_probabilities_grad_f64": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int,
                                                 c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_multinomial_risk_workspace_bytes": (c_size_t, [c_int64]),
    "gvi_multinomial_risk_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p,
                                         c_void_p]),
    "gvi_multinomial_risk_grad_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p,
                                              c_void_p]),
    "gvi_weighted_gram_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    "gvi_optimiser_step_f64": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_double, c_double,
                                       c_double, c_double, c_double, c_void_p]),
    "gvi_dot_gram_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_double, c_void_p,
                                 c_int64, c_void_p]),
    "gvi_kernel_pairs_f64": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_double,
                                     c_void_p, c_void_p]),
    "gvi_gp_factor_gram_f64": (c_int, [c_void_p, c_int64, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "gvi_gp_predict_gram_f64": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_select_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "gvi_cond_var_select_f64": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_double,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_select_record_len": (c_int64, [c_int, c_int]),
    "gvi_select_init_f64": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_double, c_void_p,
                                    c_void_p, c_void_p]),
    "gvi_select_init_gram_f64": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_double, c_void_p,
                                         c_void_p, c_void_p]),
    "gvi_select_update_gram_f64": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                           c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_select_resolve_f64": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "gvi_select_update_f64": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_fp64_tensor_peak_scratch_bytes": (c_size_t, []),
    "gvi_fp64_tensor_peak_tflops": (c_int, [c_double, c_void_p, c_void_p, c_void_p]),
    "gvi_threefry_random_bits": (c_int, [ctypes.c_uint32, ctypes.c_uint32, c_int64, c_void_p, c_void_p]),
    "gvi_comm_unique_id_h": (c_int, [c_void_p]),
    "gvi_comm_init": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int]),
    "gvi_comm_destroy": (c_int, [c_void_p]),
    "gvi_comm_rank": (c_int, [c_void_p]),
    "gvi_comm_size": (c_int, [c_void_p]),
    "gvi_comm_has_peer_window": (c_int, [c_void_p]),
    "gvi_comm_allreduce_sum_f64": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "gvi_comm_allgather_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "gvi_select_sharded_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "gvi_cond_var_select_sharded_f64": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_int,
                                                c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvi_select_virtual_shards_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "gvi_cond_var_select_virtual_shards_f64": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int,
                                                       c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                       c_void_p]),
}

# include/gvigp_test.h: exported only by libgvigp_test.so (tests and tools; never the product path)
TEST_LIB_PATH = os.path.join(_HERE, "libgvigp_test.so")
TEST_SIGNATURES = {
    "gvi_test_exp_f64": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "gvi_test_set_grad_chunk_waves": (c_int, [c_int]),
    "gvi_test_dgemm_f64": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_double,
                                   c_void_p, c_int64, c_int, c_int, c_void_p]),
    "gvi_test_dgemv_f64": (c_int, [c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_double, c_void_p, c_void_p]),
}

REG_KINDS = {
    "none": 0,
    "bhattacharyya": 1,
    "gaussian_wasserstein": 2,
    "hellinger": 3,
    "kl": 4,
    "renyi": 5,
    "gwi_no_eig": 6,
    "squared_difference": 7,
}


class GviError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load libgvigp.so and bind every symbol; raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GviError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_test_lib = None


def load_test() -> ctypes.CDLL:
    """libgvigp_test.so: the library compiled with -DGVI_TEST_HOOKS (test hooks + profiling environment switches)."""
    global _test_lib
    if _test_lib is None:
        if not os.path.exists(TEST_LIB_PATH):
            raise GviError(f"{TEST_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(TEST_LIB_PATH)
        for name, (res, args) in {**SIGNATURES, **TEST_SIGNATURES}.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _test_lib = lib
    return _test_lib


class test_library:
    """`with _lib.test_library() as lib:` - every call of the shim goes to libgvigp_test.so inside the block (tests that
    need a hook in effect while the public API runs; tools that use the profiling switches)."""

    def __enter__(self):
        global _lib
        load()
        self._prev = _lib
        _lib = load_test()
        return _lib

    def __exit__(self, *exc):
        global _lib
        _lib = self._prev
        return False


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gvi_last_error_string().decode()
        raise GviError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "libgvigp takes device pointers; got a CPU tensor"
    assert t.is_contiguous(), "libgvigp takes C-contiguous arrays"
    return t.data_ptr()


def current_stream() -> int:
    return torch.cuda.current_stream().cuda_stream
