"""CPU-only tests of the host-side mirror of the reference API: PRNG restatement, parameter containers,
assertion behaviour that fires before any launch, sharding helpers."""
import numpy as np
import pytest
import torch

from gvi_gaussian_process_b200 import distributed
from gvi_gaussian_process_b200.src.inducing_points_selection import ConditionalVarianceInducingPointsSelector
from gvi_gaussian_process_b200.src.kernels.standard import ARDKernel
from gvi_gaussian_process_b200.src.module import Module, ModuleParameters
from gvi_gaussian_process_b200.src.utils import jax_prng
from oracle import jax_prng as oracle_prng


def test_prng_matches_reference_pins():
    # tests/test_inducing_points_selection.py:128,139,151,162 of the reference
    key = jax_prng.PRNGKey(0)
    _, sub = jax_prng.split(key)
    assert list(jax_prng.permutation(sub, 4)) == [2, 3, 0, 1]
    assert list(jax_prng.permutation(sub, 3)) == [2, 0, 1]
    assert list(jax_prng.choice(key, 4, 2)) == [1, 2]
    assert list(jax_prng.choice(key, 3, 2)) == [0, 2]


@pytest.mark.parametrize("seed,n", [(0, 1), (1, 2), (7, 1625), (7, 1626), (123456789012, 5000)])
def test_prng_product_and_oracle_agree(seed, n):
    k1, k2 = jax_prng.PRNGKey(seed), oracle_prng.prng_key(seed)
    assert np.array_equal(k1, k2)
    assert np.array_equal(jax_prng.split(k1, 3), oracle_prng.split(k2, 3))
    p = jax_prng.permutation(k1, n)
    assert np.array_equal(p, oracle_prng.permutation(k2, n))
    assert sorted(p.tolist()) == list(range(n))


def test_parameters_container():
    p = ARDKernel.Parameters.construct(log_scaling=0.5, log_lengthscales=np.array([0.1, 0.2]))
    d = p.dict()
    assert set(d) == {"log_scaling", "log_lengthscales"}
    k = ARDKernel(number_of_dimensions=2)
    q = k.generate_parameters(d)
    assert isinstance(q, ARDKernel.Parameters)
    nested = ModuleParameters(kernel=p, log_observation_noise=0.0)
    assert nested.dict()["kernel"] == d
    with pytest.raises(AssertionError):
        Module.check_parameters(object(), ARDKernel.Parameters)


def test_selector_asserts_before_any_launch():
    sel = ConditionalVarianceInducingPointsSelector()
    with pytest.raises(AssertionError, match="at least 2 inducing points"):
        sel.compute_inducing_points(jax_prng.PRNGKey(0), np.zeros((4, 3)), 1, ARDKernel(3), {})
    with pytest.raises(NotImplementedError):
        sel.compute_inducing_points(jax_prng.PRNGKey(0), np.zeros((4, 3)), 2, object(), {})


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 100, 10**6 + 3):
        for w in (1, 2, 3, 8):
            cuts = [distributed.shard_bounds(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_resolve_records_tie_rules():
    rec = torch.tensor([[1.0, 5.0], [2.0, 9.0], [2.0, 3.0], [0.5, 11.0]], dtype=torch.float64)
    assert distributed.resolve_records(rec, first=True) == 2   # max d, lowest index
    assert distributed.resolve_records(rec, first=False) == 1  # max d, highest index
    rec[:, 0] = -1.0
    rec[2, 0] = 0.0
    assert distributed.resolve_records(rec, first=False) == 2  # ranks without candidates (d = -1) never win


def test_no_cpu_fallback():
    from gvi_gaussian_process_b200.src.utils import device

    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            device.device()
