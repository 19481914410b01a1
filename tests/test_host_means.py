"""CPU tests of the mean functions against the goldens of the reference's tests/test_means.py.

The means are host-framework code (they stay in the host framework, torch here) - no CUDA
kernel of this repository is involved - so their arithmetic and reshaping rules can be checked on CPU tensors.  The
product always places them on the CUDA device (`utils.device.device()` raises without one: no CPU fallback); ONLY this
test module redirects that helper to the CPU.  SVGPMean's golden needs a kernel Gram and is covered on the GPU
(tests/test_gpu_other_kernels.py::test_svgp_mean_golden)."""
import numpy as np
import pytest
import torch

from gvi_gaussian_process_b200.src.means import ConstantMean, CustomMean, MultiOutputMean
from gvi_gaussian_process_b200.src.means.base import MeanBase, MeanBaseParameters
from gvi_gaussian_process_b200.src.utils import device as device_module

X = np.array([[1.0, 2.0, 3.0], [1.5, 2.5, 3.5]])


@pytest.fixture(autouse=True)
def host_tensors(monkeypatch):
    monkeypatch.setattr(device_module, "device", lambda: torch.device("cpu"))


class MockMean(MeanBase):  # mockers/mean.py:17-43 of the reference: ones of shape (n, k)
    Parameters = MeanBaseParameters

    def generate_parameters(self, parameters):
        return MeanBaseParameters()

    def _predict(self, parameters, x):
        return torch.ones((x.shape[0], self.number_output_dimensions), dtype=torch.float64)


def test_constant_mean_goldens():
    # tests/test_means.py:14-62 (array_equal)
    mean = ConstantMean()
    out = mean.predict(mean.generate_parameters({"constant": 0}), x=X)
    assert out.shape == (2,) and np.array_equal(out.numpy(), np.zeros(2))
    mean = ConstantMean(number_output_dimensions=5)
    out = mean.predict(mean.generate_parameters({"constant": np.ones(5)}), x=X)
    assert out.shape == (5, 2) and np.array_equal(out.numpy(), np.ones((5, 2)))
    mean = ConstantMean(number_output_dimensions=5, preprocess_function=lambda x: x.reshape(x.shape[0], -1))
    out = mean.predict(mean.generate_parameters({"constant": np.ones(5)}), x=np.ones((2, 3, 3, 1)))
    assert out.shape == (5, 2) and np.array_equal(out.numpy(), np.ones((5, 2)))


def test_constant_mean_tile_then_reshape_quirk():
    # src/means/constant_mean.py:66 tiles the [K] constant n times and src/means/base.py:97-100 reshapes the flat result
    # to [K, n]: with unequal constants the rows interleave - reproduced as coded, not "fixed"
    mean = ConstantMean(number_output_dimensions=2)
    out = mean.predict(mean.generate_parameters({"constant": np.array([1.0, 2.0])}), x=np.zeros((3, 4)))
    want = np.tile(np.array([1.0, 2.0]), 3).reshape(2, -1)
    assert np.array_equal(out.numpy(), want) and np.array_equal(want, [[1.0, 2.0, 1.0], [2.0, 1.0, 2.0]])
    with pytest.raises(AssertionError):  # constant_mean.py:46-47
        mean.generate_parameters({"constant": np.ones(3)})


def test_custom_mean_golden():
    # tests/test_means.py:65-88: a network returning ones((n, 1)), parameters constructed with custom=None
    nn_mean = CustomMean(mean_function=lambda parameters, x: torch.ones((x.shape[0], 1), dtype=torch.float64))
    out = nn_mean.predict(parameters=nn_mean.Parameters.construct(custom=None), x=X)
    assert out.shape == (2,) and np.array_equal(out.numpy(), np.ones(2))


def test_multi_output_mean_golden():
    # tests/test_means.py:127-162: four mock means -> ones((4, 2))
    multi = MultiOutputMean(means=[MockMean()] * 4)
    out = multi.predict(parameters=multi.Parameters(means=[MeanBaseParameters()] * 4), x=X)
    assert out.shape == (4, 2) and np.array_equal(out.numpy(), np.ones((4, 2)))
    with pytest.raises(AssertionError):  # multi_output_mean.py: every member must be single-output
        MultiOutputMean(means=[MockMean(number_output_dimensions=2)])


def test_parameters_of_another_module_are_rejected():
    # src/module.py:92-94 via means/base.py:94
    class Other(MeanBaseParameters):
        pass

    class StrictMean(MockMean):
        class Parameters(MeanBaseParameters):
            pass

    with pytest.raises(AssertionError):
        StrictMean().predict(parameters=Other(), x=X)
