"""CPU-only tests of the host-side mirror of the reference API: PRNG restatement, parameter containers,
assertion behaviour that fires before any launch, sharding helpers."""
import os

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


def test_prng_threefry_known_answers():
    # Random123 / jax tests/random_test.py::testThreefry2x32 vectors and the documented split(PRNGKey(0))
    f = jax_prng._threefry2x32
    u = lambda *v: np.array(v, dtype=np.uint32)
    assert [int(v) for v in f(u(0, 0), u(0, 0))] == [0x6B200159, 0x99BA4EFE]
    assert [int(v) for v in f(u(0xFFFFFFFF, 0xFFFFFFFF), u(0xFFFFFFFF, 0xFFFFFFFF))] == [0x1CB996FC, 0xBB002BE7]
    assert [int(v) for v in f(u(0x13198A2E, 0x03707344), u(0x243F6A88, 0x85A308D3))] == [0xC4923A9C, 0x483DF7A0]
    assert jax_prng.split(jax_prng.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]


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
    with pytest.raises(AssertionError, match="must be a kernel"):  # any KernelBase is accepted, nothing else
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


def test_data_and_parameter_files_round_trip(tmp_path):
    """experiments/shared/data.py:25-35 (`<name>.npz` with x and y) and the parameter checkpoint tree."""
    from gvi_gaussian_process_b200.experiments.shared import Data
    from gvi_gaussian_process_b200.src.module import ModuleParameters

    rng = np.random.default_rng(0)
    a = Data(x=rng.standard_normal((5, 2)), y=rng.standard_normal(5), name="train")
    b = Data(x=rng.standard_normal((3, 2)), y=rng.standard_normal(3), name="-more")
    a.save(str(tmp_path))
    back = Data.load(str(tmp_path), "train")
    assert np.array_equal(back.x, a.x) and np.array_equal(back.y, a.y)
    assert set(np.load(tmp_path / "train.npz").files) == {"x", "y"}
    assert Data.load(str(tmp_path), "missing") is None
    both = a + b
    assert both.name == "train-more" and both.x.shape == (8, 2) and both.y.shape == (8,)
    p = ModuleParameters(log_observation_noise=np.float64(0.3), mean={"constant": np.zeros(2)},
                         kernel={"kernels": [{"log_scaling": 0.1, "log_lengthscales": np.ones(3)},
                                             {"log_scaling": 0.2, "log_lengthscales": np.zeros(3)}]})
    path = str(tmp_path / "epoch-0.ckpt")
    p.save(path)
    q = ModuleParameters().load(path)
    assert float(q.log_observation_noise) == 0.3 and np.array_equal(q.mean["constant"], np.zeros(2))
    assert float(q.kernel["kernels"][1]["log_scaling"]) == 0.2
    assert np.array_equal(q.kernel["kernels"][0]["log_lengthscales"], np.ones(3))
    # empty containers, None leaves, tuples and awkward keys survive (a CustomKernel / CustomMean with no parameters)
    odd = ModuleParameters(custom={}, nothing=None, pair=(np.ones(2), np.zeros(1)), keys={"a/b": 1.0, "#0": 2.0},
                           empty_list=[])
    odd.save(str(tmp_path / "odd"))
    back = ModuleParameters().load(str(tmp_path / "odd"))
    assert back.custom == {} and back.nothing is None and back.empty_list == []
    assert isinstance(back.pair, tuple) and np.array_equal(back.pair[0], np.ones(2))
    assert float(back.keys["a/b"]) == 1.0 and float(back.keys["#0"]) == 2.0


def test_gaussian_container_follows_the_reference():
    """src/distributions.py:14-29: `.dict()` exists (ModuleParameters in the reference) and full_covariance compares the
    ranks with != (the exact GP's [1,N] mean with an [N] diagonal reads True there)."""
    from gvi_gaussian_process_b200.src.distributions import Gaussian

    g = Gaussian(mean=np.zeros((1, 4)), covariance=np.ones(4))
    assert g.full_covariance and set(g.dict()) == {"mean", "covariance"}
    assert not Gaussian(**Gaussian(mean=np.zeros(4), covariance=np.ones(4)).dict()).full_covariance
    assert Gaussian(mean=np.zeros(4), covariance=np.eye(4)).full_covariance


def test_generate_batch_follows_the_reference_order():
    """src/utils/data.py:11-62: permutation order, floor/ceil step counts, at least one step."""
    from gvi_gaussian_process_b200.src.utils import jax_prng
    from gvi_gaussian_process_b200.src.utils.data import generate_batch

    x, y = np.arange(10.0).reshape(10, 1), np.arange(10.0)
    key = jax_prng.PRNGKey(0)
    perm = jax_prng.permutation(key, 10)
    batches = list(generate_batch(key, (x, y), batch_size=4, shuffle=True, drop_last=True))
    assert len(batches) == 2 and np.array_equal(batches[0][1], y[perm[:4]]) and np.array_equal(batches[1][0], x[perm[4:8]])
    batches = list(generate_batch(key, (x, y), batch_size=4, shuffle=False, drop_last=False))
    assert [len(b[1]) for b in batches] == [4, 4, 2] and np.array_equal(batches[2][1], y[8:])
    assert len(list(generate_batch(key, x, batch_size=64, shuffle=False, drop_last=True))) == 1


def test_module_layout_and_public_names_match_the_reference():
    """Drop-in at the import level: every module path under the reference's src/ package exists under
    gvi_gaussian_process_b200.src with every public class / function name the reference defines or re-exports there.
    tests/golden/reference_api_manifest.json holds the names (generated by tests/golden/make_api_manifest.py with `ast`
    from the reference tree, which is not available at test time)."""
    import importlib
    import json
    import os

    manifest = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_api_manifest.json")))
    assert len(manifest) >= 60
    missing = []
    for module, names in sorted(manifest.items()):
        full = "gvi_gaussian_process_b200.src" + ("." + module if module else "")
        try:
            mod = importlib.import_module(full)
        except ImportError as e:
            missing.append(f"{full}: {e}")
            continue
        missing += [f"{full}.{name}" for name in names if not hasattr(mod, name)]
    assert not missing, missing


def test_gp_and_kernel_families_share_the_reference_base_classes():
    # the bases the reference's experiments name in annotations and isinstance checks
    # (experiments/shared/resolvers/empirical_risk.py:16, regularisation.py:93-95, experiments/shared/trainers.py:20)
    from gvi_gaussian_process_b200.src.gps import (ApproximateGPClassification, ApproximateGPRegression, GPClassification,
                                                   GPRegression)
    from gvi_gaussian_process_b200.src.gps.base.approximate_base import ApproximateGPBase
    from gvi_gaussian_process_b200.src.gps.base.classification_base import GPClassificationBase
    from gvi_gaussian_process_b200.src.gps.base.exact_base import ExactGPBase
    from gvi_gaussian_process_b200.src.gps.base.regression_base import GPRegressionBase
    from gvi_gaussian_process_b200.src.kernels.approximate import (CholeskySVGPKernel, FixedSparsePosteriorKernel,
                                                                   SparsePosteriorKernel)
    from gvi_gaussian_process_b200.src.kernels.approximate.base import ApproximateBaseKernel
    from gvi_gaussian_process_b200.src.kernels.approximate.svgp.base import SVGPBaseKernel
    from gvi_gaussian_process_b200.src.kernels.standard.base import StandardKernelBase

    assert issubclass(GPRegression, (ExactGPBase)) and issubclass(GPRegression, GPRegressionBase)
    assert issubclass(ApproximateGPRegression, ApproximateGPBase) and issubclass(ApproximateGPRegression, GPRegressionBase)
    assert issubclass(GPClassification, ExactGPBase) and issubclass(GPClassification, GPClassificationBase)
    assert issubclass(ApproximateGPClassification, ApproximateGPBase)
    assert issubclass(ApproximateGPClassification, GPClassificationBase)
    assert not issubclass(GPRegression, (ApproximateGPBase, GPClassificationBase))
    assert not issubclass(ApproximateGPRegression, (ExactGPBase, GPClassificationBase))
    assert not issubclass(GPClassification, (ApproximateGPBase, GPRegressionBase))
    for kernel in (SparsePosteriorKernel, FixedSparsePosteriorKernel, SVGPBaseKernel, CholeskySVGPKernel):
        assert issubclass(kernel, ApproximateBaseKernel)
    assert issubclass(ARDKernel, StandardKernelBase) and not issubclass(ARDKernel, ApproximateBaseKernel)


@pytest.mark.parametrize("variant", ["polynomial", "table32"])
def test_device_exp_source_evaluated_on_the_host(tmp_path, variant):
    """gvi_exp (csrc/common.cuh) is the exp of every kernel evaluation on the hot path; gvi_exp_neg_t32 is the table-driven
    variant of the materialised Gram kernel (table: GVI_EXP2_TABLE in csrc/common.cuh).  Their bodies are plain IEEE-754
    double arithmetic (fma, add, compare, an integer add into the exponent field), so the DEVICE SOURCE itself is
    compiled for the host here (tests/csrc/gvi_exp_host.c supplies the three bit-cast intrinsics) and swept over 8e6
    arguments of the range the kernels use: <= 1 ulp against expl(), exact 1 at 0, exact 0 below -708 (and at -inf),
    and a NaN argument comes out as NaN (diverged hyper-parameters must surface as a NaN loss, as in the reference:
    experiments/shared/trainer.py:75-82)."""
    import re
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "gvi_gaussian_process_b200", "csrc", "common.cuh")).read()
    if variant == "polynomial":
        body = re.search(r"__device__ __forceinline__ double gvi_exp\(double x\) \{.*?\n\}\n", src, flags=re.S)
        assert body, "gvi_exp not found in common.cuh"
        text = body.group(0)
    else:
        body = re.search(r"__device__ __forceinline__ double gvi_exp_neg_t32\(double s, const double2\* tab.*?\n\}\n", src,
                         flags=re.S)
        assert body, "gvi_exp_neg_t32 not found in common.cuh"
        table = re.search(r"static __device__ const double2 GVI_EXP2_TABLE\[32\] = \{.*?\};", src, flags=re.S)
        assert table, "GVI_EXP2_TABLE not found in common.cuh"
        text = ("typedef struct { double x, y; } double2;\n"
                + table.group(0).replace("static __device__ const", "static const") + "\n#define CHECKED 1\n" + body.group(0)
                + "static inline double gvi_exp(double x) { return gvi_exp_neg_t32(-x, GVI_EXP2_TABLE, 1); }\n")
    (tmp_path / "gvi_exp_body.inc").write_text(text)
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    exe = tmp_path / "gvi_exp_host"
    build = subprocess.run([gcc, "-O2", "-ffp-contract=off", "-I", str(tmp_path),
                            os.path.join(root, "tests", "csrc", "gvi_exp_host.c"), "-o", str(exe), "-lm"],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout
    fields = dict(line.split(None, 1) for line in out.strip().splitlines())
    assert float(fields["max_ulp"].split()[0]) <= 1.0, out
    for flag in ("exp0", "below", "edge", "nan"):
        assert fields[flag].strip() == "1", out


def test_device_log_source_evaluated_on_the_host(tmp_path):
    """gvi_log (csrc/common.cuh) is the logarithm of the Gaussian NLL and of the KL / Renyi / Bhattacharyya terms in the
    risk kernel and the fused epilogue: table + function text compiled for the host and swept over 1e7 arguments (every
    binade, [0.5, 2] densely, 1 +- 2^-k, the table's interval boundaries): <= 2 ulp against logl() (1.84 next to the
    interval around 1, where log c and r cancel by one bit), log(1) = +0 exactly,
    log(0) = -inf, negative -> NaN (a negative predictive variance must surface as a NaN loss like the reference's
    jnp.log), inf, NaN and subnormal arguments through the library branch."""
    import re
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "gvi_gaussian_process_b200", "csrc", "common.cuh")).read()
    body = re.search(r"static __device__ const double GVI_LOG_TABLE\[128\]\[3\] = \{.*?__device__ __forceinline__ double gvi_log\(double x\) \{.*?\n\}\n",
                     src, flags=re.S)
    assert body, "gvi_log not found in common.cuh"
    (tmp_path / "gvi_log_body.inc").write_text(body.group(0).replace("static __device__ const", "static const"))
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    exe = tmp_path / "gvi_log_host"
    build = subprocess.run([gcc, "-O2", "-ffp-contract=off", "-I", str(tmp_path),
                            os.path.join(root, "tests", "csrc", "gvi_log_host.c"), "-o", str(exe), "-lm"],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300).stdout
    fields = dict(line.split(None, 1) for line in out.strip().splitlines())
    assert float(fields["max_ulp"].split()[0]) <= 2.0, out
    for flag in ("one", "zero", "negative", "inf", "nan", "subnormal"):
        assert fields[flag].strip() == "1", out


def _pointwise_host(tmp_path):
    """Compile the device source of the pointwise loss terms (csrc/common.cuh) for the host; returns a callable
    rows[kind, alpha, y, mp, cp, mq, cq] -> array [n, 6] = (nll, dnll/dm, dnll/dv, D0, dD0/dmq, dD0/dcq)."""
    import re
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "gvi_gaussian_process_b200", "csrc", "common.cuh")).read()
    body = re.search(r"static __device__ const double GVI_LOG_TABLE\[128\]\[3\] = \{.*?__device__ __forceinline__ double gvi_reg_term\(.*?\n\}\n",
                     src, flags=re.S)
    assert body, "pointwise loss terms not found in common.cuh"
    for name in ("gvi_log", "gvi_nll_term", "gvi_nll_grad", "gvi_reg_grad", "gvi_reg_term"):
        assert name in body.group(0)
    (tmp_path / "pointwise_body.inc").write_text(body.group(0))
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    exe = tmp_path / "pointwise_host"
    build = subprocess.run([gxx, "-O2", "-ffp-contract=off", "-I", str(tmp_path), "-I", os.path.join(root, "include"),
                            os.path.join(root, "tests", "csrc", "pointwise_host.cpp"), "-o", str(exe)],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr

    def run(rows):
        text = "".join("%d %.17g %.17g %.17g %.17g %.17g %.17g\n" % tuple(r) for r in rows)
        out = subprocess.run([str(exe)], input=text, capture_output=True, text=True, timeout=120).stdout
        return np.array([[float(v) for v in line.split()] for line in out.strip().splitlines()])

    return run


def test_device_source_of_the_pointwise_loss_terms_against_reference_goldens_and_the_oracle(tmp_path):
    """The formulas the risk / fused kernels execute per point, compiled for the host from the device source:
    the reference's golden values (tests/regularisations/projected/test_projected_*.py:120-176 regression case
    p = (1.8333.., 1/3), q = (1, 1); tests/test_empirical_risks.py), the NumPy oracle on random arguments, and the
    hand-derived derivatives (the gradient kernels' epilogues) against central differences of the terms."""
    from gvi_gaussian_process_b200._lib import REG_KINDS
    from oracle import gp_oracle as O

    run = _pointwise_host(tmp_path)
    mp, cp, mq, cq = 1.8333333333333335, 0.33333333333333326, 1.0, 1.0
    goldens = {"bhattacharyya": 0.13702468, "gaussian_wasserstein": 0.87307724, "hellinger": 0.18301035,
               "kl": 0.56319503, "renyi": 0.4042577, "squared_difference": 1.13888889}
    got = run([(REG_KINDS[k], 0.5, 0.0, mp, cp, mq, cq) for k in goldens])
    for (k, g), row in zip(goldens.items(), got):
        assert abs(row[3] - g) <= 1e-8, (k, row[3], g)
    # p == q  =>  every D0 is 0 (the reference's "same Gaussian" cases)
    same = run([(REG_KINDS[k], 0.5, 0.0, 0.7, 1.3, 0.7, 1.3) for k in goldens])
    assert np.all(np.abs(same[:, 3]) <= 1e-15)
    # Gaussian NLL golden of the approximate GP (mean 1, variance 1, y = [1.2, 4.2]): 3.48893853
    nll = run([(0, 0.5, y, 0.0, 1.0, 1.0, 1.0) for y in (1.2, 4.2)])[:, 0]
    assert abs(nll.mean() - 3.48893853) <= 1e-8
    # random arguments against the oracle
    rng = np.random.default_rng(3)
    n = 400
    y, m_p, m_q = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    c_p, c_q = np.exp(rng.uniform(-3, 2, n)), np.exp(rng.uniform(-3, 2, n))
    for name, kind in REG_KINDS.items():
        if name == "none":
            continue
        for alpha in ((0.5, 0.2, 0.9) if name == "renyi" else (0.5,)):
            rows = run([(kind, alpha, y[i], m_p[i], c_p[i], m_q[i], c_q[i]) for i in range(n)])
            if name == "gwi_no_eig":
                ref = (m_p - m_q) ** 2 + c_p + c_q
            elif name == "squared_difference":
                ref = (m_p - m_q) ** 2 + (c_p - c_q) ** 2
            elif name == "renyi":
                ref = O.d0_renyi(m_p, c_p, m_q, c_q, alpha)
            else:
                ref = O.D0[name](m_p, c_p, m_q, c_q)
            assert np.max(np.abs(rows[:, 3] - ref) / (1.0 + np.abs(ref))) <= 1e-13, name
            ref_nll = np.array([O.gaussian_nll([y[i]], [m_q[i]], [c_q[i]]) for i in range(n)])
            assert np.max(np.abs(rows[:, 0] - ref_nll) / (1.0 + np.abs(ref_nll))) <= 1e-13
            # derivatives: central differences of the terms themselves (relative steps)
            hm, hc = 1e-6, 1e-6 * c_q
            up = run([(kind, alpha, y[i], m_p[i], c_p[i], m_q[i] + hm, c_q[i]) for i in range(n)])
            dn = run([(kind, alpha, y[i], m_p[i], c_p[i], m_q[i] - hm, c_q[i]) for i in range(n)])
            uc = run([(kind, alpha, y[i], m_p[i], c_p[i], m_q[i], c_q[i] + hc[i]) for i in range(n)])
            dc = run([(kind, alpha, y[i], m_p[i], c_p[i], m_q[i], c_q[i] - hc[i]) for i in range(n)])
            for col_term, col_dm, col_dc in ((0, 1, 2), (3, 4, 5)):
                fd_m = (up[:, col_term] - dn[:, col_term]) / (2 * hm)
                fd_c = (uc[:, col_term] - dc[:, col_term]) / (2 * hc)
                scale_m = 1.0 + np.abs(fd_m)
                scale_c = 1.0 + np.abs(fd_c)
                assert np.max(np.abs(rows[:, col_dm] - fd_m) / scale_m) <= 1e-6, (name, "d/dm", col_term)
                assert np.max(np.abs(rows[:, col_dc] - fd_c) / scale_c) <= 1e-6, (name, "d/dc", col_term)
