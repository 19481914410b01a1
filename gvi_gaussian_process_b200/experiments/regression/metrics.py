"""calculate_negative_log_likelihood (reference: experiments/regression/metrics.py:14-28): mean Gaussian NLL of the
predictive on a data split plus log(y_std) (the responses were standardised).  The reference loops over the points on the
host; here the predictive comes from the CUDA path and the reduction is the risk kernel (gvi_risk_f64)."""
from __future__ import annotations

import math

from ...src import _lib_proxy as L
from ...src.utils.device import as_device


def calculate_negative_log_likelihood(data, y_std, gaussian):
    mean = as_device(gaussian.mean).reshape(-1).contiguous()
    variance = as_device(gaussian.covariance).reshape(-1).contiguous()
    y = as_device(data.y).reshape(-1).contiguous()
    out3 = L.risk(L.REG_KINDS["none"], 0.0, y, mean, variance, None, None)
    return out3[0] / out3[2] + math.log(float(y_std))


def calculate_metrics(experiment_splits, gp, gp_parameters, y_std=1.0):
    """{split name: {"negative-log-likelihood": value}} for the splits that carry responses (reference lines 31-98
    build the same table as a pandas frame)."""
    out = {}
    for name, data in experiment_splits.items():
        if data is None or data.y is None:
            continue
        gaussian = gp.predict_probability(gp_parameters, x=data.x)
        out[name] = {"negative-log-likelihood": float(calculate_negative_log_likelihood(data, y_std, gaussian))}
    return out
