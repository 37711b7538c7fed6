"""The LOESS restatement against the only fixtures that pin it: inputs (ct_mean, ct_var) and outputs
(ct_var_tech) of skmisc.loess for 2,500 genes in the reference's toydata gene-stats tables."""
import numpy as np
import pytest

import helpers
from scdrs_b200.loess1d import Loess1D


@pytest.mark.parametrize("key,tol", [("cov", 1e-10), ("nocov", 5e-6)])
def test_loess_matches_reference_gene_stats_tables(key, tol):
    gs = helpers.load_toy()[4][key]
    ok = (gs.ct_var > 0) & (gs.ct_mean > 0)
    x, y = np.log10(gs.ct_mean[ok].values), np.log10(gs.ct_var[ok].values)
    want = np.log10(gs.ct_var_tech[ok].values)
    model = Loess1D(x, y, span=0.3, degree=2).fit()
    assert len(model.vertices) == 33
    assert np.max(np.abs(model.outputs.fitted_values - want)) < tol


def test_loess_reproduces_a_quadratic_exactly():
    rng = np.random.default_rng(0)
    x = np.sort(rng.normal(size=500))
    y = 1.5 - 0.3 * x + 0.25 * x * x
    fit = Loess1D(x, y, span=0.3, degree=2).fit().outputs.fitted_values
    # local quadratics recover the curve; the cubic Hermite blend of exact values/slopes is exact too
    assert np.max(np.abs(fit - y)) < 1e-9


def test_loess_handles_ties_and_small_inputs():
    x = np.repeat(np.arange(40.0), 5)
    y = np.sin(x / 7.0)
    fit = Loess1D(x, y, span=0.5, degree=2).fit().outputs.fitted_values
    assert np.all(np.isfinite(fit)) and np.max(np.abs(fit - y)) < 0.05
