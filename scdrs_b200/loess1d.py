"""One-predictor LOESS as ``skmisc.loess.loess(x, y, span=0.3, degree=2)`` computes it.

The reference calls scikit-misc (``scdrs/pp.py:384-386``), a wrapper of the netlib ``dloess``
Fortran code by Cleveland, Grosse and Shyu; scikit-misc is not installed in this image and its
source is not under /root/reference (pinned only as ``scikit-misc>=0.1.3``, setup.py:31).  This is
a restatement of the published algorithm for the configuration scDRS uses -- one predictor,
gaussian family (no robustness iterations), ``surface="interpolate"``, ``cell=0.2``,
``statistics="approximate"``:

1. bounding box of x, widened by ``margin`` = 0.005 * range on both sides (``ehg126``);
2. k-d tree on x: a cell with more than ``fc = floor(n * span * cell)`` points is cut at the
   median order statistic ``m = (l + u) // 2`` (ties moved to the upper son) (``ehg124/ehg129``);
3. at every vertex (box ends and cuts) a local weighted quadratic fit over the
   ``q = floor(n * span + 1e-5)`` nearest points with tricube weights, bandwidth = distance of the
   q-th nearest point, giving value and slope (``ehg127``);
4. fitted values by cubic Hermite interpolation inside the enclosing cell (``ehg128``).

Pinned against the only fixtures that exist for it: the ``ct_mean/ct_var -> ct_var_tech`` columns
of ``toydata_mouse.gene_stats.{nocov,cov}.tsv`` (2,500 genes each; tests/test_loess.py).
Host-side code, O(n log n); it is not on the device hot path.
"""
import numpy as np


def _kd_cuts(xs, fc):
    """Cut values of the 1-D k-d tree over sorted xs (0-based inclusive index ranges)."""
    cuts = []
    queue = [(0, len(xs) - 1, None, None)]
    lo_box, hi_box = None, None
    head = 0
    while head < len(queue):
        l, u, vlo, vhi = queue[head]
        head += 1
        if (u - l) + 1 <= fc:
            continue
        # Fortran: m = (l + u) / 2 with 1-based l, u  ->  0-based index below
        m = (l + 1 + u + 1) // 2 - 1
        while m > l and xs[m - 1] == xs[m]:      # all ties go with the upper son
            m -= 1
        t = xs[m]
        if (vlo is not None and t == vlo) or (vhi is not None and t == vhi):
            continue
        cuts.append(t)
        queue.append((l, m, vlo, t))
        queue.append((m + 1, u, t, vhi))
    return cuts


def _local_fit(x, y, w, x0, q, degree):
    """Value and slope at x0 of the tricube-weighted local polynomial through the q nearest points."""
    d = np.abs(x - x0)
    idx = np.argpartition(d, q - 1)[:q]
    dq = d[idx]
    h = dq.max()
    if h <= 0:
        raise ValueError("loess: zero bandwidth (too many identical x values)")
    u = dq / h
    wt = np.where(dq < h, (1.0 - u ** 3) ** 3, 0.0)
    if w is not None:
        wt = wt * w[idx]
    sw = np.sqrt(wt)
    dx = x[idx] - x0
    cols = [np.ones_like(dx), dx] + ([dx * dx] if degree >= 2 else [])
    A = np.stack(cols, axis=1) * sw[:, None]
    b = y[idx] * sw
    norm = np.sqrt((A * A).sum(axis=0))
    norm[norm == 0] = 1.0
    coef, *_ = np.linalg.lstsq(A / norm, b, rcond=None)
    coef = coef / norm
    return coef[0], coef[1]


class Loess1D:
    """``model = Loess1D(x, y, span, degree); model.fit(); model.outputs.fitted_values``"""

    class _Outputs:
        fitted_values = None

    def __init__(self, x, y, span=0.75, degree=2, weights=None, cell=0.2, margin=0.005):
        self.x = np.asarray(x, dtype=np.float64).reshape(-1)
        self.y = np.asarray(y, dtype=np.float64).reshape(-1)
        assert self.x.shape == self.y.shape and self.x.shape[0] > 0
        self.span, self.degree, self.cell, self.margin = float(span), int(degree), float(cell), float(margin)
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)
        self.outputs = Loess1D._Outputs()

    def fit(self):
        x, y, n = self.x, self.y, self.x.shape[0]
        q = min(n, int(np.floor(n * self.span + 1e-5)))
        if q < self.degree + 1:
            raise ValueError("loess: span too small")
        fc = int(np.floor(n * self.span * self.cell))
        xs = np.sort(x)
        lo, hi = xs[0], xs[-1]
        mu = self.margin * max(hi - lo, 1e-10 * max(abs(lo), abs(hi)) + 1e-30)
        verts = np.array(sorted(set([lo - mu, hi + mu] + _kd_cuts(xs, fc))))
        val = np.empty(len(verts))
        slope = np.empty(len(verts))
        for i, v in enumerate(verts):
            val[i], slope[i] = _local_fit(x, y, self.weights, v, q, self.degree)
        self.vertices, self.vertex_values, self.vertex_slopes = verts, val, slope
        self.outputs.fitted_values = self.predict(x)
        return self

    def predict(self, z):
        z = np.asarray(z, dtype=np.float64)
        v = self.vertices
        # a point equal to a cut belongs to the lower cell (descent uses z <= cut)
        c = np.clip(np.searchsorted(v, z, side="left") - 1, 0, len(v) - 2)
        v0, v1 = v[c], v[c + 1]
        lg = v1 - v0
        h = (z - v0) / lg
        phi0 = (1 - h) ** 2 * (1 + 2 * h)
        phi1 = h ** 2 * (3 - 2 * h)
        psi0 = h * (1 - h) ** 2
        psi1 = -(h ** 2) * (1 - h)
        g, s = self.vertex_values, self.vertex_slopes
        return phi0 * g[c] + phi1 * g[c + 1] + (psi0 * s[c] + psi1 * s[c + 1]) * lg


def loess(x, y, span=0.75, degree=2, **kw):
    """Factory with the call signature the reference uses (``scdrs/pp.py:384``)."""
    return Loess1D(x, y, span=span, degree=degree, **kw)
