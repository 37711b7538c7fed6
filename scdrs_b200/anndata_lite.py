"""A small stand-in for ``anndata.AnnData`` / ``anndata.read_h5ad``.

The reference takes an ``anndata.AnnData`` (/root/reference/scdrs/method.py:44,
/root/reference/scdrs/util.py:46-92).  ``anndata`` is not installed in this image, so the
host side accepts *any* object exposing ``X, obs, var, uns, obs_names, var_names, shape,
copy()`` -- a real AnnData works unchanged -- and ships this minimal class for the CLI,
the tests and the benchmark.
"""
import numpy as np
import pandas as pd
from scipy import sparse

from . import h5lite


class AnnData:
    def __init__(self, X, obs=None, var=None, uns=None, obsp=None):
        if not (sparse.issparse(X) or isinstance(X, np.ndarray)):
            X = np.asarray(X)
        n_obs, n_var = X.shape
        self._X = X
        self.obs = pd.DataFrame(index=pd.RangeIndex(n_obs).astype(str)) if obs is None else obs
        self.var = pd.DataFrame(index=pd.RangeIndex(n_var).astype(str)) if var is None else var
        assert len(self.obs) == n_obs and len(self.var) == n_var, "obs/var do not match X"
        self.uns = {} if uns is None else uns
        self.obsp = {} if obsp is None else obsp      # cell x cell matrices ("connectivities")

    @property
    def X(self):
        return self._X

    @X.setter
    def X(self, value):
        assert value.shape == self._X.shape, "X.shape must not change"
        self._X = value

    @property
    def obs_names(self):
        return self.obs.index

    @property
    def var_names(self):
        return self.var.index

    @property
    def shape(self):
        return self._X.shape

    @property
    def n_obs(self):
        return self._X.shape[0]

    @property
    def n_vars(self):
        return self._X.shape[1]

    def copy(self):
        import copy as _copy

        return AnnData(self._X.copy(), self.obs.copy(), self.var.copy(), _copy.deepcopy(self.uns),
                       {k: v.copy() for k, v in self.obsp.items()})

    def __getitem__(self, key):
        """Row/column subsetting by boolean masks, integer arrays, slices or obs / var names."""
        if not isinstance(key, tuple):
            key = (key, slice(None))
        r, c = key

        def positions(k, index):
            if isinstance(k, slice):
                return k
            k = np.asarray(k)
            if k.dtype.kind in "OUT" or (k.dtype.kind not in "biu"):      # names
                pos = index.get_indexer(k)
                assert (pos >= 0).all(), "unknown names in index"
                return pos
            return k

        r, c = positions(r, self.obs.index), positions(c, self.var.index)
        X = self._X[r, :][:, c]
        obsp = {k: v[r][:, r] for k, v in self.obsp.items()}
        return AnnData(X, self.obs.iloc[r].copy(), self.var.iloc[c].copy(), dict(self.uns), obsp)

    def __repr__(self):
        return "AnnData-lite object with n_obs x n_vars = %d x %d" % self.shape


def _read_frame(g):
    """anndata dataframe group -> DataFrame (encodings 0.1.0 and 0.2.0)."""
    idx_key = g.attrs.get("_index", "index")
    if idx_key is None or idx_key not in g:
        idx_key = "_index" if "_index" in g else "index"
    index = pd.Index(g[idx_key].read().astype(str))
    order = g.attrs.get("column-order", None)
    cols = [str(c) for c in (np.atleast_1d(order) if order is not None else [])]
    df = pd.DataFrame(index=index)
    for c in cols:
        node = g[c]
        if isinstance(node, h5lite.Group):  # 0.2.0 categorical: {codes, categories}
            codes = node["codes"].read()
            cats = node["categories"].read()
            df[c] = pd.Categorical.from_codes(codes, categories=list(cats))
        else:
            v = node.read()
            if "categories" in node.attrs and "__categories" in g and c in g["__categories"]:
                cats = g["__categories"][c].read()
                v = pd.Categorical.from_codes(v, categories=list(cats))
            df[c] = v
    return df


def _read_matrix(x):
    if isinstance(x, h5lite.Group):
        enc = str(x.attrs.get("encoding-type", x.attrs.get("h5sparse_format", "csr_matrix")))
        shape = tuple(int(s) for s in x.attrs.get("shape", x.attrs.get("h5sparse_shape")))
        parts = (x["data"].read(), x["indices"].read(), x["indptr"].read())
        return (sparse.csc_matrix if enc.startswith("csc") else sparse.csr_matrix)(parts, shape=shape)
    return x.read()


def read_h5ad(path):
    """Read ``X`` (csr/csc/dense), ``obs``, ``var`` and the ``obsp`` matrices (the neighbour graph
    ``connectivities`` the downstream analyses use) of an ``.h5ad`` file."""
    f = h5lite.File(path)
    obsp = {}
    if "obsp" in f and isinstance(f["obsp"], h5lite.Group):
        for key in f["obsp"].keys():
            obsp[str(key)] = _read_matrix(f["obsp"][key])
    return AnnData(_read_matrix(f["X"]), _read_frame(f["obs"]), _read_frame(f["var"]), obsp=obsp)
