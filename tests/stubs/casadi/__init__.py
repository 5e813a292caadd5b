"""Minimal stand-in for the parts of the casadi Python API that mpc_fatigue_b200/casadi_adapter.py touches.  TEST
INFRASTRUCTURE ONLY (casadi itself is not installable in this image): it lets the adapter's code run — callback protocol,
sparsity construction, block assembly — so that its results can be compared with Function.jacobian() and the oracle.
Semantics follow CasADi >= 3.6: column-major Sparsity.triplet, DM as a sparse matrix of doubles, Callback.construct() querying
get_n_in/get_n_out/get_sparsity_*/get_name_*, calling with DM arguments, and `jacobian()` delegating to get_jacobian()."""
import numpy as np

__version__ = "0.0-stub"


class Sparsity:
    def __init__(self, nrow, ncol, mask):
        self._shape = (int(nrow), int(ncol))
        self._mask = np.asarray(mask, dtype=bool).reshape(self._shape)

    @staticmethod
    def dense(nrow, ncol=1):
        return Sparsity(nrow, ncol, np.ones((nrow, ncol), dtype=bool))

    @staticmethod
    def triplet(nrow, ncol, rows, cols):
        m = np.zeros((nrow, ncol), dtype=bool)
        m[np.asarray(rows, dtype=int), np.asarray(cols, dtype=int)] = True
        return Sparsity(nrow, ncol, m)

    def size1(self):
        return self._shape[0]

    def size2(self):
        return self._shape[1]

    def nnz(self):
        return int(self._mask.sum())

    def is_dense(self):
        return bool(self._mask.all())

    def get_triplet(self):
        c, r = np.nonzero(self._mask.T)  # column-major order, as CasADi stores non-zeros
        return r.tolist(), c.tolist()

    @property
    def shape(self):
        return self._shape


class DM:
    def __init__(self, x=None):
        if isinstance(x, Sparsity):
            self._sp, self._v = x, np.zeros(x.shape)
        elif isinstance(x, DM):
            self._sp, self._v = x._sp, x._v.copy()
        else:
            v = np.atleast_1d(np.asarray(x, dtype=np.float64))
            if v.ndim == 1:
                v = v.reshape(-1, 1)
            self._sp, self._v = Sparsity.dense(*v.shape), v.copy()

    def sparsity(self):
        return self._sp

    @property
    def shape(self):
        return self._v.shape

    def full(self):
        return np.where(self._sp._mask, self._v, 0.0)

    def nnz(self):
        return self._sp.nnz()

    def __array__(self, dtype=None, copy=None):
        return self.full() if dtype is None else self.full().astype(dtype)

    def __float__(self):
        return float(self.full().reshape(-1)[0])

    def __setitem__(self, key, val):
        mask = np.zeros(self._v.shape, dtype=bool)
        mask[key] = True
        if (mask & ~self._sp._mask).any():
            raise RuntimeError("assignment outside the sparsity pattern")  # CasADi would silently enlarge the pattern
        self._v[key] = np.asarray(val, dtype=np.float64)

    def __getitem__(self, key):
        return DM(self.full()[key])


class Callback:
    def __init__(self):
        self._constructed = False

    # defaults of casadi.Callback
    def get_n_in(self):
        return 1

    def get_n_out(self):
        return 1

    def get_sparsity_in(self, i):
        return Sparsity.dense(1, 1)

    def get_sparsity_out(self, i):
        return Sparsity.dense(1, 1)

    def get_name_in(self, i):
        return "i%d" % i

    def get_name_out(self, i):
        return "o%d" % i

    def has_jacobian(self):
        return False

    def construct(self, name, opts=None):
        self._name = name
        self._sp_in = [self.get_sparsity_in(i) for i in range(self.get_n_in())]
        self._sp_out = [self.get_sparsity_out(i) for i in range(self.get_n_out())]
        self._names_in = [self.get_name_in(i) for i in range(self.get_n_in())]
        self._names_out = [self.get_name_out(i) for i in range(self.get_n_out())]
        self._constructed = True

    # casadi.Function API subset
    def name(self):
        return self._name

    def n_in(self):
        return len(self._sp_in)

    def n_out(self):
        return len(self._sp_out)

    def name_in(self, i=None):
        return list(self._names_in) if i is None else self._names_in[i]

    def name_out(self, i=None):
        return list(self._names_out) if i is None else self._names_out[i]

    def sparsity_in(self, i):
        return self._sp_in[i]

    def sparsity_out(self, i):
        return self._sp_out[i]

    def __call__(self, *args, **kw):
        assert self._constructed, "construct() was not called"
        if kw:
            args = [kw.get(nm, 0.0) for nm in self._names_in]
        assert len(args) == len(self._sp_in), (len(args), len(self._sp_in))
        dm = []
        for a, sp in zip(args, self._sp_in):
            a = a if isinstance(a, DM) else DM(a)
            if a.shape == (1, 1) and sp.shape != (1, 1):
                a = DM(np.full(sp.shape, float(a)))
            assert a.shape == sp.shape, (a.shape, sp.shape)
            dm.append(a)
        res = self.eval(dm)
        assert len(res) == len(self._sp_out)
        for r, sp in zip(res, self._sp_out):
            assert isinstance(r, DM) and r.shape == sp.shape, (getattr(r, "shape", None), sp.shape)
            assert not (r.sparsity()._mask & ~sp._mask).any(), "result has non-zeros outside the declared sparsity"
        if kw:
            return dict(zip(self._names_out, res))
        return res[0] if len(res) == 1 else tuple(res)

    def jacobian(self):
        if not self.has_jacobian():
            raise RuntimeError("no Jacobian")
        jname = "jac_" + self._name
        inames = self._names_in + ["out_" + n for n in self._names_out]
        onames = ["jac_%s_%s" % (o, i) for o in self._names_out for i in self._names_in]
        return self.get_jacobian(jname, inames, onames, {})
