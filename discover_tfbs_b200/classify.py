"""Scaler + SVC decision function on the GPU.

Drop-in for the inference arithmetic of predict.py:286-319: the reference refits its
MinMaxScaler -> StandardScaler pipeline (cross_validate_model.py:182-189) on the training matrix,
transforms the prediction features, calls `clf.predict` / `clf.decision_function` of the
grid-searched sklearn SVC (kernels "linear" and "rbf", cross_validate_model.py:204-208) and divides
by ||coef_|| for the distance from the hyperplane.  Training stays sklearn's (north_star keeps the
classifier unchanged); this class only evaluates a FITTED scaler + SVC through `tfbs_svc_decision`,
in double, so values agree with sklearn to ~1e-12 and the labels are identical.
"""
import ctypes as C

import numpy as np

from . import _lib
from .pssm import default_context


class GpuSVC:
    """Fitted (MinMaxScaler, StandardScaler, SVC) evaluated on the GPU."""

    def __init__(self, mm_scale, mm_min, st_mean, st_scale, kernel, sv, dual, gamma, intercept, classes, ctx=None):
        self.ctx = ctx or default_context()
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        self.mm_scale, self.mm_min, self.st_mean, self.st_scale = f64(mm_scale), f64(mm_min), f64(st_mean), f64(st_scale)
        self.kernel = {"linear": 0, "rbf": 1}[kernel]
        self.sv = f64(sv).reshape(-1, self.mm_scale.size)
        self.dual = f64(dual).reshape(-1)
        self.gamma, self.intercept = float(gamma), float(intercept)
        self.classes = np.asarray(classes)

    @classmethod
    def from_sklearn(cls, scaler, clf, ctx=None):
        """scaler: the fitted Pipeline([("scale", MinMaxScaler), ("standardize", StandardScaler)]) (or a
        (MinMaxScaler, StandardScaler) pair); clf: the fitted binary sklearn.svm.SVC."""
        mm, st = (scaler.steps[0][1], scaler.steps[1][1]) if hasattr(scaler, "steps") else scaler
        if clf.kernel not in ("linear", "rbf"):
            raise ValueError(f"SVC kernel {clf.kernel!r} is not one the reference searches (linear, rbf)")
        if len(clf.classes_) != 2:
            raise ValueError("binary SVC expected (seq_type True / Not_True)")
        st_scale = st.scale_ if st.scale_ is not None else np.ones_like(mm.scale_)
        st_mean = st.mean_ if st.mean_ is not None else np.zeros_like(mm.scale_)
        if clf.kernel == "linear":
            sv, dual, gamma = clf.coef_.reshape(1, -1), np.zeros(0), 0.0
        else:
            sv, dual, gamma = clf.support_vectors_, clf.dual_coef_.reshape(-1), clf._gamma
        return cls(mm.scale_, mm.min_, st_mean, st_scale, clf.kernel, sv, dual, gamma, clf.intercept_[0], clf.classes_, ctx)

    def decision_function(self, X) -> np.ndarray:
        """SVC.decision_function(scaler.transform(X)): float64[S]."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != self.mm_scale.size:
            raise ValueError(f"expected [S][{self.mm_scale.size}] features, got {X.shape}")
        out = np.empty(X.shape[0], dtype=np.float64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _lib._check(_lib.load().tfbs_svc_decision(
            self.ctx._h, p(X), X.shape[0], X.shape[1], p(self.mm_scale), p(self.mm_min), p(self.st_mean), p(self.st_scale),
            self.kernel, p(self.sv), p(self.dual) if self.dual.size else None, int(self.dual.size), self.gamma,
            self.intercept, p(out)))
        return out

    def predict(self, X) -> np.ndarray:
        """SVC.predict: classes_[1] where the decision value is positive."""
        return self.classes[(self.decision_function(X) > 0).astype(np.intp)]

    def hyperplane_distance(self, X) -> np.ndarray:
        """predict.py:314-319: decision_function / ||coef_|| (linear kernel); the reference falls back to
        the plain decision value when coef_ does not exist (rbf), and so does this."""
        d = self.decision_function(X)
        return d / np.linalg.norm(self.sv) if self.kernel == 0 else d
