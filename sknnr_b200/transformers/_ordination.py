"""Fit-time ordination math (cold path, NumPy/LAPACK; out of the hot-path scope but needed so
that the estimators are usable stand-alone).  Written from the published algorithms:

* canonical correspondence analysis after ter Braak (1986) as implemented by vegan's
  ``ordConstrained`` (the formulation the reference follows in
  ref:src/sknnr/transformers/_cca.py:74-203), and
* canonical correlation analysis in the SVD form of statsmodels' ``CanCorr`` with yaImpute's
  coefficient scaling and F-test for the number of significant axes
  (ref:src/sknnr/transformers/_ccora.py:5-136).

Only the quantities the query-time path needs are produced: the centre and the projector.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_SQRT_EPS = float(np.sqrt(2.220446e-16))


@dataclass
class CCAResult:
    env_center: np.ndarray      # [d] row-weighted mean of the environmental matrix
    coefficients: np.ndarray    # [d, rank] regression coefficients of the site scores
    eigenvalues: np.ndarray     # [rank]
    rank: int

    @property
    def max_components(self) -> int:
        return self.rank

    def projector(self, n_components: int) -> np.ndarray:
        """coefficients[:, :n] scaled by sqrt(eigenvalue share) (ref:_cca.py:187-203)."""
        share = np.sqrt(self.eigenvalues / self.eigenvalues.sum())
        return self.coefficients[:, :n_components] * share[None, :n_components]


def fit_cca(X: np.ndarray, Y: np.ndarray) -> CCAResult:
    """Canonical correspondence analysis of species matrix Y constrained by X."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if np.any(Y.sum(axis=1) <= 0.0):
        raise ValueError("All row sums must be greater than 0")
    Y = Y[:, Y.sum(axis=0) > 0.0]

    # chi-square standardisation of the species table (vegan initCA)
    P = Y / Y.sum()
    rw = P.sum(axis=1)
    cw = P.sum(axis=0)
    expected = np.outer(rw, cw)
    Ybar = (P - expected) / np.sqrt(expected)

    # weighted centring and row-weighting of the constraints
    center = np.average(X, axis=0, weights=rw)
    Xw = (X - center) * np.sqrt(rw)[:, None]

    # weighted regression of Ybar on Xw through QR, then SVD of the fitted values
    Q, R = np.linalg.qr(Xw)
    beta, _, ls_rank, _ = np.linalg.lstsq(R, Q.T @ Ybar, rcond=None)
    fitted = Xw @ beta
    U, s, _ = np.linalg.svd(fitted, full_matrices=False)
    rank = int(min(ls_rank, int(np.sum(s > _SQRT_EPS))))
    U = U[:, :rank]
    eig = np.square(s)[:rank]

    coef = np.linalg.lstsq(R, Q.T @ U, rcond=None)[0]
    return CCAResult(env_center=center, coefficients=coef, eigenvalues=eig, rank=rank)


@dataclass
class CCorAResult:
    x_coef: np.ndarray          # [d, k] canonical coefficients of X (yaImpute scaling)
    cancorr: np.ndarray         # canonical correlations, clipped to [0, 1]
    n_significant: int

    @property
    def max_components(self) -> int:
        return self.n_significant

    def projector(self, n_components: int) -> np.ndarray:
        return self.x_coef[:, :n_components] * self.cancorr[None, :n_components]


def _svd_tol(A: np.ndarray, tol: float):
    u, s, vt = np.linalg.svd(A, full_matrices=False)
    keep = s > tol
    return u[:, keep], s[keep], vt[keep][:, keep]


def _rao_f_pvalues(p: int, q: int, n: int, cor: np.ndarray) -> np.ndarray:
    """Rao's F approximation to Wilks' lambda for each successive set of canonical
    correlations (yaImpute ``ftest.cor``; ref:_ccora.py:5-41)."""
    from scipy.stats import f as f_dist

    s = min(p, q)
    k = np.arange(1, s + 1)
    wilks = np.array([np.prod(1.0 - np.square(cor[i:s])) for i in range(s)])
    r = (n - s - 1) - ((abs(p - q) + 1) / 2.0)
    a, b = (p - k + 1), (q - k + 1)
    ndf = a * b
    u = (ndf - 2) / 4.0
    denom = np.square(a) + np.square(b) - 5
    t = np.where(denom > 0, np.sqrt(np.where(denom > 0, (np.square(a) * np.square(b) - 4) / np.where(denom > 0, denom, 1), 0.0)), 0.0)
    ok = t > 0
    t, wilks, u, ndf = t[ok], wilks[ok], u[ok], ndf[ok]
    lam = np.power(wilks, 1.0 / t)
    ddf = r * t - 2 * u
    bad = (ddf < 1.0) | (ndf < 1)
    first = np.flatnonzero(bad)
    if len(first):
        bad[first[0]:] = True
    F = ((1.0 - lam) / lam) * (ddf / ndf)
    pvals = np.array([1.0 - f_dist.cdf(F[i], ndf[i], ddf[i]) for i in range(len(F))])
    return pvals[~bad & ~np.isnan(pvals)]


def fit_ccora(X: np.ndarray, Y: np.ndarray, tol: float = 1e-8, p_val: float = 0.05) -> CCorAResult:
    """Canonical correlation analysis between (already standardised) X and Y."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    k = min(X.shape[1], Y.shape[1])
    Xc = X - X.mean(axis=0)
    Yc = Y - Y.mean(axis=0)
    ux, sx, vx = _svd_tol(Xc, tol)
    uy, sy, vy = _svd_tol(Yc, tol)
    u, s, vt = np.linalg.svd(ux.T @ uy, full_matrices=False)
    cancorr = np.clip(s, 0.0, 1.0)

    vx_ds = vx.T / sx
    vy_ds = vy.T / sy
    # yaImpute rescales all coefficients by the sd of the first canonical variate
    first_variate = Xc @ (vx_ds @ u[:, 0])
    cscal = 1.0 / np.std(first_variate, ddof=1)
    x_coef = vx_ds @ u[:, :k] * cscal
    y_coef = vy_ds @ vt.T[:, :k] * cscal

    pvals = _rao_f_pvalues(y_coef.shape[0], x_coef.shape[0], Yc.shape[0], cancorr)
    n_sig = max(1, len(pvals) - int(np.sum(pvals > p_val)))
    return CCorAResult(x_coef=x_coef, cancorr=cancorr, n_significant=n_sig)
