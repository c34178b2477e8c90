'''
CPU ORACLE for the `yanocomp gmmtest` hot path -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy/SciPy restatement of the reference algorithm (all
citations are file:line under /root/reference/yanocomp/).  It is the checker
the CUDA path is compared with.  Only `tests/`, `__graft_entry__.smoke()` and
the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the
product package `yanocomp_b200` never does.

Parity status
-------------
* KS stage (stats.py:20-73): the arithmetic the reference itself uses
  (`np.linalg.svd`, `np.sort`, `np.searchsorted`, `scipy.stats.kstwobign.sf`)
  is installed here and is called the same way, so this stage IS the
  reference's arithmetic.  Pinned by tests/golden/fixture_ks.npz (generated
  from the reference's own HDF5 fixtures by tests/golden/make_golden.py).
* G-test / log-odds stage (stats.py:302-355): `scipy.stats.chi2_contingency`
  is called exactly as the reference does; statsmodels `Table2x2` is absent
  and restated in closed form (A8).
* GMM stage (stats.py:76-299): **parity unpinned**.  The arithmetic lives in
  pomegranate (>=0.13, setup.py:30; conda pin >=0.13.3, yanocomp.yml:12), a
  Cython dependency that is neither vendored in /root/reference nor
  installable offline, and the reference's only golden file (one stale BED
  line) was written by an older code version with an unseeded RNG.  The
  published pomegranate 0.13/0.14 algorithm is restated below; every place
  where its behaviour had to be assumed is listed in ASSUMPTIONS and marked
  `# A<n>` in the code.

Determinism: the reference never seeds pomegranate's k-means++ RNG
(stats.py:108-113) and consumes its sub-sampling RNG in Python-set iteration
order per worker chunk (gmmtest.py:35, :148-149), so it is not reproducible
run to run nor across `-p`.  The oracle (and the CUDA path, identically) uses
a counter-based generator keyed on (seed, feature id, position, stream), so
results are independent of visiting order and of the number of shards.
'''
import dataclasses
import gzip
import json
import math
import os

import numpy as np
from scipy import stats as sstats

N_COMPONENTS = 2   # stats.py:17

ASSUMPTIONS = {
    'A1': 'pm.Kmeans.distance returns SQUARED Euclidean distances',
    'A2': 'pm.Kmeans.fit: per restart k-means++ (first centre uniform, next '
          'with probability proportional to squared distance), Lloyd updates '
          'with the pomegranate loop shape (iteration 0 assigns only; stop '
          'when the decrease of the total squared distance <= stop_threshold '
          'or after max_iterations+1 passes); best restart = smallest final '
          'total distance; empty clusters keep their centroid',
    'A3': 'MultivariateGaussianDistribution: ML mean S1/S0; covariance '
          '(S2 - S1 u^T - u S1^T + S0 u u^T)/S0 with u the NEW mean; '
          'from_summaries(inertia) blends old*inertia + new*(1-inertia) for '
          'mean and covariance and is skipped when S0 < 1e-7 (summaries kept)',
    'A4': 'a covariance whose Cholesky factorisation fails (non-positive or '
          'non-finite pivot) makes the reference raise LinAlgError (directly '
          'or through the complex-cast bug of stats.py:245-249): the position '
          'is dropped.  An initial cluster of 1..d rows has an exactly singular '
          'covariance: its pivot is pure rounding noise in any implementation, '
          'so the exact-arithmetic outcome (failure) is used',
    'A5': 'UniformDistribution: inclusive bounds, log density -log(hi-lo) '
          'inside and -inf outside; its min/max summaries cannot leave the '
          'initial box so the box is constant during EM; a zero-width box '
          'gives non-finite likelihoods and is treated as a failed fit',
    'A6': 'GeneralMixtureModel stores log weights; summarize returns '
          'sum_i log sum_j w_j p_j(x_i) (sequential pairwise log-sum-exp) and '
          'adds the responsibilities to summaries[j]',
    'A7': 'GeneralMixtureModel(dists, weights) renormalises weights to sum 1',
    'A8': 'statsmodels Table2x2(shift_zeros=True): +0.5 on every cell if any '
          'cell is 0; log odds = ln f00 - ln f01 - ln f10 + ln f11; CI = '
          'lor -/+ norm.ppf(1-alpha/2) * sqrt(sum 1/f)',
    'A10': '--covariance-type diag: IndependentComponentsDistribution of '
           'NormalDistribution(mean, std): log density -log(sigma sqrt(2 pi)) '
           '- (x-mu)^2 / (2 sigma^2) per column (sigma = 0 gives +inf); '
           'summaries sum w, sum w x, sum w x^2; from_summaries(inertia) is '
           'skipped (summaries kept) when sum w < 1e-8, else mu = S1/S0, '
           'sigma = sqrt(S2/S0 - S1^2/S0^2) (min_std 0) and BOTH mu and sigma '
           'are blended old*inertia + new*(1-inertia); a non-finite likelihood '
           'is a failed fit as in A5',
    'A11': '--outlier-method dbscan: sklearn DBSCAN(eps, min_samples = '
           '0.25 n as a float, euclidean): a point is a core point when at '
           'least min_samples points (itself included) lie within eps, and an '
           'outlier (label -1) when it is neither a core point nor within eps '
           'of one (recent scikit-learn rejects a float min_samples; the '
           'pinned environment of the reference accepted it)',
    'A9': 'a window whose pooled data has a constant column makes '
          'np.linalg.svd raise inside the reference (stats.py:26, outside any '
          'try block) and abort the run; here the position is skipped',
}

# Every assumption whose alternative is plausible sits behind a named switch.  The defaults are what
# the CUDA path implements; `assume(...)` flips switches for a sensitivity run
# (oracle/assumption_sensitivity.py -> DESIGN.md section 4: how many significant calls move).
SWITCHES = {
    'A1_kmeans_distance': 'squared',     # | 'euclidean': Kmeans.distance returns sqrt(d2) (changes the MAD rule only)
    'A2_kmeans_stop': 'total',           # | 'mean': stop_threshold applies to the mean, not the total, squared distance
    'A2_kmeans_n_init': 3,               # | 1: n_init ignored, a single k-means++ start
    'A3_cov_centre': 'new_mean',         # | 'blended_mean': covariance centred on the inertia-blended mean
    'A4_cholesky_failure': 'drop',       # | 'eig_floor': eigenvalues floored at min_covar = 1e-5 and a retry
                                         # | 'eig_shift': cov -= min(eigenvalue) I, cov += min_covar I and a retry
                                         #   (two readings of pomegranate's fallback path; in both the position is
                                         #   dropped only if the retry fails too)
    'A4_rank_deficient_init': 'fail',    # | 'noise': an initial cluster of 1..d rows (exactly singular covariance) is
                                         #   left to the sign of the rounding noise in the Cholesky pivot
    'A5_zero_width_box': 'fail',         # | 'widen': a zero-width column of the uniform box is ignored
}
_DEFAULT_SWITCHES = dict(SWITCHES)


class assume:
    '''context manager: `with assume(A4_cholesky_failure='eig_floor'): ...` (test infrastructure)'''

    def __init__(self, **kw):
        unknown = set(kw) - set(SWITCHES)
        if unknown:
            raise KeyError(f'unknown assumption switches {sorted(unknown)}')
        self.kw = kw

    def __enter__(self):
        self.saved = dict(SWITCHES)
        SWITCHES.update(self.kw)
        return self

    def __exit__(self, *exc):
        SWITCHES.clear()
        SWITCHES.update(self.saved)


# position status codes shared with the CUDA path (include/yanocomp_b200.h)
ST_NOT_A_WINDOW = 0    # fails the depth / window filters (gmmtest.py:24-61)
ST_KS_FILTERED = 1     # entered position_stats, failed the gate (stats.py:465)
ST_TESTED = 2          # fitted and tested (was_tested=True)
ST_FIT_FAILED = 3      # LinAlgError path (stats.py:478-479)
ST_DEGENERATE = 4      # A9
ST_KS_ONLY = 5         # --no-gmm (stats.py:460-462)


# ---------------------------------------------------------------------------
# counter-based random numbers (shared definition with csrc/yc_rng.cuh)
# ---------------------------------------------------------------------------
_M64 = (1 << 64) - 1
STREAM_KS_CNTRL = 0
STREAM_KS_TREAT = 1
STREAM_FIT_SAMPLE0 = 16     # + sample index
STREAM_KMEANS0 = 64         # + restart index


def mix64(z):
    '''splitmix64 output function on (z + golden gamma)'''
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def position_key(seed, feature_key, pos):
    h = mix64(int(seed) & _M64)
    h = mix64(h ^ (int(feature_key) & _M64))
    return mix64(h ^ (int(pos) & _M64))


def rng_u64(pkey, stream, counter):
    return mix64(pkey ^ ((int(stream) << 32) | int(counter)))


def rng_uniform(pkey, stream, counter):
    return (rng_u64(pkey, stream, counter) >> 11) * (2.0 ** -53)


def feature_key(feature_id):
    '''FNV-1a 64 of the utf-8 feature id'''
    h = 0xCBF29CE484222325
    for b in str(feature_id).encode('utf-8'):
        h = ((h ^ b) * 0x100000001B3) & _M64
    return h


def _mix64_vec(z):
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def subsample_index(rows, max_size, pkey, stream):
    '''
    Counter-based replacement for stats.py:52-59 (`subsample`).  `rows` are
    the packed read indices (row numbers on the feature's read axis, control
    reads first) of the candidate rows; the `max_size` rows with the smallest
    (hash32(row), row) keys are kept, in their original order.  Returns None
    when nothing is dropped.
    '''
    n = len(rows)
    if n <= max_size:
        return None
    ctr = np.asarray(rows, dtype=np.uint64)
    with np.errstate(over='ignore'):
        h = _mix64_vec(np.uint64(pkey) ^ ((np.uint64(stream) << np.uint64(32)) | ctr))
    keys = ((h >> np.uint64(32)) << np.uint64(32)) | ctr
    kth = np.partition(keys, max_size - 1)[max_size - 1]
    return np.flatnonzero(keys <= kth)


# ---------------------------------------------------------------------------
# KS stage  (stats.py:20-73)
# ---------------------------------------------------------------------------
def pca_comp1(X):
    '''stats.py:20-31'''
    mu = X.mean(axis=0)
    sig = np.sqrt(((X - mu) ** 2.0).mean(0))
    Z = (X - mu) / sig
    _, s, vt = np.linalg.svd(Z, full_matrices=False)
    v = vt.T[:, np.argmax(s ** 2.0), np.newaxis]
    return Z.dot(v).ravel()


def ks_d(data1, data2, pooled):
    '''stats.py:34-43'''
    n1, n2 = data1.shape[0], data2.shape[0]
    s1, s2 = np.sort(data1), np.sort(data2)
    cdf1 = np.searchsorted(s1, pooled, side='right') / (1.0 * n1)
    cdf2 = np.searchsorted(s2, pooled, side='right') / (1.0 * n2)
    d = np.max(np.absolute(cdf1 - cdf2))
    en = np.sqrt(n1 * n2 / float(n1 + n2))
    return d, en


def ks_pvalue(d, en):
    '''stats.py:46-49'''
    return sstats.kstwobign.sf((en + 0.12 + 0.11 / en) * d)


def pca_kstest(cntrl, treat, max_size, pkey, cntrl_rows=None, treat_rows=None):
    '''
    stats.py:62-73.  Returns (status, ks, p).  A9: constant columns are
    reported instead of raising.
    '''
    if cntrl_rows is None:
        cntrl_rows = np.arange(len(cntrl))
    if treat_rows is None:
        treat_rows = len(cntrl) + np.arange(len(treat))
    ic = subsample_index(cntrl_rows, max_size, pkey, STREAM_KS_CNTRL)
    it = subsample_index(treat_rows, max_size, pkey, STREAM_KS_TREAT)
    if ic is not None:
        cntrl = cntrl[ic]
    if it is not None:
        treat = treat[it]
    n_c = len(cntrl)
    pooled = np.concatenate([cntrl, treat])
    if n_c == 0 or len(treat) == 0 or np.any(pooled.max(0) == pooled.min(0)):
        return ST_DEGENERATE, np.nan, np.nan
    comps = pca_comp1(pooled)
    d, en = ks_d(comps[:n_c], comps[n_c:], comps)
    return ST_KS_FILTERED, d, ks_pvalue(d, en)


# ---------------------------------------------------------------------------
# pomegranate stand-ins (a9 in SURVEY.md section 8)
# ---------------------------------------------------------------------------
class FitError(np.linalg.LinAlgError):
    pass


def cholesky_lower(a):
    '''unblocked Cholesky-Crout, row by row; A4 failure rule'''
    d = a.shape[0]
    L = np.zeros((d, d))
    for i in range(d):
        for j in range(i + 1):
            s = a[i, j]
            for k in range(j):
                s -= L[i, k] * L[j, k]
            if i == j:
                if not (s > 0.0) or not math.isfinite(s):
                    raise FitError('covariance is not positive definite')
                L[i, i] = math.sqrt(s)
            else:
                L[i, j] = s / L[j, j]
    return L


def lower_inverse(L):
    d = L.shape[0]
    Li = np.zeros((d, d))
    for c in range(d):
        Li[c, c] = 1.0 / L[c, c]
        for r in range(c + 1, d):
            s = 0.0
            for k in range(c, r):
                s -= L[r, k] * Li[k, c]
            Li[r, c] = s / L[r, r]
    return Li


_LOG_2PI = math.log(2.0 * math.pi)


class MVNormal:
    '''A3/A4: pm.MultivariateGaussianDistribution'''

    def __init__(self, mu, cov):
        self.mu = np.array(mu, dtype=np.float64)
        self.cov = np.array(cov, dtype=np.float64)
        self.d = self.mu.shape[0]
        self._factor()
        self.clear()

    @classmethod
    def from_samples(cls, X):
        '''blank N(0, I) then one fit at inertia 0 (stats.py:150)'''
        d = X.shape[1]
        if 0 < len(X) <= d and SWITCHES['A4_rank_deficient_init'] == 'fail':
            # 1..d rows: the ML covariance is exactly singular; whether LAPACK's pivot is +1e-13 or -1e-13 is noise
            raise FitError('rank-deficient initial cluster')
        dist = cls(np.zeros(d), np.eye(d))
        dist.summarize(X, np.ones(len(X)))
        dist.from_summaries(0.0)
        return dist

    def _factor(self):
        try:
            L = cholesky_lower(self.cov)
        except FitError:
            mode = SWITCHES['A4_cholesky_failure']
            if mode == 'drop' or not np.isfinite(self.cov).all():
                raise
            # alternatives to A4: regularise the spectrum and retry once
            ev, vec = np.linalg.eigh((self.cov + self.cov.T) / 2)
            if mode == 'eig_floor':
                self.cov = (vec * np.maximum(ev, 1e-5)).dot(vec.T)
            else:
                self.cov = self.cov + (1e-5 - min(float(ev.min()), 0.0)) * np.eye(self.d)
            L = cholesky_lower(self.cov)
        self.linv = lower_inverse(L)
        self.logdet = 2.0 * float(np.sum(np.log(np.diagonal(L))))

    def clear(self):
        self.s0 = 0.0
        self.s1 = np.zeros(self.d)
        self.s2 = np.zeros((self.d, self.d))

    def log_probability(self, X):
        y = (X - self.mu).dot(self.linv.T)
        return -0.5 * (self.d * _LOG_2PI + np.sum(y * y, axis=1)) - 0.5 * self.logdet

    def summarize(self, X, w):
        self.s0 += float(np.sum(w))
        self.s1 += w.dot(X)
        self.s2 += (X * w[:, np.newaxis]).T.dot(X)

    def from_summaries(self, inertia):
        if self.s0 < 1e-7:
            return
        u = self.s1 / self.s0
        if SWITCHES['A3_cov_centre'] == 'blended_mean':
            c = self.mu * inertia + u * (1.0 - inertia)
            cov = (self.s2 - np.outer(self.s1, c) - np.outer(c, self.s1)
                   + self.s0 * np.outer(c, c)) / self.s0
        else:
            cov = (self.s2 - np.outer(self.s1, u) - np.outer(u, self.s1)
                   + self.s0 * np.outer(u, u)) / self.s0
        self.mu = self.mu * inertia + u * (1.0 - inertia)
        self.cov = self.cov * inertia + cov * (1.0 - inertia)
        self._factor()
        self.clear()


class DiagNormal:
    '''A10: pm.IndependentComponentsDistribution([pm.NormalDistribution(mu_j, sigma_j)])  (stats.py:153-166)'''

    def __init__(self, mu, sigma):
        self.mu = np.array(mu, dtype=np.float64)
        self.sigma = np.array(sigma, dtype=np.float64)
        self.d = self.mu.shape[0]
        self.clear()

    @classmethod
    def from_samples(cls, X):
        '''stats.py:155-166: N(0, 1) per column for an empty cluster, else column mean / population std'''
        d = X.shape[1]
        if len(X) == 0:
            return cls(np.zeros(d), np.ones(d))
        return cls(X.mean(axis=0), X.std(axis=0))

    @property
    def cov(self):
        '''the evident intent of stats.py:395-412 (np.diag(std ** 2); the reference raises there)'''
        return np.diag(self.sigma ** 2)

    def clear(self):
        self.s0 = np.zeros(self.d)
        self.s1 = np.zeros(self.d)
        self.s2 = np.zeros(self.d)

    def log_probability(self, X):
        with np.errstate(divide='ignore', invalid='ignore'):
            log_norm = -np.log(self.sigma * math.sqrt(2.0 * math.pi))
            inv2 = np.where(self.sigma > 0, 1.0 / (2.0 * self.sigma ** 2), 0.0)
            return np.sum(log_norm - (X - self.mu) ** 2 * inv2, axis=1)

    def summarize(self, X, w):
        self.s0 += float(np.sum(w))
        self.s1 += w.dot(X)
        self.s2 += w.dot(X * X)

    def from_summaries(self, inertia):
        if self.s0[0] < 1e-8:
            return
        mu = self.s1 / self.s0
        with np.errstate(invalid='ignore'):
            sigma = np.sqrt(self.s2 / self.s0 - self.s1 ** 2.0 / self.s0 ** 2.0)
        self.mu = self.mu * inertia + mu * (1.0 - inertia)
        self.sigma = self.sigma * inertia + sigma * (1.0 - inertia)
        self.clear()


class UniformBox:
    '''A5: IndependentComponentsDistribution of UniformDistributions'''

    def __init__(self, lo, hi):
        self.lo = np.array(lo, dtype=np.float64)
        self.hi = np.array(hi, dtype=np.float64)
        width = self.hi - self.lo
        if np.any(width <= 0):
            if SWITCHES['A5_zero_width_box'] != 'widen':
                raise FitError('zero-width uniform component')
            width = np.where(width > 0, width, 1.0)
        self.logp = -float(np.sum(np.log(width)))

    def log_probability(self, X):
        inside = np.all((X >= self.lo) & (X <= self.hi), axis=1)
        return np.where(inside, self.logp, -np.inf)

    def summarize(self, X, w):
        pass

    def from_summaries(self, inertia):
        pass


def _pair_lse(x, y):
    '''A6: pomegranate's pairwise log-sum-exp, vectorised'''
    hi = np.maximum(x, y)
    lo = np.minimum(x, y)
    with np.errstate(invalid='ignore'):
        out = hi + np.log(np.exp(lo - hi) + 1.0)
    return np.where(np.isneginf(hi), -np.inf, out)


class Mixture:
    '''A6: one pm.GeneralMixtureModel sharing `dists` with its siblings'''

    def __init__(self, dists, weights):
        w = np.asarray(weights, dtype=np.float64)
        self.dists = dists
        with np.errstate(divide='ignore'):
            self.logw = np.log(w / w.sum())                       # A7
        self.summaries = np.zeros(len(dists))

    def _joint(self, X):
        return np.stack([d.log_probability(X) for d in self.dists]) + self.logw[:, None]

    def summarize(self, X):
        joint = self._joint(X)
        total = np.full(X.shape[0], -np.inf)
        for j in range(len(self.dists)):
            total = _pair_lse(total, joint[j])
        resp = np.exp(joint - total)
        for j, d in enumerate(self.dists):
            d.summarize(X, resp[j])
        self.summaries += resp.sum(axis=1)
        return float(np.sum(total))

    def predict_proba(self, X):
        joint = self._joint(X)
        total = np.full(X.shape[0], -np.inf)
        for j in range(len(self.dists)):
            total = _pair_lse(total, joint[j])
        return np.exp(joint - total).T


def _sqdist(X, c):
    diff = X[:, np.newaxis, :] - c[np.newaxis, :, :]
    return np.sum(diff * diff, axis=2)


def kmeans_fit(X, pkey, k=N_COMPONENTS, n_init=3, max_iterations=100,
               stop_threshold=0.5):
    '''A2: pm.Kmeans(k, init='kmeans++', n_init=3).fit (stats.py:108-113)'''
    n = X.shape[0]
    best_c, best_total = None, np.inf
    n_init = min(n_init, SWITCHES['A2_kmeans_n_init'])
    stop_scale = float(n) if SWITCHES['A2_kmeans_stop'] == 'mean' else 1.0
    for restart in range(n_init):
        stream = STREAM_KMEANS0 + restart
        i0 = min(int(rng_uniform(pkey, stream, 0) * n), n - 1)
        c = np.empty((k, X.shape[1]))
        c[0] = X[i0]
        mind = np.full(n, np.inf)
        for m in range(1, k):
            mind = np.minimum(mind, _sqdist(X, c[m - 1:m])[:, 0])
            cdf = np.cumsum(mind)
            u = rng_uniform(pkey, stream, m) * cdf[-1]
            c[m] = X[min(int(np.searchsorted(cdf, u, side='right')), n - 1)]
        iteration, improvement = 0, np.inf
        sums = counts = None
        last = total = np.inf
        while improvement > stop_threshold and iteration < max_iterations + 1:
            if sums is not None:
                for j in range(k):
                    if counts[j] > 0:
                        c[j] = sums[j] / counts[j]
            d2 = _sqdist(X, c)
            lab = np.argmin(d2, axis=1)
            total = float(np.sum(d2[np.arange(n), lab]))
            sums = np.stack([X[lab == j].sum(axis=0) for j in range(k)])
            counts = np.bincount(lab, minlength=k)
            if iteration > 0:
                improvement = (last - total) / stop_scale
            iteration += 1
            last = total
        if total < best_total:
            best_c, best_total = c.copy(), total
    if best_c is None:        # every restart ended non-finite
        raise FitError('k-means failed')
    return best_c


def median_absolute_deviation(arr):
    '''stats.py:76-78'''
    return np.median(np.abs(arr - np.median(arr)))


def dbscan_outliers(X, eps=5.0, min_samples=None):
    '''A11: DBSCAN(eps, min_samples='auto' -> 0.25 n).fit_predict(X) == -1  (stats.py:91-99)'''
    n = len(X)
    if min_samples is None:
        min_samples = 0.25 * n
    near = _sqdist(X, X) <= eps * eps                      # exact for binary16 inputs; sqrt is monotone
    core = near.sum(axis=1) >= min_samples
    return ~core & ~near[:, core].any(axis=1)


def kmeans_init_clusters(X, pkey, outlier_factor=0.5, detect_outliers='mad', dbscan_eps=5.0):
    '''stats.py:81-124; detect_outliers in (None, 'mad', 'dbscan') as position_stats passes it (stats.py:472)'''
    if detect_outliers is True:
        detect_outliers = 'mad'
    elif detect_outliers is False:
        detect_outliers = None
    outlier_mask = None
    X_fit = X
    if detect_outliers == 'dbscan':
        outlier_mask = dbscan_outliers(X, dbscan_eps)
        if outlier_mask.all():
            return np.full(len(X), N_COMPONENTS, dtype=np.int64)
        X_fit = X[~outlier_mask]                           # stats.py:104
    centroids = kmeans_fit(X_fit, pkey)
    d2 = _sqdist(X, centroids)                                     # A1
    init_pred = np.argmin(d2, axis=1)
    if detect_outliers is None:
        return init_pred
    if detect_outliers == 'mad':
        dists = np.min(d2, axis=1)
        if SWITCHES['A1_kmeans_distance'] == 'euclidean':
            dists = np.sqrt(dists)
        mad = median_absolute_deviation(dists)
        outlier_mask = dists > outlier_factor * mad
    init_pred[outlier_mask] = N_COMPONENTS
    return init_pred


def initialise_gmms(samples, pkey, outlier_factor=0.5, pseudocount=0.5, add_uniform=True,
                    covariance='full', outlier_method='mad', dbscan_eps=5.0):
    '''stats.py:127-191: full / diag covariance; MAD or DBSCAN outliers + uniform component unless --no-uniform'''
    sizes = [len(s) for s in samples]
    pooled = np.concatenate(samples)
    init_pred = kmeans_init_clusters(pooled, pkey, outlier_factor,
                                     outlier_method if add_uniform else None, dbscan_eps)
    if covariance == 'full':
        dists = [MVNormal.from_samples(pooled[init_pred == i]) for i in range(N_COMPONENTS)]
    elif covariance == 'diag':
        dists = [DiagNormal.from_samples(pooled[init_pred == i]) for i in range(N_COMPONENTS)]
    else:
        raise ValueError('Only full and diag covariances are supported')
    ncomp = N_COMPONENTS + int(add_uniform)
    per_samp = np.array_split(init_pred, np.cumsum(sizes)[:-1])
    weights = [(np.bincount(p, minlength=ncomp) + pseudocount)
               / (len(p) + pseudocount * ncomp) for p in per_samp]
    if add_uniform:
        outliers = pooled[init_pred == N_COMPONENTS]
        if len(outliers) == 0:
            outliers = pooled
        dists.append(UniformBox(outliers.min(axis=0), outliers.max(axis=0)))
    return [Mixture(dists, w) for w in weights], dists, init_pred


def em(samples, gmms, dists, max_iterations=1e3, stop_threshold=0.1,
       inertia=0.01, lr_decay=1e-3, pseudocount=0.5, trace=None):
    '''stats.py:194-225, control flow kept exactly'''
    iteration, improvement = 0, np.inf
    last = None
    while improvement > stop_threshold and iteration < max_iterations + 1:
        step = 1 - ((1 - inertia) * (2 + iteration) ** -lr_decay)
        if iteration:
            for d in dists:
                d.from_summaries(step)
            for g in gmms:
                if g.summaries.sum():
                    s = g.summaries + pseudocount
                    s /= s.sum()
                    g.logw[:] = np.log(s)
                    g.summaries[:] = 0
        ll = 0.0
        for g, samp in zip(gmms, samples):
            ll += g.summarize(samp)
        if iteration:
            improvement = ll - last
        iteration += 1
        last = ll
        if trace is not None:
            trace.append(ll)
    return iteration


@dataclasses.dataclass
class GMMFit:
    dists: list
    combined_weights: np.ndarray
    per_samp_weights: np.ndarray
    cntrl_preds: np.ndarray
    treat_preds: np.ndarray
    iterations: int
    init_pred: np.ndarray
    ll_trace: list


def fit_gmm(cntrl, treat, pkey, outlier_factor=0.5, max_fit_depth=1000,
            sample_ids=None, sample_rows=None, add_uniform=True,
            covariance='full', outlier_method='mad', dbscan_eps=5.0):
    '''
    stats.py:228-299 (`full`, `mad`); cntrl / treat are lists of (n_s, d).
    sample_ids[s] = packed sample index (selects the sub-sampling stream),
    sample_rows[s] = packed read indices of that sample's rows.
    '''
    n_cntrl = len(cntrl)
    all_samples = list(cntrl) + list(treat)
    orig_sizes = np.array([len(s) for s in all_samples])
    if sample_ids is None:
        sample_ids = list(range(len(all_samples)))
    if sample_rows is None:
        ends = np.cumsum(orig_sizes)
        sample_rows = [np.arange(e - n, e) for e, n in zip(ends, orig_sizes)]
    samples = []
    for sid, rows, samp in zip(sample_ids, sample_rows, all_samples):
        idx = subsample_index(rows, max_fit_depth, pkey, STREAM_FIT_SAMPLE0 + sid)
        samples.append(samp if idx is None else samp[idx])
    trace = []
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        gmms, dists, init_pred = initialise_gmms(samples, pkey, outlier_factor, add_uniform=add_uniform,
                                                 covariance=covariance, outlier_method=outlier_method,
                                                 dbscan_eps=dbscan_eps)
        iters = em(samples, gmms, dists, trace=trace)
    if not all(np.isfinite(trace)):
        raise FitError('non-finite likelihood')                       # A5
    fit_sizes = [len(s) for s in samples]
    per_samp = np.array([np.exp(g.logw) for g in gmms])
    combined = np.average(per_samp, axis=0, weights=fit_sizes)
    preds = np.round(per_samp * orig_sizes[:, np.newaxis])
    return GMMFit(dists, combined, per_samp, preds[:n_cntrl], preds[n_cntrl:],
                  iters, init_pred, trace)


# ---------------------------------------------------------------------------
# G-tests and fractions (stats.py:302-355)
# ---------------------------------------------------------------------------
def two_cond_g_test(counts):
    '''stats.py:302-313'''
    try:
        g, p, *_ = sstats.chi2_contingency(counts, lambda_='log-likelihood')
        return g, p
    except ValueError:
        return 0.0, 1.0


def gmm_g_test(cntrl_preds, treat_preds, p_val_threshold=0.05):
    '''stats.py:316-338'''
    with np.errstate(invalid='raise'):
        try:
            het_g, p_val = two_cond_g_test([
                cntrl_preds[:, :2].sum(0), treat_preds[:, :2].sum(0)])
        except FloatingPointError:
            het_g, p_val = np.nan, 1
    if p_val < p_val_threshold:
        hom_g = two_cond_g_test(cntrl_preds)[0] + two_cond_g_test(treat_preds)[0]
        if hom_g >= het_g:
            p_val = 1
    else:
        hom_g = np.nan
    return het_g, hom_g, p_val


def log_odds_ci(table, ci=95):
    '''A8: statsmodels Table2x2(table, shift_zeros=True) (stats.py:352-354)'''
    f = np.asarray(table, dtype=np.float64).ravel()
    if np.any(f == 0):
        f = f + 0.5
    lor = float(np.dot(np.log(f), np.r_[1, -1, -1, 1]))
    se = math.sqrt(float(np.sum(1.0 / f)))
    z = -sstats.norm.ppf((1 - ci / 100) / 2)
    return lor, lor - z * se, lor + z * se


def calculate_fractional_stats(cntrl_preds, treat_preds):
    '''stats.py:341-355'''
    c = cntrl_preds.sum(0)
    t = treat_preds.sum(0)
    with np.errstate(invalid='ignore', divide='ignore'):
        c_frac = c[1] / c[:N_COMPONENTS].sum()
        t_frac = t[1] / t[:N_COMPONENTS].sum()
    lor, lo, hi = log_odds_ci([c[:N_COMPONENTS], t[:N_COMPONENTS]])
    return c_frac, t_frac, lor, lo, hi


# ---------------------------------------------------------------------------
# position_stats (stats.py:423-513)
# ---------------------------------------------------------------------------
@dataclasses.dataclass
class Options:
    '''the subset of GMMTestOpts (gmmtest.py:255-316) the hot path reads'''
    window_size: int = 3
    min_read_depth: int = None
    max_fit_depth: int = 1000
    min_ks: float = 0.2
    fdr_threshold: float = 0.05
    outlier_factor: float = 0.5
    add_uniform: bool = True
    covariance_type: str = 'full'
    outlier_method: str = 'mad'
    dbscan_eps: float = 5.0
    gmm: bool = True
    generate_sm_preds: bool = False
    test_level: str = 'gene'
    random_seed: int = 0

    def __post_init__(self):
        if self.min_read_depth is None:
            self.min_read_depth = max(2 * self.window_size, 5)   # gmmtest.py:245-251


@dataclasses.dataclass
class PositionResult:
    '''GMMTestResults (stats.py:423-442) + bookkeeping'''
    status: int
    kmers: list
    centre: int
    ks_stat: float = np.nan
    ks_p_val: float = np.nan
    log_odds: float = np.nan
    log_odds_lower_ci: float = np.nan
    log_odds_upper_ci: float = np.nan
    p_val: float = 1.0
    fdr: float = 1.0
    cntrl_frac_mod: float = np.nan
    treat_frac_mod: float = np.nan
    g_stat: float = np.nan
    hom_g_stat: float = np.nan
    dist1_means: np.ndarray = None
    dist1_stds: np.ndarray = None
    dist2_means: np.ndarray = None
    dist2_stds: np.ndarray = None
    fit: GMMFit = None
    sm_preds: dict = None

    @property
    def kmer(self):
        return self.kmers[self.centre]

    @property
    def was_tested(self):
        return self.status in (ST_TESTED, ST_KS_ONLY)


@dataclasses.dataclass
class CondWindow:
    '''complete-window rows of one condition at one position (gmmtest.py:51-54)'''
    X: np.ndarray          # (n, w) float64
    rep: np.ndarray        # 1-based replicate number per row
    rows: np.ndarray       # packed read index per row (control reads first)
    read_ids: list = None


def _split_by_replicate(cw):
    '''groupby('replicate', sort=False): groups in order of first appearance'''
    order = []
    for r in cw.rep.tolist():
        if r not in order:
            order.append(r)
    return order, [cw.X[cw.rep == r] for r in order], [cw.rows[cw.rep == r] for r in order]


def format_model(fit):
    '''stats.py:382-419'''
    w = fit.combined_weights / fit.combined_weights.sum()            # A7
    model = {
        'unmod': {'mu': fit.dists[0].mu.tolist(), 'cov': fit.dists[0].cov.tolist()},
        'mod': {'mu': fit.dists[1].mu.tolist(), 'cov': fit.dists[1].cov.tolist()},
    }
    if len(fit.dists) == N_COMPONENTS + 1:
        model['outlier'] = {'bounds': [[float(a), float(b)] for a, b in
                                       zip(fit.dists[2].lo, fit.dists[2].hi)]}
    model['weights'] = np.exp(np.log(w)).tolist()
    return model


def format_sm_preds(prob, outlier, cw):
    '''stats.py:358-379'''
    out = {}
    ids = cw.read_ids if cw.read_ids is not None else [''] * len(cw.X)
    for r_id, r, ev, p, o in zip(ids, cw.rep.tolist(), cw.X.tolist(),
                                 prob.tolist(), outlier.tolist()):
        entry = out.setdefault(r, {'read_ids': [], 'events': [], 'preds': [],
                                   'outlier_preds': []})
        entry['read_ids'].append(r_id)
        entry['events'].append(ev)
        entry['preds'].append(p)
        entry['outlier_preds'].append(o)
    return out


def position_stats(cntrl, treat, kmers, opts, pkey, n_cntrl_files=None):
    '''
    stats.py:445-513.  cntrl / treat are CondWindow; n_cntrl_files fixes the
    packed sample index of treatment replicates (Rc + rep - 1).
    '''
    w = len(kmers)
    r = PositionResult(status=ST_KS_FILTERED, kmers=list(kmers), centre=w // 2)
    status, r.ks_stat, r.ks_p_val = pca_kstest(
        cntrl.X, treat.X, opts.max_fit_depth, pkey, cntrl.rows, treat.rows)
    if status == ST_DEGENERATE:
        r.status = ST_DEGENERATE
        return r
    if not opts.gmm:
        r.p_val = r.ks_p_val
        r.status = ST_KS_ONLY
        return r
    if not (r.ks_stat >= opts.min_ks and r.ks_p_val < opts.fdr_threshold):
        return r
    c_reps, cntrl_fit, c_rows = _split_by_replicate(cntrl)
    t_reps, treat_fit, t_rows = _split_by_replicate(treat)
    if n_cntrl_files is None:
        n_cntrl_files = max(c_reps)
    sample_ids = [rep - 1 for rep in c_reps] + [n_cntrl_files + rep - 1 for rep in t_reps]
    try:
        fit = fit_gmm(cntrl_fit, treat_fit, pkey, opts.outlier_factor,
                      opts.max_fit_depth, sample_ids, c_rows + t_rows, opts.add_uniform,
                      opts.covariance_type, opts.outlier_method, opts.dbscan_eps)
    except np.linalg.LinAlgError:
        r.status = ST_FIT_FAILED
        return r
    r.status = ST_TESTED
    r.fit = fit
    r.dist1_means, r.dist2_means = fit.dists[0].mu, fit.dists[1].mu
    # stats.py:277-291: sqrt of the covariance diagonal (full) / the NormalDistribution sigmas (diag)
    r.dist1_stds, r.dist2_stds = [d.sigma if isinstance(d, DiagNormal) else np.sqrt(np.diagonal(d.cov))
                                  for d in fit.dists[:2]]
    r.g_stat, r.hom_g_stat, r.p_val = gmm_g_test(
        fit.cntrl_preds, fit.treat_preds, opts.fdr_threshold)
    (r.cntrl_frac_mod, r.treat_frac_mod, r.log_odds,
     r.log_odds_lower_ci, r.log_odds_upper_ci) = calculate_fractional_stats(
         fit.cntrl_preds, fit.treat_preds)
    if r.p_val < opts.fdr_threshold and opts.generate_sm_preds:
        gmm = Mixture(fit.dists, fit.combined_weights)
        cp = gmm.predict_proba(np.concatenate(cntrl_fit))[:, 1:].T
        tp = gmm.predict_proba(np.concatenate(treat_fit))[:, 1:].T
        if not opts.add_uniform:
            # stats.py:500-502 sets the outlier predictions to zeros_like of a (1, n) array and then
            # zips it with the reads, which keeps a single entry; the evident intent (one prediction
            # per read, outlier probability 0) is restated here instead of that bug.
            cp = np.vstack([cp[0], np.zeros_like(cp[0])])
            tp = np.vstack([tp[0], np.zeros_like(tp[0])])
        r.sm_preds = {
            'kmers': list(kmers),
            'model': format_model(fit),
            'cntrl': format_sm_preds(cp[0], cp[1], cntrl),
            'treat': format_sm_preds(tp[0], tp[1], treat),
        }
    return r


# ---------------------------------------------------------------------------
# ingest semantics (io.py:296-343) and position enumeration (gmmtest.py:24-117)
# ---------------------------------------------------------------------------
@dataclasses.dataclass
class CondEvents:
    '''read x position matrix of one condition (the concat of io.py:337-341)'''
    X: np.ndarray            # (reads, positions) float64, NaN = no event
    positions: np.ndarray    # sorted int64
    rep: np.ndarray          # 1-based replicate number per read row
    read_ids: list


def events_to_matrix(events, read_ids, rep):
    '''
    One file's part of load_gene_events (io.py:306-332) for one feature:
    keep-first dedupe on (read_idx, pos), f16 -> f64, drop dwell > 1 s
    (log10(duration) <= 0 kept; NaN / negative durations drop), rows sorted by
    read id string as `unstack` does.
    '''
    key = events['read_idx'].astype(np.int64) << 32 | events['pos'].astype(np.int64)
    _, first = np.unique(key, return_index=True)
    e = events[np.sort(first)]
    dur = e['duration'].astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        keep = np.log10(dur) <= 0
    e = e[keep]
    ids = np.asarray(read_ids).astype('U32')[e['read_idx']]
    uniq_ids, row = np.unique(ids, return_inverse=True)
    positions, col = np.unique(e['pos'].astype(np.int64), return_inverse=True)
    X = np.full((len(uniq_ids), len(positions)), np.nan)
    X[row, col] = e['mean'].astype(np.float64)
    return CondEvents(X, positions, np.full(len(uniq_ids), rep, dtype=np.int64),
                      uniq_ids.tolist())


def concat_replicates(parts):
    positions = np.unique(np.concatenate([p.positions for p in parts])) \
        if parts else np.zeros(0, np.int64)
    n = sum(len(p.rep) for p in parts)
    X = np.full((n, len(positions)), np.nan)
    r0 = 0
    for p in parts:
        X[r0:r0 + len(p.rep), np.searchsorted(positions, p.positions)] = p.X
        r0 += len(p.rep)
    return CondEvents(
        X, positions,
        np.concatenate([p.rep for p in parts]) if parts else np.zeros(0, np.int64),
        [i for p in parts for i in p.read_ids])


def load_gene_events(gene_id, files, by_transcript=False):
    '''io.py:296-343 on h5lite/h5py-like handles'''
    feats = {}
    for rep, f in enumerate(files, 1):
        g = f[gene_id]
        events = g['events'][:]
        read_ids = g['read_ids'][:]
        if by_transcript:
            t_ids = [str(t) for t in g['transcript_ids'][:]]
            # dedupe / dwell filter happen before the transcript split
            for t_idx, t_id in enumerate(t_ids):
                m = events_to_matrix_masked(events, read_ids, rep, t_idx)
                if m is not None:
                    feats.setdefault(t_id, []).append(m)
        else:
            feats.setdefault(gene_id, []).append(events_to_matrix(events, read_ids, rep))
    return {k: concat_replicates(v) for k, v in feats.items()}


def events_to_matrix_masked(events, read_ids, rep, t_idx):
    key = events['read_idx'].astype(np.int64) << 32 | events['pos'].astype(np.int64)
    _, first = np.unique(key, return_index=True)
    e = events[np.sort(first)]
    e = e[e['transcript_idx'] == t_idx]
    if len(e) == 0:
        return None
    # already de-duplicated: events_to_matrix's dedupe is now a no-op
    return events_to_matrix(e, read_ids, rep)


def load_gene_kmers(gene_id, files):
    '''io.py:270-281'''
    kmers = {}
    for f in files:
        k = f[gene_id]['kmers'][:]
        for pos, kmer in zip(k['pos'].tolist(), k['kmer'].tolist()):
            kmers[pos] = kmer.decode() if isinstance(kmer, bytes) else str(kmer)
    return kmers


def get_valid_pos(ev, min_read_depth):
    '''gmmtest.py:24-30'''
    ok = np.ones(len(ev.positions), dtype=bool)
    for r in np.unique(ev.rep):
        ok &= (~np.isnan(ev.X[ev.rep == r])).sum(axis=0) >= min_read_depth
    return set(ev.positions[ok].tolist())


def _depth_ok(rep, min_read_depth):
    '''gmmtest.py:57-61: groupby().size() only lists replicates with rows'''
    return all((rep == r).sum() >= min_read_depth for r in np.unique(rep))


def create_positional_data(cntrl, treat, kmers, reverse, min_read_depth, window_size):
    '''gmmtest.py:33-78; yields (pos, kmers, CondWindow, CondWindow) by ascending pos'''
    valid = get_valid_pos(cntrl, min_read_depth) & get_valid_pos(treat, min_read_depth)
    w = window_size // 2
    n_c_rows = len(cntrl.rep)
    for pos in sorted(valid):
        win = np.arange(pos - w, pos + w + 1)
        if reverse:
            win = win[::-1]
        if not valid.issuperset(win.tolist()):
            continue
        cx = cntrl.X[:, np.searchsorted(cntrl.positions, win)]
        tx = treat.X[:, np.searchsorted(treat.positions, win)]
        ck = ~np.isnan(cx).any(axis=1)
        tk = ~np.isnan(tx).any(axis=1)
        if _depth_ok(cntrl.rep[ck], min_read_depth) and _depth_ok(treat.rep[tk], min_read_depth):
            yield (pos, [kmers[p] for p in win.tolist()],
                   CondWindow(cx[ck], cntrl.rep[ck], np.flatnonzero(ck),
                              [i for i, k in zip(cntrl.read_ids, ck) if k]),
                   CondWindow(tx[tk], treat.rep[tk], n_c_rows + np.flatnonzero(tk),
                              [i for i, k in zip(treat.read_ids, tk) if k]))


def get_shared_keys(files):
    '''io.py:259-267'''
    genes = set(files[0].keys())
    for f in files[1:]:
        genes &= set(f.keys())
    return sorted(genes)


def load_gene_attrs(gene_id, files):
    '''io.py:284-293'''
    a = files[0][gene_id].attrs
    return str(a['chrom']), str(a['strand'])


def benjamini_hochberg(p):
    '''statsmodels multipletests(method='fdr_bh') (gmmtest.py:223)'''
    p = np.asarray(p, dtype=np.float64)
    n = len(p)
    order = np.argsort(p)
    raw = p[order] / (np.arange(1, n + 1) / float(n))
    adj = np.minimum.accumulate(raw[::-1])[::-1]
    adj[adj > 1] = 1
    out = np.empty(n)
    out[order] = adj
    return out


def test_features(cntrl_files, treat_files, opts, gene_ids=None):
    '''
    test_chunk + parallel_test (gmmtest.py:139-224) over open file handles.
    Returns (records, all_results): records = tested positions as dicts with
    the DataFrame columns of gmmtest.py:216-220 (fdr filled by BH over rows
    with a non-NaN p_val); all_results = every PositionResult keyed by
    (feature_id, pos), including filtered / failed ones.
    '''
    if gene_ids is None:
        gene_ids = get_shared_keys(cntrl_files + treat_files)
    by_t = opts.test_level == 'transcript'
    records, everything = [], {}
    for gene_id in gene_ids:
        chrom, strand = load_gene_attrs(gene_id, cntrl_files)
        c_feats = load_gene_events(gene_id, cntrl_files, by_t)
        t_feats = load_gene_events(gene_id, treat_files, by_t)
        kmers = load_gene_kmers(gene_id, cntrl_files + treat_files)
        for feat in sorted(set(c_feats) & set(t_feats)):
            fkey = feature_key(feat)
            for (pos, pk, cw, tw) in create_positional_data(
                    c_feats[feat], t_feats[feat], kmers, strand == '-',
                    opts.min_read_depth, opts.window_size):
                res = position_stats(cw, tw, pk, opts,
                                     position_key(opts.random_seed, fkey, pos),
                                     len(cntrl_files))
                everything[(feat, int(pos))] = res
                if res.was_tested:
                    records.append(dict(chrom=chrom, pos=int(pos), gene_id=feat,
                                        strand=strand, result=res))
    ok = [r for r in records if not np.isnan(r['result'].p_val)]
    if ok:
        fdr = benjamini_hochberg([r['result'].p_val for r in ok])
        for r, q in zip(ok, fdr):
            r['result'].fdr = float(q)
    return ok, everything


# ---------------------------------------------------------------------------
# host tail: assign_modified_distribution (stats.py:516-599), writers
# ---------------------------------------------------------------------------
def load_model_priors(model_fn):
    '''io.py:352-362 -> {kmer: level_mean}'''
    means = {}
    with open(model_fn) as fh:
        header = None
        for line in fh:
            if line.startswith('#') or not line.strip():
                continue
            fields = line.rstrip('\n').split('\t')
            if header is None:
                header = fields
                continue
            means[fields[header.index('kmer')]] = float(fields[header.index('level_mean')])
    return means


def assign_modified_distribution(records, model_means):
    '''stats.py:516-599 on the record dicts of test_features (in place)'''
    if not records:
        return records
    lower, upper = {}, {}
    for rec in records:
        r = rec['result']
        for k, lm, um in zip(r.kmers, r.dist1_means, r.dist2_means):
            if lm > um:
                lm, um = um, lm
            exp = model_means[k]
            lower.setdefault(k, []).append(abs(lm - exp))
            upper.setdefault(k, []).append(abs(um - exp))
    shift = {k: 1 if np.median(upper[k]) > np.median(lower[k]) else 0 for k in lower}
    for rec in records:
        r = rec['result']
        s1 = s2 = 0
        for k, d in zip(r.kmers, r.dist1_means - r.dist2_means):
            if shift[k] ^ (d < 0):
                s1 += abs(d)
            else:
                s2 += abs(d)
        if s1 > s2:
            r.cntrl_frac_mod = 1 - r.cntrl_frac_mod
            r.treat_frac_mod = 1 - r.treat_frac_mod
            r.log_odds = -r.log_odds
            # stats.py:577-579 assigns a label-indexed Series through .loc: pandas aligns it on the labels, so
            # each bound is negated in place (the interval comes out inverted); reproduced, not repaired
            r.log_odds_lower_ci, r.log_odds_upper_ci = (-r.log_odds_lower_ci,
                                                        -r.log_odds_upper_ci)
            r.dist1_means, r.dist2_means = r.dist2_means, r.dist1_means
            r.dist1_stds, r.dist2_stds = r.dist2_stds, r.dist1_stds
            if r.sm_preds is not None:
                m = r.sm_preds['model']
                m['unmod'], m['mod'] = m['mod'], m['unmod']
                for cond in ('cntrl', 'treat'):
                    for rep in r.sm_preds[cond].values():
                        rep['preds'] = [1 - (p + o) for p, o in
                                        zip(rep['preds'], rep['outlier_preds'])]
    return records


def bed_line(rec):
    '''io.py:372-398'''
    r = rec['result']
    c = r.centre
    if r.dist1_means is None:
        um = us = mm = ms = np.nan
    else:
        um, us, mm, ms = r.dist1_means[c], r.dist1_stds[c], r.dist2_means[c], r.dist2_stds[c]
    with np.errstate(divide='ignore'):
        score = int(round(min(-np.log10(r.fdr), 100)))
    pos = rec['pos']
    return (f'{rec["chrom"]:s}\t{pos - 2:d}\t{pos + 3:d}\t'
            f'{rec["gene_id"]}:{r.kmer}\t{score:d}\t{rec["strand"]:s}\t'
            f'{r.log_odds:.2f}[{r.log_odds_lower_ci:.2f},{r.log_odds_upper_ci:.2f}]\t'
            f'{r.p_val:.2g}\t{r.fdr:.2g}\t'
            f'{r.cntrl_frac_mod:.2f}\t{r.treat_frac_mod:.2f}\t'
            f'{r.g_stat:.2f}\t{r.hom_g_stat:.2f}\t'
            f'{um:.2f}\t{us:.2f}\t{mm:.2f}\t{ms:.2f}\t'
            f'{r.ks_stat:.2f}\n')


def gmmtest(cntrl_fns, treat_fns, opts, model_fn, output_bed_fn=None,
            output_sm_preds_fn=None, open_file=None):
    '''gmm_test (gmmtest.py:317-362) end to end; returns the BED text'''
    if open_file is None:
        from yanocomp_b200 import h5lite
        open_file = h5lite.File
    cf = [open_file(fn) for fn in cntrl_fns]
    tf = [open_file(fn) for fn in treat_fns]
    try:
        records, _ = test_features(cf, tf, opts)
    finally:
        for f in cf + tf:
            f.close()
    sig = [r for r in records if r['result'].fdr < opts.fdr_threshold]
    if opts.gmm:
        assign_modified_distribution(sig, load_model_priors(model_fn))
    sig.sort(key=lambda r: (r['chrom'], r['pos']))
    bed = ''.join(bed_line(r) for r in sig)
    if output_bed_fn:
        with open(output_bed_fn, 'w') as fh:
            fh.write(bed)
    if output_sm_preds_fn and opts.gmm and opts.generate_sm_preds:
        sm = {}
        for r in sig:
            if r['result'].sm_preds is not None:
                sm.setdefault(r['gene_id'], {})[r['pos']] = r['result'].sm_preds
        doc = {'input_fns': {'cntrl': dict(enumerate(cntrl_fns)),
                             'treat': dict(enumerate(treat_fns))},
               'single_molecule_predictions': sm}
        if os.path.splitext(output_sm_preds_fn)[1] == '.gz':
            fh = gzip.open(output_sm_preds_fn, 'wt', encoding='ascii')
        else:
            fh = open(output_sm_preds_fn, 'w')
        with fh:
            json.dump(doc, fh)
    return bed
