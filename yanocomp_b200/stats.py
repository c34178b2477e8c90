'''
Host tail of gmmtest: multiple-testing correction and the global
modified-distribution assignment.  O(#tested positions) NumPy; mirrors the
reference's post-processing (yanocomp/gmmtest.py:216-242,
yanocomp/stats.py:516-599) on structure-of-arrays results.
'''
import numpy as np


def benjamini_hochberg(p):
    '''
    statsmodels.stats.multitest.multipletests(p, method='fdr_bh')[1]
    (reference yanocomp/gmmtest.py:223): sorted p * n / rank, made monotone
    from the right, clipped at 1, returned in the input order.
    '''
    p = np.asarray(p, dtype=np.float64)
    n = p.shape[0]
    if n == 0:
        return p.copy()
    order = np.argsort(p, kind='stable')
    ecdf = np.arange(1, n + 1, dtype=np.float64) / float(n)
    raw = p[order] / ecdf
    adj = np.minimum.accumulate(raw[::-1])[::-1]
    adj[adj > 1] = 1
    out = np.empty(n, dtype=np.float64)
    out[order] = adj
    return out
