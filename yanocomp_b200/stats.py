'''
Host tail of gmmtest: multiple-testing correction and the global
modified-distribution assignment.  O(#tested positions) NumPy; mirrors the
reference's post-processing (yanocomp/gmmtest.py:216-242,
yanocomp/stats.py:516-599) on structure-of-arrays results.
'''
import numpy as np


def _argsort(p):
    '''
    Ascending order of the p-values (or of unique integer keys).  Tied p-values receive the same adjusted value whatever
    their relative order (the running minimum from the right equalises them), so the sort need not be stable; a large
    family is sorted on the GPU the process already owns (2 M p-values: ~1 ms instead of ~0.2 s of host time per job).
    '''
    if p.shape[0] >= 32768:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.argsort(torch.from_numpy(np.ascontiguousarray(p)).cuda()).cpu().numpy()
        except ImportError:
            pass
    return np.argsort(p, kind='stable')


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
    order = _argsort(p)
    ecdf = np.arange(1, n + 1, dtype=np.float64) / float(n)
    raw = p[order] / ecdf
    adj = np.minimum.accumulate(raw[::-1])[::-1]
    adj[adj > 1] = 1
    out = np.empty(n, dtype=np.float64)
    out[order] = adj
    return out


# ---------------------------------------------------------------------------
# single-molecule prediction dictionaries (reference yanocomp/stats.py:358-419)
# ---------------------------------------------------------------------------
def format_model(mu, cov, box, cw):
    '''JSON-serialisable GMM parameters (reference stats.py:382-419)'''
    w = np.asarray(cw, dtype=np.float64)
    model = {
        'unmod': {'mu': mu[0].tolist(), 'cov': cov[0].tolist()},
        'mod': {'mu': mu[1].tolist(), 'cov': cov[1].tolist()},
    }
    if box is not None:
        model['outlier'] = {'bounds': [[float(lo), float(hi)] for lo, hi in box]}
    else:
        w = w[:2]                                   # --no-uniform: two components
    w = w / w.sum()
    model['weights'] = np.exp(np.log(w)).tolist()
    return model


def format_sm_preds(preds, outlier, events, reps, read_ids):
    '''
    per-replicate read ids / events / predictions (reference stats.py:358-379).  The packed layout
    keeps the reads of a replicate together, so a replicate is normally one slice of each list; reads
    of one replicate scattered over several runs take the reference's read-by-read route.
    '''
    reps = list(reps)
    n = len(reps)
    out = {}
    if n:
        r = np.asarray(reps)
        starts = [0] + (np.flatnonzero(r[1:] != r[:-1]) + 1).tolist() + [n]
        if len({reps[a] for a in starts[:-1]}) == len(starts) - 1:
            for a, b in zip(starts[:-1], starts[1:]):
                out[reps[a]] = {'read_ids': list(read_ids[a:b]), 'events': list(events[a:b]),
                                'preds': list(preds[a:b]), 'outlier_preds': list(outlier[a:b])}
            return out
    for r_id, rep, ev, p, o in zip(read_ids, reps, events, preds, outlier):
        entry = out.get(rep)
        if entry is None:
            entry = out[rep] = {'read_ids': [], 'events': [], 'preds': [], 'outlier_preds': []}
        entry['read_ids'].append(r_id)
        entry['events'].append(ev)
        entry['preds'].append(p)
        entry['outlier_preds'].append(o)
    return out


# ---------------------------------------------------------------------------
# modified-distribution assignment (reference yanocomp/stats.py:516-599)
# ---------------------------------------------------------------------------
def calculate_kmer_shift_directions(results, expected_model):
    '''
    Per kmer: is the upper or the lower fitted component further from the
    expected ONT level, over all significant records (reference stats.py:516-534).
    expected_model: {kmer: (level_mean, level_stdv)}.
    '''
    lower, upper = {}, {}
    for kmers, m1, m2 in zip(results['kmers'], results['dist1_means'], results['dist2_means']):
        for k, lm, um in zip(kmers, m1, m2):
            if lm > um:
                lm, um = um, lm
            exp = expected_model[k][0]
            lower.setdefault(k, []).append(abs(lm - exp))
            upper.setdefault(k, []).append(abs(um - exp))
    return {k: 1 if np.median(upper[k]) > np.median(lower[k]) else 0 for k in lower}


def dist1_is_modified(kmers, dist1_means, dist2_means, kmer_shift_dirs):
    '''reference stats.py:537-554'''
    s1 = s2 = 0
    for k, d in zip(kmers, np.asarray(dist1_means) - np.asarray(dist2_means)):
        if kmer_shift_dirs[k] ^ (d < 0):
            s1 += abs(d)
        else:
            s2 += abs(d)
    return s1 > s2


def _stack(column):
    '''(n, w) matrix of a DataFrame column whose cells are length-w arrays'''
    return np.array(column.tolist())


def assign_modified_distribution(results, sm_preds, model):
    '''
    Decide for every significant record which fitted component is the modified
    one and flip the record (and its single-molecule predictions) when it is
    dist1 (reference stats.py:557-599).  `results` is the DataFrame of
    gmmtest.parallel_test; modified in place and returned.

    Evaluated on whole columns: the per-kmer medians of calculate_kmer_shift_directions
    over a flattened (records x window) table, the vote of dist1_is_modified as w masked
    column additions in window order (the same additions, in the same order, as the
    per-record loop), the flip as masked column assignments.
    '''
    n = len(results)
    if not n:
        return results, sm_preds
    kmers = _stack(results['kmers']).astype(str)
    m1, m2 = _stack(results['dist1_means']).astype(np.float64), _stack(results['dist2_means']).astype(np.float64)
    w = kmers.shape[1]
    # ---- shift direction per kmer (stats.py:516-534)
    uniq, code = np.unique(kmers.ravel(), return_inverse=True)
    expected = np.array([model[k][0] for k in uniq.tolist()])[code]
    lower = np.abs(np.minimum(m1, m2).ravel() - expected)
    upper = np.abs(np.maximum(m1, m2).ravel() - expected)
    order = np.argsort(code, kind='stable')
    cuts = np.flatnonzero(np.diff(code[order])) + 1
    dirs = np.array([1 if np.median(u) > np.median(lo) else 0
                     for u, lo in zip(np.split(upper[order], cuts), np.split(lower[order], cuts))])
    # ---- which records have dist1 as the modified component (stats.py:537-554)
    d = m1 - m2
    vote1 = dirs[code].reshape(n, w).astype(bool) ^ (d < 0)
    s1, s2 = np.zeros(n), np.zeros(n)
    for j in range(w):
        ad = np.abs(d[:, j])
        s1 = s1 + np.where(vote1[:, j], ad, 0.0)
        s2 = s2 + np.where(vote1[:, j], 0.0, ad)
    flip = s1 > s2
    if not flip.any():
        return results, sm_preds
    # ---- flip them (stats.py:571-599)
    idx = results.index[flip]
    results.loc[idx, 'cntrl_frac_mod'] = 1 - results.loc[idx, 'cntrl_frac_mod'].to_numpy()
    results.loc[idx, 'treat_frac_mod'] = 1 - results.loc[idx, 'treat_frac_mod'].to_numpy()
    results.loc[idx, 'log_odds'] = np.negative(results.loc[idx, 'log_odds'].to_numpy())
    lo, hi = results.loc[idx, 'log_odds_lower_ci'].to_numpy(), results.loc[idx, 'log_odds_upper_ci'].to_numpy()
    # stats.py:577-579 assigns np.negative(results.loc[i, [upper, lower]]), a label-indexed Series, through .loc:
    # pandas aligns on the labels, so each bound is negated IN PLACE and the interval of a flipped row comes out
    # inverted (lower > upper) in the reference's BED.  Reproduced here: the output contract is the reference's.
    results.loc[idx, 'log_odds_lower_ci'], results.loc[idx, 'log_odds_upper_ci'] = np.negative(lo), np.negative(hi)
    for c1, c2 in (('dist1_means', 'dist2_means'), ('dist1_stds', 'dist2_stds')):
        a, b = results[c1].to_numpy(dtype=object, copy=True), results[c2].to_numpy(dtype=object, copy=True)
        a[flip], b[flip] = b[flip], a[flip].copy()
        results[c1], results[c2] = a, b
    if sm_preds:
        for gene_id, pos in zip(results.loc[idx, 'gene_id'].tolist(), results.loc[idx, 'pos'].tolist()):
            try:
                pos_sm_preds = sm_preds[gene_id][pos]
            except KeyError:
                continue
            m = pos_sm_preds['model']
            m['unmod'], m['mod'] = m['mod'], m['unmod']
            for cond in ('cntrl', 'treat'):
                for rep_sm_preds in pos_sm_preds[cond].values():
                    rep_sm_preds['preds'] = [1 - (p + o) for p, o in
                                             zip(rep_sm_preds['preds'], rep_sm_preds['outlier_preds'])]
    return results, sm_preds
