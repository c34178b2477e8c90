'''
Feeds the packed features of the product path to the CPU oracle, so both see
exactly the same inputs (test infrastructure).
'''
import numpy as np

import yanocomp_oracle as yo


def cond_events(feature, s0, s1, rep0=1):
    '''samples [s0, s1) of a packed Feature as an oracle CondEvents'''
    ss = feature.samp_start
    a, e = int(ss[s0]), int(ss[s1])
    X = feature.X[:, a:e].astype(np.float64).T           # reads x positions
    rep = np.zeros(e - a, dtype=np.int64)
    for s in range(s0, s1):
        rep[ss[s] - a:ss[s + 1] - a] = s - s0 + rep0
    positions = np.asarray(feature.positions, dtype=np.int64)
    ids = feature.read_ids[a:e] if feature.read_ids is not None else [''] * (e - a)
    return yo.CondEvents(X, positions, rep, list(ids))


def run_feature(feature, n_samples, n_cntrl, opts):
    '''oracle results of every window of one packed feature: {pos: PositionResult}'''
    cntrl = cond_events(feature, 0, n_cntrl)
    treat = cond_events(feature, n_cntrl, n_samples)
    fkey = yo.feature_key(feature.feature_id)
    kmers = feature.kmers
    if kmers is None:
        kmers = {int(p): 'NNNNN' for p in cntrl.positions}
    out = {}
    for pos, pk, cw, tw in yo.create_positional_data(
            cntrl, treat, kmers, feature.strand == '-', opts.min_read_depth, opts.window_size):
        out[int(pos)] = yo.position_stats(
            cw, tw, pk, opts, yo.position_key(opts.random_seed, fkey, pos), n_cntrl)
    return out


def run_features(features, n_samples, n_cntrl, opts):
    return {f.feature_id: run_feature(f, n_samples, n_cntrl, opts) for f in features}
