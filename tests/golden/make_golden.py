#!/usr/bin/env python
'''
Generates the golden fixtures under tests/golden/ (run in the build container,
where /root/reference is mounted; the GPU box never sees /root/reference).

  hek293t_wt.collapsed.hdf5, hek293t_mettl3.collapsed.hdf5
        the reference's own test inputs (tests/test_data/, prep outputs), copied
        byte for byte -- test DATA, not source code
  fixture_ks.npz
        every position of those files that enters position_stats with window 3
        (3 444 of them): feature, position, KS D, KS p, gate status.  The KS
        stage of the oracle is the reference's own arithmetic (numpy svd /
        sort / searchsorted, scipy kstwobign) -> these values PIN the oracle
        and the CUDA path.
  fixture_gmm.npz
        the 213 gated positions: oracle fit (labels, iterations, parameters,
        statistics).  Pomegranate is not available, so these pin the oracle's
        *restatement* (regression guard), not pomegranate itself.
  fixture_w3.bed, fixture_w3_sm.json
        oracle end-to-end output (BED text, single-molecule JSON).
  scipy_kats.npz
        known answers from SciPy for the scalar routines the kernels
        re-implement: kolmogorov sf, chi2 sf (1 dof), chi2_contingency G-tests.
'''
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]

import yanocomp_oracle as yo          # noqa: E402
from yanocomp_b200 import h5lite      # noqa: E402

REF = '/root/reference/tests/test_data'
FILES = ('hek293t_wt.collapsed.hdf5', 'hek293t_mettl3.collapsed.hdf5')
MODEL = '/root/reference/yanocomp/data/r9.4_70bps.u_to_t_rna.5mer.template.model'


def main():
    for fn in FILES:
        shutil.copyfile(os.path.join(REF, fn), os.path.join(HERE, fn))
    cf = [h5lite.File(os.path.join(HERE, FILES[0]))]
    tf = [h5lite.File(os.path.join(HERE, FILES[1]))]
    opts = yo.Options(window_size=3, random_seed=0, generate_sm_preds=True)
    records, everything = yo.test_features(cf, tf, opts)
    keys = sorted(everything)
    feat = np.array([k[0] for k in keys])
    pos = np.array([k[1] for k in keys], dtype=np.int64)
    res = [everything[k] for k in keys]
    np.savez_compressed(
        os.path.join(HERE, 'fixture_ks.npz'), feature=feat, pos=pos,
        ks_stat=np.array([r.ks_stat for r in res]), ks_p=np.array([r.ks_p_val for r in res]),
        status=np.array([r.status for r in res], dtype=np.int8))
    gated = [(k, r) for k, r in zip(keys, res) if r.status in (yo.ST_TESTED, yo.ST_FIT_FAILED)]
    fitted = [(k, r) for k, r in gated if r.fit is not None]
    np.savez_compressed(
        os.path.join(HERE, 'fixture_gmm.npz'),
        gated_feature=np.array([k[0] for k, _ in gated]),
        gated_pos=np.array([k[1] for k, _ in gated], dtype=np.int64),
        gated_status=np.array([r.status for _, r in gated], dtype=np.int8),
        feature=np.array([k[0] for k, _ in fitted]),
        pos=np.array([k[1] for k, _ in fitted], dtype=np.int64),
        iterations=np.array([r.fit.iterations for _, r in fitted]),
        label_counts=np.array([np.bincount(r.fit.init_pred, minlength=3) for _, r in fitted]),
        mu=np.array([[d.mu for d in r.fit.dists[:2]] for _, r in fitted]),
        cov=np.array([[d.cov for d in r.fit.dists[:2]] for _, r in fitted]),
        box=np.array([np.stack([r.fit.dists[2].lo, r.fit.dists[2].hi], axis=1) for _, r in fitted]),
        cw=np.array([r.fit.combined_weights for _, r in fitted]),
        stats=np.array([[r.g_stat, r.hom_g_stat, r.p_val, r.cntrl_frac_mod, r.treat_frac_mod,
                         r.log_odds, r.log_odds_lower_ci, r.log_odds_upper_ci] for _, r in fitted]),
        loglik=np.array([r.fit.ll_trace[-1] for _, r in fitted]))
    for f in cf + tf:
        f.close()
    bed = yo.gmmtest([os.path.join(HERE, FILES[0])], [os.path.join(HERE, FILES[1])], opts, MODEL,
                     output_bed_fn=os.path.join(HERE, 'fixture_w3.bed'),
                     output_sm_preds_fn=os.path.join(HERE, 'fixture_w3_sm.json'))
    # make the JSON independent of the directory it was generated in
    with open(os.path.join(HERE, 'fixture_w3_sm.json')) as fh:
        doc = json.load(fh)
    doc['input_fns'] = {'cntrl': {'0': FILES[0]}, 'treat': {'0': FILES[1]}}
    with open(os.path.join(HERE, 'fixture_w3_sm.json'), 'w') as fh:
        json.dump(doc, fh)
    print(f'{len(keys)} positions, {len(gated)} gated, {len(fitted)} fitted, {bed.count(chr(10))} BED lines')

    # ---- SciPy known answers
    from scipy import special, stats
    rng = np.random.default_rng(20261019)
    x = np.concatenate([np.linspace(0.02, 4.0, 797), rng.uniform(0.3, 2.5, 1200)])
    g = np.concatenate([np.linspace(0, 150, 301), rng.uniform(0, 30, 700)])
    tables, answers = [], []
    for t in range(600):
        rc, rt = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        pr = np.round(rng.uniform(0, 80, (rc + rt, 3)) * (rng.uniform(size=(rc + rt, 3)) > 0.12))
        if t % 40 == 0:
            pr[:] = 0
        if t % 13 == 0:
            pr[:, 1] = 0
        if t % 17 == 0:
            pr[:, 2] = 0
        with np.errstate(all='ignore'):
            het, hom, p = yo.gmm_g_test(pr[:rc], pr[rc:], 0.05)
            fs = yo.calculate_fractional_stats(pr[:rc], pr[rc:])
        padded = np.full((6, 3), -1.0)
        padded[:rc + rt] = pr
        tables.append(np.concatenate([[rc, rt], padded.ravel()]))
        answers.append([het, hom, p, *fs])
    np.savez_compressed(os.path.join(HERE, 'scipy_kats.npz'),
                        kolmogorov_x=x, kolmogorov_sf=special.kolmogorov(x),
                        chi2_g=g, chi2_sf=stats.chi2.sf(g, 1),
                        gtest_tables=np.array(tables), gtest_answers=np.array(answers),
                        z_ci95=-stats.norm.ppf((1 - 95 / 100) / 2))
    print('scipy', __import__('scipy').__version__, 'numpy', np.__version__)


if __name__ == '__main__':
    main()
