'''
CPU: the C-ABI library loads, exports every symbol include/yanocomp_b200.h
declares, and its host-callable scalar routines (the same __host__ __device__
code the kernels run) match SciPy / NumPy known answers.
'''
import ctypes as C
import os
import re

import numpy as np
import pytest

import yanocomp_oracle as yo
from yanocomp_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, 'include', 'yanocomp_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(yc_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 12
    lib = N.lib()
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    assert b'sm_100a' in lib.yc_version()


def test_kolmogorov_sf_matches_scipy():
    k = np.load(os.path.join(GOLD, 'scipy_kats.npz'))
    lib = N.lib()
    mine = np.array([lib.yc_kolmogorov_sf_host(float(x)) for x in k['kolmogorov_x']])
    ref = k['kolmogorov_sf']
    ulps = np.abs(mine - ref) / np.spacing(np.maximum(ref, 1e-300))
    assert ulps.max() <= 2, ulps.max()
    # the gate compares p with 0.05: the decision must agree everywhere
    assert np.array_equal(mine < 0.05, ref < 0.05)


def test_chi2_sf_matches_scipy():
    k = np.load(os.path.join(GOLD, 'scipy_kats.npz'))
    lib = N.lib()
    mine = np.array([lib.yc_chi2_sf_1dof_host(float(g)) for g in k['chi2_g']])
    np.testing.assert_allclose(mine, k['chi2_sf'], rtol=1e-12, atol=0)


def test_gtest_fractions_match_scipy_and_table2x2():
    k = np.load(os.path.join(GOLD, 'scipy_kats.npz'))
    lib = N.lib()
    assert abs(float(k['z_ci95']) - 1.959963984540054) == 0
    for row, ans in zip(k['gtest_tables'], k['gtest_answers']):
        rc, rt = int(row[0]), int(row[1])
        pr = np.ascontiguousarray(row[2:].reshape(6, 3)[:rc + rt])
        out = np.zeros(8)
        lib.yc_gtest_fractions_host(_pd(pr), rc, rt, 0.05, _pd(out))
        np.testing.assert_allclose(out, ans, rtol=1e-12, atol=1e-300, equal_nan=True)


@pytest.mark.parametrize('d', [1, 3, 5, 7, 9])
def test_pca_axis_has_lapack_orientation(d):
    lib = N.lib()
    rng = np.random.default_rng(d)
    for t in range(1500):
        n = int(rng.integers(max(12, 2 * d + 2), 300))
        X = np.round(rng.normal(100, 5, (n, d)) * 16) / 16
        if t % 3 == 0 and d > 1:
            X[:, 1] += np.round(rng.uniform(-1, 1) * (X[:, 0] - 100) * 16) / 16
        mu = X.mean(0); sig = np.sqrt(((X - mu) ** 2.0).mean(0)); Z = (X - mu) / sig
        _, s, vt = np.linalg.svd(Z, full_matrices=False)
        G = np.ascontiguousarray(Z.T @ Z); T = np.ascontiguousarray(Z[:d]); v = np.zeros(d)
        assert lib.yc_first_right_singular_vector_host(d, _pd(G), _pd(T), _pd(v)) == 0
        assert np.abs(v - vt[np.argmax(s ** 2.0)]).max() < 1e-9


def test_rng_shared_with_oracle():
    lib = N.lib()
    for z in (0, 1, 0xdeadbeef, (1 << 64) - 1):
        assert lib.yc_mix64_host(z) == yo.mix64(z)
    fk = yo.feature_key('ENST00000351111')
    assert lib.yc_position_key_host(7, fk, 870) == yo.position_key(7, fk, 870)
    from yanocomp_b200.batch import feature_key
    assert feature_key('ENST00000351111') == fk


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = N.lib()
    b = N.Batch(n_features=1, n_samples=40, n_cntrl=1, max_reads=10, max_cond_reads=5, total_cols=1)
    p = N.Params(window_size=3, min_read_depth=6, max_fit_depth=1000, gmm=1)
    out = (C.c_int64 * 4)()
    assert lib.yc_smem_bytes(C.byref(b), C.byref(p), out) == N.YC_EINVAL
    assert b'n_samples' in lib.yc_last_error()
    b.n_samples = 4
    p.window_size = 4
    assert lib.yc_smem_bytes(C.byref(b), C.byref(p), out) == N.YC_EINVAL
    p.window_size = 3
    b.max_reads = 1 << 22
    p.max_fit_depth = 1 << 20
    assert lib.yc_smem_bytes(C.byref(b), C.byref(p), out) == N.YC_ECAPACITY
    b.max_reads = 400
    p.max_fit_depth = 1000
    assert lib.yc_smem_bytes(C.byref(b), C.byref(p), out) == 0 and max(out) < 232448


def test_two_component_gtest_matches_scipy():
    '''--no-uniform: R x 2 homogeneity tables get Yates' correction when they are 2 x 2 (dof 1)'''
    lib = N.lib()
    rng = np.random.default_rng(3)
    for t in range(400):
        rc, rt = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        pr = np.zeros((rc + rt, 3))
        pr[:, :2] = np.round(rng.uniform(0, 60, (rc + rt, 2)) * (rng.uniform(size=(rc + rt, 2)) > 0.1))
        if t % 25 == 0:
            pr[:] = 0
        with np.errstate(all='ignore'):
            het, hom, pv = yo.gmm_g_test(pr[:rc, :2], pr[rc:, :2], 0.05)
            fs = yo.calculate_fractional_stats(pr[:rc, :2], pr[rc:, :2])
        out = np.zeros(8)
        lib.yc_gtest_fractions2_host(_pd(np.ascontiguousarray(pr)), rc, rt, 0.05, _pd(out))
        np.testing.assert_allclose(out, [het, hom, pv, *fs], rtol=1e-12, atol=1e-300, equal_nan=True)
