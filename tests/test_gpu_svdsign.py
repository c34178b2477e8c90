'''
The PCA axis must carry the orientation numpy.linalg.svd (LAPACK dgesdd)
gives it, because the rounded KS statistic depends on it (csrc/yc_svdsign.cuh).
Device emulation vs numpy on random standardised windows.
'''
import ctypes as C

import numpy as np
import pytest

from yanocomp_b200 import _native as N

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('d', [1, 3, 5, 7, 9])
def test_device_axis_matches_numpy(d):
    import torch
    rng = np.random.default_rng(100 + d)
    count = 4000
    G = np.zeros((count, d, d)); T = np.zeros((count, d, d)); ref = np.zeros((count, d))
    for t in range(count):
        n = int(rng.integers(max(12, 2 * d + 2), 300))
        X = np.round(rng.normal(100, 5, (n, d)) * 16) / 16
        if t % 3 == 0 and d > 1:
            X[:, 1] += np.round(rng.uniform(-1, 1) * (X[:, 0] - 100) * 16) / 16
        mu = X.mean(0); sig = np.sqrt(((X - mu) ** 2.0).mean(0)); Z = (X - mu) / sig
        _, s, vt = np.linalg.svd(Z, full_matrices=False)
        ref[t] = vt[np.argmax(s ** 2.0)]
        G[t] = Z.T @ Z; T[t] = Z[:d]
    Gd = torch.from_numpy(G).cuda(); Td = torch.from_numpy(T).cuda()
    vd = torch.zeros(count, d, dtype=torch.float64, device='cuda')
    N.check(N.lib().yc_first_right_singular_vector_device(
        d, Gd.data_ptr(), Td.data_ptr(), vd.data_ptr(), count, torch.cuda.current_stream().cuda_stream))
    v = vd.cpu().numpy()
    err = np.abs(v - ref).max(axis=1)
    assert (err < 1e-9).all(), (int((err >= 1e-9).sum()), v[err.argmax()], ref[err.argmax()])
