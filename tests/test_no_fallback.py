'''
CPU: the product path has no CPU fallback and never touches the oracle.
'''
import ast
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'yanocomp_b200')


def test_product_never_imports_the_oracle():
    for fn in os.listdir(PKG):
        if not fn.endswith('.py'):
            continue
        tree = ast.parse(open(os.path.join(PKG, fn)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or '']
            for n in names:
                assert 'oracle' not in n, f'{fn} imports {n}'
        assert 'yanocomp_oracle' not in open(os.path.join(PKG, fn)).read()


def test_engine_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from yanocomp_b200.engine import Engine, EngineParams
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Engine(EngineParams())
    from yanocomp_b200 import synth
    from yanocomp_b200.batch import HostBatch, DeviceBatch
    hb = HostBatch(synth.make_dataset(1, 1, 10, 8), 4, 2, pin=False)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        DeviceBatch(hb)


def test_missing_library_is_a_hard_error(monkeypatch):
    from yanocomp_b200 import _native
    monkeypatch.setattr(_native, '_lib', None)
    monkeypatch.setattr(_native, 'LIB_PATH', '/nonexistent/libyanocomp_b200.so')
    with pytest.raises(ImportError, match='no CPU fallback'):
        _native.lib()
