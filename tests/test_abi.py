"""CPU-only: the C-ABI library loads and exports every symbol include/parl_b200.h declares."""
import os
import re

from parl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'parl_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(rl_[a-z0-9_]+)\s*\(', src)))


def test_library_loads_and_exports_header_symbols():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), 'libparl_b200.so does not export %s' % n
    assert lib.rl_abi_version() == 1


def test_ctypes_signatures_cover_header():
    assert sorted(_lib.SIGNATURES.keys()) == _declared()


def test_workspace_size_query():
    lib = _lib.load()
    assert lib.rl_loss_workspace_bytes(4096) >= 4096 * 16


def test_cpu_tensor_is_rejected_loudly():
    import pytest
    import torch
    from parl_b200 import kernels
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError):
        kernels.sample_categorical(x, 0, 0)


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under parl_b200/ may import it (static scan of every module)."""
    import ast
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'parl_b200')
    offenders = []
    for d, _, files in os.walk(root):
        for f in files:
            if not f.endswith('.py'):
                continue
            tree = ast.parse(open(os.path.join(d, f)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or '']
                if any(n == 'oracle' or n.startswith('oracle.') for n in names):
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders
