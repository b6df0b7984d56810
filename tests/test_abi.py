"""CPU: the C-ABI library loads, exports every symbol include/vited_b200.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from tests.conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'vited_b200.h')).read()
    return sorted(set(re.findall(r'VITED_API\s+[\w\s\*]+?\b(vited_\w+)\s*\(', text)))


def test_header_symbols_are_exported_and_bound():
    import vited_b200
    from vited_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 18
    assert sorted(_lib.SIGNATURES) == declared, 'ctypes table and header disagree'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} not exported'


def test_library_is_sm100a_with_tcgen05_and_tma():
    from vited_b200 import _lib
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in sass.stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM'):
        assert mnemonic in sass.stdout, f'{mnemonic} missing: the GEMM is not a tcgen05/TMA kernel'


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback():
    import vited_b200
    from vited_b200 import _lib
    cfg = _lib.Config(64, 8, 3, 4, 384, 8, 8, 12, 4.0, 1)
    handle = ctypes.c_void_p()
    status = _lib.lib.vited_create(ctypes.byref(cfg), 0, ctypes.byref(handle))
    assert status != 0 and 'no CUDA device' in _lib.last_error()
    model = vited_b200.build_model(vited_b200.get_config('test'))
    with pytest.raises(vited_b200.VitedError):
        model(torch.zeros(1, 2, 3, 64, 64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'vit-ed_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.sh')):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f'{f} imports the oracle'
