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


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """The boundary is a C ABI: include/vited_b200.h compiles as C99 (no C++, no torch / CUDA types in the signatures)
    and a C program linked against the shared library -- no Python, no torch -- gets the same loud failure without a
    GPU (and a handle with one)."""
    from vited_b200 import _lib
    src = tmp_path / 'abi.c'
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "vited_b200.h"
int main(void) {
  vited_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.img_size = 64; cfg.patch_size = 8; cfg.in_chans = 3; cfg.num_classes = 4; cfg.embed_dim = 384;
  cfg.depth = 8; cfg.c_depth = 8; cfg.num_heads = 12; cfg.mlp_ratio = 4.0f; cfg.qkv_bias = 1;
  vited_engine* e = NULL;
  int rc = vited_create(&cfg, 0, &e);
  if (rc != 0) { printf("create failed: %s\\n", vited_last_error()); return 0; }
  printf("created: %d weights expected, operand type %d\\n", vited_num_weights_expected(e), vited_act_dtype());
  vited_destroy(e);
  return 0;
}
''')
    inc = os.path.join(ROOT, 'include')
    exe = tmp_path / 'abi'
    syntax = subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-fsyntax-only', '-I', inc, str(src)],
                            capture_output=True, text=True)
    assert syntax.returncode == 0, syntax.stderr
    libdir = os.path.dirname(_lib.LIB_PATH)
    link = subprocess.run(['gcc', '-std=c99', '-I', inc, str(src), '-o', str(exe), '-L', libdir, '-lvited_b200',
                           f'-Wl,-rpath,{libdir}'], capture_output=True, text=True)
    assert link.returncode == 0, link.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stderr
    if torch.cuda.is_available():
        assert 'created: ' in run.stdout and 'weights expected' in run.stdout
    else:
        assert 'create failed' in run.stdout and 'no CUDA device' in run.stdout
