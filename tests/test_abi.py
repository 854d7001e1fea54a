"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/dccf_b200.h declares.
No compute calls here (no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from dccf_b200 import _lib, build


@pytest.fixture(scope='module')
def lib_path():
    return build.build()


def test_header_symbols_are_exported_and_bound(lib_path):
    header = open(os.path.join(ROOT, 'include', 'dccf_b200.h')).read()
    body = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(dccf_[a-z0-9_]+)\s*\(', body))
    assert declared, 'no declarations found'
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.exported_symbols())


def test_library_loads_and_reports_abi(lib_path):
    lib = _lib.load()
    assert lib.dccf_abi_version() == _lib.ABI_VERSION
    header = open(os.path.join(ROOT, 'include', 'dccf_b200.h')).read()
    assert '#define DCCF_ABI_VERSION %d' % _lib.ABI_VERSION in header


def test_argument_errors_are_reported_without_a_gpu(lib_path):
    """Argument validation happens before any CUDA call."""
    lib = _lib.load()
    rc = lib.dccf_rank_eval(None, None, None, None, None, 1, 5, None, None, None, None)
    assert rc == -1
    assert b'null buffer' in lib.dccf_last_error()
    rc = lib.dccf_noise_fill(None, 4, 768, 0.1, 1, 1, 0, None)
    assert rc == -1 and b'null output' in lib.dccf_last_error()
    assert lib.dccf_bwd_splits(5632) >= 1
    rc = lib.dccf_score_gather(None, None, None, None, None, None, None, 4, None, None, None)
    assert rc == -1 and b'dccf_score_gather' in lib.dccf_last_error()
    dims = _lib.Dims(n_users=4, n_items=4, dim=32, feat_dim=64, n_samples=1, n_attr=1)
    expo = _lib.Expo(mode=0)
    rc = lib.dccf_score_gather(ctypes.byref(dims), None, None, None, ctypes.byref(expo), None, None, 4, None, None, None)
    assert rc == -1 and b'dim=32' in lib.dccf_last_error()
    state = (ctypes.c_uint32 * 624)()
    left, nxt = ctypes.c_int32(0), ctypes.c_int32(0)
    out = (ctypes.c_int64 * 4)()
    rc = lib.dccf_confounder_draw(state, ctypes.byref(left), ctypes.byref(nxt), 10, 4, out)
    assert rc == -1 and b'corrupt generator state' in lib.dccf_last_error()
    left.value = 1
    assert lib.dccf_confounder_draw(state, ctypes.byref(left), ctypes.byref(nxt), 0, 4, out) == -1
    assert lib.dccf_confounder_draw(state, ctypes.byref(left), ctypes.byref(nxt), 10, 0, None) == 0


def test_sass_is_sm100a(lib_path):
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '-lelf', lib_path], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_lib.DccfError, match='no CPU fallback'):
        _lib.load(str(tmp_path / 'nope.so'))


def test_training_kernels_use_tcgen05_and_bulk_copies(lib_path):
    """The two contractions of the training step and the evaluation scorer are tcgen05 kernels: their SASS holds
    the tensor-core MMA (UTCHMMA), TMEM loads (LDTM), TMEM allocation (UTCATOMSWS), the MMA-completion barrier
    (UTCBAR); the forward and the scorer also stream their W operand with bulk copies (UBLKCP)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    want = {'k_train_fwd_tc': ('UTCHMMA', 'LDTM', 'UTCATOMSWS', 'UTCBAR', 'UBLKCP'),
            'k_train_bwd_tc': ('UTCHMMA', 'LDTM', 'UTCATOMSWS', 'UTCBAR'),
            'k_row_scores_tc': ('UTCHMMA', 'LDTM', 'UTCATOMSWS', 'UTCBAR', 'UBLKCP')}
    sass = subprocess.run([cuobjdump, '-sass', lib_path], capture_output=True, text=True).stdout
    chunks = sass.split('Function : ')
    for kernel, mnemonics in want.items():
        body = [c for c in chunks if kernel in c.split('\n', 1)[0]]
        assert body, kernel
        for m in mnemonics:
            assert any(m in c for c in body), (kernel, m)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """Every ctypes mirror in dccf_b200/_lib.py has the size and the member offsets the C compiler gives the struct of
    include/dccf_b200.h (the header is compiled as plain C: it is the boundary a C / cgo / JNI host would bind)."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc') or shutil.which('cc')
    if gcc is None:
        pytest.skip('no C compiler')
    pairs = {'dccf_dims': _lib.Dims, 'dccf_expo': _lib.Expo, 'dccf_rng': _lib.Rng, 'dccf_adam': _lib.Adam,
             'dccf_adam_table': _lib.AdamTable, 'dccf_adam_tensor': _lib.AdamTensor, 'dccf_dp_channel': _lib.DpChannel,
             'dccf_dp_sync': _lib.DpSync, 'dccf_link_extra': _lib.LinkExtra, 'dccf_batch_ref': _lib.BatchRef}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dccf_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append('  printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for field, _ in cls._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, field, cname, field))
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run([gcc, '-std=c99', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {}
    for line in out.splitlines():
        cname, field, value = line.split()
        got[(cname, field)] = int(value)
    for cname, cls in pairs.items():
        assert got[(cname, 'size')] == ctypes.sizeof(cls), cname
        for field, _ in cls._fields_:
            assert got[(cname, field)] == getattr(cls, field).offset, (cname, field)
