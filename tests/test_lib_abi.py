"""CPU: the C-ABI library loads and exports every symbol include/brcnn.h
declares; the pure-host workspace queries answer without a GPU."""
import ctypes
import os
import re

import pytest

from boosting_rcnn_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'brcnn.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(brcnn_\w+)\s*\(', src)))


def test_header_declares_what_python_binds():
    names = header_functions()
    assert len(names) >= 18
    assert sorted(_lib.SIGNATURES) == names


def test_library_loads_and_exports_every_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f'{name} missing from libbrcnn.so'
    lib = _lib.load()
    assert b'sm_100a' in lib.brcnn_version()
    assert lib.brcnn_launch_count() >= 0


def test_workspace_queries_are_host_only():
    p = ops.make_rpn_params(16, [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)],
                            [8, 16, 32, 64, 128], 9, 1000, 256, 0.7, 0)
    lay = ops.rpn_workspace_layout(p)
    assert lay.cand_cap == 1000 and lay.keep_cap == 256
    assert 0 < lay.total_bytes == _lib.load().brcnn_rpn_workspace_bytes(p)
    rp = ops.make_rcnn_params(16, 256, 4, 0.05, 0.7, 100)
    rl = ops.rcnn_workspace_layout(rp)
    assert 0 < rl.total_bytes == _lib.load().brcnn_rcnn_workspace_bytes(rp)
    assert _lib.load().brcnn_nms_workspace_bytes(5000) > 0


def test_bad_arguments_are_rejected_not_crashed():
    lib = _lib.load()
    p = ops.make_rpn_params(0, [(4, 4)], [8], 3, 10, 10, 0.7, 0)  # batch 0
    assert lib.brcnn_rpn_workspace_bytes(p) == 0
    lay = _lib.RpnWsLayout()
    assert lib.brcnn_rpn_workspace_layout(p, lay) == -1  # BRCNN_ERR_ARG
    with pytest.raises(RuntimeError):
        _lib.check(-2, 'x')


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.load()


def _build_c_consumer(tmp_path):
    """gcc-compile tests/c_abi/abi_smoke.c against include/brcnn.h + libbrcnn.so."""
    import shutil
    import subprocess
    src = os.path.join(ROOT, 'tests', 'c_abi', 'abi_smoke.c')
    exe = str(tmp_path / 'abi_smoke')
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    if shutil.which('gcc') is None or not os.path.isdir(os.path.join(cuda, 'include')):
        pytest.skip('gcc / CUDA toolkit headers not available')
    cmd = ['gcc', '-O1', '-Wall', '-Werror', src, '-I' + os.path.join(ROOT, 'include'),
           '-I' + os.path.join(cuda, 'include'), '-L' + os.path.dirname(_lib.LIB_PATH),
           '-L' + os.path.join(cuda, 'lib64'), '-lbrcnn', '-lcudart', '-lm', '-o', exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    env = dict(os.environ)
    env['LD_LIBRARY_PATH'] = os.pathsep.join(
        [os.path.dirname(_lib.LIB_PATH), os.path.join(cuda, 'lib64'), env.get('LD_LIBRARY_PATH', '')])
    return exe, env


def test_plain_c_consumer_compiles_and_links(tmp_path):
    """The boundary is a C ABI: a C translation unit that includes only brcnn.h and the CUDA
    runtime compiles with -Wall -Werror, links against libbrcnn.so and loads (exit code 77 =
    'no CUDA device', the expected answer in the CPU-only build container)."""
    import subprocess
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    exe, env = _build_c_consumer(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert res.returncode in (0, 77), (res.returncode, res.stdout, res.stderr)
    assert 'sm_100a' in res.stdout or 'abi_smoke ok' in res.stdout


@pytest.mark.gpu
def test_plain_c_consumer_runs_on_gpu(tmp_path, cuda):
    import subprocess
    exe, env = _build_c_consumer(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, env=env)
    if res.returncode == 77:
        pytest.skip('abi_smoke: no CUDA device')
    assert res.returncode == 0, (res.stdout, res.stderr)
    assert 'abi_smoke ok' in res.stdout
