"""ASan + UBSan over the host-side native code of the hot path (the
reference leans on CompressAI's C++ coder and zarr's chunk store; their replacements here are
rans_host.cpp and host_io.cpp).  Builds tests/native/sanitize_host.cpp with the two sources under
-fsanitize=address,undefined and runs it: any heap overflow, use after free, misaligned access or
signed overflow in a round trip, a truncated stream, a ragged tile gather or the threaded file
I/O fails the test."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'cnn_autoencoder_b200', 'csrc')


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
def test_host_coder_and_io_under_asan_ubsan(tmp_path):
    exe = tmp_path / 'sanitize_host'
    cmd = ['g++', '-std=c++17', '-O1', '-g', '-fno-omit-frame-pointer', '-fsanitize=address,undefined',
           '-fno-sanitize-recover=undefined', '-pthread',
           os.path.join(ROOT, 'tests', 'native', 'sanitize_host.cpp'),
           os.path.join(CSRC, 'rans_host.cpp'), os.path.join(CSRC, 'host_io.cpp'), '-o', str(exe)]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and 'asan' in (build.stderr or '').lower():
        pytest.skip('libasan is not installed: ' + build.stderr.splitlines()[-1])
    assert build.returncode == 0, build.stderr
    work = tmp_path / 'files'
    work.mkdir()
    env = dict(os.environ, ASAN_OPTIONS='detect_leaks=1:abort_on_error=0', UBSAN_OPTIONS='print_stacktrace=1')
    env.pop('LD_PRELOAD', None)
    run = subprocess.run([str(exe), str(work)], capture_output=True, text=True, env=env, timeout=300)
    assert run.returncode == 0 and 'sanitize_host: ok' in run.stdout, run.stdout + run.stderr
