"""AddressSanitizer + UBSan (alignment, shifts, overflow) over the kernel bodies (SURVEY.md §5 aux subsystems: the reference has no race / memory checking at
all).  The per-lane bodies of every kernel are compiled for the host with -fsanitize=address and driven through the C ABI
on odd batch sizes, the over-budget contact paths and both fused steps; compute-sanitizer does the same on the GPU box."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kernel_bodies_are_asan_clean(tmp_path):
    asan = sorted(glob.glob('/usr/lib/gcc/x86_64-linux-gnu/*/libasan.so'))
    if not asan:
        pytest.skip('libasan not installed')
    so = str(tmp_path / 'libagphys_asan.so')
    subprocess.check_call(['/usr/bin/g++', '-x', 'c++', '-std=c++17', '-O1', '-g', '-fPIC', '-shared', '-DAG_CPU_EMU', '-fsanitize=address,undefined', '-fno-sanitize-recover=undefined',
                           '-fno-omit-frame-pointer', '-Wno-unused-function', '-o', so, os.path.join(ROOT, 'assistive_gym_b200', 'csrc', 'agphys.cu')])
    ubsan = sorted(glob.glob('/usr/lib/gcc/x86_64-linux-gnu/*/libubsan.so'))
    env = dict(os.environ, LD_PRELOAD=' '.join([asan[-1]] + ubsan[-1:]), ASAN_OPTIONS='detect_leaks=0')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'kernel_harness', 'asan_workload.py'), so, ROOT], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and 'ASAN-WORKLOAD-OK' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    assert 'AddressSanitizer' not in r.stderr and 'runtime error' not in r.stderr
