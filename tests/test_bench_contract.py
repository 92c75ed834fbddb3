"""bench.py's contract pieces that need no GPU: the reference arm's JSON line and the clock sampler.

The driver runs `bench.py --impl reference` beside our arm and computes the ratio itself, so the
line must carry the same metric / unit / config keys, `impl`, `cpu_baseline` and `e2e`.
"""

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '1'], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    assert d['metric'] == 'Msamples/s spectrogram+filter+envelope' and d['unit'] == 'Msamples/s'
    assert d['higher_is_better'] is True and d['dtype'] == 'f64' and d['vs_baseline'] is None
    assert d['value'] > 0 and d['ms_per_step'] > 0
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value']
    assert cb['single_thread_value'] > 0 and 'sample' in cb
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit']
    assert e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_stdout_carries_the_json_line_only():
    """Whatever libraries print on file descriptor 1 while the bench runs goes to stderr (NCCL's
    version banner used to land in front of the JSON line): stdout is exactly one line."""
    code = ('import os, sys; sys.argv = ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"];'
            'import bench; real = bench.run_reference;'
            'bench.run_reference = lambda a: (os.write(1, b"banner from a library\\n"), real(a));'
            'bench.main()')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1 and lines[0].startswith('{'), out.stdout[:300]
    assert json.loads(lines[0])['impl'] == 'reference'
    assert 'banner from a library' in out.stderr


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--gpus', '2', '--steps', '1', '--warmup', '1'], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0
    assert not [l for l in out.stdout.splitlines() if l.startswith('{')]


def test_clock_sampler_without_a_gpu_reports_no_samples():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0, None)
    s.start()
    s.wait_ready()
    t0 = time.perf_counter()
    time.sleep(0.05)
    r = s.stop(t0, time.perf_counter())
    assert set(r) >= {'sm_mhz', 'sm_max_mhz', 'reasons', 'samples'}
    assert r['samples'] == 0 or r['sm_mhz'] > 0
