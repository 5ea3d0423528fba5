"""bench.py's JSON contract, checked on the CPU through the reference arm (the only arm that runs without a GPU):
one JSON line with the driver's keys, the same `config` / `metric` / `unit` as the b200 arm would print, a
`cpu_baseline` describing the run and a zero-copy `e2e` object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--size', '64', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ['impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e']:
        assert key in d, key
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['value'] > 0 and d['unit'] == 'images/s'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    sys.path.insert(0, ROOT)
    import bench
    cfg = bench.workload_config(64, True, 16, 1)
    assert d['metric'] == bench.METRIC
    for k, v in cfg.items():                       # the reference arm reports the b200 arm's config
        assert d['config'][k] == v


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--size', '64',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
