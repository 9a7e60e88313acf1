"""bench.py contract checks that run without a GPU: the reference arm (`--impl reference`) prints exactly one JSON line
with the keys the driver reads, and ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['value'] > 0 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and d['gpu_launches'] == 0


def test_reference_arm_other_ranks_are_silent():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == []
