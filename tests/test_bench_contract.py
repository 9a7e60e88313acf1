"""bench.py contract checks that run without a GPU: the reference arm (`--impl reference`) prints exactly one JSON line
with the keys the driver reads, and ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, extra=('--no-full',)):
    env = dict(os.environ)
    env.update(env_extra or {})
    # small frames keep the CPU suite short; the driver runs the default 512x512
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1',
                        '--size', '64'] + list(extra), capture_output=True, text=True, env=env, timeout=900)
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
    assert d['steps'] == 2 and d['warmup'] == 1          # the arm runs exactly the steps it is asked for


def test_reference_arm_config_is_the_product_arms_config():
    """The driver compares the two arms' `config`: both come from bench.workload_config."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    a = argparse.Namespace(batch=4, unroll=8, size=512, mode='all')
    c = bench.workload_config(a, 1)
    assert 'C2' in c['workload'] and '512x512' in c['workload'] and 'T=8' in c['workload'] and c['global_batch'] == 4
    assert 'l2_policy' in c


def test_reference_arm_full_shape_and_train_leg():
    """Without --no-full the line also carries ONE step at the workload's true shape and the C1 train step
    (train2D.py:87-93 at 128x128, T=4, batch 2) of the CPU arm."""
    lines = _run(extra=('--batch', '1', '--unroll', '1'))
    d = json.loads(lines[0])
    assert d['full_shape']['value'] > 0 and 'true C2 shape' in d['full_shape']['shape']
    assert d['train']['value'] > 0 and 'fwd+loss+bwd+Adam' in d['train']['sample']


def test_reference_arm_other_ranks_are_silent():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == []
