"""bench.py's CPU arm (`--impl reference`: the oracle on the host cores) runs without a GPU and prints ONE JSON line with
the contract's keys; under a torchrun-style environment (OMP_NUM_THREADS=1, RANK > 0) only rank 0 works and it still
uses every host core."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'c1',
                        '--steps', '2', '--warmup', '1', *args], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_json_line():
    out = _run({'OMP_NUM_THREADS': '1'})
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['dtype'] == 'f64' and d['value'] > 0 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cores = len(os.sched_getaffinity(0))
    assert d['cpu_baseline']['cores'] == cores      # torchrun's OMP_NUM_THREADS=1 does not throttle the CPU arm


def test_reference_arm_other_ranks_exit_quietly():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}, '--gpus', '2').strip() == ''
