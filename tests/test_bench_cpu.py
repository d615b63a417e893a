"""bench.py without a GPU: the reference arm (CPU oracle) prints a valid JSON line, non-zero ranks of the reference arm
exit without work, and -- statically -- no collective-bearing call of the GPU arm sits in a rank-specific branch (a
rank-0-only step() once deadlocked the N > 1 runs: step() holds the NCCL all-reduce)."""
import ast
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, 'bench.py')


def _run(args, env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run(['--impl', 'reference', '--workload', 'cfg1', '--steps', '1', '--warmup', '0', '--gpus', '1'], {'RANK': '0', 'WORLD_SIZE': '1'})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['impl'] == 'reference' and d['unit'] == 'MPix/s' and d['higher_is_better'] is True
    assert d['metric'] == 'MPix/s encode+probclass fwd' and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'MPix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'] == 'cfg1' and d['n_gpus'] == 1 and d['steps'] == 1


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'], {'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''


COLLECTIVE_CALLS = {'step', 'barrier', 'timed', 'e2e_run'}


def _mentions_rank(node):
    return any(isinstance(n, ast.Name) and n.id == 'rank' for n in ast.walk(node))


def test_no_collective_in_a_rank_specific_branch():
    tree = ast.parse(open(BENCH).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ('main', 'measure_encode_pc')]
    assert len(fns) == 2
    bad = []
    for node in [m for f in fns for m in ast.walk(f)]:
        if isinstance(node, ast.If) and _mentions_rank(node.test):
            for sub in node.body + node.orelse:
                for c in ast.walk(sub):
                    if isinstance(c, ast.Call):
                        f = c.func
                        if isinstance(f, ast.Name) and f.id in COLLECTIVE_CALLS:
                            bad.append((node.lineno, f.id))
                        if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == 'dist' and \
                                f.attr in ('all_reduce', 'barrier', 'broadcast', 'all_gather'):
                            bad.append((node.lineno, 'dist.' + f.attr))
    assert not bad, 'collective-bearing calls inside rank-specific branches: %s' % bad
    # and the step itself all-reduces on every rank when world > 1
    src = open(BENCH).read()
    assert 'dist.all_reduce(metric)' in src
