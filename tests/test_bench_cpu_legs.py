"""bench.py's CPU legs are bounded by wall clock (a host busy with other jobs must not stall the GPU numbers): the SIGALRM limiter,
the shared budget and the degraded results.  CPU only."""
import importlib.util
import os
import sys
import time

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def bench():
    argv = sys.argv
    sys.argv = ['bench.py']
    try:
        spec = importlib.util.spec_from_file_location('bench_under_test', os.path.join(REPO, 'bench.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    yield mod
    mod.wall_clock_limit.pool = None


def test_wall_clock_limit_interrupts_python_loops(bench):
    bench.wall_clock_limit.pool = None
    t0 = time.perf_counter()
    with pytest.raises(bench.CpuLegTimeout):
        with bench.wall_clock_limit(1):
            x = 0
            while True:
                x += 1
    assert 0.9 < time.perf_counter() - t0 < 3.0
    with bench.wall_clock_limit(5):          # a leg that finishes in time leaves no alarm behind
        pass
    time.sleep(0.05)


def test_shared_budget_is_charged_and_refuses_when_spent(bench):
    bench.wall_clock_limit.pool = 3.0
    with bench.wall_clock_limit(60):
        time.sleep(1.2)
    assert 1.5 < bench.wall_clock_limit.pool < 1.9
    with pytest.raises(bench.CpuLegTimeout):
        with bench.wall_clock_limit(60):     # < 2 s left: refused before any work starts
            pass
    bench.wall_clock_limit.pool = None


def test_cpu_baseline_block_degrades_to_a_note_when_the_budget_is_spent(bench):
    bench.wall_clock_limit.pool = 0.5
    t0 = time.perf_counter()
    r = bench.cpu_baseline_block(None, None, None, 300, 20)
    assert 'skipped' in r and time.perf_counter() - t0 < 1.0
    bench.wall_clock_limit.pool = None


def test_thread_sweep_never_tries_all_threads_on_a_wide_host(bench, monkeypatch):
    tried = []

    class Port(bench.CpuPort):
        def __init__(self):
            pass

        def time_iterations(self, n, threads=None):
            tried.append(threads)
            return [{8: 0.10, 16: 0.06, 32: 0.09}.get(threads, 5.0)] * n

    monkeypatch.setattr(bench, 'host_threads', lambda: 128)
    bench.wall_clock_limit.pool = None
    best, res = Port().sweep_threads(warm=1, probe=2)
    assert best == 16 and 128 not in tried            # 32 is slower than 16 x 1.25: the sweep stops there
    assert set(res) == {8, 16, 32}
