"""Import shim for MeshFEM's `benchmark` module (BENCHMARK_* timers, used by python/LayerByLayerObjective.py:5 as
decorators and python/CoarseningLevelBenchmark.py:11): a small wall-clock timer tree with the same entry points."""
import functools
import time

_timers = {}


def reset(): _timers.clear()


def start_timer(name): _timers.setdefault(name, [0.0, 0, None])[2] = time.perf_counter()


def stop_timer(name):
    t = _timers.get(name)
    if t and t[2] is not None:
        t[0] += time.perf_counter() - t[2]; t[1] += 1; t[2] = None


def report():
    for k, (tot, n, _) in sorted(_timers.items(), key=lambda kv: -kv[1][0]):
        print("%-50s %10.4f s  (%d calls)" % (k, tot, n))


def benchmarkit_customname(name):
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*a, **k):
            start_timer(name)
            try:
                return fn(*a, **k)
            finally:
                stop_timer(name)
        return wrapped
    return deco


def benchmarkit(fn): return benchmarkit_customname(fn.__qualname__)(fn)
