"""Import shim for MeshFEM's `benchmark` module (3rdParty/MeshFEM/python/benchmark.py over python_bindings/benchmark.cc:9-13):
reset / start_timer_section / stop_timer_section / start_timer / stop_timer / report(include_messages) and the benchmarkit
decorators.  The timers are the library's own section tree (voxelfem_b200/csrc/vf_trace.cu), so Python sections nest with the
sections the C++ host code opens under the reference's names ("CG Iterations", "OC step", "Build load", ...).  Sections drain
the device when they close, which is only done while timing is enabled: `enable(True)` (or VF_BENCHMARK=1) turns it on."""
import ctypes as C
import functools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # repo root
from voxelfem_b200 import capi  # noqa: E402


def _L(): return capi.lib()


def enable(on=True): _L().vf_benchmark_enable(int(bool(on)))
def enabled(): return bool(_L().vf_benchmark_enabled())
def reset(): _L().vf_benchmark_reset()
def start_timer_section(name): _L().vf_benchmark_start_timer_section(name.encode())
def stop_timer_section(name): _L().vf_benchmark_stop_timer_section(name.encode())
def start_timer(name): _L().vf_benchmark_start_timer(name.encode())
def stop_timer(name): _L().vf_benchmark_stop_timer(name.encode())


def report_string(include_messages=False):
    L = _L()
    n = L.vf_benchmark_report(int(bool(include_messages)), None, 0)
    buf = C.create_string_buffer(n + 1)
    L.vf_benchmark_report(int(bool(include_messages)), buf, n + 1)
    return buf.value.decode()


def report(include_messages=False):
    sys.stdout.write(report_string(include_messages))


def to_dict():
    """{section path: (seconds, {timer: seconds})} (benchmark.cc:22-32)."""
    out = {}
    for line in report_string().splitlines():
        path, secs, _ = line.strip().split("\t")
        out[path] = (float(secs), {})
    return out


def benchmarkit(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        start_timer_section(fn.__name__)
        res = fn(*args, **kwargs)
        stop_timer_section(fn.__name__)
        return res
    return wrapper


def benchmarkit_customname(name):
    def named_benchmarkit(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            start_timer_section(name)
            res = fn(*args, **kwargs)
            stop_timer_section(name)
            return res
        return wrapper
    return named_benchmarkit
