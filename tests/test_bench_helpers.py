"""CPU checks of bench.py's host-side helpers (no GPU, no nvidia-smi): the clock sampler keeps the samples whose
timestamps fall inside the timed region and falls back to the warm-up's when the region was shorter than the
sampling period; the reference arm and the bench line share one metric name per workload."""
import datetime
import importlib.util
import os
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Proc:
    def terminate(self):
        pass

    def wait(self, timeout=None):
        pass


def _sampler(bench, rows, t_begin):
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.idx, s.p, s.t_begin, s.t_end = 0, _Proc(), t_begin, None
    s.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    for t, clk, power, slow in rows:
        stamp = datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        s.f.write(f"{stamp}, 0, {clk}, 1965, {power}, 0x0000000000000000, {slow}, Not Active, Not Active, Not Active\n")
    return s


def test_clock_sampler_keeps_only_the_timed_region(monkeypatch):
    bench = _bench()
    now = time.time()
    rows = [(now - 0.50, 1200, 100.0, "Active"),      # idle, before the region: must not count
            (now - 0.10, 1900, 300.0, "Not Active"),
            (now + 0.05, 1965, 350.0, "Not Active"),
            (now + 0.10, 1950, 360.0, "Not Active"),
            (now + 0.30, 1500, 120.0, "Not Active")]  # after the region
    s = _sampler(bench, rows, now)
    monkeypatch.setattr(bench.time, "time", lambda: now + 0.2)
    out = s.stop()
    assert out["samples"] == 2 and out["window"] == "timed region"
    assert out["sm_max_mhz"] == 1965.0 and out["reasons"] == []
    assert out["sm_mhz"] in (1950.0, 1965.0) and out["power_w_max"] == 360.0


def test_clock_sampler_falls_back_to_the_warm_up_for_a_short_region(monkeypatch):
    bench = _bench()
    now = time.time()
    rows = [(now - 3.0, 600, 80.0, "Not Active"), (now - 0.30, 1965, 340.0, "Not Active"),
            (now - 0.25, 1965, 345.0, "Not Active")]
    s = _sampler(bench, rows, now)
    monkeypatch.setattr(bench.time, "time", lambda: now + 0.01)
    out = s.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1965.0
    assert "warm-up" in out["window"]


def test_clock_sampler_without_nvidia_smi_reports_no_samples():
    bench = _bench()
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.idx, s.p, s.t_begin, s.t_end = 0, None, None, None
    s.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    s.start()
    assert s.stop() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    assert not os.path.exists(s.f.name)
