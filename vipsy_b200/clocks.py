"""Sample SM clocks / throttle reasons while a timed region runs (bench hygiene)."""
from __future__ import annotations

import statistics
import threading
import time


class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int = 0, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.power = [], set(), []
        self._stop = threading.Event()
        self._thr = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:          # no NVML: report nulls rather than fail the bench
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if r & bit and name not in ("gpu_idle",):
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "power_w_max": max(self.power) if self.power else None,
            "samples": len(self.samples),
        }
