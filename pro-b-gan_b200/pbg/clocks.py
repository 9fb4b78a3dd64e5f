"""SM-clock / throttle-reason sampling during a timed region (bench.py's ``clocks`` key).

NVML is polled from a background thread every few milliseconds (the timed regions here last tens of
milliseconds, far below nvidia-smi's -lms floor); falls back to one nvidia-smi query if NVML is missing.
"""
from __future__ import annotations

import statistics
import subprocess
import threading
import time

_REASONS = {
    0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
    0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
    0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
}


class ClockSampler:
    def __init__(self, device_index: int = 0, period_s: float = 0.002):
        self.idx, self.period = device_index, period_s
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.power_w, self.power_limit_w, self.power_kind = [], None, "1 s average"
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            try:
                self.power_limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(self._h) / 1000.0
            except Exception:
                self.power_limit_w = None
        except Exception:
            self._nvml = None

    def _sample(self):
        n = self._nvml
        try:
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
            try:   # board power: the instantaneous field where the driver has it (nvmlDeviceGetPowerUsage is a 1 s average)
                fv = n.nvmlDeviceGetFieldValues(self._h, [n.NVML_FI_DEV_POWER_INSTANT])[0]
                if fv.nvmlReturn != 0:
                    raise RuntimeError
                self.power_w.append(fv.value.uiVal / 1000.0)
                self.power_kind = "instant"
            except Exception:
                try:
                    self.power_w.append(n.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                except Exception:
                    pass
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(
                n, "nvmlDeviceGetCurrentClocksEventReasons") else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            for bit, name in _REASONS.items():
                if mask & bit and name != "gpu_idle":
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            time.sleep(self.period)

    def __enter__(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        elif self._nvml is None:
            self._smi_once()
        return False

    def _smi_once(self):
        try:
            out = subprocess.run(
                ["nvidia-smi", f"--id={self.idx}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            self.sm.append(float(out[0])); self.sm_max = float(out[1])
        except Exception:
            pass

    def summary(self) -> dict:
        return {
            "sm_mhz": statistics.median(self.sm) if self.sm else None,
            "sm_max_mhz": self.sm_max,
            "reasons": sorted(self.reasons),
            "samples": len(self.sm),
            "power_w": round(statistics.median(self.power_w), 1) if self.power_w else None,
            "power_w_max": round(max(self.power_w), 1) if self.power_w else None,
            "power_limit_w": self.power_limit_w,
            "power_kind": self.power_kind if self.power_w else None,
        }
