"""NVML clock record for scratch benchmark lines (bench.py has its own sampler thread)."""
import threading
import time


class Clocks:
    NAMES = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample_while(self, fn):
        """run fn() while a thread samples the SM clock every 5 ms; returns (fn result, clock record)"""
        if self.nv is None:
            return fn(), None
        samples, reasons, stop = [], set(), threading.Event()

        def loop():
            while not stop.is_set():
                try:
                    samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    for k, bit in self.NAMES.items():
                        if r & bit:
                            reasons.add(k)
                except Exception:
                    pass
                time.sleep(0.005)
        th = threading.Thread(target=loop, daemon=True)
        th.start()
        try:
            res = fn()
        finally:
            stop.set()
            th.join()
        s = sorted(samples)
        return res, {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(s)}
