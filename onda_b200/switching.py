"""Host-side switch statistics: sliding-window monitor and the static/dynamic selectors.

This is the control logic that reads the per-step confidence scalars produced by the CUDA
kernels.  It stays on the host, with the reference's semantics:

* ``Monitor``       -- framework/utils/monitoring.py:7-96
* ``HybridSelect``  -- model_select in prototypes_hybrid_switch.py:5-34
* ``DevSelect``     -- model_select in prototypes_vswitch.py:5-25
* ``static_share``  -- the h-switch ramp, prototypes_hswitch.py:45-55

Values handed to ``Monitor.add`` by this package are Python floats (the reference's
v/h-switch variants append 0-dim CUDA tensors, which breaks ``dev_avg`` once the window is
full -- SURVEY.md appendix B).
"""
from __future__ import annotations

from statistics import median

import numpy as np


class Monitor:
    """Per-key window of the last ``limit`` samples plus an exponential average.

    ``avg`` is the window *median*, ``exp`` the EMA with constant ``exp_const`` and
    ``dev_avg`` a weighted first difference that stays 0 until the window is full.
    Missing keys read as 1 (avg/exp) or 0 (dev_avg); ``add`` is a no-op while frozen.
    """

    def __init__(self, limit=None, exp_const=0.01, dev_func="hamming"):
        self.limit = limit
        self.exp_const = exp_const
        self.freeze = False
        self.current_dict = {}
        self.exp_dict = {}
        self.signal = np.hamming(limit - 1) if limit is not None else None
        self.signal_sum = float(np.sum(self.signal)) if limit is not None else None
        if dev_func == "median":
            self.mean_func = median
        elif dev_func == "mean":
            self.mean_func = lambda x: np.mean(np.array(x))
        elif dev_func == "hamming":
            self.mean_func = lambda x: np.sum(self.signal * np.array(x)) / self.signal_sum
        else:
            self.mean_func = None  # the reference leaves it undefined for other names

    def eval(self):
        self.freeze = True

    def train(self):
        self.freeze = False

    def add(self, values, reset=False):
        if self.freeze:
            return 0
        a = self.exp_const
        for key, val in values.items():
            window = self.current_dict.get(key)
            if window is None or reset:
                self.current_dict[key] = [val]
                self.exp_dict[key] = val
                continue
            window.append(val)
            if self.limit is not None and len(window) > self.limit:
                window.pop(0)
            self.exp_dict[key] = (1 - a) * self.exp_dict[key] + a * val

    def _dev_avg(self, item):
        window = self.current_dict.get(item)
        if window is None or len(window) < self.limit:
            return 0
        return self.mean_func(window[1:]) - self.mean_func(window[:-1])

    def dev_avg(self, item=None):
        if item is not None:
            return self._dev_avg(item)
        return {key: self._dev_avg(key) for key in self.current_dict}

    def exp(self, item=None):
        if item is None:
            return self.exp_dict
        return self.exp_dict.get(item, 1)

    def avg(self, item=None):
        if item is None:
            return {key: median(vals) for key, vals in self.current_dict.items()}
        window = self.current_dict.get(item)
        return median(window) if window is not None else 1

    def reset(self):
        self.current_dict = {}


class HybridSelect:
    """Confidence + confidence-derivative selector with a gray area."""

    static = 0
    dynamic = 1

    def __init__(self, start=0, gray_area=(0.84, 0.88), dev_threshold=0.0002):
        self.current = start
        self.current_dev = start
        self.freeze = False
        self.gray_area = gray_area
        self.dev_threshold = dev_threshold

    def eval(self):
        self.freeze = True

    def train(self):
        self.freeze = False

    def evaluate(self, confidence, dev_value):
        if self.freeze:
            return
        if dev_value > self.dev_threshold:
            self.current_dev = self.static
        elif dev_value < -self.dev_threshold:
            self.current_dev = self.dynamic
        low, high = self.gray_area[0], self.gray_area[1]
        if confidence < low:
            self.current = self.dynamic
        elif confidence > high:
            self.current = self.static
        else:
            self.current = self.current_dev


class DevSelect:
    """Confidence-derivative selector."""

    static = 0
    dynamic = 1

    def __init__(self, start=0, threshold_c=0.00028):
        self.current = start
        self.freeze = False
        self.threshold = threshold_c

    def eval(self):
        self.freeze = True

    def train(self):
        self.freeze = False

    def evaluate(self, dev_value):
        if self.freeze:
            return
        if dev_value > self.threshold:
            self.current = self.static
        elif dev_value < -self.threshold:
            self.current = self.dynamic


def static_share(median_static, soft_trans, switch_prior_thresh=0.0):
    """Fraction of the prior taken from the static model by the h-switch (prototypes_hswitch.py:45-55).

    The reference's Monitor holds 0-dim float32 tensors there, so its ramp ``vl * (25/3) - 41/6`` and its comparison
    with the threshold are float32 operations; reproduced with numpy float32 scalars."""
    v = np.float32(median_static)
    if soft_trans:
        ramp = np.float32(v * np.float32(25.0 / 3)) - np.float32(41.0 / 6)
        return float(max(min(ramp, np.float32(1)), np.float32(0)))
    return int(v > np.float32(switch_prior_thresh))


def below_f32(value, threshold):
    """``value < threshold`` the way the reference evaluates it on a 0-dim float32 tensor (prototypes.py:231-233)."""
    return bool(np.float32(value) < np.float32(threshold))
