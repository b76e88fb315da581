"""LR lambda used by bbox.yaml's scheduler_config (reference ldm/lr_scheduler.py:80-98). Training is out of scope for
the B200 hot path; the class exists so the config block instantiates unchanged."""
import numpy as np


class LambdaLinearScheduler:
    def __init__(self, warm_up_steps, f_min, f_max, f_start, cycle_lengths, verbosity_interval=0):
        assert len(warm_up_steps) == len(f_min) == len(f_max) == len(f_start) == len(cycle_lengths)
        self.lr_warm_up_steps, self.f_start, self.f_min, self.f_max = warm_up_steps, f_start, f_min, f_max
        self.cycle_lengths = cycle_lengths
        self.cum_cycles = np.cumsum([0] + list(cycle_lengths))
        self.last_f = 0.0

    def find_in_interval(self, n):
        for i, cl in enumerate(self.cum_cycles[1:]):
            if n <= cl:
                return i
        return len(self.cycle_lengths) - 1

    def schedule(self, n, **kwargs):
        c = self.find_in_interval(n)
        n = n - self.cum_cycles[c]
        if n < self.lr_warm_up_steps[c]:
            f = (self.f_max[c] - self.f_start[c]) / self.lr_warm_up_steps[c] * n + self.f_start[c]
        else:
            f = self.f_min[c] + (self.f_max[c] - self.f_min[c]) * (self.cycle_lengths[c] - n) / self.cycle_lengths[c]
        self.last_f = f
        return f

    def __call__(self, n, **kwargs):
        return self.schedule(n, **kwargs)
