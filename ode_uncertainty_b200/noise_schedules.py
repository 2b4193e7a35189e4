"""Process-noise tempering schedules: gamma per tempering stage, a host scalar handed to the
kernel as `gamma_sqrt` (mirror of the reference's src/noise_schedules.py:9-130; same class
names, constructor arguments and `step(idx)` meaning)."""
from __future__ import annotations

import math


class NoiseSchedule:
    """src/noise_schedules.py:9-32."""

    def __init__(self, init_noise_log: float) -> None:
        self.init_noise_log = init_noise_log

    def step(self, idx: int) -> float:
        raise NotImplementedError


class LinearDecaySchedule(NoiseSchedule):
    """gamma = 10^(init - idx * rate)  (src/noise_schedules.py:35-61)."""

    def __init__(self, init_noise_log: float = 0.0, decay_rate: float = 1.0) -> None:
        super().__init__(init_noise_log)
        self.decay_rate = decay_rate

    def step(self, idx: int) -> float:
        return 10.0 ** (self.init_noise_log - idx * self.decay_rate)


class ExponentialDecaySchedule(NoiseSchedule):
    """gamma = 10^(init - rate * log10(idx + 1))  (src/noise_schedules.py:64-90)."""

    def __init__(self, init_noise_log: float = 0.0, decay_rate: float = 8.0) -> None:
        super().__init__(init_noise_log)
        self.decay_rate = decay_rate

    def step(self, idx: int) -> float:
        return 10.0 ** (self.init_noise_log - self.decay_rate * math.log10(idx + 1))


class CosineAnnealingSchedule(NoiseSchedule):
    """Cosine annealing with restarts in log10 space (src/noise_schedules.py:93-130)."""

    def __init__(self, init_noise_log: float = 0.0, min_noise_log: float = -10.0,
                 cycle_length: int = 4) -> None:
        super().__init__(init_noise_log)
        self.min_noise_log = min_noise_log
        self.cycle_length = cycle_length

    def step(self, idx: int) -> float:
        i = idx % self.cycle_length
        return 10.0 ** (self.min_noise_log + 0.5 * (self.init_noise_log - self.min_noise_log)
                        * (1.0 + math.cos(i / (self.cycle_length - 1) * math.pi)))
