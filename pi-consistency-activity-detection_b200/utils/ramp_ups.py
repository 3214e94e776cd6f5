"""Host-side scalar schedules with the names the reference scripts use (``utils/ramp_ups.py``; main_ucf101.py:419
calls ``ramp_ups.exp_rampup(N_EPOCHS)`` and evaluates the returned callable once per epoch).  Zero GPU work.

Every factory returns ``f(epoch) -> float`` in [0, 1]."""
import math


def _unit(x: float) -> float:
    return 0.0 if x < 0.0 else (1.0 if x > 1.0 else float(x))


def exp_rampup(rampup_length):
    """exp(-5 (1 - t)^2) with t = epoch / rampup_length clipped to [0, 1] (Laine & Aila, arXiv:1610.02242)."""
    def ramp(epoch):
        if rampup_length <= 0 or epoch >= rampup_length:
            return 1.0
        t = _unit(epoch / float(rampup_length))
        return math.exp(-5.0 * (1.0 - t) ** 2)
    return ramp


def linear_rampup(rampup_length):
    def ramp(epoch):
        if rampup_length <= 0 or epoch >= rampup_length:
            return 1.0
        return _unit(epoch / float(rampup_length))
    return ramp


def pseudo_rampup(T1, T2):
    """0 up to epoch T1, linear to 1 at T2, 1 afterwards."""
    def ramp(epoch):
        if epoch <= T1:
            return 0.0
        return _unit((epoch - T1) / float(T2 - T1))
    return ramp


def exp_rampdown(rampdown_length, num_epochs):
    def ramp(epoch):
        start = num_epochs - rampdown_length
        if epoch < start:
            return 1.0
        half = 0.5 * (epoch - start)          # the reference's schedule runs at half speed past `start`
        return math.exp(-(half * half) / float(rampdown_length))
    return ramp


def cosine_rampdown(rampdown_length, num_epochs):
    def ramp(epoch):
        start = num_epochs - rampdown_length
        if epoch < start:
            return 1.0
        half = 0.5 * (epoch - start)
        return 0.5 * (math.cos(math.pi * half / float(rampdown_length)) + 1.0)
    return ramp


def exp_warmup(rampup_length, rampdown_length, num_epochs):
    up, down = exp_rampup(rampup_length), exp_rampdown(rampdown_length, num_epochs)
    return lambda epoch: up(epoch) * down(epoch)
