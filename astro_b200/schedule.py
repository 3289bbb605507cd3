"""reload / t schedule of a game (reference: astro/core.py:257-280,302).

In the reference both are Python-float accumulators advanced by `+ dt` every tick, so the tick
on which ships fire and the tick on which the game times out are a pure function of
(config, starting reload, starting t) and the tick index.  They are evaluated here once, with
the reference's own float64 operations in the reference's order, and handed to the device as a
bit mask + a tick index (astro_set_schedule); the device keeps only an integer tick per game.
"""
import numpy as np

from . import _native


class Schedule:
    def __init__(self, config, reload0=0.0, t0=0.0, max_ticks=_native.MAX_TICKS):
        dt, max_time, reload_time = config.dt, config.max_time, config.reload_time
        reload_, t = reload0, t0
        reloads, ts, fire = [], [], []
        k = 0
        while True:
            reloads.append(reload_)
            ts.append(t)
            if max_time <= t + dt:            # core.py:257 (checked before the reload clock moves)
                break
            if k >= max_ticks:
                raise ValueError('max_time / dt exceeds %d ticks' % max_ticks)
            next_reload = reload_ + dt        # core.py:263
            if reload_time <= next_reload:    # core.py:267
                fire.append(k)
                next_reload -= reload_time    # core.py:280
            reload_ = next_reload
            t = t + dt                        # core.py:302
            k += 1
        self.timeout_tick = k
        self.n_ticks = k + 1
        self.fire_ticks = fire
        self.reload = np.array(reloads, dtype=np.float64)
        self.t = np.array(ts, dtype=np.float64)
        bits = np.zeros((self.n_ticks + 31) // 32, dtype=np.uint32)
        for f in fire:
            bits[f >> 5] |= np.uint32(1 << (f & 31))
        self.fire_bits = bits

    def tick_of(self, reload_, t):
        """Tick index whose (reload, t) equal the given values bit for bit, or None."""
        idx = np.nonzero((self.t == t) & (self.reload == reload_))[0]
        return int(idx[0]) if idx.size else None
