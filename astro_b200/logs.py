"""Game logs of batched rollouts in the reference's JSONL format.

The reference records a game as `Game(config, winner, ticks)` with one `Tick(state, control,
reward, bot_data)` per step, `state` being the PRE-step state (core.play, core.py:377-410), and
writes it with `save_log` (core.py:413-427; util.to_jsonable, util.py:13-30): a header line
{config, winner}, then one Tick per line, arrays as {_values, _shape}, namedtuples tagged
`_type: astro.core:<Name>` — the format `static/astro.js:82-112` replays.

`GameRecorder` follows a handful of game slots of a `BatchedGames` rollout: before every tick it
gathers their states from HBM (a few hundred bytes per followed game, nothing else leaves the
GPU), after the tick their rewards; a followed game that ends becomes a `core.Game` (and, with a
folder given, a .jsonl file), and with auto-reset the slot's next episode starts a new one.
"""
import os

import numpy as np

from . import core


class GameRecorder:
    def __init__(self, games, indices, folder=None, prefix='game', bot_data=None):
        self.games = games
        self.indices = [int(i) for i in indices]
        self.folder = folder
        self.prefix = prefix
        self.bot_data = bot_data if bot_data is not None else [None] * games.S
        self.finished = []          # completed core.Game objects, in completion order
        self.paths = []             # files written (when folder is given)
        self._ticks = {i: [] for i in self.indices}
        self._pending = None
        self._episode = {i: 0 for i in self.indices}

    def before_step(self, actions=None):
        """Call right before games.step(actions): records the pre-step state and the controls.
        actions -- what step() receives: [n, S] array / cuda tensor of control codes, or None for
        the device counter stream (the controls are then regenerated on the host, rng.actions)."""
        g = self.games
        states = g.get_states(self.indices)
        if actions is None:
            from . import rng
            ctl = rng.actions(g.seed, g.first_game + np.asarray(self.indices), g.step_index, g.S)
        else:
            a = actions[self.indices] if isinstance(actions, np.ndarray) else actions[self.indices].cpu().numpy()
            ctl = np.asarray(a, dtype=np.int64).reshape(len(self.indices), g.S)
        self._pending = (states, ctl)

    def after_step(self, reward, done):
        """Call right after games.step(...) with its (reward, done) results."""
        states, ctl = self._pending
        self._pending = None
        r = reward[self.indices].cpu().numpy() if hasattr(reward, 'cpu') else np.asarray(reward)[self.indices]
        d = done[self.indices].cpu().numpy() if hasattr(done, 'cpu') else np.asarray(done)[self.indices]
        for j, i in enumerate(self.indices):
            if states[j] is None:       # the slot held a finished game: nothing happened
                continue
            rew = np.asarray(r[j], dtype=np.float32)
            self._ticks[i].append(core.Tick(state=states[j], control=ctl[j].copy(), reward=rew, bot_data=list(self.bot_data)))
            if d[j]:
                winner = None if np.max(rew) < 1 else int(np.argmax(rew))   # core.py:409
                game = core.Game(config=self.games.config, winner=winner, ticks=self._ticks[i])
                self.finished.append(game)
                if self.folder is not None:
                    path = os.path.join(self.folder, '%s_%d_%d.jsonl' % (self.prefix, i, self._episode[i]))
                    core.save_log(path, game)
                    self.paths.append(path)
                self._ticks[i] = []
                self._episode[i] += 1

    def step(self, actions=None, **kw):
        """before_step + games.step + after_step in one call; returns what games.step returns."""
        self.before_step(actions)
        out = self.games.step(actions, **kw)
        self.after_step(out[0], out[1])
        return out
