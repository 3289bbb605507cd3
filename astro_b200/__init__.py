"""astro_b200 — B200-native batched implementation of the Astro game tick
(DouglasOrr/Astro: astro/core.py step + astro/rl.py feature extraction).

    from astro_b200 import core          # drop-in for astro.core (create/step/roll_ships/play…)
    from astro_b200 import rl            # get_features / to_batch / get_features_batch
    from astro_b200.batched import BatchedGames   # N games in HBM, one CUDA launch per tick
"""
__version__ = '0.1.0'
