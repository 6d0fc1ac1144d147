"""Hot-path constants (values from the reference's ffsubsync/constants.py:7-19)."""
from typing import List

import numpy as np

SAMPLE_RATE: int = 100                      # 10 ms windows (constants.py:7)
FRAMERATE_RATIOS: List[float] = [24.0 / 23.976, 25.0 / 23.976, 25.0 / 24.0]  # constants.py:9
DEFAULT_FRAME_RATE: int = 48000             # ffmpeg decode rate (constants.py:11)
DEFAULT_NON_SPEECH_LABEL: float = 0.0       # constants.py:12
DEFAULT_START_SECONDS: int = 0
DEFAULT_SCALE_FACTOR: float = 1
DEFAULT_MAX_OFFSET_SECONDS: int = 60        # constants.py:18
DEFAULT_VAD: str = "energy_zcr"             # the detector this package implements

# energy / zero-crossing detector defaults: auditok's energy_threshold=50 dB
# (speech_transformers.py:125) is mean(x^2) >= 1e5
DEFAULT_ENERGY_THRESHOLD: int = 100000


def framerate_ratios_to_try(no_fix_framerate: bool = False, gss: bool = False) -> list:
    """The candidate list try_sync builds (ffsubsync/ffsubsync.py:131-142): the ratios and their
    inverses as float64; ``None`` stands for the golden-section search entry."""
    if no_fix_framerate:
        return []
    r = np.array(FRAMERATE_RATIOS)
    out = list(np.concatenate([r, 1.0 / r]))
    if gss:
        out.append(None)
    return out
