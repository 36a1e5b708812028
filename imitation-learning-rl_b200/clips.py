"""Motion-clip tables (the CSVs read at REF low_level_env.py:58-71 / hier_env.py:61-80), packed by tools/gen_clips.py."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CLIP_NAMES = ["motion02_04", "motion08_03", "motion09_03", "motion13_13"]
_cache = {}


def load_clip(name):
    """-> dict(pos, rel, vel, ep: float32 C-contiguous; max_frame).

    max_frame = len(JointPosRad) - 1 as at REF low_level_env.py:80-82, clamped to the rows that exist in the velocity
    table: motion13_13 ships 120 velocity rows for 220 position rows, so the reference raises IndexError beyond frame
    119 (declared divergence)."""
    if name not in _cache:
        if name not in CLIP_NAMES:
            raise KeyError("unknown motion clip %r (have %s)" % (name, CLIP_NAMES))
        with np.load(os.path.join(HERE, "data", "clips.npz")) as z:
            d = {k: np.ascontiguousarray(z["%s_%s" % (name, k)], dtype=np.float32) for k in ("pos", "rel", "vel", "ep")}
        d["max_frame"] = int(min(len(d["pos"]) - 1, len(d["vel"])))
        _cache[name] = d
    return _cache[name]
