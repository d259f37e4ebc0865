#!/usr/bin/env python
"""Generate tests/golden/track_global.npz: Map.getGlobalPosition (src/fnc/simulator/Track.py:135-189) of the REAL reference at
probe points (s, ey), including s beyond one track length (the wrap at Track.py:141-142) and the segment boundaries.
Run in the build container (imports /root/reference/src/fnc/simulator/Track.py; NumPy only)."""
import os
import sys
import numpy as np

sys.path.insert(0, "/root/reference/src/fnc/simulator")
from Track import Map      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
m = Map(0.4)
rng = np.random.default_rng(12)
L = m.TrackLength
s = np.concatenate([np.linspace(0.0, 2.0 * L, 500)[:-1], m.PointAndTangent[:, 3] + 1e-9, m.PointAndTangent[:, 3] + m.PointAndTangent[:, 4] - 1e-9,
                    rng.uniform(0.0, 2.0 * L, 300)])
s = s[(np.mod(s, L) > 1e-12) & (s != L) & (s != 2 * L)]
ey = np.concatenate([np.zeros(200), rng.uniform(-0.4, 0.4, s.shape[0] - 200)])
xy = np.array([m.getGlobalPosition(float(a), float(b)) for a, b in zip(s, ey)])
np.savez_compressed(os.path.join(OUT, "track_global.npz"), s=s, ey=ey, xy=xy, table=m.PointAndTangent, track_length=np.array(L))
print("points", s.shape[0], "xy range", xy.min(axis=0), xy.max(axis=0))
