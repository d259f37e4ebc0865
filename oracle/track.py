"""Oracle restatement of the race-track table and curvature lookup.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows src/fnc/simulator/Track.py:
  * segment table construction   Track.py:31-133  (L-shaped track, halfWidth fixed 0.4)
  * ``curvature(s)``             Track.py:292-310
  * ``global_position(s, ey)``   Track.py:135-189  (curvilinear -> inertial frame; presentation only)
Only columns 3:6 of the table (cumulative s, length, signed curvature) and
``TrackLength`` are consumed by the hot path (PredictiveModel.py:95-96).
"""
import numpy as np


def _wrap(a):
    # Track.py:361-369
    if a < -np.pi:
        return 2 * np.pi + a
    if a > np.pi:
        return a - 2 * np.pi
    return a


def _sgn(a):
    # Track.py:371-377
    return 1 if a >= 0 else -1


class TrackTable:
    """Rows = [x_end, y_end, psi_end, s_start, length, curvature] per segment."""

    def __init__(self, half_width=0.4):
        # Track.py:31 — the constructor argument is ignored by the reference.
        self.halfWidth = 0.4
        self.slack = 0.45
        lc = 4.5
        segs = [(1.0, 0.0), (lc, lc / np.pi), (lc / 2, -lc / np.pi),
                (lc, lc / np.pi), (lc / np.pi * 2, 0.0), (lc / 2, lc / np.pi)]
        tab = np.zeros((len(segs) + 1, 6))
        for i, (length, radius) in enumerate(segs):
            if i == 0:
                ang0, px, py, s0 = 0.0, 0.0, 0.0, tab[0, 3]
            else:
                ang0, px, py = tab[i - 1, 2], tab[i - 1, 0], tab[i - 1, 1]
                s0 = tab[i - 1, 3] + tab[i - 1, 4]
            if radius == 0.0:                      # straight (Track.py:56-74)
                x = px + length * np.cos(ang0)
                y = py + length * np.sin(ang0)
                tab[i] = [x, y, ang0, s0, length, 0.0]
            else:                                  # arc (Track.py:75-114)
                direction = 1 if radius >= 0 else -1
                cx = px + np.abs(radius) * np.cos(ang0 + direction * np.pi / 2)
                cy = py + np.abs(radius) * np.sin(ang0 + direction * np.pi / 2)
                span = length / np.abs(radius)
                psi = _wrap(ang0 + span * np.sign(radius))
                normal = _wrap(direction * np.pi / 2 + ang0)
                a = -(np.pi - np.abs(normal)) * _sgn(normal)
                x = cx + np.abs(radius) * np.cos(a + direction * span)
                y = cy + np.abs(radius) * np.sin(a + direction * span)
                tab[i] = [x, y, psi, s0, length, 1.0 / radius]
        # closing straight back to the origin (Track.py:118-129)
        xs, ys = tab[-2, 0], tab[-2, 1]
        closing = np.sqrt((0.0 - xs) ** 2 + (0.0 - ys) ** 2)
        tab[-1] = [0.0, 0.0, 0.0, tab[-2, 3] + tab[-2, 4], closing, 0.0]
        self.PointAndTangent = tab
        self.TrackLength = tab[-1, 3] + tab[-1, 4]

    def seg_table(self):
        """(nseg,3) array [s_start, length, curvature] — what the GPU gets."""
        return np.ascontiguousarray(self.PointAndTangent[:, 3:6])

    def curvature(self, s):
        """Track.py:292-310: wrap by repeated subtraction, first matching segment."""
        tab = self.PointAndTangent
        L = tab[-1, 3] + tab[-1, 4]
        while s > L:
            s = s - L
        hit = np.logical_and(s >= tab[:, 3], s < tab[:, 3] + tab[:, 4])
        idx = np.where(hit)[0]
        # the reference does int(np.where(...)[0]) which raises unless exactly one hit
        if idx.shape[0] != 1:
            raise ValueError("curvature: s=%r matches %d segments" % (s, idx.shape[0]))
        return tab[int(idx[0]), 5]

    # Track.py:135-189
    def global_position(self, s, ey):
        while s > self.TrackLength:
            s = s - self.TrackLength
        pt = self.PointAndTangent
        hit = np.where((s >= pt[:, 3]) & (s < pt[:, 3] + pt[:, 4]))[0]
        i = int(hit[0])                     # the reference takes int() of a single hit and raises otherwise
        if pt[i, 5] == 0.0:                 # straight segment: interpolate between its end points
            xf, yf, psi = pt[i, 0], pt[i, 1], pt[i, 2]
            xs, ys = pt[i - 1, 0], pt[i - 1, 1]
            rel = (s - pt[i, 3]) / pt[i, 4]
            return ((1 - rel) * xs + rel * xf + ey * np.cos(psi + np.pi / 2),
                    (1 - rel) * ys + rel * yf + ey * np.sin(psi + np.pi / 2))
        r = 1 / pt[i, 5]
        ang = pt[i - 1, 2]
        d = 1 if r >= 0 else -1
        cx = pt[i - 1, 0] + np.abs(r) * np.cos(ang + d * np.pi / 2)
        cy = pt[i - 1, 1] + np.abs(r) * np.sin(ang + d * np.pi / 2)
        span = (s - pt[i, 3]) / (np.pi * np.abs(r)) * np.pi
        an = _wrap(d * np.pi / 2 + ang)
        angle = -(np.pi - np.abs(an)) * _sgn(an)
        return (cx + (np.abs(r) - d * ey) * np.cos(angle + d * span), cy + (np.abs(r) - d * ey) * np.sin(angle + d * span))

