"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Kneedle knee detector.

PARITY UNPINNED.  The reference picks its OI regularisation factor with the
third-party package `kneed` (pinned `kneed==0.8.3`, /root/reference/requirements.txt:9),
called at /root/reference/oisatgmi/optimal_interpolation.py:37-39 as
`KneeLocator(x, y, direction='increasing')` (defaults: S=1.0, curve='concave',
interp_method='interp1d', online=False).  kneed is not vendored in the
reference tree, is not installed in this image and cannot be fetched (no
network); the reference has no test or golden vector for the call.  This file
therefore restates the published algorithm (Satopaa et al. 2011, "Finding a
'Kneedle' in a Haystack", as implemented by kneed 0.8.x) and parity for the
knee index is anchored only on the reference's own call site.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU baseline may
import this module.
"""
from __future__ import annotations

import numpy as np
from scipy import interpolate
from scipy.signal import argrelextrema


class KneeLocator:
    """Restates kneed.KneeLocator for the one configuration the reference uses
    (and the other curve/direction combinations, for completeness)."""

    def __init__(self, x, y, S: float = 1.0, curve: str = "concave",
                 direction: str = "increasing", interp_method: str = "interp1d",
                 online: bool = False, polynomial_degree: int = 7):
        self.x = np.array(x)
        self.y = np.array(y)
        self.curve = curve
        self.direction = direction
        self.N = len(self.x)
        self.S = S
        self.online = online
        self.all_knees = set()
        self.all_norm_knees = set()

        # step 1: "smoothing" -- interp1d evaluated at its own knots is the identity
        if interp_method == "interp1d":
            self.Ds_y = interpolate.interp1d(self.x, self.y)(self.x)
        elif interp_method == "polynomial":
            self.Ds_y = np.poly1d(np.polyfit(self.x, self.y, polynomial_degree))(self.x)
        else:
            raise ValueError("interp_method must be 'interp1d' or 'polynomial'")

        # step 2: normalise both axes to [0, 1]
        self.x_normalized = self._normalize(self.x)
        self.y_normalized = self._normalize(self.Ds_y)

        # step 3: turn every (curve, direction) case into concave/increasing
        self.y_normalized = self._transform_y(self.y_normalized, direction, curve)
        self.y_difference = self.y_normalized - self.x_normalized
        self.x_difference = self.x_normalized.copy()

        # step 4: local extrema of the difference curve (plateaus count)
        self.maxima_indices = argrelextrema(self.y_difference, np.greater_equal)[0]
        self.x_difference_maxima = self.x_difference[self.maxima_indices]
        self.y_difference_maxima = self.y_difference[self.maxima_indices]
        self.minima_indices = argrelextrema(self.y_difference, np.less_equal)[0]

        # step 5: thresholds
        self.Tmx = self.y_difference_maxima - (
            self.S * np.abs(np.diff(self.x_normalized).mean()))

        # step 6
        self.knee, self.norm_knee = self._find_knee()
        self.elbow = self.knee

    @staticmethod
    def _normalize(a):
        return (a - min(a)) / (max(a) - min(a))

    @staticmethod
    def _transform_y(y, direction, curve):
        if direction == "decreasing":
            if curve == "concave":
                y = np.flip(y)
            elif curve == "convex":
                y = y.max() - y
        elif direction == "increasing" and curve == "convex":
            y = np.flip(y.max() - y)
        return y

    def _find_knee(self):
        if not self.maxima_indices.size:
            return None, None
        maxima_threshold_index = 0
        threshold = None
        threshold_index = None
        knee = norm_knee = None
        for i, x in enumerate(self.x_difference):
            if i < self.maxima_indices[0]:
                continue
            j = i + 1
            if x == 1.0:
                break
            if (self.maxima_indices == i).any():
                threshold = self.Tmx[maxima_threshold_index]
                threshold_index = i
                maxima_threshold_index += 1
            if (self.minima_indices == i).any():
                threshold = 0.0
            if self.y_difference[j] < threshold:
                if self.curve == "convex":
                    if self.direction == "decreasing":
                        knee = self.x[threshold_index]
                        norm_knee = self.x_normalized[threshold_index]
                    else:
                        knee = self.x[-(threshold_index + 1)]
                        norm_knee = self.x_normalized[threshold_index]
                else:
                    if self.direction == "decreasing":
                        knee = self.x[-(threshold_index + 1)]
                        norm_knee = self.x_normalized[threshold_index]
                    else:
                        knee = self.x[threshold_index]
                        norm_knee = self.x_normalized[threshold_index]
                self.all_knees.add(knee)
                self.all_norm_knees.add(norm_knee)
                if self.online is False:
                    return knee, norm_knee
        if self.all_knees == set():
            return None, None
        return knee, norm_knee
