"""Host-side mirror of the reference's plain data classes (names and argument meaning as in
/root/reference/envs/atc/model.py:132-145 SimParameters, :309-315 EntryPoint).  No simulation logic lives here:
the aircraft model runs in csrc/atc_kernels.cu."""
from typing import List


class SimParameters(object):
    """model.py:132-145.  `precision` is accepted and unused, exactly like the reference."""

    def __init__(self, timestep: float, precision: float = 0.5, reward_shaping: bool = True,
                 normalize_state: bool = True, discrete_action_space: bool = False):
        if not timestep > 0:
            raise ValueError("timestep must be > 0")
        self.timestep = timestep
        self.precision = precision
        self.reward_shaping = reward_shaping
        self.normalize_state = normalize_state
        self.discrete_action_space = discrete_action_space


class EntryPoint(object):
    """model.py:309-315"""

    def __init__(self, x: float, y: float, phi: int, levels: List[int]):
        self.x = x
        self.y = y
        self.phi = phi
        self.levels = list(levels)


class MinimumVectoringAltitude(object):
    """model.py:260-268 — `area` is the closed ring [[x, y], ...] in the order given (no shapely needed)."""

    def __init__(self, area, height: int):
        ring = [(float(p[0]), float(p[1])) for p in area]
        if ring[0] != ring[-1]:
            ring.append(ring[0])
        self.area_as_list = ring
        self.height = height
        xs = [p[0] for p in ring]
        ys = [p[1] for p in ring]
        self.outer_bounds = (min(xs), min(ys), max(xs), max(ys))


class Runway(object):
    """model.py:234-257 (data only; the corridor geometry is derived in sector.py)."""

    def __init__(self, x, y, h, phi):
        self.x = x
        self.y = y
        self.h = h
        self.phi_from_runway = phi
        self.phi_to_runway = (phi + 180) % 360


# aircraft performance limits (model.py:14, 45-50) — the kernel hard-codes the same numbers
H_MIN, H_MAX, V_MIN, V_MAX = 0, 38000, 100, 300
