"""Integral diagnostics of a displacement run.

The reference records only iteration counts, dt, mass / boundary-velocity / divergence errors and the cumulative
produced mass (src/Impec2DSolver.h:16-25).  The north-star additionally compares fingering runs on breakthrough time,
recovered mass and mixing-zone length; those are DEFINED here, once, and computed by this one post-processor on both
the oracle's and the GPU path's fields so that the comparison cannot be skewed by two implementations.

  produced_fraction(c, q)   |q|-weighted mean concentration over the production cells (q < 0): what the wells produce
  breakthrough_time         first time produced_fraction reaches `threshold` (default 1 %), linearly interpolated
                            between the two bracketing steps
  recovered_mass            injected mass minus produced mass minus nothing else: simply sum(c) dx dy (what is still
                            in place) reported next to the produced mass accumulator of the solver
  mixing_zone_length        area of {lo <= c <= hi} (default 0.1..0.9) divided by the injector-producer distance:
                            the mean width of the transition zone across the displacement direction
"""
import numpy as np


def produced_fraction(c_int, q):
    q = np.asarray(q).reshape(c_int.shape)
    w = np.where(q < 0, -q, 0.0)
    s = w.sum()
    return float((w * c_int).sum() / s) if s > 0 else 0.0


def mixing_zone_length(c_int, dx, dy, lx, ly, lo=0.1, hi=0.9):
    area = float(((c_int >= lo) & (c_int <= hi)).sum()) * dx * dy
    return area / float(np.hypot(lx, ly))


def mass_in_place(c_int, dx, dy):
    return float(c_int.sum()) * dx * dy


class RunDiagnostics:
    """Accumulates the per-step series; feed it the interior concentration after every step."""

    def __init__(self, q, dx, dy, lx, ly, threshold=0.01):
        self.q = np.asarray(q); self.dx = dx; self.dy = dy; self.lx = lx; self.ly = ly; self.threshold = threshold
        self.t = []; self.frac = []; self.mix = []; self.mass = []

    def record(self, t, c_int):
        self.t.append(float(t)); self.frac.append(produced_fraction(c_int, self.q))
        self.mix.append(mixing_zone_length(c_int, self.dx, self.dy, self.lx, self.ly)); self.mass.append(mass_in_place(c_int, self.dx, self.dy))

    def record_device(self, t, ctx):
        """The same three series from the DEVICE reductions of impec_integral_diagnostics: nothing but four scalars crosses PCIe."""
        d = ctx.integral_diagnostics()
        self.t.append(float(t)); self.frac.append(d.produced_fraction); self.mix.append(d.mixing_zone_length); self.mass.append(d.mass_in_place)

    def breakthrough_time(self):
        f = np.asarray(self.frac); t = np.asarray(self.t)
        k = np.nonzero(f >= self.threshold)[0]
        if len(k) == 0:
            return None
        k = int(k[0])
        if k == 0:
            return float(t[0])
        return float(t[k - 1] + (self.threshold - f[k - 1]) / (f[k] - f[k - 1]) * (t[k] - t[k - 1]))

    def summary(self):
        return {"breakthrough_time": self.breakthrough_time(), "mixing_zone_length": self.mix[-1] if self.mix else None,
                "mass_in_place": self.mass[-1] if self.mass else None, "produced_fraction": self.frac[-1] if self.frac else None}
