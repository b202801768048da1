"""Closed-form areas used by the reference's analytic known-answer tests
(reference tests/test_freesasa.c:27-43): two intersecting spheres."""
import math


def hidden_two_spheres(r1, r2, d):
    if d > r1 + r2:
        return 0.0
    if r1 + d < r2:
        return 4 * math.pi * r1 * r1
    if r2 + d < r1:
        return 4 * math.pi * r2 * r2
    return math.pi / d * (r1 * (r2 * r2 - (d - r1) ** 2) + r2 * (r1 * r1 - (d - r2) ** 2))


def surface_two_spheres(x1, x2, r1, r2, probe):
    d = math.dist(x1, x2)
    R1, R2 = r1 + probe, r2 + probe
    return 4 * math.pi * (R1 * R1 + R2 * R2) - hidden_two_spheres(R1, R2, d)


def rel_err(a, b):
    return abs(a - b) / (abs(a) + abs(b))


TWO_SPHERE_CASES = [
    ([0.0, 0, 0], [2.0, 0, 0]),
    ([0.0, 0, 0], [0.0, 2, 0]),
    ([0.0, 0, 0], [0.0, 0, 2]),
]

_S = math.sqrt(2.0)
# four coplanar spheres: original, translated, rotated 90 deg about z, -45 deg about z, 90 deg about x
FOUR_SPHERE_RADII = [1.0, 1.0, 2.0, 1.0]
FOUR_SPHERE_POSES = [
    [0, 0, 0, 1, 0, 0, 0, 1, 0, 1, 1, 0],
    [1, 1, 1, 2, 1, 1, 1, 2, 1, 2, 2, 1],
    [0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 0, 0],
    [-1 / _S, 1 / _S, 0, 0, 0, 0, 0, _S, 0, 1 / _S, 1 / _S, 0],
    [-1 / _S, 0, 1 / _S, 0, 0, 0, 0, 0, _S, 1 / _S, 0, 1 / _S],
]
