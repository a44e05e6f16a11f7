"""TEST INFRASTRUCTURE -- the synthetic-input generator lives in tools/synth.py (bench.py uses it too, and nothing
under oracle/ may be on the bench's measured path); the tests keep importing it under this name."""
from tools.synth import *  # noqa: F401,F403
from tools.synth import chord_lines, knn_triplets, make_pair, random_rotation, surface_points  # noqa: F401
