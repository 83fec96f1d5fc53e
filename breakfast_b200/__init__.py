"""breakfast_b200 — B200-native (sm_100a) implementation of breakfast's distance-and-clustering hot path.

Python host code mirrors the reference's interface (breakfast.console / breakfast.breakfast /
breakfast.cache); the pairwise-distance, thresholding and connected-components work runs in
hand-written CUDA kernels behind the C ABI of include/breakfast_b200.h (libbreakfast_b200.so).
There is no CPU fallback.
"""

__version__ = "0.1.0"
