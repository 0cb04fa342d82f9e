"""yolo_ad_refine_b200: B200-native (sm_100a) implementation of the YOLO-AD-Refine detection hot path.

Host code is Python/PyTorch plumbing over libyad.so (hand-written CUDA, C ABI in include/yad.h).  No CPU fallback.
"""


__version__ = "0.1.0"
