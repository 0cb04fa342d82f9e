"""TEST INFRASTRUCTURE -- re-export of the deterministic synthetic weight / input generator (it lives in the product package because
bench.py needs random-init weights of the reference architecture without touching oracle/)."""
from yolo_ad_refine_b200.synth import *  # noqa: F401,F403
from yolo_ad_refine_b200.synth import SPEC_PATH, _draw, _rs  # noqa: F401
