#!/usr/bin/env python
"""Regenerates tests/golden/window_tiny_v1.blob (the byte-for-byte pin of the window wire format, version 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mc_slam_b200 import synth  # noqa: E402

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "window_tiny_v1.blob")
with open(out, "wb") as f:
    f.write(synth.make_config("tiny").to_bytes())
print(out, os.path.getsize(out), "bytes")
