#!/usr/bin/env python3
"""tests/golden/unittest_vectors.json: what the reference's own unit-test programs compute on their own inputs
(oracle/_ref/unittest_probe, built by `make -C oracle ref` from /root/reference/unittest/*.cpp where they lie).
Usage: python tools/make_unittest_vectors.py   (only where /root/reference is mounted)"""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "unittest_probe")], check=True, capture_output=True, text=True).stdout
vec = json.loads(out)
vec["source"] = "unittest/test_tp_algos.cpp:12-32,104-121 and unittest/test_effective_sinr.cpp:227-234 of the reference"
path = os.path.join(ROOT, "tests", "golden", "unittest_vectors.json")
json.dump(vec, open(path, "w"), indent=1)
print("wrote", path)
