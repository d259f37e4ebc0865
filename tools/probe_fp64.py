#!/usr/bin/env python
"""tools/probe_fp64.py — print the measured fp64 numbers of cuda:0 (csrc/probe.cuh via lmpc_probe_fp64) as one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racinglmpc_b200 import _native
print(json.dumps(_native.probe_fp64(int(sys.argv[1]) if len(sys.argv) > 1 else 0)))
