#!/usr/bin/env python3
"""sha256 of `phantom.make_pair_exact` for the fixture shapes, on one device:  python tools/check_exact_phantom.py cpu|cuda [ref.json]
With a reference JSON (written by an earlier run on another device) the hashes are compared."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dicomautomaton_b200 import phantom  # noqa: E402

dev = sys.argv[1]
out = {}
for shape in ((128, 128, 128), (128, 256, 256), (256, 512, 512)):
    f, m = phantom.make_pair_exact(shape, device=dev)
    out["x".join(map(str, shape))] = [hashlib.sha256(np.ascontiguousarray(t.cpu().numpy()).view(np.uint8).reshape(-1)).hexdigest()
                                      for t in (f, m)]
print(json.dumps(out))
if len(sys.argv) > 2:
    ref = json.load(open(sys.argv[2]))
    ok = all(ref[k] == v for k, v in out.items())
    print("MATCH" if ok else "MISMATCH", {k: ref[k] == v for k, v in out.items()})
    sys.exit(0 if ok else 1)
