#!/usr/bin/env python3
"""Build a tuning variant of the library:  python tools/build_variant.py <tag> DEFINE[=VALUE] ..."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dicomautomaton_b200 import build
print(build.build_variant(sys.argv[1], sys.argv[2:], verbose=True))
