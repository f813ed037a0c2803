"""Build experimental variants of libxcb200 (extra -D flags) for A/B timing on the GPU box.
usage: python scripts/build_variants.py name1:D1,D2 name2:D3 ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xcontour_b200 import build
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    print(build.build(variant=name, defines=[d for d in defs.split(",") if d]))
