#!/usr/bin/env python
"""Print selected keys of the last JSON line of a bench.py log (development aid)."""
import json
import sys

path, keys = sys.argv[1], sys.argv[2:]
line = [x for x in open(path) if x.lstrip().startswith("{")][-1]
d = json.loads(line)
for k in keys:
    v = d
    for part in k.split("."):
        v = v.get(part) if isinstance(v, dict) else None
    print(k, json.dumps(v)[:3000])
