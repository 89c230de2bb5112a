#!/usr/bin/env python
"""Per-kernel SASS statistics of libsplendor_b200.so (instruction count, local/global/shared memory ops)."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "splendor_gym_b200/libsplendor_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for p in re.split(r"\n\s*Function : ", txt)[1:]:
    name = p.split("\n", 1)[0]
    n = len(re.findall(r"/\*[0-9a-f]{4,5}\*/\s+[@A-Z]", p))
    c = lambda k: len(re.findall(r"\b" + k, p))
    print(f"{name[:58]:58s} inst={n:5d} STL={c('STL'):3d} LDL={c('LDL'):3d} LDG={c('LDG'):3d} STG={c('STG'):3d} LDS={c('LDS'):3d} STS={c('STS'):3d} BRA={c('BRA'):4d} CALL={c('CALL')}")
