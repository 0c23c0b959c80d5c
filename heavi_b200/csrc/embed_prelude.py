"""Writes build/hvb_prelude.inc: plan_format.h + hvb_device.cuh as one C++ raw string literal
(include lines and `#pragma once` dropped: NVRTC compiles the text without a file system)."""
import sys

out, srcs = sys.argv[1], sys.argv[2:]
text = []
for path in srcs:
    for line in open(path):
        s = line.strip()
        if s.startswith("#include") or s.startswith("#pragma once"):
            continue
        text.append(line)
body = "".join(text)
assert ')HVBPRE"' not in body
open(out, "w").write('R"HVBPRE(\n' + body + ')HVBPRE"\n')
