"""Summarise `nvcc -Xptxas -v` logs (build/*.ptxas.log): registers, stack, spills per kernel."""
import glob, re, sys, os
here = os.path.dirname(os.path.abspath(__file__))
rows = []
for path in sorted(glob.glob(os.path.join(here, "build", "*.ptxas.log"))):
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?Function properties for \1\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt, re.S):
        name = m.group(1)
        dm = re.match(r"_Z\d+([a-z_0-9]+?)(?:ILi(\d+)E(?:Li(\d+)E)?E?)?(?:v|$)", name)
        label = dm.group(1) + ("<%s%s>" % (dm.group(2), "," + dm.group(3) if dm.group(3) else "") if dm.group(2) else "")
        rows.append((label, int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
print("%-28s %5s %6s %8s %8s" % ("kernel", "regs", "stack", "spill_st", "spill_ld"))
for r in rows:
    print("%-28s %5d %6d %8d %8d" % r)
