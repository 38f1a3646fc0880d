"""Print registers / spills / stack per kernel from the ptxas logs of the last build."""
import glob
import os
import re

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = "".join(open(f).read() for f in sorted(glob.glob(os.path.join(root, "symphonypy_b200/_C/*.ptxas.log"))))
pat = (r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
       r"(\d+) bytes spill loads\n.*?Used (\d+) registers")
for m in re.finditer(pat, txt):
    print(f"{m.group(1)[:72]:72s} regs={m.group(5):>3s} stack={m.group(2):>4s} spill={m.group(3)}/{m.group(4)}")
