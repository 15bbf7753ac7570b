"""Static SASS instruction count per kernel of an object file / library.
    python scripts/sass_count.py advchain_b200/csrc/_obj/advk_chain.o [filter]"""
import re
import subprocess
import sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
name, cnt, ops = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        cnt[name] = 0
        ops[name] = {}
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        cnt[name] += 1
        op = m.group(1).split(".")[0]
        ops[name][op] = ops[name].get(op, 0) + 1
names = subprocess.run(["c++filt"] + list(cnt), capture_output=True, text=True).stdout.splitlines()
for mangled, dem in sorted(zip(cnt, names), key=lambda t: cnt[t[0]]):
    dem = re.sub(r"\(.*", "", dem)
    if flt in dem:
        top = sorted(ops[mangled].items(), key=lambda kv: -kv[1])[:10]
        print("%6d  %s   %s" % (cnt[mangled], dem, " ".join("%s:%d" % kv for kv in top)))
