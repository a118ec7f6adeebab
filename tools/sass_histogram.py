"""Per-kernel instruction histogram of the shipped library (cuobjdump -sass): the mnemonics that prove the tcgen05 /
TMA / TMEM path, plus code size.  Writes profiles/<tag>_sass_histogram.txt.
    python tools/sass_histogram.py r02"""
import collections, os, re, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(REPO, "pixel_nerf_multiscale_b200", "libpixelnerf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT", "HMMA",
         "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "FFMA", "BAR", "WARPSYNC", "CCTL", "ERRBAR", "MEMBAR"]
per = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        per[name]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                per[name][w] += 1
        if op.startswith("UTCHMMA"):
            per[name]["variants:" + op] += 1
        if op.startswith("UTMALDG") or op.startswith("UTCBAR") or op.startswith("LDTM") or op.startswith("STTM"):
            per[name]["variants:" + op] += 1
out = ["# cuobjdump -sass %s -- per-kernel instruction counts (static code, 16 B per instruction)" % os.path.relpath(lib, REPO), ""]
for k, c in sorted(per.items(), key=lambda kv: -kv[1]["_total"]):
    out.append("%s: %d instructions (%.1f KB)" % (k, c["_total"], c["_total"] * 16 / 1024.0))
    line = "   " + "  ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w])
    out.append(line)
    var = ["%s x%d" % (v[len("variants:"):], n) for v, n in sorted(c.items()) if v.startswith("variants:")]
    if var:
        out.append("   " + ", ".join(var))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
path = os.path.join(REPO, "profiles", "%s_sass_histogram.txt" % tag)
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
