"""Address-ordered source map of one kernel's SASS: which source lines own which instruction range.
usage: nvdisasm --print-line-info x.cubin > x.dis; python tools/sass_map.py x.dis <kernel substring> [chunk]"""
import re, sys, collections
lines = open(sys.argv[1]).read().split('\n'); key = sys.argv[2]; CH = int(sys.argv[3]) if len(sys.argv) > 3 else 250
pat = re.compile(r'//## File "([^"]+)", line (\d+)')
seq = []; cur = None; on = False
for l in lines:
    if l.startswith('.text.'): on = key in l; continue
    if not on: continue
    m = pat.search(l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', l): seq.append(cur)
print(len(seq), "instructions,", len(seq) * 16 // 1024, "KB")
for i in range(0, len(seq), CH):
    c = collections.Counter((x[0].split('.')[0][:6], x[1] // 10 * 10) for x in seq[i:i + CH] if x)
    print(i, ' '.join('%s:%d(%d)' % (k[0], k[1], v) for k, v in sorted(c.items(), key=lambda t: -t[1])[:6]))
