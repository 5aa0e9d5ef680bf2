"""Executed instructions and stall samples per SOURCE line of one profiled kernel.
   [NCU_KERNEL=<regex>] python tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel-name-substring> [source.cu]
(NCU_KERNEL selects the kernel inside a report that holds several)
Joins ncu's SASS-level source page (instruction order) with nvdisasm -g's line table of the same object."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, pat = sys.argv[1:4]
src = sys.argv[4] if len(sys.argv) > 4 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
line_of = []; cur = None; inside = False
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        inside = pat in m.group(1); continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        line_of.append(cur)
flt = ["--kernel-name", "regex:" + os.environ["NCU_KERNEL"], "--launch-skip", os.environ.get("NCU_SKIP", "0"), "--launch-count", "1"] if os.environ.get("NCU_KERNEL") else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
if len(body) != len(line_of):
    print("warning: %d profiled instructions vs %d in the object" % (len(body), len(line_of)))
ex = collections.Counter(); sm = collections.Counter()
for r, ln in zip(body, line_of):
    ex[ln] += int(r[ci]); sm[ln] += int(r[cs])
te, ts = sum(ex.values()), sum(sm.values())
lines = open(src).read().splitlines() if src else None
print("%-22s %5s %12s %6s %8s %6s" % ("file", "line", "inst_exec", "%", "samples", "%"))
for ln in sorted(ex, key=lambda x: (x[0], x[1])):
    text = lines[ln[1] - 1].strip()[:90] if lines and ln[0] == os.path.basename(src) else ""
    print("%-22s %5d %12d %6.1f %8d %6.1f  %s" % (ln[0], ln[1], ex[ln], 100.0 * ex[ln] / te, sm[ln], 100.0 * sm[ln] / max(ts, 1), text))
print("total inst_exec %d, samples %d" % (te, ts))
