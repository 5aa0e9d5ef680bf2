"""Instructions per source line of one kernel:  python tools/sass_lines.py <object.o> <kernel-name-substring> [source.cu]
(nvdisasm -g on the embedded cubin; needs -lineinfo, which build.py passes)."""
import collections, os, re, subprocess, sys, tempfile

obj, pat = sys.argv[1], sys.argv[2]
src = sys.argv[3] if len(sys.argv) > 3 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
cnt = collections.Counter(); cur = None; inside = False; total = 0
for l in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        inside = pat in m.group(1); continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        cnt[cur] += 1; total += 1
lines = open(src).read().splitlines() if src else None
for (f, n), v in sorted(cnt.items(), key=lambda x: (x[0][0], x[0][1])):
    text = lines[n - 1].strip()[:100] if lines and f == os.path.basename(src) else ""
    print("%-22s %5d  %4d  %s" % (f, n, v, text))
print("total", total)
