#!/usr/bin/env python
"""tools/prof_lines.py REPORT.ncu-rep MANGLED_PREFIX SOURCE_FILE [marker ...] — warp-stall samples and executed instructions of one
kernel of an `ncu --set full --import-source on` report, per source REGION (lines between the given `line:label` markers) and per
source line, joined through `nvdisasm -g` of the in-tree library (the .so must be the one that was profiled)."""
import csv, subprocess, re, collections, os, sys
rep, func, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
marks = []
for m in sys.argv[4:]:
    ln, lab = m.split(":", 1); marks.append((int(ln), lab))
marks.sort()
ksel = ["-k", "regex:" + os.environ["PROF_KERNEL"]] if os.environ.get("PROF_KERNEL") else []     # reports that hold several kernels
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + ksel, capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if '# Samples' in r][0]
H = rows[hi]; ci = H.index('# Samples'); ce = H.index('Instructions Executed'); data = rows[hi + 1:]
os.makedirs('/tmp/cub', exist_ok=True)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.system('cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all %s/racinglmpc_b200/liblmpc_b200.so >/dev/null 2>&1 && nvdisasm -g -c *.cubin > all.sass 2>/dev/null' % root)
lines = open('/tmp/cub/all.sass').read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + func)][0]
cur = None; insts = []
for l in lines[start + 1:]:
    if l.startswith('//---------------------'): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m2 = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m2: insts.append((cur, m2.group(2)))
n = min(len(insts), len(data))
tot = sum(int(data[k][ci]) for k in range(n)); tote = sum(int(data[k][ce]) for k in range(n))
print("instructions (static)", len(insts), "rows", len(data), "samples", tot, "executed", tote)
base = os.path.basename(srcfile)
reg = collections.OrderedDict(); byline = collections.defaultdict(lambda: [0, 0])
def region(c):
    if not c or c[0] != base: return "other:" + (c[0] if c else "?")
    r = "<top>"
    for ln, lab in marks:
        if ln <= c[1]: r = lab
    return r
for k in range(n):
    c = insts[k][0]; a = reg.setdefault(region(c), [0, 0]); a[0] += int(data[k][ci]); a[1] += int(data[k][ce])
    byline[c][0] += int(data[k][ci]); byline[c][1] += int(data[k][ce])
for k, a in sorted(reg.items(), key=lambda kv: -kv[1][1]): print('%5.1f%% smp %5.1f%% inst  %s' % (100 * a[0] / tot, 100 * a[1] / tote, k))
srcl = open(srcfile).read().split('\n')
for c, (s_, e_) in sorted(byline.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get('TOPLINES', '25'))]:
    t = srcl[c[1] - 1].strip()[:100] if c and c[0] == base else str(c)
    print('%5s %5.1f%% %5.1f%%  %s' % (c[1] if c else '?', 100 * s_ / tot, 100 * e_ / tote, t))
