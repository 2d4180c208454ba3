"""Development aid: attribute ncu warp-stall samples (SASS level) to CUDA source lines.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [lib.so]
Needs ncu + cuobjdump + nvdisasm (no GPU).  The library must be the build the report was taken from.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         'hybrid-drt_b200', '_lib', 'libhybdrt_b200.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith('.cubin') or 'sm_100a' not in f:
        continue
    txt = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infunc, cur = False, None
    for ln in txt.splitlines():
        if ln.startswith('//---') and '.text.' in ln:
            infunc = kname in ln
            continue
        if not infunc:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(\S.*);', ln)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isamp, iexec = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = None
per_line = collections.Counter()
per_line_exec = collections.Counter()
per_line_stall = collections.defaultdict(collections.Counter)
tot = 0
for r in rows[2:]:
    if len(r) <= isamp:
        continue
    addr = int(r[ia], 16)
    if base is None:
        base = addr
    key = line_of.get(addr - base, (None, ''))[0]
    s = int(r[isamp] or 0)
    per_line[key] += s
    per_line_exec[key] += int(r[iexec] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            per_line_stall[key][hdr[i]] += v
    tot += s
src_cache = {}
print(f'total samples {tot}')
for key, s in per_line.most_common(int(os.environ.get("NCU_LINES_TOP", "45"))):
    text = ''
    if key:
        fn = os.path.join(os.path.dirname(lib), '..', 'csrc', key[0])
        if fn not in src_cache and os.path.exists(fn):
            src_cache[fn] = open(fn).read().splitlines()
        if fn in src_cache and key[1] <= len(src_cache[fn]):
            text = src_cache[fn][key[1] - 1].strip()[:90]
    top = ', '.join(f'{k[6:]}:{v}' for k, v in per_line_stall[key].most_common(3))
    print(f'{100.0 * s / tot:5.1f}%  exec {per_line_exec[key]:>12}  {key}  {text}   [{top}]')

# ---- per-function summary (device functions located by scanning the source for their first lines)
import bisect
fn = os.path.join(os.path.dirname(lib), '..', 'csrc', 'qphb_kernel.cu')
if os.path.exists(fn):
    src = open(fn).read().splitlines()
    marks = []
    for i, l in enumerate(src, 1):
        m = re.match(r'\s*(?:template.*\n)?__device__.*?(\w+)\(', l) or re.match(r'^qphb_kernel\(|^__global__.*?(\w+)\(', l)
        if l.startswith('__device__') or l.startswith('__global__') or l.startswith('qphb_kernel('):
            name = re.findall(r'(\w+)\(', l)
            marks.append((i, name[0] if name else l[:30]))
    starts = [m[0] for m in marks]
    agg_s, agg_e = collections.Counter(), collections.Counter()
    for key, s in per_line.items():
        if key and key[0] == 'qphb_kernel.cu':
            idx = bisect.bisect_right(starts, key[1]) - 1
            nm = marks[idx][1] if idx >= 0 else 'header'
        else:
            nm = str(key[0]) if key else 'unknown'
        agg_s[nm] += s
        agg_e[nm] += per_line_exec[key]
    te = sum(agg_e.values())
    print('\nper function: samples%  instr%  (instr)')
    for nm, s in agg_s.most_common():
        print(f'  {nm:28s} {100.0*s/tot:5.1f}%  {100.0*agg_e[nm]/te:5.1f}%  {agg_e[nm]}')
