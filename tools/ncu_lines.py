"""Attribute an ncu SASS source page (CSV) to CUDA source lines using nvdisasm -g line info.

    python tools/ncu_lines.py <source_page.csv> <nvdisasm -g -c output> <mangled kernel substring> [top]
"""
import csv, re, sys
from collections import Counter, defaultdict

src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ie = hdr.index('Instructions Executed'); ss = hdr.index('Warp Stall Sampling (All Samples)')
data = []
for r in rows[2:]:
    try: data.append((int(r[0], 16), r[1].strip(), int(r[ss]), int(r[ie])))
    except Exception: pass
base = data[0][0]
# parse nvdisasm: instruction lines look like "        /*0010*/   OPCODE ... ;"
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l)
cur = None; inl = None
line_of = {}
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//------'):
        if line_of: break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = int(m.group(2)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
inst = Counter(); stall = Counter()
for a, s, st, n in data:
    ln = line_of.get(a - base)
    inst[ln] += n; stall[ln] += st
ti, ts = sum(inst.values()), sum(stall.values())
srcfile = open('gmpnp_b200/csrc/edl1d.cu').read().split('\n') if len(sys.argv) < 6 else open(sys.argv[5]).read().split('\n')
print(f"total inst {ti:.3e} samples {ts}")
for ln, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    txt = srcfile[ln - 1].strip()[:90] if ln else ''
    print(f"{str(ln):>5s} inst {n/ti*100:5.2f}% stall {stall[ln]/ts*100:5.2f}%  {txt}")
