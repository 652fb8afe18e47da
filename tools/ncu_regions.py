"""Summarise an ncu report of newton1d_kernel: key raw metrics, instruction mix, per-region instruction and stall shares.

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep <mangled kernel substring> <warp_rows>
"""
import csv, io, os, re, subprocess, sys
from collections import Counter

rep, kern = sys.argv[1], sys.argv[2]
warp_rows = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
d = dict(zip(raw[0], raw[2]))
for k in ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
          'smsp__sass_inst_executed_op_local_ld.sum']:
    print(f"{k:75s} {d.get(k)}")
for k in raw[0]:
    if 'issue_stalled' in k and k.endswith('ratio') and 'not_issued' not in k:
        try:
            if float(d[k]) > 0.1: print(f"{k:75s} {d[k]}")
        except Exception: pass

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv")))); hdr = rows[1]
ie = hdr.index('Instructions Executed'); ss = hdr.index('Warp Stall Sampling (All Samples)')
data = []
for r in rows[2:]:
    try: data.append((int(r[0], 16), r[1].strip(), int(r[ss]), int(r[ie]), r))
    except Exception: pass
base = data[0][0]
os.makedirs('/tmp/cub', exist_ok=True)
subprocess.run("cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all %s/gmpnp_b200/libgmpnp.so >/dev/null && nvdisasm -g -c edl1d.sm_100a.cubin > edl1d.sass" % ROOT, shell=True)
lines = open('/tmp/cub/edl1d.sass').read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l)
cur = None; line_of = {}
for l in lines[start + 1:]:
    if l.startswith('.text.') and line_of: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', l)
    if m: line_of[int(m.group(1), 16)] = cur
# function line ranges from the source
src = open(os.path.join(ROOT, 'gmpnp_b200/csrc/edl1d.cu')).read().split('\n')
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:template <[^>]*>\s*)?(?:__device__|__global__|static|int|struct)\b.*?(\w+)\s*\(', l)
    if m and not l.startswith(' '): marks.append((i, m.group(1)))
def region(fl):
    if fl is None: return 'none'
    f, l = fl
    if f != 'edl1d.cu': return f
    name = 'top'
    for i, nme in marks:
        if i <= l: name = nme
        else: break
    return name
inst = Counter(); stall = Counter(); ops = {}; mix = Counter()
for a, s, st, n, r in data:
    rg = region(line_of.get(a - base)); inst[rg] += n; stall[rg] += st
    t = [x for x in s.split() if not x.startswith('@')]
    o = t[0].split('.')[0] if t else '?'
    if o == 'IMAD' and 'MOV' in t[0]: o = 'IMAD.MOV'
    ops.setdefault(rg, Counter())[o] += n; mix[o] += n
ti, ts = sum(inst.values()), sum(stall.values())
print("mix:", ", ".join(f"{o} {n/ti*100:.1f}%" for o, n in mix.most_common(12)))
for rg, n in inst.most_common(14):
    print(f"{rg:22s} inst {n/ti*100:5.1f}% ({n/warp_rows:7.0f}/warp-row) stall {stall[rg]/ts*100:5.1f}%  ",
          [(o, round(c / warp_rows)) for o, c in ops[rg].most_common(6)])
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
top = sorted(data, key=lambda d: -d[2])[:14]
for a, s, st, n, r in top:
    reasons = {hdr[i]: int(r[i]) for i in cols if r[i] not in ('', '0')}
    big = sorted(reasons.items(), key=lambda kv: -kv[1])[:2]
    print(f"{st/ts*100:5.2f}% {n:10d} {line_of.get(a-base)} {s[:60]:60s} {big}")
