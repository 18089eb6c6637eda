"""Join `ncu --page source --csv` (SASS rows) with `nvdisasm -g -c` line info of lz.sm_100a.cubin (extracted with
cuobjdump -xelf all libjdeflate.so into /tmp/cub, disassembled to /tmp/cub/lz.dis) and aggregate instructions and stall
samples of a kernel by source line.  usage: ncu_by_line.py <source.csv> [kernel substring] [min pct] [file.cu]"""
import re, csv, collections, sys
kern = sys.argv[2] if len(sys.argv) > 2 else 'lz_kernel'
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
srcfile = sys.argv[4] if len(sys.argv) > 4 else 'lz.cu'
cur = None; line = None; addr2line = {}
for l in open('/tmp/cub/%s.dis' % srcfile.split('.')[0]):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\.text\.(\S+):', l)
    if m: cur = m.group(1); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m and cur and kern in cur:
        addr2line[int(m.group(1), 16)] = line
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]; ci = H.index("# Samples"); ie = H.index("Instructions Executed"); it = H.index("Thread Instructions Executed")
base = None; agg = collections.Counter(); aggi = collections.Counter(); aggt = collections.Counter()
for r in rows[hi + 1:]:
    try: a = int(r[0], 16)
    except Exception: continue
    if base is None: base = a
    ln = addr2line.get(a - base)
    agg[ln] += int(r[ci] or 0); aggi[ln] += int(r[ie] or 0); aggt[ln] += int(r[it] or 0)
ts = sum(agg.values()); ti = sum(aggi.values())
import os
src = open(os.environ.get('NCU_SRC') or '/root/repo/jdeflate_b200/csrc/device/' + srcfile).read().split('\n')   # NCU_SRC: the source as it was when the capture was taken
print("total warp-instructions %d, samples %d" % (ti, ts))
for ln in sorted(aggi, key=lambda x: (str(x[0]), x[1]) if x else ("", 0)):
    c = aggi[ln]
    if 100 * c / ti < minpct and 100 * agg[ln] / max(ts, 1) < minpct: continue
    txt = src[ln[1] - 1].strip()[:90] if ln and ln[0] == srcfile else str(ln)
    print("%5s %5.1f%% ins %5.1f%% smp lanes %4.1f  %s" % (ln[1] if ln else None, 100 * c / ti, 100 * agg[ln] / ts, aggt[ln] / max(1, c), txt))
