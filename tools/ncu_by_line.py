"""Join `ncu --page source --csv` (SASS rows) with `nvdisasm -g -c` line info of lz.sm_100a.cubin (extracted with
cuobjdump -xelf all libjdeflate.so into /tmp/cub, disassembled to /tmp/cub/lz.dis) and aggregate instructions and stall
samples of lz_kernel<false> by source line and by region.  usage: ncu_by_line.py <source.csv> [top lines]"""
import re,csv,collections,sys
cur=None; line=None; addr2line={}
for l in open('/tmp/cub/lz.dis'):
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',l)
    if m: line=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\.text\.(\S+):',l)
    if m: cur=m.group(1); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/',l)
    if m and cur and 'lz_kernelILb0' in cur:
        addr2line[int(m.group(1),16)]=line
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
H=rows[hi]; ci=H.index("# Samples"); ie=H.index("Instructions Executed"); it=H.index("Thread Instructions Executed")
base=None; agg=collections.Counter(); aggi=collections.Counter(); aggt=collections.Counter()
for r in rows[hi+1:]:
    try: a=int(r[0],16)
    except: continue
    if base is None: base=a
    ln=addr2line.get(a-base)
    agg[ln]+=int(r[ci] or 0); aggi[ln]+=int(r[ie] or 0); aggt[ln]+=int(r[it] or 0)
ts=sum(agg.values()); ti=sum(aggi.values())
src=open('/root/repo/jdeflate_b200/csrc/device/lz.cu').read().split('\n')
# regions
regions=[(404,466,"stage"),(466,552,"pass 1 / rounds setup (ROUNDS only)"),(552,568,"WALK"),(568,606,"COMPARE + mode ballots"),(606,704,"FETCH (position hand-out)"),(704,742,"3-byte probes"),(742,744,"parse call"),(744,830,"token emission"),(296,404,"parse_phase")]
reg=collections.Counter(); regs=collections.Counter()
other=0
for ln,c in aggi.items():
    if ln and ln[0]=='lz.cu':
        for a,b,nm in regions:
            if a<=ln[1]<b: reg[nm]+=c; regs[nm]+=agg[ln]; break
        else: reg["lz.cu other"]+=c; regs["lz.cu other"]+=agg[ln]
    else: reg[str(ln)]+=c; regs[str(ln)]+=agg[ln]
for k,v in reg.most_common(16): print("%-40s %5.1f%% ins %5.1f%% smp"%(k,100*v/ti,100*regs[k]/ts))
print()
for ln,c in aggi.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 25):
    txt = src[ln[1]-1].strip()[:80] if ln and ln[0]=='lz.cu' else str(ln)
    print(ln[1] if ln else None, "%5.1f%% ins %5.1f%% smp lanes %.1f"%(100*c/ti,100*agg[ln]/ts, aggt[ln]/max(1,c)), txt)
