import sys, time, zlib, os, random
import pathlib; R=pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0,str(R)); sys.path.insert(0,str(R/'tests'))
from support import *
from jdeflate_b200 import api
lib = api.JDeflateLib(sys.argv[1])
ref = api.JDeflateLib("oracle/_ref/libjdeflate_ref.so")
o = Oracle(); c = Corpus()
random.seed(int(sys.argv[2]) if len(sys.argv)>2 else 7)
sizes = [0, 1, 100, 5000, 70000]
if len(sys.argv)>3: sizes=[int(x) for x in sys.argv[3].split(',')]
t=time.time(); nbad=0; ncase=0
def cmp(tag, r1, r2, strict_out=True):
    global nbad, ncase
    ncase+=1
    ok = (r1[0],r1[1])==(r2[0],r2[1])
    if r2[0]==0: ok = ok and r1[2]==r2[2] and r1[3]==r2[3]
    elif r2[0] in (1,2): ok = ok and (r1[2]==r2[2] if strict_out else r2[2].startswith(r1[2]) or r1[2].startswith(r2[2]))
    if not ok:
        nbad+=1; print("MISMATCH", tag, r1[0:2], r2[0:2], len(r1[2]), len(r2[2]), r1[3], r2[3])
for kind in range(5):
    for n in sizes:
        d = c.fill(kind, n, offset=777)
        for lvl in (0,1,6,9):
            for ci, comp in enumerate((zlib_raw(d,lvl), ref.deflate_bytes(d, level=lvl))):
                tag=(KIND_NAMES[kind],n,lvl,ci)
                tail = os.urandom(5)
                cmp(tag+("valid",), lib.inflate_bytes(comp+tail, len(d)+10), o.inflate(comp+tail, len(d)+10))
                # streaming windows
                r = lib.inflate_bytes(comp+tail, len(d), window=random.choice([1,7,100,1000,4096]), feed=random.choice([1,3,50,333,5000]) if n<=5000 else random.choice([50,333,5000]))
                ncase+=1
                if r[0]!=0 or r[2]!=d or r[3]!=len(comp): nbad+=1; print("MISMATCH stream", tag, r[0:2], len(r[2]), r[3], len(comp))
                if len(comp) > 4:
                    k = random.randrange(1, len(comp))
                    for fin in (True, False):
                        cmp(tag+("trunc",fin,k), lib.inflate_bytes(comp[:k], len(d)+10, final=fin), o.inflate(comp[:k], len(d)+10, final=fin), strict_out=False)
                if len(d) > 10:
                    cmp(tag+("smalltgt",), lib.inflate_bytes(comp, len(d)//2), o.inflate(comp, len(d)//2))
                for _ in range(4):
                    b = bytearray(comp)
                    if not b: break
                    i = random.randrange(len(b)); b[i] ^= 1 << random.randrange(8)
                    cmp(tag+("corrupt",i), lib.inflate_bytes(bytes(b), len(d)+1000), o.inflate(bytes(b), len(d)+1000), strict_out=False)
print("cases", ncase, "bad", nbad, "%.1fs" % (time.time()-t))
