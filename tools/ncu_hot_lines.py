"""hot.py <ncu-rep> <kernel-regex> [topn] — per-source-line stall samples by joining ncu SASS rows with nvdisasm -g line info."""
import csv,sys,subprocess,re,collections,os,glob
rep=sys.argv[1]; kern=sys.argv[2]; topn=int(sys.argv[3]) if len(sys.argv)>3 else 25
# kern is matched against the MANGLED section name in the cubin (e.g. k4_nmsILi4E for k4_nms<4>); the ncu filter
# uses the plain function name in front of the template suffix
ncu_kern=re.sub(r"I(Li\d+|f|6__half)E.*$", "", kern)
so="/root/repo/sar-yolo_b200/libsarpost.so"
wd="/tmp/probe/cubin"; os.makedirs(wd,exist_ok=True)
for f in glob.glob(wd+"/*.cubin"): os.remove(f)
subprocess.run(["cuobjdump","-xelf","all",so],cwd=wd,capture_output=True)
cubin=glob.glob(wd+"/*.cubin")[0]
dis=subprocess.run(["nvdisasm","-g",cubin],capture_output=True,text=True).stdout.splitlines()
# locate .text section of kernel
start=None
for i,l in enumerate(dis):
    if l.startswith("//--------------------- .text.") and re.search(kern,l): start=i; break
assert start is not None, "kernel not found in cubin"
lines=[]; cur=("?",0)
for l in dis[start+1:]:
    if l.startswith("//--------------------- "): break
    m=re.match(r'\s*//## File "(.*)", line (\d+)(.*)',l)
    if m: cur=(os.path.basename(m.group(1)),int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: lines.append((int(m.group(1),16),cur,m.group(2)))
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+ncu_kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
# may contain multiple kernel instances; take the first
hdr=None; inst=[]; ninst=0
for r in rows:
    if r and r[0]=="Kernel Name":
        ninst+=1
        if ninst>1: break
        continue
    if r and r[0]=="Address": hdr=r; continue
    if hdr and len(r)>=len(hdr): inst.append(dict(zip(hdr,r)))
print(f"sass instrs: ncu={len(inst)} nvdisasm={len(lines)}")
n=min(len(inst),len(lines))
stalls=[k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg=collections.defaultdict(lambda: collections.defaultdict(float))
tot=0
for i in range(n):
    d=inst[i]; s=float(d["# Samples"] or 0); tot+=s
    key=lines[i][1]
    agg[key]["samples"]+=s
    agg[key]["exec"]+=float(d["Instructions Executed"] or 0)
    for k in stalls: agg[key][k]+=float(d[k] or 0)
print("total samples",tot)
mix=collections.defaultdict(float)
for key,v in agg.items():
    for k in stalls: mix[k]+=v[k]
print("stall mix:",{k[6:]:round(v/tot*100,1) for k,v in sorted(mix.items(),key=lambda kv:-kv[1])[:8]})
srcs={}
def src(f,ln):
    if f not in srcs:
        p=[x for x in glob.glob("/root/repo/sar-yolo_b200/csrc/*")+glob.glob("/root/repo/include/*") if os.path.basename(x)==f]
        srcs[f]=open(p[0]).read().splitlines() if p else []
    L=srcs[f]; return L[ln-1].strip()[:95] if 0<ln<=len(L) else ""
for key,v in sorted(agg.items(),key=lambda kv:-kv[1]["samples"])[:topn]:
    top=sorted(((v[k],k[6:]) for k in stalls),reverse=True)[:2]
    print(f"{v['samples']/tot*100:5.1f}% {key[0]}:{key[1]:<4} ex={int(v['exec']):>8} {top[0][1]:>10}/{top[1][1]:<10} {src(*key)}")
