"""cuobjdump -sass of libsarpost.so -> per kernel: SASS instruction count and the mnemonics that prove what the kernel uses
(TMA loads UTMALDG, mbarrier SYNCS, cluster barriers UCGABAR, distributed-shared-memory stores, MUFU, votes, shuffles ...).
    python tools/sass_summary.py > profiles/r2/sass_summary.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sar-yolo_b200", "libsarpost.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keep = re.compile(r"^(UTMALDG|UTMAPF|UCGABAR|SYNCS|MUFU|REDUX|CREDUX|ATOMS|ATOMG|RED|LDGSTS|VOTE|VOTEU|MATCH|SHFL|MEMBAR|NANOSLEEP|FFMA$|LDS|STS|ST\b.*CLUSTER|MAPA|ERRBAR|CCTL|FMNMX$|BAR)")
fn, c = None, collections.OrderedDict()
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1); c[fn] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][\w.]*)", l)
    if fn and m:
        op = m.group(1); c[fn]["_total"] += 1
        base = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("MUFU", "SHFL", "ATOMS", "ATOMG", "REDUX", "CREDUX", "VOTE", "VOTEU", "MATCH", "MEMBAR", "BAR", "SYNCS", "UTMALDG", "RED") else op.split(".")[0]
        if keep.match(op): c[fn][base] += 1
for f, cnt in c.items():
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    print(f"{name}\n    {cnt['_total']} SASS instructions; " + ", ".join(f"{k} x{v}" for k, v in sorted(cnt.items()) if k != "_total"))
