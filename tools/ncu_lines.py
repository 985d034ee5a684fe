"""Attributes the warp-stall samples of one kernel in an ncu report (--set full --import-source on) to source lines and to
the device functions they were inlined from (ncu's CSV source page has SASS only; nvdisasm -gi gives the inline chains).
  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <mangled .text symbol prefix> [cubin object: tree|net]
Run in the dev container (needs the same build of kami_b200/libkami_b200.so that was profiled)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, rx, sym = sys.argv[1], sys.argv[2], sys.argv[3]
obj = sys.argv[4] if len(sys.argv) > 4 else "tree"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "kami_b200", "libkami_b200.so")], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, obj + ".sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text." + sym)][0]
end = [i for i, l in enumerate(dis) if i > start and l.startswith("\t.section")][0]
seq, pend, chain = [], [], []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pend.append(m)
        continue
    m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m2:
        if pend:
            chain = [(os.path.basename(pend[0].group(1)), int(pend[0].group(2)))]
            chain += [(os.path.basename(p.group(3)), int(p.group(4))) for p in pend if p.group(3)]
            pend = []
        seq.append((int(m2.group(1), 16), list(chain)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = [r for r in rows if r and r[0] == "Address"][0]
body = [r for r in rows if r and r[0].startswith("0x")]
ia, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
first = int(body[0][0], 16)
# several launches of the kernel repeat the address range: fold them
bymap = dict(seq)
srcs = {}
def text(f, n):
    if f not in srcs:
        p = os.path.join(ROOT, "kami_b200", "csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    L = srcs[f]
    return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""
funcs = {}
def func_of(f, line):
    if f not in funcs:
        text(f, 1)
        lst = []
        for i, l in enumerate(srcs[f], 1):
            if l and not l[0].isspace() and ("__device__" in l or "__global__" in l or "KB_HD" in l):
                t = re.sub(r"__launch_bounds__\([^)]*\)", "", l).replace("__forceinline__", "")
                m = re.search(r"\b([a-zA-Z_]\w*)\s*\(", t)
                if m:
                    lst.append((i, m.group(1)))
        funcs[f] = lst
    name = "?"
    for s, n in funcs[f]:
        if s <= line:
            name = n
        else:
            break
    return name
by_fn, by_line, fn_inst = collections.Counter(), collections.Counter(), collections.Counter()
tot = toti = 0
for r in body:
    off = int(r[0], 16) - first
    s, e = int(r[ia] or 0), int(r[ie] or 0)
    tot += s
    toti += e
    ch = bymap.get(off) or []
    names = [func_of(f, ln) for f, ln in ch]
    # the innermost frame that is a "phase" function of tree.cu, else the innermost function
    key = names[0] if names else "?"
    for nme, (f, ln) in zip(names, ch):
        if f in ("tree.cu", "arena.inl") and not nme.startswith(("lane_id", "shfl_", "splitmix", "u01_", "raise", "tree_nodes", "tree_meta", "root_turn")):
            key = nme
            break
    by_fn[key] += s
    fn_inst[key] += e
    if ch:
        by_line[ch[0]] += s
print("samples %d, warp instructions %d" % (tot, toti))
print("--- by function (innermost tree.cu frame)")
for k, v in by_fn.most_common(16):
    print("%5.1f%% samples %5.1f%% instr  %s" % (100.0 * v / tot, 100.0 * fn_inst[k] / max(1, toti), k))
print("--- by innermost source line")
for (f, ln), v in by_line.most_common(30):
    print("%5.1f%%  %s:%d  %s" % (100.0 * v / tot, f, ln, text(f, ln)))
