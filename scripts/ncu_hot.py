"""Hot source lines of a kernel from an ncu report: python scripts/ncu_hot.py <report.ncu-rep> <libscpb.so> <kernel mangled substring> [top]
Maps the SASS addresses of the ncu source page to (file, line) through nvdisasm -g and aggregates the stall samples."""
import sys, os, re, csv, collections, subprocess, tempfile, glob
rep, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
sass = None
for f in glob.glob(os.path.join(tmp, "*.cubin")):
    out = subprocess.run(["nvdisasm", "-g", f], capture_output=True, text=True).stdout
    if kern in out:
        sass = out
        break
assert sass, "kernel not found"
lines = sass.split("\n")
start = next(i for i, l in enumerate(lines) if (".text." in l and kern in l and l.strip().startswith(".section")))
addr2line, cur = {}, None
for l in lines[start + 1:]:
    if l.strip().startswith(".section") and addr2line:
        break
    m2 = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m2:
        cur = (os.path.basename(m2.group(1)), int(m2.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
keys = ("stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_lg", "stall_mio", "stall_math", "stall_no_inst",
        "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_dispatch")
ist = {k: hdr.index(k) for k in keys}
base, tot = None, 0
per, inst, stall, totst = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
perfile = collections.Counter()
for r in rows[2:]:
    if len(r) <= ia:
        continue
    try:
        a = int(r[ia], 16)
    except Exception:
        continue
    if base is None:
        base = a
    ln = addr2line.get(a - base)
    s = int(r[isamp] or 0); tot += s
    per[ln] += s; inst[ln] += int(r[iinst] or 0)
    perfile[ln[0] if ln else None] += s
    for k, i in ist.items():
        v = int(r[i] or 0)
        stall[ln][k] += v; totst[k] += v
print("total samples", tot, {k: round(100 * v / max(tot, 1), 1) for k, v in totst.most_common(8)})
print("by file", {k: round(100 * v / max(tot, 1), 1) for k, v in perfile.most_common()})
srcdir = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc")
cache = {}
for ln, s in per.most_common(top):
    st = [(k.replace("stall_", ""), round(100 * v / max(s, 1))) for k, v in stall[ln].most_common(2)]
    txt = "?"
    if ln:
        fn = os.path.join(srcdir, ln[0])
        if fn not in cache:
            cache[fn] = open(fn).read().split("\n") if os.path.exists(fn) else []
        if 0 < ln[1] <= len(cache[fn]):
            txt = cache[fn][ln[1] - 1].strip()[:100]
    print(f"{100 * s / max(tot, 1):5.1f}%  {ln}  inst {inst[ln] / 1e6:7.1f}M  {st}  | {txt}")
