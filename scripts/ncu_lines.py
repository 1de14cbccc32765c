"""Aggregate an ncu SASS source page (csv) per CUDA source line using nvdisasm -g line info."""
import re, csv, collections, sys
sass, srccsv, kernel, srcfile = sys.argv[1:5]
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if ('.text.' + kernel) in l and l.strip().startswith('.section'))
addr2line, cur = {}, None
for l in lines[start + 1:]:
    if l.strip().startswith('.section') and addr2line:
        break
    m2 = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m2:
        cur = int(m2.group(2)) if srcfile in m2.group(1) else None
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
keys = ("stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_lg", "stall_mio", "stall_math", "stall_not_selected", "stall_selected")
ist = {k: hdr.index(k) for k in keys}
base, tot = None, 0
per, inst, stall = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
totst = collections.Counter()
for r in rows[2:]:
    try:
        a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    except Exception:
        continue
    if base is None:
        base = a
    ln = addr2line.get(a - base)
    s = int(r[isamp] or 0); tot += s
    per[ln] += s; inst[ln] += int(r[iinst] or 0)
    for k, i in ist.items():
        stall[ln][k] += int(r[i] or 0); totst[k] += int(r[i] or 0)
src = open(sys.argv[5]).read().split('\n')
print("total samples", tot, {k: round(100 * v / tot, 1) for k, v in totst.most_common(6)})
for ln, s in per.most_common(int(sys.argv[6]) if len(sys.argv) > 6 else 25):
    st = [(k, round(100 * v / max(s, 1))) for k, v in stall[ln].most_common(2)]
    print(f"{100*s/tot:5.1f}%  L{ln}  inst {inst[ln]/1e6:8.1f}M  {st}  | {src[ln-1].strip()[:90] if ln else '?'}")
