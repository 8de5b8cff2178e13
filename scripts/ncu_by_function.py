"""Per-function breakdown of an ncu report of the throughput kernel: executed warp instructions and stall samples per
source function of cuipm_fast_core.h.  ncu's source page exports SASS rows only for code in headers, so the rows are
joined with nvdisasm's line table of the same cubin (extracted from libcuipm.so) by instruction index.
usage: python scripts/ncu_by_function.py <report.ncu-rep> <kernel-substring e.g. Li21ELi3E> [--lines N]"""
import bisect, collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, ksub = sys.argv[1], sys.argv[2]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "acados_b200", "csrc", "libcuipm.so")], cwd=tmp, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, "cuipm_fast.sm_100a.cubin")], stdout=subprocess.PIPE, text=True).stdout
src = open(os.path.join(ROOT, "acados_b200", "csrc", "cuipm_fast_core.h")).read().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r"FK_DEV\s+[\w:<>&\s\*]*?\b(\w+)\s*\(", l)
    if m and "define" not in l:
        marks.append((i, m.group(1)))
starts = [m[0] for m in marks]
# instruction index -> (file, line) for the kernel
loc, kern, cf, cl = [], None, None, None
for l in dis.split("\n"):
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cf, cl = m.group(1), int(m.group(2)); continue
    m = re.match(r"\.text\.(\S+):", l)
    if m:
        kern = m.group(1); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l) and kern and ksub in kern:
        loc.append((cf, cl, l.strip()))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(csvtxt.split("\n")))
hdr = next(r for r in rows if r and r[0] == "Address")
body = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
assert len(body) == len(loc), (len(body), len(loc))
ex, st, byline, stline = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
opmix = collections.Counter()
for r, (f, ln, txt) in zip(body, loc):
    e, s = int(r[ci["Instructions Executed"]]), int(r[ci["Warp Stall Sampling (All Samples)"]])
    if f and f.endswith("cuipm_fast_core.h"):
        j = bisect.bisect_right(starts, ln) - 1
        fn = marks[j][1] if j >= 0 else "?"
        byline[ln] += e; stline[ln] += s
    else:
        fn = "other:" + os.path.basename(str(f))
    ex[fn] += e; st[fn] += s
    op = re.sub(r"^(@!?U?P\d+\s+)?", "", r[ci["Source"]].strip()).split()[0].split(".")[0]
    opmix[op] += e
te, ts = sum(ex.values()), sum(st.values())
print(f"executed warp instructions {te:.4g}, stall samples {ts}")
print(f"{'function':22s} {'instr %':>8s} {'stall %':>8s}")
for fn, e in ex.most_common(25):
    print(f"{fn:22s} {100*e/te:8.2f} {100*st[fn]/max(ts,1):8.2f}")
print("instruction mix:", ", ".join(f"{o} {100*c/te:.1f}%" for o, c in opmix.most_common(16)))
if nlines:
    print("hottest lines (instr %, stall %):")
    for ln, e in byline.most_common(nlines):
        print(f"  {ln:5d} {100*e/te:6.2f} {100*stline[ln]/max(ts,1):6.2f}  {src[ln-1].strip()[:110]}")
    print("lines with most stall samples (instr %, stall %):")
    for ln, e in stline.most_common(nlines):
        print(f"  {ln:5d} {100*byline[ln]/te:6.2f} {100*e/max(ts,1):6.2f}  {src[ln-1].strip()[:110]}")
