"""Per-function / per-line / per-opcode shares of executed warp instructions and stall samples from an ncu report that was
captured with --import-source on:  python scripts/ncu_hotspots.py gpurun_out/r01_full.ncu-rep > profiles/r01_hotspots.md"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
def page(src):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", src, "--csv"], stdout=subprocess.PIPE, text=True).stdout
    return list(csv.reader(out.splitlines()))
full = page("cuda")
src = {int(r[0]): r[1] for r in full[2:] if r and r[0].isdigit()}
funcs = []
for ln in sorted(src):
    m = re.search(r'__device__ (?:__noinline__ |__forceinline__ )?(?:constexpr )?(?:[\w:<> \*&]+?) (\w+)\(', src[ln])
    if m and not src[ln].strip().startswith('//'): funcs.append((ln, m.group(1)))
funcs.append((10 ** 9, 'end'))
rows = page("cuda,sass")
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No' and len(r) > 10]
hdr = rows[hi[0]]
ci, cs = hdr.index('Instructions Executed'), hdr.index('# Samples')
cl, cn = hdr.index('stall_long_sb'), hdr.index('stall_no_inst')
inst, smp, lsb, noi = (collections.Counter() for _ in range(4))
cur = None
for r in rows[hi[0] + 1:(hi[1] - 2 if len(hi) > 1 else len(rows))]:
    if r[0] != '': cur = int(r[0])
    try:
        inst[cur] += int(r[ci]); smp[cur] += int(r[cs]); lsb[cur] += int(r[cl]); noi[cur] += int(r[cn])
    except Exception: pass
tot, ts = sum(inst.values()), sum(smp.values())
fi, fs = collections.Counter(), collections.Counter()
for ln, v in inst.items():
    name = [f for (a, f), (b, _) in zip(funcs, funcs[1:]) if a <= ln < b]
    fi[name[0] if name else 'other'] += v; fs[name[0] if name else 'other'] += smp[ln]
print(f"# hot spots of cuipm_solve_kernel<1,21,3> ({rep.split('/')[-1]}; {tot:.3e} warp instructions, {ts} stall samples)\n")
print("## by function (share of executed warp instructions / of stall samples)\n\n| function | instr | samples |\n|---|---|---|")
for k, v in fi.most_common(14): print(f"| `{k}` | {100*v/tot:.1f}% | {100*fs[k]/ts:.1f}% |")
print("\n## by source line (top 25 by stall samples; lsb = long scoreboard, noi = no instruction)\n\n| instr | samples | lsb | noi | line |\n|---|---|---|---|---|")
for ln, v in sorted(smp.items(), key=lambda x: -x[1])[:25]:
    print(f"| {100*inst[ln]/tot:.1f}% | {100*v/ts:.1f}% | {100*lsb[ln]/ts:.1f} | {100*noi[ln]/ts:.1f} | `{ln}: {src.get(ln, '').strip()[:100]}` |")
sass = page("sass")
h = sass[1]; ci2, cs2 = h.index('Instructions Executed'), h.index('# Samples')
mix, ms = collections.Counter(), collections.Counter(); ex = []
for r in sass[2:]:
    if len(r) <= ci2 or not r[ci2].isdigit(): continue
    op = re.sub(r'^@!?U?P\d+\s+', '', r[1].strip()).split()[0].split('.')[0] if r[1].strip() else '?'
    mix[op] += int(r[ci2]); ms[op] += int(r[cs2]); ex.append(int(r[ci2]))
t2 = sum(mix.values())
print("\n## by opcode\n\n| opcode | instr | samples |\n|---|---|---|")
for k, v in mix.most_common(22): print(f"| {k} | {100*v/t2:.1f}% | {100*ms[k]/max(1,sum(ms.values())):.1f}% |")
ex.sort(reverse=True); cum = 0; marks = {}
for i, e in enumerate(ex):
    cum += e
    for f in (0.5, 0.9, 0.99):
        if f not in marks and cum >= f * t2: marks[f] = i + 1
print(f"\nstatic footprint: 50% of the executed instructions come from {marks[0.5]} static ones ({marks[0.5]*16//1024} KB), "
      f"90% from {marks[0.9]} ({marks[0.9]*16//1024} KB), 99% from {marks[0.99]} ({marks[0.99]*16//1024} KB); kernel total {len(ex)} ({len(ex)*16//1024} KB)")
