"""Static SASS size per source function / line for one kernel instantiation of cuipm_kernel.cu.
usage: python scripts/sass_size.py 'ILi1ELi21ELi3E' """
import re, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(root, "acados_b200/csrc/build/cuipm_kernel.o")
os.makedirs("/tmp/sass_size", exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd="/tmp/sass_size", stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir("/tmp/sass_size") if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", os.path.join("/tmp/sass_size", cubin)], stdout=subprocess.PIPE, text=True).stdout
want = sys.argv[1] if len(sys.argv) > 1 else "ILi1ELi21ELi3E"
src = open(os.path.join(root, "acados_b200/csrc/cuipm_kernel.cu")).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.search(r'__device__ (?:__noinline__ |__forceinline__ )?(?:constexpr )?(?:[\w:<> \*&]+?) (\w+)\(', l)
    if m and not l.strip().startswith('//'): funcs.append((i, m.group(1)))
funcs.append((10**9, 'end'))
inside = False; cur = None; agg = {}; lines = {}
for l in txt.split("\n"):
    if l.startswith(".text."):
        inside = want in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if 'cuipm_kernel.cu' in m.group(1):
            ln = int(m.group(2)); name = [f for (a, f), (b, _) in zip(funcs, funcs[1:]) if a <= ln < b]
            cur = (name[0] if name else 'other', ln)
        else: cur = ('hdr', 0)
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l) and cur:
        agg[cur[0]] = agg.get(cur[0], 0) + 1; lines[cur] = lines.get(cur, 0) + 1
tot = sum(agg.values()); print('total', tot)
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:16]: print(f'{k:20s} {v:6d} {100*v/tot:5.1f}%')
print('top lines')
for k, v in sorted(lines.items(), key=lambda x: -x[1])[:30]: print(k, v, src[k[1]-1].strip()[:100] if k[1] else '')
