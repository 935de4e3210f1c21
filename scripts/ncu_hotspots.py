"""Join an `ncu --page source --csv --print-source sass` export with `nvdisasm -gi` line tables and
aggregate executed instructions / stall samples by call site (the line of physics_step /
rollout_warp an instruction was inlined into) and by innermost function.
    python scripts/ncu_hotspots.py <src.csv> <disasm -gi txt> <device header at the profiled commit> <physics steps per launch>"""
import collections
import csv
import re
import sys

src_csv, dis_txt, header, steps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
hdr_lines = open(header).read().splitlines()
# function spans of the header: name of the enclosing DEV function per line
func_of = {}
cur = "?"
for i, l in enumerate(hdr_lines, 1):
    m = re.match(r"^(?:template <[^>]*>\s*)?(?:DEV|HD|static inline|__global__)?\s*[\w:<>\*&\s]*?\b(\w+)\s*\([^;]*\)\s*\{\s*$", l)
    if (l.startswith("DEV ") or l.startswith("HD ")) and "(" in l:
        m2 = re.search(r"\b(\w+)\s*\(", l[4:])
        if m2:
            cur = m2.group(1)
    func_of[i] = cur
# disassembly: per instruction (in order) the inline chain [(file, line), ...] innermost first
# (nvdisasm prints one "//## File A, line n inlined at B, line m" per level on consecutive lines,
# innermost first; the group applies to the instructions that follow until the next group)
chains, chain, in_group = [], [], False
for l in open(dis_txt):
    s = l.strip()
    if s.startswith("//## File"):
        parts = re.findall(r'"([^"]+)", line (\d+)', s)
        loc = (parts[0][0].split("/")[-1], int(parts[0][1]))
        if not in_group:
            chain, in_group = [], True
        chain.append(loc)
    elif re.match(r"^/\*[0-9a-f]{4,}\*/", s):
        in_group = False
        chains.append(chain)
rows = list(csv.reader(open(src_csv)))
h = rows[1]
data = rows[2:]
iI, iS, iT = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
iW = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None
assert len(data) == len(chains), (len(data), len(chains))
by_site, by_func = collections.defaultdict(lambda: [0, 0, 0, 0]), collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0, 0]
for r, ch in zip(data, chains):
    v = [int(r[iI]), int(r[iS]), int(r[iT]), int(r[iW]) if iW is not None and r[iW].isdigit() else 0]
    inner = ch[0] if ch else ("?", 0)
    # call site: the outermost dial_device.cuh frame that lies in physics_step / rollout_warp
    site = None
    for want in ("physics_step", "rollout_warp"):     # the statement of physics_step if there is one, else of rollout_warp
        for f, n in reversed(ch):
            if f == "dial_device.cuh" and func_of.get(n) == want:
                site = (want, n)
                break
        if site:
            break
    site = site or (inner[0], inner[1])
    fn = func_of.get(inner[1], "?") if inner[0] == "dial_device.cuh" else inner[0]
    for acc in (by_site[site], by_func[fn], tot):
        for k in range(4):
            acc[k] += v[k]
print(f"total: {tot[0]/steps:.0f} warp inst / physics step, {tot[2]/max(tot[0],1):.1f} lanes, {tot[1]} samples")
print("\n| call site | source | time | inst/step | inst | smem | lanes |\n|---|---|---|---|---|---|---|")
for site, v in sorted(by_site.items(), key=lambda kv: -kv[1][1])[:28]:
    line = hdr_lines[site[1] - 1].strip()[:70] if site[0] in ("physics_step", "rollout_warp") else ""
    print(f"| `{site[0]}:{site[1]}` | `{line}` | {100*v[1]/tot[1]:.1f} % | {v[0]/steps:.0f} | {100*v[0]/tot[0]:.1f} % | {100*v[3]/max(tot[3],1):.1f} % | {v[2]/max(v[0],1):.1f} |")
print("\n| innermost function | time | inst/step | inst | smem | lanes |\n|---|---|---|---|---|---|")
for fn, v in sorted(by_func.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"| `{fn}` | {100*v[1]/tot[1]:.1f} % | {v[0]/steps:.0f} | {100*v[0]/tot[0]:.1f} % | {100*v[3]/max(tot[3],1):.1f} % | {v[2]/max(v[0],1):.1f} |")
