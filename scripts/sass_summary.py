"""Instruction-mix summary of the cubins inside fabind_b200/libfabind_b200.so (cuobjdump -sass): per kernel the instruction count and the
Blackwell-specific mnemonics that show what it runs on -- UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA loads / stores), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (mma.sync), MUFU, STG/LDG.  Run on the CPU: python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "fabind_b200", "libfabind_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "ATOM", "RED"]
fn, stats = None, collections.OrderedDict()
arch = set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn).replace("void ", "")
        stats[fn] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        op = m.group(1)
        stats[fn]["_n"] += 1
        for k in KEYS:
            if op.startswith(k):
                stats[fn][k] += 1
print(f"{os.path.relpath(lib, ROOT)}: cubin targets {sorted(arch)}; {len(stats)} kernels")
tot = collections.Counter()
for c in stats.values():
    tot.update(c)
print("library totals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
print()
print(f"{'kernel':84s} {'instr':>6s}  " + " ".join(f"{k:>7s}" for k in KEYS[:8]))
for fn, c in sorted(stats.items(), key=lambda kv: -(kv[1]['UTCHMMA'] * 100000 + kv[1]['UTMALDG'] * 1000 + kv[1]['_n'])):
    if c["_n"] == 0:
        continue
    print(f"{fn[:84]:84s} {c['_n']:6d}  " + " ".join(f"{c[k]:7d}" for k in KEYS[:8]))
