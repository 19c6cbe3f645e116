"""static SASS evidence of the shipped library: per kernel, how many TMA loads (UTMALDG), mbarrier operations (SYNCS), FFMA / DFMA,
shared-memory loads / stores and tensor-core instructions (none expected: the path is a 1-channel stencil) it holds.
usage: python scratch/sass_static_hist.py [lib] > profiles/r2_sass_hist.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "scene-net_b200/libscenenet_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTMALDG", "SYNCS", "FFMA", "DFMA", "LDS", "STS", "ATOMS", "RED", "ATOMG", "LDG", "STG", "UTCHMMA", "UTCQMMA", "HMMA", "LDTM", "STTM", "WARPSYNC", "BAR"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        per[cur][op] += 1
        per[cur]["_total"] += 1
def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
tot = collections.Counter()
print(f"# {lib}: {len(per)} sm_100a kernels")
print("kernel".ljust(78), "total".rjust(7), *[w.rjust(8) for w in WANT])
for k, c in per.items():
    for w in WANT + ["_total"]:
        tot[w] += c[w]
    name = demangle(k)
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    print(name[:78].ljust(78), str(c["_total"]).rjust(7), *[str(c[w]).rjust(8) for w in WANT])
print("ALL".ljust(78), str(tot["_total"]).rjust(7), *[str(tot[w]).rjust(8) for w in WANT])
