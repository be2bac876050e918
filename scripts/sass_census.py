"""profiles/r2_sass_census.txt: per-kernel SASS opcode census of the shipped libc2v_b200.so (cuobjdump -sass), so the
tensor-core / TMA / multicast instructions can be checked without rebuilding: UTCHMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UBLKCP (cp.async.bulk), LDGSTS (cp.async), LDGMC / STG...MC (multimem), SYNCS (mbarrier), REDUX ..."""
import collections, os, re, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(R, "code2vec_b200", "libc2v_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = re.compile(r"^(UTC\w*MMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|LDGSTS|LDGMC|STGMC|REDGMC|SYNCS|ARRIVES|USETMAXREG|REDUX|ATOMG|REDG|RED|ATOMS|MUFU|HMMA|FFMA|LDG|STG|LDS|STS|SHFL|BAR|NANOSLEEP|ACQBULK|UCGABAR)")
fn, census = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", name)
        if "<" in name:
            fn = re.sub(r"\(.*", "", name.split("(")[0]) if "(" in name else name
        census[fn] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_x]+)*)", line)
    if m and fn:
        op = m.group(1)
        census[fn]["_total"] += 1
        if KEY.match(op):
            full = op + (m.group(2) if op in ("LDGMC", "STG", "LDGSTS", "UBLKCP", "MUFU", "REDG", "ATOMG") else "")
            census[fn][full if op != "STG" or ".MC" in m.group(2) or "MMIO" in m.group(2) else op] += 1
print(f"# SASS opcode census of {os.path.basename(lib)} (cuobjdump -sass; counts are static instruction counts per kernel)")
for f, c in census.items():
    if c["_total"] < 50:
        continue
    items = ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda kv: (-kv[1], kv[0])) if k != "_total")
    print(f"{f}\n    total {c['_total']}: {items}")
