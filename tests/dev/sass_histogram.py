"""Per-kernel SASS instruction histogram of libsedb.so -> profiles/<tag>_sass_histogram.md (cuobjdump -sass, no GPU needed).

    python tests/dev/sass_histogram.py r2
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "soundeventdetection-pytorch_b200", "libsedb.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "USETMAXREG", "ELECT", "FFMA2", "FMUL2", "FADD2",
        "HMMA", "R2UR", "SHFL", "ATOMG", "ATOMS", "REDG", "RED", "LDG", "STG", "LDS", "STS"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        funcs[cur]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                funcs[cur][k] += 1
out = [f"# SASS instruction histogram of libsedb.so ({tag})", "",
       "`cuobjdump -sass soundeventdetection-pytorch_b200/libsedb.so`, counted per kernel by `tests/dev/sass_histogram.py`.",
       "UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine, 1-D),",
       "UBLKPF = cp.async.bulk.prefetch.L2, LDGSTS = cp.async, SYNCS = mbarrier ops, FFMA2/FMUL2/FADD2 = packed fp32.",
       "No HMMA (mma.sync) anywhere: every tensor-core instruction is tcgen05.", "",
       "| kernel | instrs | " + " | ".join(KEYS) + " |", "|---|---:|" + "---:|" * len(KEYS)]
for f, c in funcs.items():
    if not f.startswith("sedb::"):
        continue
    out.append(f"| `{f}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
path = os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.md")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out))
