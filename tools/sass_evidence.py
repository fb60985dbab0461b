"""Instruction-mnemonic counts per kernel of libuwr_b200.so (cuobjdump -sass), the 'what proves a Blackwell-native
kernel' table of /opt/skills/guides/B200_PROFILING.md.  usage: python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "underwater-image-restoration_b200", "csrc", "libuwr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cols = [("UTC*MMA", r"^UTC\w*MMA"), ("LDTM", r"^LDTM"), ("UTMALDG", r"^UTMALDG"), ("SYNCS", r"^SYNCS"), ("UTCBAR", r"^UTCBAR"),
        ("HMMA", r"^HMMA"), ("LDGSTS", r"^LDGSTS"), ("FFMA2", r"^FFMA2|^FMUL2|^FADD2")]
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        for label, pat in cols:
            if re.match(pat, m.group(1)):
                counts[name][label] += 1
names = list(counts)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
def short(d):
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    return re.sub(r"\(.*$", "", d)
print("# SASS evidence (cuobjdump -sass underwater-image-restoration_b200/csrc/libuwr_b200.so, sm_100a), instruction counts per kernel.")
print("# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA (cp.async.bulk.tensor) -> UTMALDG/UTMASTG, mbarrier -> SYNCS,")
print("# tcgen05.commit -> UTCBAR, legacy mma.sync -> HMMA, cp.async -> LDGSTS, packed fp32 -> FFMA2/FMUL2/FADD2")
print("# gemm_tcgen05_kernel<BN, LAY (0 NT, 1 NN, 2 TN), EPI, HF (fp16 C / R), CL (cluster size), CV (virtual im2col operand)>")
print(f"{'kernel':78s} " + " ".join(f"{c[0]:>7s}" for c in cols))
rows = sorted((short(d), counts[n]) for n, d in zip(names, dem))
for k, c in rows:
    if any(c.values()):
        print(f"{k[:78]:78s} " + " ".join(f"{c[l[0]]:7d}" for l in cols))
