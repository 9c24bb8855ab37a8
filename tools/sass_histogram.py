"""SASS evidence for the built library: per kernel, the count of the Blackwell-native opcodes (B200_PROFILING.md:
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit/alloc -> UTCBAR/UTCATOM...,
mbarrier -> SYNCS) next to the legacy tensor path (HMMA) that must NOT appear.  Runs without a GPU:
    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200", "lib", "libnfdpm_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "HGMMA",
         "LDGSTS", "FFMA", "MUFU"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = {}
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        hist[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                hist[cur][w] += 1
names = list(hist)
dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a), {len(names)} kernels; opcode counts per kernel")
print(f"# {'kernel':<100} " + " ".join(f"{w:>8}" for w in ["instr"] + WATCH))
tot = collections.Counter()
for n, d in zip(names, dem):
    h = hist[n]
    i = d.find(">(")                                   # keep the template arguments, drop the parameter list
    short = (d[:i + 1] if i >= 0 else d.split("(")[0]).replace("void nfdpm::", "").replace("nfdpm::", "")
    print(f"{short[:100]:<102} " + " ".join(f"{h[w]:>8}" for w in ["_total"] + WATCH))
    tot.update(h)
print(f"{'TOTAL':<102} " + " ".join(f"{tot[w]:>8}" for w in ["_total"] + WATCH))
assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor-core opcodes present"
