"""SASS opcode histogram per kernel of libupgpt_b200.so (cuobjdump -sass): the evidence that the contractions are tcgen05 / TMEM / TMA
(UTC*MMA, LDTM / STTM, UTMALDG / UTMASTG / UBLKCP) and that no legacy tensor path (HMMA) is present. Runs without a GPU.
usage: python tools/sass_histogram.py [lib.so] > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "upgpt_b200", "_lib", "libupgpt_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "UTCBAR", "UTCATOM", "SYNCS", "HMMA", "HGMMA", "LDGSTS",
       "MUFU.EX2", "MUFU.RCP", "ACQBULK", "UCGABAR", "CGABAR", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "BAR", "FFMA", "HFMA2", "F2FP", "SHFL")
kern, hist = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["_total"] += 1
        for k in KEY:
            if op == k or op.startswith(k + "."):
                hist[kern][k] += 1
print("SASS opcode counts per kernel of %s (sm_100a)" % os.path.relpath(lib, ROOT))
print("tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA tensor loads/stores -> UTMALDG/UTMASTG, 1-D bulk copies -> UBLKCP; HMMA would be a legacy mma.sync path\n")
for k in sorted(hist, key=lambda n: -hist[n]["_total"]):
    c = hist[k]
    cols = ", ".join("%s %d" % (n, c[n]) for n in KEY if c[n])
    print("%-40s %6d instr: %s" % (k[:40], c["_total"], cols))
