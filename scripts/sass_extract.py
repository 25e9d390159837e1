"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the built library (no GPU needed):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), LDGSTS (cp.async), REDG / RED (global reductions).
usage: python scripts/sass_extract.py [lib.so] > profiles/rNN_sass_extract.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "optical-flow-dnn_b200", "libraftcorr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "LDGSTS", "REDG", "RED."]
rows, cur, it = [], None, iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = collections.Counter(); rows.append((next(it), cur)); continue
    if cur is None: continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        cur["_total"] += 1
        for w in WANT:
            if m.group(1).startswith(w): cur[w.rstrip(".")] += 1
cols = [w.rstrip(".") for w in WANT]
print(f"{'kernel':72s} {'SASS':>6s} " + " ".join(f"{c:>8s}" for c in cols))
tot = collections.Counter()
for name, c in sorted(rows, key=lambda r: r[0]):
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).replace("raftcorr::", "").replace("void ", "")
    print(f"{short[:72]:72s} {c['_total']:6d} " + " ".join(f"{c[k]:8d}" for k in cols))
    tot.update(c)
print(f"{'TOTAL (' + str(len(rows)) + ' kernels)':72s} {tot['_total']:6d} " + " ".join(f"{tot[k]:8d}" for k in cols))
