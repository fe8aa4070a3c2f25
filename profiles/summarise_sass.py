"""SASS opcode summary of the built library (works without a GPU): python profiles/summarise_sass.py > profiles/sass_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "multimodal_dataset_distillation_b200", "libvldd_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
WANT = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "SYNCS", "ELECT", "BRA.U.ANY", "ATOMG", "REDG", "MUFU.EX2"]
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + "."):
                per[cur][w] += 1
def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"# SASS opcode summary of multimodal_dataset_distillation_b200/libvldd_b200.so (cuobjdump -sass, sm_100a), round 2, final state")
print("# tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, TMA tensor loads -> UTMALDG, TMA tensor stores -> UTMASTG, cp.async.bulk -> UBLKCP,")
print("# tcgen05.ld/st -> LDTM/STTM, mbarrier -> SYNCS.*, elect.sync -> ELECT, divergence waterfall around uniform-datapath instructions ->")
print("# BRA.U.ANY (0 in every tc_gemm_kernel instance)")
print("total:", dict(tot))
hot = [(n, c) for n, c in per.items() if any(c[w] for w in ("UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP"))]
print(f"functions: {len(per)}; with tensor-core or bulk-copy instructions: {len(hot)}")
for n, c in hot:
    d = re.sub(r"\(.*", "", demangle(n))
    print(f"{d[:110]:110s} " + " ".join(f"{w}={c[w]}" for w in WANT if c[w]))
