"""Per-kernel counts of the Blackwell opcodes in libvidchap.so (cuobjdump -sass):  python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vidchapters_b200", "libvidchap.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = {}
cur = None
counts = collections.OrderedDict()
want = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|LDTM|STTM|UTCATOMSWS|UTCBAR|UTMALDG|UTMASTG|UTMAREDG|UTMACCTL|UTMAPF|SYNCS|HMMA|LDSM|MUFU|ACQBULK|UCGABAR_ARV|MEMBAR)\b(\.2CTA)?")
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur:
        m = want.search(line)
        if m and "/*" in line:
            counts[cur][m.group(1) + (m.group(2) or "")] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("""# SASS evidence of the Blackwell-native path: `cuobjdump -sass vidchapters_b200/libvidchap.so` (built by __graft_entry__.build(),
# nvcc 12.9 -gencode arch=compute_100a,code=sm_100a), opcode counts per kernel (tools/sass_opcodes.py).  tcgen05.mma -> UTCHMMA
# (.2CTA = cta_group::2), tcgen05.ld/st -> LDTM/STTM, tcgen05.alloc/commit -> UTCATOMSWS/UTCBAR, TMA loads/stores/reduce-adds ->
# UTMALDG/UTMASTG/UTMAREDG, mbarrier -> SYNCS.  HMMA/LDSM (mma.sync) appear ONLY in decode_linear_kernel (the M <= 64-row decode
# linears, DESIGN.md §4).
""")
for (k, c), d in zip(counts.items(), dem):
    if c:
        print(d[:150]); print("    " + "  ".join(f"{o}:{n}" for o, n in sorted(c.items())))
