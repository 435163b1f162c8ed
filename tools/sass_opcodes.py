"""Per-kernel SASS opcode census of libmvlt_b200.so (cuobjdump -sass): which kernels carry tcgen05 / TMEM / TMA instructions.
   python tools/sass_opcodes.py > profiles/sass_opcodes_rNN.txt     (runs without a GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "medical_vision_langauge_transformer_b200", "libmvlt_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTMACCTL", "SYNCS", "HMMA", "LDGSTS",
         "MUFU", "FFMA2", "ERRBAR", "UCGABAR", "MAPA", "ELECT"]
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); total[kern] = 0; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if kern and m:
        op = m.group(1); total[kern] += 1
        for w in WATCH:
            if op.startswith(w): counts[kern][w] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode census of {os.path.basename(lib)} (sm_100a), `cuobjdump -sass`; columns: instructions, then watched opcode counts")
print("# UTCHMMA = tcgen05.mma (bf16), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG / UTMAREDG = TMA tensor")
print("# load / store / reduce, UTMAPF = TMA prefetch to L2, SYNCS = mbarrier ops, HMMA = mma.sync (legacy warp-level tensor op)")
for k, name in zip(counts, demangle):
    short = re.sub(r"\(.*", "", name)
    c = counts[k]
    print(f"{short:90s} {total[k]:6d}  " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
