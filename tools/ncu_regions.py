"""Instruction / stall-sample shares of the cluster kernel by code region (source line ranges of
orz_cluster_kernels.cuh found by their marker comments), from an ncu report with --import-source on.
usage: python tools/ncu_regions.py <report.ncu-rep> <kernel-substring>"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, kern = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "orz_cluster_kernels.cuh"
src = open(os.path.join(root, "rasterizer_b200", "csrc", SRC)).read().splitlines()
def line_of(marker):
    for i, l in enumerate(src):
        if marker in l: return i + 1
    raise SystemExit("marker not found: " + marker)
marks = [
    ("step_chain (iterated add chains)", "__device__ __forceinline__ void step_chain("),
    ("tile_prim: HiZ test + chain set-up", "__device__ __forceinline__ void tile_prim("),
    ("tile_prim: coverage (edge masks)", "// ---- coverage (Rasterizer.cpp:1155-1239)"),
    ("tile_prim: depth chains set-up", "const uint32_t updMask = "),
    ("tile_prim: depth rows + merge + HiZ", "// ---- depth rows, merge, HiZ (Rasterizer.cpp:1241-1290)"),
    ("kernel prologue (tables, clear)", "k_raster_views_cluster(const FrameParams p) {"),
    ("pre-announce loop", "// ---- candidates whose rectangle does not touch my tiles"),
    ("walk: slot bookkeeping", "  uint4 infoNext = recInfo[0]"),
    ("decision words (ld_flag / votes: polling)", "__device__ __forceinline__ uint32_t ld_flag("),
    ("gate test: one block (query_block_h)", "// one block of query2D (Rasterizer.cpp:305-343)"),
    ("gate test on my tiles", "// ---- gate: query2D (Rasterizer.cpp:283-349) on the part"),
    ("decision wait (spin)", "// visible as soon as ONE warp says so"),
    ("occluder prologue (info, box)", "// ---- rasterize<clipped>(occluder): the records k_setup_views wrote"),
    ("flush: gather + tile loop", "auto flush = [&]() {"),
    ("tile_hits / tile_loop (which record on which tile, tile order)", "__device__ __forceinline__ uint32_t tile_hits("),
    ("flush: tile open (load)", "// open the tile: its depth goes to shared memory"),
    ("flush: tile close (store)", "if (dirty) {  // close"),
    ("header scan + staging", "for (uint32_t r0 = 0; r0 < cnt; r0 += 32u) {"),
    ("epilogue (zero fill, final barrier, outputs)", "cluster.sync();  // no CTA may leave"),
]
bounds = sorted((line_of(m), name) for name, m in marks)
def region(f, n):
    if f != SRC: return "inlined helpers (" + f + ")"
    name = "setup kernel / helpers above step_chain (k_setup_views, query_block_h, decision words)"
    for ln, nm in bounds:
        if n >= ln: name = nm
    return name
so = os.path.join(root, "rasterizer_b200", "librasterizer_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("orz_kernels.sm")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
addr_line, cur, infn = {}, None, False
for l in dis:
    if l.startswith("//--------------------- .text."):
        infn = kern in l; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # innermost frame of an inlined chain names the helper; keep the OUTERMOST line of the cluster file for the region
        f, n, rest = os.path.basename(m.group(1)), int(m.group(2)), m.group(3)
        if "inlined at" in rest:
            mm = re.findall(r'"([^"]+)", line (\d+)', rest)
            outer = [(os.path.basename(a), int(b)) for a, b in mm if os.path.basename(a) == SRC]
            cur = outer[0] if (f != SRC and outer) else (f, n)
        else:
            cur = (f, n)
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', l)
    if m: addr_line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(collections.Counter); base = None
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    k = addr_line.get(a - base) or ("?", 0)
    g = agg[region(*k)]
    g["inst"] += int(r[ix["Instructions Executed"]] or 0)
    g["thr"] += int(r[ix["Predicated-On Thread Instructions Executed"]] or 0)
    g["smp"] += int(r[ix["# Samples"]] or 0)
ti = sum(g["inst"] for g in agg.values()); ts = sum(g["smp"] for g in agg.values())
print(f"{'region':60s} {'inst%':>7s} {'lanes':>6s} {'samples%':>9s}")
for name, g in sorted(agg.items(), key=lambda kv: -kv[1]["inst"]):
    print(f"{name[:60]:60s} {100*g['inst']/ti:7.2f} {g['thr']/max(1,g['inst']):6.1f} {100*g['smp']/max(1,ts):9.2f}")
print(f"total warp instructions {ti:,}, samples {ts:,}")
