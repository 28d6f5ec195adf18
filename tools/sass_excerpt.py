"""SASS evidence for the lowering the design relies on: per kernel, how often the mnemonics that matter occur
(packed u16 min / max, packed f32 add, cp.async = LDGSTS, bulk copy = UBLKCP + SYNCS mbarrier ops, cluster barrier,
DSMEM mapa, shared-memory atomic min), plus the update pass of the cluster kernel (the loop around ATOMS.MIN) in full.
usage: python tools/sass_excerpt.py > profiles/r2_cluster_sass_excerpt.txt"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "rasterizer_b200", "librasterizer_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ("VIMNMX.U16x2", "VIMNMX3.U16x2", "FADD2", "LDGSTS", "UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK", "UCGABAR", "MAPA", "ATOMS.MIN",
        "I2IP.U16.S32.SAT", "PRMT", "REDUX", "VOTE", "STS.128", "LDS.128")
WANT = ("k_raster_views_cluster", "k_raster_tiles", "k_rasterize_call", "k_query_chain", "k_setup_views", "k_query_views")
parts = re.split(r"\n\s*Function : ", txt)
print("cuobjdump -sass rasterizer_b200/librasterizer_b200.so (sm_100a), mnemonic counts per kernel\n")
shown = False
for part in parts[1:]:
    name = part.split("\n", 1)[0].strip()
    if not any(w in name for w in WANT):
        continue
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", part, re.M)
    counts = {k: sum(1 for o in ops if o.startswith(k)) for k in KEYS}
    print(f"{dem}\n   {len(ops)} instructions; " + ", ".join(f"{k} {v}" for k, v in counts.items() if v))
    if "k_raster_views_cluster<(int)2, (unsigned int)4>" in dem:
        lines = part.split("\n")
        idx = [i for i, l in enumerate(lines) if "ATOMS.MIN" in l]
        if idx and not shown:
            shown = True
            lo, hi = max(0, idx[0] - 95), min(len(lines), idx[0] + 12)
            keep = [l for l in lines[lo:hi] if re.match(r"^\s+/\*[0-9a-f]{4}\*/", l)]
            print("\n   --- update pass (eight lanes per block: LDS.128 of the item, packed-u16 build, VIMNMX.U16x2 merge, STS.128, ATOMS.MIN HiZ) ---")
            for l in keep:
                print("   " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip())
            print()
