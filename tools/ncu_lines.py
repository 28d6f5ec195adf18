"""Join an ncu report's per-SASS-instruction counters with nvdisasm -g line info and print the
hottest source lines (instructions executed, stall samples by reason).
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [top]"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "rasterizer_b200", "librasterizer_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("orz_kernels.sm")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# locate function
line_of = {}
cur = None
infn = False
for l in dis:
    if l.startswith("//--------------------- .text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
kfilter = os.environ.get("NCU_KERNEL")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + kfilter] if kfilter else []), capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
base = None
tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    key = line_of.get(addr - base)
    key = (key[0], key[1]) if key else ("?", 0)
    a = agg[key]
    a["inst"] += int(r[ix["Instructions Executed"]] or 0)
    a["samples"] += int(r[ix["# Samples"]] or 0)
    for s in stall_cols:
        a[s] += int(r[ix[s]] or 0)
    a["n_sass"] += 1
for k, a in agg.items():
    for kk, v in a.items():
        tot[kk] += v
src = {}
def srcline(f, n):
    p = os.path.join(root, "rasterizer_b200", "csrc", f)
    if p not in src:
        try: src[p] = open(p).read().splitlines()
        except OSError: src[p] = []
    return src[p][n - 1].strip()[:90] if 0 < n <= len(src[p]) else ""
print(f"total inst {tot['inst']:,}  samples {tot['samples']:,}")
print("stalls:", ", ".join(f"{s[6:]}={tot[s]*100//max(1,tot['samples'])}%" for s in sorted(stall_cols, key=lambda s: -tot[s])[:9]))
print(f"{'file:line':22s} {'inst%':>6s} {'smp%':>6s} {'sass':>5s}  top stalls / source")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(stall_cols, key=lambda s: -a[s])[:3]
    sts = " ".join(f"{s[6:]}:{a[s]*100//max(1,a['samples'])}" for s in st if a[s])
    print(f"{k[0]+':'+str(k[1]):22s} {a['inst']*100/max(1,tot['inst']):6.2f} {a['samples']*100/max(1,tot['samples']):6.2f} {a['n_sass']:5d}  [{sts}] {srcline(*k)}")
