"""profiles/traffic.json from an ncu per-launch CSV (--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum ...): DRAM bytes of ONE bench step = the launches between two k_prepare_views."""
import csv, json, sys, collections
src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
launch = collections.OrderedDict()
for r in rows:
    launch.setdefault(int(r[0]), {"kernel": r[4]})[r[-3]] = float(r[-1])
ids = [i for i, l in launch.items() if "k_prepare_views" in l["kernel"]]
start = ids[0]
step = [l for i, l in launch.items() if i >= start and (len(ids) < 2 or i < ids[1]) and any(k in l["kernel"] for k in ("k_prepare_views", "k_sort_views", "k_render_views", "k_query_views", "k_setup_views", "k_raster_views_cluster"))]
tot = lambda k: sum(l.get(k, 0.0) for l in step)
out = {
    "source": src, "launches_in_step": len(step),
    "dram_bytes_read": tot("dram__bytes_read.sum"), "dram_bytes_write": tot("dram__bytes_write.sum"),
    "dram_bytes_per_launch": tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum"),
    "note": "sum over the launches of one 1024-view step (k_prepare_views, k_sort_views, k_setup_views, k_raster_views_cluster, k_query_views), ncu per-launch counters",
    "per_kernel": [{"kernel": l["kernel"][:60], "ms": l.get("gpu__time_duration.sum", 0) / 1e6,
                    "dram_read": l.get("dram__bytes_read.sum"), "dram_write": l.get("dram__bytes_write.sum"),
                    "warp_inst": l.get("smsp__inst_executed.sum"), "issue_active_pct": l.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "warps_active_pct": l.get("sm__warps_active.avg.pct_of_peak_sustained_active")} for l in step],
}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({k: out[k] for k in ("launches_in_step", "dram_bytes_read", "dram_bytes_write", "dram_bytes_per_launch")}))
