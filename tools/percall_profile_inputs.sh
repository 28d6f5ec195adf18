set -e
cd /root/repo
python - <<'P'
import os, numpy as np, subprocess, sys
sys.path.insert(0, '.')
from rasterizer_b200 import workloads as wl, camera as cam
from oracle import port_oracle as po
ps = wl.load_scene("castle"); w, h = 1920, 1080
po.set_tables()
baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
c = ps.camera
mvp = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)
order = cam.front_to_back_order(np.stack([b[1] for b in baked]), c["pos"])
mvp.tofile("/tmp/mvp.bin"); order.tofile("/tmp/order.bin")
print(wl.prepared_path("castle"))
P
