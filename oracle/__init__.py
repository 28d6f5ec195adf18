"""TEST INFRASTRUCTURE ONLY: CPU oracles for the occlusion-culling hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (rasterizer_b200) never does.
"""
