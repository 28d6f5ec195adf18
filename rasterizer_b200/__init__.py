"""B200-native occlusion-culling rasterizer (hot path of rawrunprotected/rasterizer)."""
