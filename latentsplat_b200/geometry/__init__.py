"""Camera geometry of the hot path (mirror of /root/reference/src/geometry, GPU-friendly: no host syncs)."""
