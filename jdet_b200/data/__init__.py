"""jdet.data mirror: only what sits either side of the oriented-box hot path — the tile -> image merge."""
