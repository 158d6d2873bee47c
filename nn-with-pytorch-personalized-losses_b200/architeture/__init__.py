"""Top-level drop-in for the reference's `architeture` package (only the ranker with a fused B200 path)."""
