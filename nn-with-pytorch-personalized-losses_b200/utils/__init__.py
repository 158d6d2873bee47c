"""Top-level drop-in for the reference's `utils` package (only the hot-path module `metrics`)."""
