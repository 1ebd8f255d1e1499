"""fn: surface-normal estimation model (drop-in for the reference's `fn` package on the inference path)."""
