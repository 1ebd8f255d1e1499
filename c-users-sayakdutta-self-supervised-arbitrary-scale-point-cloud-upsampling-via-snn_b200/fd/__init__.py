"""fd: distance estimation model (drop-in for the reference's `fd` package on the inference path)."""
