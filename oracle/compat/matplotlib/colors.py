"""Stub: the oracle never plots (reference ``plot=False`` paths only)."""
