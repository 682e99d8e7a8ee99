"""Empty stand-in so the untouched reference's ``import matplotlib.pyplot`` succeeds (oracle only)."""
