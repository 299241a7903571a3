"""Empty stand-in for matplotlib (rendering is never exercised by the oracle)."""
