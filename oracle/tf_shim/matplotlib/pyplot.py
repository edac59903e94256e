"""Empty stand-in for matplotlib.pyplot (never called by the golden-vector generator)."""
