"""Empty stand-in for matplotlib.pyplot (golden generation only)."""
