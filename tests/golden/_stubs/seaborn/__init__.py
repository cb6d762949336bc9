"""Empty stand-in so the reference's workers.py imports in a container without seaborn
(golden generation only; never on a product path)."""
def heatmap(*args, **kwargs):
    return None
