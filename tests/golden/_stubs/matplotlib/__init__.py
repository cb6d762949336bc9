"""Empty stand-in so the reference's workers.py imports in a container without matplotlib
(golden generation only; never on a product path)."""
def use(*args, **kwargs):
    return None
