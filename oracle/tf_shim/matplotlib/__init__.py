"""Empty stand-in: the reference's normalising_flow.py imports pyplot at module level (plotting is out of scope)."""
