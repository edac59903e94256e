"""Empty stand-in: imported at module level by the reference's plotting_utils.py (plotting is out of scope)."""
