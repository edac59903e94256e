"""Empty stand-in: the reference's normalising_flow.py / plotting_utils.py import matplotlib at module level
(plotting is out of scope); only the names touched at import time exist."""


class figure:  # noqa: N801  (matplotlib.figure.Figure appears in a dataclass annotation)
    class Figure:
        pass
