"""Import alias: ``import oak_b200`` resolves to ``orthogonal-additive-gaussian-processes_b200/``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "orthogonal-additive-gaussian-processes_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
