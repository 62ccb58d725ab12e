"""Import shim: the product package lives in ``gnn-recommendations_b200/`` (the
directory name the build contract fixes); a hyphen is not importable, so this
module points its ``__path__`` at that directory and executes its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "gnn-recommendations_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), __file__, "exec"), globals())
