"""Import shim: the package directory is named `davinci-node_b200/` (not a valid Python
identifier), so `import davinci_node_b200` resolves to it through this module."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "davinci-node_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
