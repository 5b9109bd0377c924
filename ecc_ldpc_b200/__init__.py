"""Import alias: the package lives in the directory `ecc-ldpc_b200/` (not a valid Python identifier),
so `import ecc_ldpc_b200` forwards there."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ecc-ldpc_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
