"""Import shim: the product package lives in the directory `particleincellcodegolf.jl_b200/`
(a dotted directory name cannot be imported directly), so `import particleincellcodegolf.jl_b200`
is wired to it here."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "particleincellcodegolf.jl_b200")
_name = __name__ + ".jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    jl_b200 = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = jl_b200
    _spec.loader.exec_module(jl_b200)
else:
    jl_b200 = sys.modules[_name]
