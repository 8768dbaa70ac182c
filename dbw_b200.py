"""Import alias: the product package lives in `differentiable-blocksworld_b200/` (a directory name Python cannot
import directly because of the hyphen).  `import dbw_b200` loads that directory as the package `dbw_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'differentiable-blocksworld_b200')
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, '__init__.py'),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
