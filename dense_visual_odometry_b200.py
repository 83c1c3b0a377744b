"""Import shim: the package lives in `dense-visual-odometry_b200/` (a directory name Python cannot
import directly).  `import dense_visual_odometry_b200` from the repo root loads that directory as a
regular package under this name."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "dense-visual-odometry_b200"
_spec = importlib.util.spec_from_file_location(
    __name__, _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
