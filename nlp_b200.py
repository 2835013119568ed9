"""Import shim: the package directory is named after the reference
(``neighborhood-link-prediction-openmp_b200``), which is not a Python identifier."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neighborhood-link-prediction-openmp_b200")
_spec = importlib.util.spec_from_file_location("nlp_b200_pkg", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["nlp_b200_pkg"] = _pkg
_spec.loader.exec_module(_pkg)

build = _pkg.build
binding = _pkg.binding
graphs = _pkg.graphs
distributed = _pkg.distributed
Predictor = _pkg.Predictor
MEASURES = _pkg.MEASURES
UNBOUNDED = _pkg.UNBOUNDED
NlpError = _pkg.NlpError
