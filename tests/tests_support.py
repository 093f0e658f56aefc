"""Shared deterministic input recipes for tests (mirrors tests/golden/make_golden.py)."""
import importlib.util
import os

_here = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("_make_golden", os.path.join(_here, "golden", "make_golden.py"))
_mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mg)

dcn_cases = _mg.dcn_cases
dcn_inputs = _mg.dcn_inputs
SEED = _mg.SEED
