"""TEST INFRASTRUCTURE: the oracle's local-regression model under the reference's module / class name (see
PredictiveControllers.py in this directory)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle.ltv_model import LocalLTVModel as PredictiveModel      # noqa: E402,F401
