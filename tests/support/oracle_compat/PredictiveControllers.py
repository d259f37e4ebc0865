"""TEST INFRASTRUCTURE: the oracle's controllers under the reference's module / class names, so that the reference's own
`main.py` can be run against the ORACLE (CPU, reference arithmetic + OSQP-algorithm port at 1e-9) exactly the way it is run
against racinglmpc_b200/compat — the two printouts are then compared (tools/run_reference_main.py)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle import ftocp, osqp_port      # noqa: E402

MPCParams = ftocp.FTOCPParams


class MPC(ftocp.OracleMPC):
    def __init__(self, mpcParameters, predictiveModel=[]):
        super().__init__(mpcParameters, predictiveModel if predictiveModel != [] else None, qp=osqp_port.tight_qp)


class LMPC(ftocp.OracleLMPC):
    def __init__(self, numSS_Points, numSS_it, QterminalSlack, mpcPrameters, predictiveModel, dt=0.1):
        super().__init__(numSS_Points, numSS_it, QterminalSlack, mpcPrameters, predictiveModel, qp=osqp_port.tight_qp)
