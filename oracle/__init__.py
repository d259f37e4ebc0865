"""CPU oracle for the LMPC finite-time optimal control hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product path
(``racinglmpc_b200``) never imports this package and fails loudly when its CUDA
library is missing.

What is restated here (fp64, NumPy + one C file), with the reference file:line
each piece follows:

* ``track.py``      – ``Map.__init__`` table + ``Map.curvature``
                      (src/fnc/simulator/Track.py:10-133, 292-310)
* ``vehicle.py``    – ``Simulator.sim/dynModel``, ``PID``, ``Regression``
                      (src/fnc/simulator/SysModel.py:22-147, src/fnc/Utilities.py:5-68);
                      only used to *generate* benchmark/test inputs
* ``ltv_model.py``  – ``PredictiveModel`` (src/fnc/controller/PredictiveModel.py:11-197)
* ``ftocp.py``      – ``MPC`` / ``LMPC`` matrix assembly and controller state
                      machine (src/fnc/controller/PredictiveControllers.py:56-514)
* ``osqp_port.c``   – the OSQP algorithm (Stellato et al., Math. Prog. Comp. 2020)
                      restated from the paper; the ``osqp`` wheel is an un-vendored,
                      unpinned pip dependency of the reference (README.md:18) and is
                      not installable here
* ``kkt.py``        – solver-independent fp64 KKT residual checker

Parity status: the assembly / regression / safe-set pieces are pinned against
the *real* reference modules imported in the build container (golden vectors in
``tests/golden``, generator ``tests/golden/make_golden.py``).  The QP solver is
"parity unpinned" against the real OSQP binary (absent); it is pinned instead by
the KKT checker and by agreement between two independent algorithms (this ADMM
port and the primal-dual interior-point model in ``pdip_model.py``).
"""
