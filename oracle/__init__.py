"""CPU oracle for the flex_provision hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU and in fp64, the algorithm of the reference's
batched-env hot path (kosmylo/Safe-MARL):

  * utils/create_net.py:8-38            -> oracle/ieee33.py
  * utils/pf.py:10-113, 115-192         -> oracle/pf_ref.py  (Newton + sweep)
  * madrl/environments/flex_provision/
      flexibility_provision_env.py      -> oracle/env_ref.py
  * safety_signal/*.py, safemaddpg.py   -> oracle/predictor_ref.py
  * utils/replay_buffer.py:3-30         -> oracle/replay_ref.py
  * a C mirror of the kernel's exact operation order (bit-exact masks and the
    CPU baseline)                       -> oracle/c/flex_oracle.c

PARITY: PINNED ON THE REFERENCE'S OWN PYTHON CODE, IPOPT EXCEPTED.  The
reference has no tests and no golden vectors for this path (its only would-be
vectors, data/net_power_inputs.csv and data/bus_voltages_outputs.csv, are
Git-LFS pointers), and its power-flow root finding lives in IPOPT (unpinned
version, driven through Pyomo==6.7.1), which is not installed here.  What CAN
be run here is all of the reference's Python: tests/golden/make_ref_golden.py
imports utils/create_net.py, utils/pf.py, utils/util.py::translate_action and
FlexibilityProvisionEnv from the reference checkout and executes them unchanged
on top of oracle/pyomo_shim.py, a stand-in for Pyomo whose "ipopt" is the
oracle's Newton solver followed by an evaluation of EVERY constraint rule of the
reference's model on the solution (max residual < 1e-12 in the fixtures).  The
committed fixtures tests/golden/ref_*.npz are therefore outputs of the
reference itself (network dict, power_flow_solver / _simplified results, two
full 95-step episodes of reset/step/get_obs/get_state, translate_action), and
tests/test_oracle_*.py check the restatements below against them (the env
restatement reproduces the reference's episodes to 1e-16).  Not pinned: which
root IPOPT lands on from its default start (the high-voltage root is assumed)
and its 1e-8 termination tolerance.  In addition the oracle is anchored on
  (K1) the literature IEEE 33-bus base case (Baran & Wu 1989):
       V_min = 0.913090 p.u. at bus 18, P_loss = 202.6771 kW, Q_loss = 135.1410 kvar,
  (K2) the code-defined operating point of run_pf.py:36-57,
  (K3) residuals of the equations utils/pf.py:65-98 evaluated on the outputs,
and on two independent solvers (dense Newton and backward/forward sweep) that
must agree to 1e-10.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import anything from here.  The product (safe-marl_b200/)
never does.
"""
