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

PARITY UNPINNED.  The reference has no tests and no golden vectors for this
path (its only would-be vectors, data/net_power_inputs.csv and
data/bus_voltages_outputs.csv, are Git-LFS pointers), and its power-flow
arithmetic lives in IPOPT (unpinned version, driven through Pyomo==6.7.1),
which is not installed here.  The oracle is therefore anchored on
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
