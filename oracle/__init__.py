"""CPU oracle for the actor-learner hot path (TEST INFRASTRUCTURE — not product code).

Everything under ``oracle/`` is a CPU restatement (numpy / torch-CPU) of the
reference algorithms on the path SURVEY.md §8 names.  It exists only so that
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg can check and time the reference semantics on the host.
Nothing in ``parl_b200/`` imports it: the product path fails loudly when the
CUDA library is missing.

Pinning status:
  * V-trace: pinned by the reference's own known-answer test
    (parl/algorithms/paddle/impala/tests/vtrace_test_paddle.py:33-144),
    vectors committed in tests/golden/vtrace_kat.npz.
  * GAE / A2C / PPO / DQN / DDQN / PG losses, RolloutStorage.compute_returns,
    ReplayMemory, SumTree / ProportionalPER: the reference holds no fixtures;
    pinned by outputs of the reference itself imported in the build container
    (tests/golden/make_golden.py -> tests/golden/*.npz).
  * Synthetic envs / Philox RNG / CartPole physics: no reference code to pin
    against (reference mock gym uses unseeded numpy RNG; gym physics is third
    party and absent) -> "parity unpinned"; Philox itself is pinned by the
    Random123 known-answer vectors.
"""
