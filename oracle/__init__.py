"""CPU oracle for the flocking-GNN rollout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline legs do.

Parity status
-------------
* learner side (``oracle.learner``): PINNED against the reference's own
  ``learner/actor.py`` + ``learner/state_with_delay.py`` + the shipped
  checkpoint via ``tests/golden/*.npz`` (made by ``oracle/gen_golden.py``).
* env side (``oracle.flock_env``): **parity unpinned** -- the arithmetic lives
  in the third-party ``gym_flock`` package (github.com/katetolstaya/gym-flock,
  no version pinned by the reference, not vendored, not installed here).  The
  restatement follows the published algorithm as recorded in SURVEY.md
  Appendix B and is anchored on the reference call sites
  (``learner/state_with_delay.py:22-26``, ``learner/gnn_dagger.py:150-163``).
"""
