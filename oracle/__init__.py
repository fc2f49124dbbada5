"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference hot path (DidItWork/Sigma-Zero self-play MCTS) used as the
checker for the CUDA product in `sigma-zero_b200/`.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import anything from here.
"""
