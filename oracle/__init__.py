"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's operator semantics (DataFusion 53.1 / arrow 58.1 as used by
lakehq/sail), pinned against the reference's own TPC-H golden snapshots.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import or execute
anything in this directory.  The product path (sail_b200/) never does and fails loudly without its
CUDA library.
"""
