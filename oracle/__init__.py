"""ORACLE -- CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product (goldfish_b200) never does.
PARITY UNPINNED: the reference's arithmetic lives in un-vendored, unpinned,
uninstalled dependencies and its tests assert nothing.
"""
