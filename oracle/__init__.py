"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the reference's hot-path algorithms.

Nothing in the product package (e-osvos_b200/) may import this; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do, and only as the checker or the timed baseline.
"""
