"""ORACLE -- CPU restatements of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import anything from this package; the product
(realtime-pose-estimation_b200/, import name ``rtpe_b200``) never does.
"""
