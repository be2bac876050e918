"""CPU oracle for the code2vec path-attention hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  code2vec_b200/ never does.
"""
