"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle_np.py / oracle.c headers).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.
"""
