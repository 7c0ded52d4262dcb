"""CPU oracle for lvc_b200 -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
