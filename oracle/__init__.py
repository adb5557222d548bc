"""CPU oracle for the pyfastx hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package, and only as the checker.  Nothing under pyfastx_b200/ imports it.
"""
