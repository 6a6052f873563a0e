"""CPU oracle of beta9's task fan-out hot path (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (beta9_b200/) never does.
"""
