"""CPU oracle for the MEMFlow phase-space hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. memflow_b200/ (the product) never does.
"""
