"""CPU oracle for the wass_stereo hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  Nothing under wass_b200/ imports it.
"""
