"""TEST INFRASTRUCTURE: CPU oracle for the AFEC low-level descriptor path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (afec_b200) never does.
"""
