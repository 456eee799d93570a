"""TEST INFRASTRUCTURE — CPU oracles for the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product package
(handwriting_line_generation_b200) never does.
"""
