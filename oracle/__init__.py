"""TEST INFRASTRUCTURE ONLY: CPU oracles restating the reference hot path (see each module's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this package.
"""
