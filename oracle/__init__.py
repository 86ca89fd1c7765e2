"""TEST INFRASTRUCTURE ONLY.

CPU restatements (oracles) of the reference algorithms on the hot path. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package — as the checker or the timed CPU baseline,
never as a fallback for the CUDA path.
"""
