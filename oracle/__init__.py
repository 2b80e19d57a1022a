"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's arithmetic for the hot path.

Nothing under oracle/ may be imported by the product (octave_b200/, architectures/); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
"""
