"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements (numpy / torch fp32) of the reference's algorithm for the image -> LLM-embedding
path, each function citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package, and only as the checker
(or the timed CPU baseline) -- never as part of the product path in vision-zephyr_b200/.

Pinning: the reference ships no tests or golden vectors of its own, so the
oracle is pinned against outputs of the reference ITSELF, generated in the build container by
oracle/gen_golden.py (which imports /root/reference) and committed under tests/golden/, plus
bit-exact checks against Pillow for the integer image operations.
"""
