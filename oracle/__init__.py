"""TEST INFRASTRUCTURE ONLY — CPU parity oracle for the render hot path.

Nothing under minecraftskin_raytracer_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
