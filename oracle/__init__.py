"""TEST INFRASTRUCTURE.  CPU oracle for the patch-parallel UNet hot path of mit-han-lab/distrifuser.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product package `distrifuser_b200` never does (tests/test_boundary.py greps for it).
See oracle/README.md for what is pinned against the real reference and what is not."""
