"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference hot path (flatland_oracle.c) plus the harness that runs the
unmodified reference in the build container (ref_harness.py).  Only tests/, __graft_entry__.smoke()
and the cpu_baseline / --impl reference legs of bench.py may import this package; the product
package never does.
"""
