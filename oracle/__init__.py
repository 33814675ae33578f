"""oracle/ — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; discover_tfbs_b200/ never does (tests/test_no_oracle_in_product.py
enforces it).  See the header of oracle/tfbs_oracle.c for the parity status of every part.
"""
