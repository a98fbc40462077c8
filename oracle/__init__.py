"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (multi-modal-gnn_b200/) may import from here.  Allowed importers:
tests/, __graft_entry__.smoke() (as the checker) and bench.py's cpu_baseline / --impl reference legs.
"""
