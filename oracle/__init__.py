"""Parity checker for cdae_b200 — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  See oracle/cdae_oracle.h.
"""
