"""CPU oracle (test infrastructure only — see ca_oracle.cpp header).  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import this package."""
