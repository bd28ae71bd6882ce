"""TEST INFRASTRUCTURE ONLY — see oracle/oracle.c.  Nothing under ceng795_b200/
may import this package; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do."""
