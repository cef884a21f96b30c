"""CPU oracle = test infrastructure.  Import only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs (never from vitmarl_b200)."""
