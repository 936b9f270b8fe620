"""CPU oracle for the RBM CD-k hot path.  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the product package mdbn_b200/."""
