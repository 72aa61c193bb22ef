"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference algorithm for the D-MPNN hot path.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package;
nothing under chemprop_b200/ does."""
