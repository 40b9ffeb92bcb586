"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the BESO denoiser hot path (see oracle/beso_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  Nothing under beso_b200/ does.
"""
