"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/srt_oracle.cpp header).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
