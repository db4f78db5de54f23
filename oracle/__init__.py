"""CPU oracle (test infrastructure only) -- see oracle/spaa_oracle.py."""
