"""CPU oracle (test infrastructure only - see mpc_oracle.py header)."""
