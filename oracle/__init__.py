"""CPU oracle (test infrastructure only) -- see gccnmf_oracle.py for the rules."""
