"""CPU oracle (test infrastructure only; see outgrid_oracle.py)."""
