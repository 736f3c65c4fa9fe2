"""CPU oracle (test infrastructure only; see pcaa_oracle.py header)."""
