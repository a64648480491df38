"""CPU oracles for the PanTax hot path (test infrastructure only; see pantax_oracle.py)."""
