"""CPU oracle for the DLA model-selection hot path. TEST INFRASTRUCTURE ONLY (see dla_oracle.py)."""
