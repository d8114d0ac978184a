"""CPU restatement of the reference algorithm: TEST INFRASTRUCTURE ONLY (see pointflow_oracle.py)."""
