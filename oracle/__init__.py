"""CPU oracle (test infrastructure only) — see protograd_oracle.py."""
