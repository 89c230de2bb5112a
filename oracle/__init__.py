"""CPU parity oracle package (test infrastructure only; see splendor_oracle.c)."""
