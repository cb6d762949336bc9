"""CPU oracle package -- test infrastructure only (see oracle/gmz_oracle.c header)."""
