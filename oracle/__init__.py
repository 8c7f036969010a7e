"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/ss_oracle.h)."""
