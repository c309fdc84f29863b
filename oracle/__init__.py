"""CPU oracle for the ruzstd decode path -- TEST INFRASTRUCTURE ONLY (see ruzstd_oracle.h)."""
