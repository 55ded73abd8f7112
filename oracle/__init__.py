"""CPU oracle (test infrastructure only).  See oracle/mobilevit_oracle.c."""
