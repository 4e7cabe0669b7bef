"""CPU oracle (test infrastructure only). See oracle/hydro_oracle.h."""
