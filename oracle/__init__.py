"""CPU oracle (test infrastructure only; see oracle/goal_oracle.h)."""
