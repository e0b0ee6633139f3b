"""Test infrastructure only (CPU oracle for the GLL hot path). See oracle/gll_oracle.py."""
