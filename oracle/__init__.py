"""CPU oracles for the putative-matching path.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py)."""
