"""CPU oracle (test infrastructure only -- never imported by kgwas_b200)."""
