"""TEST INFRASTRUCTURE ONLY: CPU oracle of the Lennard-Jones MD step (see oracle/ljmd_oracle.c)."""
