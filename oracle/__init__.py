"""TEST INFRASTRUCTURE — CPU oracle of the VTMAE hot path (see oracle/vtmae_oracle.py).
Never imported by the product package m3l_b200/."""
