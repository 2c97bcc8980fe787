"""CPU oracle (test infrastructure only; see oracle/tnqs_oracle.py header).  Never imported by the product."""
