"""CPU oracle (test infrastructure only) -- see mel_oracle.py.  Never imported by the product."""
