"""CPU oracle — TEST INFRASTRUCTURE ONLY (see mha_oracle.py).  Never imported by perceiver_io_b200."""
