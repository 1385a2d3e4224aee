"""Test-infrastructure oracle (see cldm_oracle.py header). Never imported by the product."""
