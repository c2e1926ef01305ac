"""Mirror of the reference's ``evaluation`` package (evaluation/criteria_new.py) on the B200 kernels."""
